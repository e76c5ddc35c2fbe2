// One-pass fused update of one view: F step AND G step with a SINGLE read of X per update-iteration.
//
// Why: the two streaming kernels of rn_kernels.cuh each read X once (B_alg = 16 n p bytes per iteration) and run
// at 0.85-0.92 of the HBM roofline, so the only way left to go faster is to move fewer bytes.  The G step needs
// T = X' F_new, and F_new[r,] depends on the WHOLE row r of X (P = X G) -- but only on that row.  So a group of
// 8 rows that is resident on chip can do both: P for its rows, the F update (update_f, R/update_steps.r:141-165),
// and then its contribution X[rows,]' F_new[rows,] to T (update_g, :180-207), before the rows are dropped.
// 8 rows x p columns of FP64 do not fit one SM (256 KB at p = 4000), so a CLUSTER of C = 1..8 CTAs splits the
// columns (1008 per CTA) and exchanges the 8 x 8 partial of P through distributed shared memory.
//
//   X8 layout (HBM): X8[row group (8 rows)][column pair q][8 x 16 B]; the 16 B piece of row r holds
//   (X[r][2q], X[r][2q+1]) and sits at piece position r ^ 2(q & 3).  A warp's share of a group (64 column pairs =
//   8 KB) is one contiguous run = one bulk copy, and a verbatim copy in shared memory is bank-conflict-free for the
//   LDS.128 fragment reads of both phases (each LDS.128 feeds two m8n8k4 MMAs).
//   Roles per CTA (384 threads = 12 warps).  Warps 3, 7 and 11 -- all on SM sub-partition 3 -- are the producer warp
//   (bulk copies into a ring of 3 row groups x 63 KB: a warp's slot is read twice, F phase and G phase, and then
//   released) and the epilogue warp (sums the 9 warp partials of P, exchanges the CTA partial with the cluster
//   peers by st.async + mbarrier complete_tx, runs the F update for the 8 rows redundantly in every CTA with the
//   8 x 8 products as DMMAs, writes F_new to shared memory for the G phase and -- rank 0 -- to HBM, accumulates F'F
//   and colSums(F)) and the auxiliary warp (everything the F update needs that does not depend on P: it prefetches
//   the old F rows and, for phi-coupled views, forms the coupling sum, two row groups ahead of the epilogue warp,
//   which is instruction-bound -- one warp of dependent code per row group -- and must only carry the critical
//   path).  The other 9 warps (3 per sub-partition 0..2) are consumers: consumer c owns the 7 blocks of
//   16 columns 7c..7c+6 of the CTA's 63 for every row group; the G fragments and the T accumulators of those
//   columns live in registers for the whole kernel.  Keeping sub-partition 3 free of consumers matters: DMMA and
//   scalar FP64 share one pipe per sub-partition, and an epilogue warp that queues behind DMMA streams becomes the
//   critical path of the whole kernel (measured on C2, k = 8: epilogue sharing a sub-partition with two consumers
//   178 us per iteration, with one light consumer 162 us, alone 146 us).
//   phi coupling (star_prod_relevant, R/utils.r:63-78) gathers rows of the partner views' F through int32 row maps:
//   two dependent global loads per partner and row, which must not sit on the epilogue warp's critical path.  They
//   run as a cp.async pipeline in the auxiliary warp: map entries of row group i+4, partner rows of group i+2 (using
//   the entries fetched two iterations earlier), coupling sum of group i from shared memory, handed to the epilogue
//   warp through a 4-slot mbarrier ring together with the old F rows.
//   Software pipeline of a consumer warp: F phase of group i+1, then G phase of group i, so the exchange and the
//   F update of group i+1 overlap the G phase MMAs of group i.
//   Tail: every cluster publishes its T partial [pp8][8]; after a grid-wide arrival counter the 64-column groups
//   are dealt round-robin to the CTAs, each sums the cluster partials in cluster order and runs the same G-update
//   epilogue as rn_g_step_tma (update_g, G'G, A = T'G, colSums(G); last CTA: update_s, update_lm, error,
//   bookkeeping).  All summation orders are fixed: results are bit-reproducible run to run.
#pragma once
#include "rn_kernels.cuh"

#ifndef RN_FU_TG
#define RN_FU_TG 64  // data columns per column group of the tail (G update, G'G | A partials); 4000 columns: 63 groups
#endif
#define RN_FU_NCW 9                                       // consumer warps (3 per sub-partition 0..2)
#define RN_FU_THREADS 384                                 // 12 warps: 9 consumers, producer (3), epilogue (7), auxiliary (11)
#define RN_FU_NB 7                                        // 16-column blocks per consumer warp and row group
#define RN_FU_NCT (32 * RN_FU_NCW)                        // consumer threads (named barrier 1)
#define RN_FU_CBLOCKS 63                                  // 16-column blocks per CTA
#define RN_FU_CCOLS (16 * RN_FU_CBLOCKS)                  // data columns per CTA (1008)
#define RN_FU_GROUP_BYTES (RN_FU_CBLOCKS * 1024)          // a CTA's share of one row group (63 KB)
#define RN_FU_NSLOT (3 * RN_FU_NCW)                       // (row group in the ring, consumer warp) slots
#define RN_FU_RING_BYTES (3 * RN_FU_GROUP_BYTES)          // 189 KB = 3 row groups
#define RN_FU_MAXC 8                                      // largest (portable) cluster: columns <= 8064
// Early probes: a satisfied mbarrier.try_wait still costs ~100 cycles of latency (measured: the consumer warps spend
// ~105 cycles per row group in the wait for X although X is always there), and the three consumer warps of a
// sub-partition reach their waits together, so the FP64 pipe idles behind them.  A non-blocking test_wait issued a few
// MMAs before the phase boundary lets that latency pass under the MMAs; the blocking wait is only taken when the probe
// came back negative.  (Also measured and dropped: splitting a warp's 7 KB share of a row group into 2 / 3 pieces with
// their own barriers so that a piece is reloaded earlier -- 149 / 155.5 / 184 us per update-iteration with 1 / 2 / 3
// pieces: nobody waits for X, every piece only adds a barrier wait -- and one producer lane per consumer warp instead
// of the sequential producer loop: +4 %, the spinning lanes take issue slots from the epilogue warp next to them.)
#ifndef RN_FU_PROBE
#define RN_FU_PROBE 1
#endif
#ifndef RN_FU_PROBE_LEAD
#define RN_FU_PROBE_LEAD 8  // measured 4 / 8: -0.9 % / -2.1 % against no probes on one box. F phase: MMA pairs (of 14) / G phase: half as many blocks (of 7) before the phase ends
#endif

// doubles of shared memory behind the ring (see the carve-up in the kernel)
#define RN_FU_MAXPART 7                                   // most phi partners of a view on the fused path
// Pw | Pex | Fp | Ps (= Wsm during set-up) | Fo | Msm | Ssm | lamh | muh | partner table | SrcIdx | Fg | Pcn
#define RN_FU_AUX_DOUBLES                                                                                 \
  (2 * RN_FU_NCW * 64 + 2 * RN_FU_MAXC * 64 + 2 * 64 + 64 + 4 * 64 + 64 + 64 + 8 + 8 + 48 + 6 * 8 * 8 / 2 + \
   3 * RN_FU_MAXPART * 64 + 4 * 64)
static inline size_t rn_fused_smem() {
  return (size_t)RN_FU_RING_BYTES + (size_t)RN_FU_AUX_DOUBLES * 8 + (2 * RN_FU_NSLOT + 6 + 8) * 8 + 16;
}

// position (in doubles) of X[r][j] in the X8 layout with pp8 (even) columns
__device__ __forceinline__ int64_t rn_x8idx(int64_t r, int64_t j, int64_t pp8) {
  const int64_t q = j >> 1;
  return (((r >> 3) * (pp8 >> 1) + q) << 4) + 2 * ((int)(r & 7) ^ (2 * (int)(q & 3))) + (j & 1);
}

// panel layout -> X8 layout (once per plan; one thread per 16-byte piece, grid-stride)
__global__ void __launch_bounds__(256) rn_panels_to_x8(const RnView vw) {
  const int64_t qrow = vw.pp8 >> 1;
  const int64_t pieces = (vw.ldx >> 3) * qrow * 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < pieces; i += (int64_t)gridDim.x * 256) {
    const int pos = (int)(i & 7);
    const int64_t q = (i >> 3) % qrow;
    const int64_t grp = (i >> 3) / qrow;
    const int64_t r = grp * 8 + (pos ^ (2 * (int)(q & 3)));
    double2 val = make_double2(0.0, 0.0);
    const int64_t j = 2 * q;
    if (j < vw.pp)
      val.x = vw.X[((r >> 6) * vw.pp + j) * RN_ROW_TILE + 2 * ((int)((r & 63) >> 1) ^ rn_sigma(j)) + (r & 1)];
    if (j + 1 < vw.pp)
      val.y = vw.X[((r >> 6) * vw.pp + j + 1) * RN_ROW_TILE + 2 * ((int)((r & 63) >> 1) ^ rn_sigma(j + 1)) + (r & 1)];
    *reinterpret_cast<double2*>(vw.X8 + (i << 1)) = val;
  }
}

// ---- cluster / DSMEM primitives ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t rn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t rn_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void rn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t rn_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 16-byte store into a peer CTA's shared memory; completion is counted (in bytes) on the peer's mbarrier
__device__ __forceinline__ void rn_st_async2(uint32_t raddr, double a, double b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
               "d"(a), "d"(b), "r"(rbar)
               : "memory");
}
// wait on a local mbarrier whose phase is completed by peer CTAs (cluster-scope acquire)
__device__ __forceinline__ void rn_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RN_WAITC_%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RN_DONEC_%=;\n"
      "bra RN_WAITC_%=;\n"
      "RN_DONEC_%=:\n"
      "}\n" ::"r"(rn_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Non-blocking probe of an mbarrier phase.  Deliberately NOT volatile and without a memory clobber: the compiler
// and ptxas may schedule it among the MMAs around it, and the predicate is only consumed at the phase boundary.  The
// caller orders the reads that depend on a positive probe behind an explicit compiler barrier.
__device__ __forceinline__ bool rn_mbar_probe(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm("{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(rn_smem_u32(bar)), "r"(parity));
  return ok != 0;
}
__device__ __forceinline__ void rn_cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(rn_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void rn_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void rn_cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rn_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void rn_cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void rn_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// sum of p[0], p[stride], ..., p[(n-1)*stride] in index order, all loads of a 40-batch in flight at once (one L2 round
// trip per batch instead of one per 8): the cross-cluster / cross-group reductions of the tail are latency-bound
__device__ __forceinline__ double rn_sum_wide(const double* p, int64_t stride, int n) {
  double s = 0.0;
  for (int i0 = 0; i0 < n; i0 += 40) {
    double v[40];
#pragma unroll
    for (int q = 0; q < 40; ++q) v[q] = (i0 + q < n) ? __ldcg(p + (int64_t)(i0 + q) * stride) : 0.0;
#pragma unroll
    for (int q = 0; q < 40; ++q) s += v[q];
  }
  return s;
}
// developer timeline: stamp slot `which` of this CTA with the global nanosecond timer (no-op unless enabled)
__device__ __forceinline__ void rn_fu_stamp(const RnView& vw, int which) {
  if (vw.fu_timeline) {
    long long tns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
    vw.fu_timeline[(int64_t)blockIdx.x * 12 + which] = tns;
  }
}
// developer trace of CTA 0: stamp `which` of local row group i (consumer warp 0: 0 F phase starts, 1 F phase MMAs
// issued, 2 F_new awaited, 3 G phase done; epilogue warp: 4 all warp partials in, 5 cluster partials in, 6 F_new out)
__device__ __forceinline__ void rn_fu_trace(const RnView& vw, int i, int which) {
  if (vw.fu_trace && blockIdx.x == 0) {
    long long tns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
    vw.fu_trace[i * 32 + which] = tns;
  }
}
__device__ __forceinline__ void rn_fu_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RN_FU_NCT) : "memory"); }

template <int K>
__global__ void __launch_bounds__(RN_FU_THREADS, 1) rn_fused_step(const RnView vw, const RnFit ft, const int v,
                                                                  const int fuse_finish) {
  constexpr int KP = 8, KK = K * K, NFF = KK + K, NOUT = 2 * KK + K;
  constexpr int NB = RN_FU_NB, NCW = RN_FU_NCW, NCT = RN_FU_NCT, NSLOT = RN_FU_NSLOT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const bool is_consumer = (warp & 3) != 3;
  const int ci = warp - (warp >> 2);  // consumer index 0..8 (warps 0,1,2, 4,5,6, 8,9,10)
  const int ctid = ci * 32 + lane;    // consumer thread index
  constexpr int nb = NB;              // column blocks of this consumer (uniform; the loops below allow nb < NB)
  const int boff = NB * ci;
  // Programmatic dependent launch: the next launch of the stream (the next view / update-iteration) may be placed on
  // the SMs as this grid drains -- all but the finishing CTA leave ~6 us before the grid completes -- and set itself
  // up (barriers, first row groups of X in flight) while it waits for this grid's results (rn_griddep_wait below).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) rn_fu_stamp(vw, 0);

  extern __shared__ __align__(128) unsigned char rn_smem[];
  unsigned char* ring = rn_smem;
  double* Pw = reinterpret_cast<double*>(rn_smem + RN_FU_RING_BYTES);  // [2][NCW][64]  warp partials of P
  double* Pex = Pw + 2 * NCW * 64;                                      // [2][MAXC][64] CTA partials (peers write)
  double* Fp = Pex + 2 * RN_FU_MAXC * 64;                               // [2][8 rows][8] F_new of a group
  double* Ps = Fp + 2 * 64;                                             // [8][8]        P of the current group
  double* Fo = Ps + 64;                                                 // [4][8][8]     old F rows (cp.async)
  double* Msm = Fo + 4 * 64;                                            // [8][8]        M = S W (zero padded)
  double* Ssm = Msm + 64;                                               // [K*K] (64 reserved)
  double* Wsm = Ps;  // only needed to form M during the set-up, before the epilogue warp first writes Ps
  double* lamh = Ssm + 64;
  double* muh = lamh + 8;
  double* cpl_ph = muh + 8;                                                // phi partner table (epilogue warp)
  double* cpl_nw = cpl_ph + 8;
  const double** cpl_F = reinterpret_cast<const double**>(cpl_nw + 8);
  const int32_t** cpl_map = reinterpret_cast<const int32_t**>(cpl_nw + 16);
  int* cpl_kp = reinterpret_cast<int*>(cpl_nw + 24);                       // [8] kp, then [8] = number of partners
  int32_t* SrcIdx = reinterpret_cast<int32_t*>(cpl_nw + 40);               // [6 groups][8 partners][8 rows]
  double* Fg = cpl_nw + 40 + 6 * 8 * 8 / 2;                                // [3 groups][MAXPART][8 rows][8]
  double* Pcn = Fg + 3 * RN_FU_MAXPART * 64;                               // [4 groups][8 rows][8] coupling sums / n
  uint64_t* full = reinterpret_cast<uint64_t*>(Pcn + 4 * 64);
  uint64_t* empty = full + NSLOT;
  uint64_t* pw_full = empty + NSLOT;
  uint64_t* pex_full = pw_full + 2;
  uint64_t* fp_full = pex_full + 2;
  uint64_t* aux_full = fp_full + 2;   // [4] old F rows + coupling sum of a row group are in shared memory
  uint64_t* aux_empty = aux_full + 4;  // [4] the epilogue warp is done with them
  int* s_flag = reinterpret_cast<int*>(aux_empty + 4);

  const uint32_t rank = rn_cluster_rank(), csize = rn_cluster_size();
  const int64_t n_clusters = gridDim.x / csize, cid = blockIdx.x / csize;
  const int64_t NGT = (vw.n + 7) >> 3;  // row groups that hold data
  const RnSplit gsp(NGT, n_clusters);
  const int64_t g0 = gsp.begin(cid);
  const int NGL = (int)(gsp.begin(cid + 1) - g0);
  const int64_t qrow = vw.pp8 >> 1;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      rn_mbar_init(&full[i], 1);
      rn_mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      rn_mbar_init(&pw_full[i], NCW);
      rn_mbar_init(&pex_full[i], 1);
      rn_mbar_init(&fp_full[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      rn_mbar_init(&aux_full[i], 1);
      rn_mbar_init(&aux_empty[i], 1);
    }
    rn_mbar_init_fence();
  }
  __syncthreads();

  // bulk copies of local row group i into its ring slots, one per consumer warp, each as soon as that warp has
  // released the slot (executed by the whole producer warp)
  auto produce = [&](int i) {
    const double* src = vw.X8 + (((g0 + i) * qrow + (int64_t)rank * (RN_FU_CCOLS / 2)) << 4);
    const int gs = i % 3;
    const uint32_t ph = (uint32_t)((i / 3) & 1);
#pragma unroll 1
    for (int w = 0; w < NCW; ++w) {
      const int wnb = NB;
      const int wboff = NB * w;
      const int st = gs * NCW + w;
      rn_mbar_wait(&empty[st], ph ^ 1u);
      if (lane == 0) {
        rn_mbar_expect_tx(&full[st], (uint32_t)wnb * 1024u);
        rn_bulk_g2s(ring + gs * RN_FU_GROUP_BYTES + wboff * 1024, src + wboff * 128, (uint32_t)wnb * 1024u, &full[st]);
      }
      __syncwarp();
    }
  };
  // X is never written by a kernel: the ring is filled before the previous launch has finished
  const int npre = NGL < 3 ? NGL : 3;
  if (warp == 3)
    for (int i = 0; i < npre; ++i) produce(i);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads what the previous launch wrote
  if (ft.ctrl->done) {  // uniform over the grid; the copies in flight must land before the CTA may exit
    if (warp == 3)
      for (int i = 0; i < npre; ++i)
        for (int w = 0; w < NCW; ++w) rn_mbar_wait(&full[i * NCW + w], 0u);
    return;
  }
  if (tid < 64) {
    Ssm[tid] = (tid < KK) ? vw.S[tid] : 0.0;
    Wsm[tid] = 0.0;
    Msm[tid] = 0.0;
  }
  if (tid < 8) {
    lamh[tid] = (tid < K) ? 0.5 * vw.lam[tid] : 0.0;
    muh[tid] = (tid < K) ? 0.5 * vw.mu[tid] : 0.0;
  }
  __syncthreads();
  if (tid < KK) {  // W = crossprod(G) %*% t(S)
    const int a = tid % K, b = tid / K;
    double s = 0.0;
    for (int c = 0; c < K; ++c) s = fma(vw.GtG[a + c * K], Ssm[b + c * K], s);
    Wsm[a + b * K] = s;
  }
  __syncthreads();
  if (tid < KK) {  // M = S W: the denominator (F S) W of update_f is evaluated as F (S W)
    const int b = tid % K, c = tid / K;
    double s = 0.0;
    for (int a = 0; a < K; ++a) s = fma(Ssm[b + a * K], Wsm[a + c * K], s);
    Msm[b * 8 + c] = s;
  }
  __syncthreads();
  rn_cluster_sync();  // peers' mbarriers are initialised before anyone stores into them
  if (tid == 0) rn_fu_stamp(vw, 1);

  double tacc[2 * NB][2];  // consumer warps: T accumulators of the warp's 112 columns (tile 2b+e: columns 16b+2g+e)
#pragma unroll
  for (int s = 0; s < 2 * NB; ++s) tacc[s][0] = tacc[s][1] = 0.0;
  const int64_t colbase = (int64_t)rank * RN_FU_CCOLS + 16 * boff;

  if (warp == 3) {
    // ---- producer warp -------------------------------------------------------------------------------
    for (int i = npre; i < NGL; ++i) produce(i);
  } else if (warp == 11) {
    // ---- auxiliary warp: lane (g,t) <-> row g, factor columns 2t, 2t+1; runs two row groups ahead of the epilogue --
    const int V = ft.n_views;
    const int kp = vw.kp;
    const int c0 = 2 * t, c1 = 2 * t + 1;
    double phisum = 0.0;
    for (int w = 0; w < V; ++w) phisum += ft.phi[w + v * V];
    const double nv = (double)vw.n_glob;
    // phi partners in view order (zero phi and NA pairs are skipped, as in star_prod_relevant)
    const bool coupled = phisum != 0.0;
    if (coupled) {
      if (lane == 0) {
        int np_ = 0;
        for (int w = 0; w < V && np_ < RN_FU_MAXPART; ++w) {
          const double phw = ft.phi[w + v * V];
          if (phw == 0.0) continue;
          const int mode = ft.rowmode[w + v * V];
          if (mode == RN_MODE_NA) continue;
          const RnView* ow = ft.views + w;
          cpl_ph[np_] = phw;
          cpl_nw[np_] = (double)ow->n_glob;
          cpl_F[np_] = ow->F;
          cpl_kp[np_] = ow->kp;
          cpl_map[np_] = (mode == RN_MODE_MAP) ? ft.rowmap[w + v * V] : nullptr;  // NULL pair: nothing overwritten
          ++np_;
        }
        cpl_kp[8] = np_;
      }
      __syncwarp();
    }
    const int np = coupled ? cpl_kp[8] : 0;
    auto prefetch_f = [&](int i) {  // old F rows of local group i -> Fo[i & 3]
      if (i < NGL) {
        const int64_t r = (g0 + i) * 8 + g;
        rn_cp_async8(Fo + (i & 3) * 64 + g * 8 + c0, vw.F + rn_fidx(r, c0, kp));
        rn_cp_async8(Fo + (i & 3) * 64 + g * 8 + c1, vw.F + rn_fidx(r, c1, kp));
      }
    };
    auto prefetch_idx = [&](int i) {  // row-map entries of local group i (lane t of a row: partners t, t+4)
      if (i < NGL) {
        const int64_t r = (g0 + i) * 8 + g;
        if (r < vw.n)
          for (int pi = t; pi < np; pi += 4) {
            const int32_t* mp = cpl_map[pi];
            if (mp) rn_cp_async4(SrcIdx + ((i % 6) * 8 + pi) * 8 + g, mp + r);
          }
      }
    };
    auto prefetch_gather = [&](int i) {  // partner rows of local group i (its map entries are in shared memory)
      if (i < NGL) {
        const int64_t r = (g0 + i) * 8 + g;
        if (r < vw.n)
          for (int pi = 0; pi < np; ++pi) {
            if (!cpl_map[pi]) continue;
            const int src = SrcIdx[((i % 6) * 8 + pi) * 8 + g];
            if (src < 0) continue;
            double* dst = Fg + ((i % 3) * RN_FU_MAXPART + pi) * 64 + g * 8;
            rn_cp_async8(dst + c0, cpl_F[pi] + rn_fidx(src, c0, cpl_kp[pi]));
            rn_cp_async8(dst + c1, cpl_F[pi] + rn_fidx(src, c1, cpl_kp[pi]));
          }
      }
    };
    if (coupled) {
      for (int i = 0; i < 4; ++i) prefetch_idx(i);
      rn_cp_async_commit();
      rn_cp_async_wait_all();
      __syncwarp();
    }
    for (int i = 0; i < 2; ++i) {  // one cp.async batch per row group: the batch of group i is complete at iteration i
      prefetch_f(i);
      if (coupled) prefetch_gather(i);
      rn_cp_async_commit();
    }
    auto aux_step = [&](int i) {
      // the slots written below held group i-2 (old F rows) and i-4 (coupling sum): the epilogue is done with them
      if (i >= 2) rn_mbar_wait(&aux_empty[(i - 2) & 3], (uint32_t)(((i - 2) >> 2) & 1));
      rn_cp_async_wait1();  // batch i: old F rows + partner rows of this group, map entries of group i+2
      __syncwarp();
      prefetch_f(i + 2);
      if (coupled) {
        prefetch_gather(i + 2);
        prefetch_idx(i + 4);
      }
      rn_cp_async_commit();
      if (coupled) {
        const double2 fmine = *reinterpret_cast<const double2*>(Fo + (i & 3) * 64 + g * 8 + c0);
        double pc0 = 0.0, pc1 = 0.0;
        for (int pi = 0; pi < np; ++pi) {
          const int src = cpl_map[pi] ? SrcIdx[((i % 6) * 8 + pi) * 8 + g] : -1;
          double m0 = fmine.x, m1 = fmine.y;  // row not shared with this partner: the view's own row (utils.r:69-73)
          if (src >= 0) {
            const double2 mm = *reinterpret_cast<const double2*>(Fg + ((i % 3) * RN_FU_MAXPART + pi) * 64 + g * 8 + c0);
            m0 = mm.x;
            m1 = mm.y;
          }
          pc0 += (cpl_ph[pi] * m0) * cpl_nw[pi];
          pc1 += (cpl_ph[pi] * m1) * cpl_nw[pi];
        }
        *reinterpret_cast<double2*>(Pcn + (i & 3) * 64 + g * 8 + c0) = make_double2(pc0 / nv, pc1 / nv);
      }
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&aux_full[i & 3]);
    };
    for (int i = 0; i < NGL; ++i) aux_step(i);
  } else if (warp == 7) {
    // ---- epilogue warp: lane (g,t) owns row g, factor columns 2t and 2t+1 of every row group; critical path only --
    const int V = ft.n_views;
    const int kp = vw.kp;
    const int c0 = 2 * t, c1 = 2 * t + 1;
    // B operands of the 8x8 products: N = P t(S)  (B[a][c] = S[c, a]),  D = F M  (B[b][c] = M[b, c]); lane (g,t)
    // supplies B[t][g] and B[t+4][g]
    const double bs0 = (g < K && t < K) ? Ssm[g + t * K] : 0.0;
    const double bs1 = (g < K && t + 4 < K) ? Ssm[g + (t + 4) * K] : 0.0;
    const double bm0 = Msm[t * 8 + g], bm1 = Msm[(t + 4) * 8 + g];
    const double lam0 = lamh[c0], lam1 = lamh[c1];
    double phisum = 0.0;
    for (int w = 0; w < V; ++w) phisum += ft.phi[w + v * V];
    const bool coupled = phisum != 0.0;
    double ff0 = 0.0, ff1 = 0.0, cs0 = 0.0, cs1 = 0.0;
    const uint32_t my_pex = rn_smem_u32(Pex + rank * 64 + 2 * lane);
    const int sig0 = rn_sigma(c0), sig1 = rn_sigma(c1);
    for (int i = 0; i < NGL; ++i) {
      const int sl = i & 1;
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      if (lane == 0) rn_mbar_expect_tx(&pex_full[sl], csize * 512u);
      // old F rows (+ coupling sum) of this group are in shared memory -- in particular before any peer can be
      // released to overwrite those rows in HBM
      rn_mbar_wait(&aux_full[i & 3], (uint32_t)((i >> 2) & 1));
      rn_mbar_wait(&pw_full[sl], ph);
      if (lane == 0) rn_fu_trace(vw, i, 4);
      double2 acc = *reinterpret_cast<const double2*>(Pw + (sl * NCW) * 64 + 2 * lane);
#pragma unroll
      for (int w = 1; w < NCW; ++w) {
        const double2 x = *reinterpret_cast<const double2*>(Pw + (sl * NCW + w) * 64 + 2 * lane);
        acc.x += x.x;
        acc.y += x.y;
      }
      for (uint32_t rr = 0; rr < csize; ++rr)
        rn_st_async2(rn_mapa(my_pex + sl * (RN_FU_MAXC * 64 * 8), rr), acc.x, acc.y,
                     rn_mapa(rn_smem_u32(&pex_full[sl]), rr));
      // operands that do not depend on the exchange, loaded while it is in flight
      const double* fo = Fo + (i & 3) * 64 + g * 8;
      const double fa0 = fo[t], fa1 = fo[t + 4];
      const double2 fmine = *reinterpret_cast<const double2*>(fo + c0);
      double2 pcn = make_double2(0.0, 0.0);
      if (coupled) pcn = *reinterpret_cast<const double2*>(Pcn + (i & 3) * 64 + g * 8 + c0);
      double N0 = 0.0, N1 = 0.0, D0 = 0.0, D1 = 0.0;
      rn_dmma(D0, D1, fa0, bm0);
      rn_dmma(D0, D1, fa1, bm1);
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&aux_empty[i & 3]);
      rn_mbar_wait_cluster(&pex_full[sl], ph);
      if (lane == 0) rn_fu_trace(vw, i, 5);
      double2 tot = *reinterpret_cast<const double2*>(Pex + (sl * RN_FU_MAXC) * 64 + 2 * lane);
      for (uint32_t rr = 1; rr < csize; ++rr) {
        const double2 x = *reinterpret_cast<const double2*>(Pex + (sl * RN_FU_MAXC + rr) * 64 + 2 * lane);
        tot.x += x.x;
        tot.y += x.y;
      }
      *reinterpret_cast<double2*>(Ps + 2 * lane) = tot;
      __syncwarp();
      rn_dmma(N0, N1, Ps[g * 8 + t], bs0);
      rn_dmma(N0, N1, Ps[g * 8 + t + 4], bs1);
      const double f0 = fmine.x, f1 = fmine.y;
      const int64_t grp = g0 + i;
      const int64_t r = grp * 8 + g;
      double o0 = 0.0, o1 = 0.0;
      if (r < vw.n) {
        // padding columns (c >= K) would divide 0 by 0 and take the slow exact path on the critical warp: the
        // denominators of those lanes are replaced by 1, their outputs are zero anyway
        const bool v0 = c0 < K, v1 = c1 < K;
        if (!coupled) {  // update_steps.r:152-155
          double q0 = rn_fast_div(N0, v0 ? D0 + lam0 : 1.0), q1 = rn_fast_div(N1, v1 ? D1 + lam1 : 1.0);
          if (isnan(q0)) q0 = 1.0;
          if (isnan(q1)) q1 = 1.0;
          o0 = fabs(f0 * q0);
          o1 = fabs(f1 * q1);
        } else {  // update_steps.r:156-163 with star_prod_relevant (utils.r:63-78)
          o0 = fabs(f0 * rn_fast_div(N0 + pcn.x, v0 ? (D0 + phisum * f0) + lam0 : 1.0));
          o1 = fabs(f1 * rn_fast_div(N1 + pcn.y, v1 ? (D1 + phisum * f1) + lam1 : 1.0));
        }
        if (!v0) o0 = 0.0;
        if (!v1) o1 = 0.0;
      }
      *reinterpret_cast<double2*>(Fp + sl * 64 + g * 8 + c0) = make_double2(o0, o1);
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&fp_full[sl]);
      if (lane == 0) rn_fu_trace(vw, i, 6);
      if (rank == 0) {  // off the critical path: F_new to HBM, F'F and colSums(F) of this cluster's rows
        if (r < vw.n) {
          // rn_fidx(r, c, kp) with r = 8 grp + g: 8 row groups per 64-row panel of F
          double* fpan = vw.F + (grp >> 3) * kp * RN_ROW_TILE + (g & 1);
          const int piece = (((int)(grp & 7)) << 2) | (g >> 1);
          if (c0 < K) fpan[c0 * RN_ROW_TILE + 2 * (piece ^ sig0)] = o0;
          if (c1 < K) fpan[c1 * RN_ROW_TILE + 2 * (piece ^ sig1)] = o1;
        }
        const double x1 = Fp[sl * 64 + t * 8 + g], x2 = Fp[sl * 64 + (t + 4) * 8 + g];
        rn_dmma(ff0, ff1, x1, x1);
        rn_dmma(cs0, cs1, 1.0, x1);
        rn_dmma(ff0, ff1, x2, x2);
        rn_dmma(cs0, cs1, 1.0, x2);
      }
    }
    if (rank == 0) {
      double* mine = vw.FFpart + cid * NFF;
      if (g < K) {
        if (c0 < K) mine[g + c0 * K] = ff0;
        if (c1 < K) mine[g + c1 * K] = ff1;
      }
      if (g == 0) {
        if (c0 < K) mine[KK + c0] = cs0;
        if (c1 < K) mine[KK + c1] = cs1;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(&vw.misc_ticket[2], 1);
    }
  } else {
    // ---- consumer warps --------------------------------------------------------------------------------
    double gfr[2 * NB][2];  // G fragments of the warp's columns: B operand of the F phase
#pragma unroll
    for (int s = 0; s < 2 * NB; ++s)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t col = colbase + 8 * s + 2 * t + e;
        gfr[s][e] = (s < 2 * nb && col < vw.pp) ? vw.G[col * KP + g] : 0.0;
      }
    const int ra = (t & 1) + 4 * (t >> 1);  // G-phase K slot t <-> rows {0,1,4,5} (first MMA), {2,3,6,7} (second)
    const int rb = ra + 2;
    const uint32_t off1 = (uint32_t)(t * 128 + ((g ^ (2 * t)) * 16));
    const uint32_t off2a = (uint32_t)(g * 128 + ((ra ^ (2 * (g & 3))) * 16));
    const uint32_t off2b = (uint32_t)(g * 128 + ((rb ^ (2 * (g & 3))) * 16));

    // One loop iteration = F phase of group i+1, its publication, then G phase of group i.  The partial is published
    // the moment the F phase ends: F_new of group i+1 is needed one G phase + one F phase later and the chain behind
    // the publication (9 warps in, exchange, F update) takes most of that (trace: RESNMTF_FU_TIMELINE=1) -- publishing
    // after the first G-phase MMAs instead (tried: no pipe drain at the phase boundary) cost 350 ns of that slack.
    double pe0 = 0.0, pe1 = 0.0, po0 = 0.0, po1 = 0.0;  // two accumulator chains: even / odd column of a pair
    const bool tr_waits = vw.fu_waits != nullptr;        // developer switch: cycles this warp waited for X / F_new
    long long wait_x = 0, wait_f = 0;
    bool x_ready = false, fn_ready = false;              // outcome of the early probes (RN_FU_PROBE)
    auto f_phase = [&](int i) {
      const int gs = i % 3;
      if (tid == 0) rn_fu_trace(vw, i, 0);
      {
        long long c0_ = 0;
        if (tr_waits) c0_ = clock64();
        if (!(RN_FU_PROBE && x_ready)) rn_mbar_wait(&full[gs * NCW + ci], (uint32_t)((i / 3) & 1));
        asm volatile("" ::: "memory");  // nothing below is read before the (probed or awaited) phase completion
        if (tr_waits) wait_x += clock64() - c0_;
      }
      if (tid == 0) rn_fu_trace(vw, i, 7);
      const unsigned char* xs = ring + gs * RN_FU_GROUP_BYTES + boff * 1024 + off1;
      pe0 = pe1 = po0 = po1 = 0.0;
#pragma unroll
      for (int s = 0; s < 2 * NB; ++s) {
        // F_new of the group whose G phase follows this F phase (group i - 1), probed a few MMAs early
        if (RN_FU_PROBE && s == 2 * NB - RN_FU_PROBE_LEAD)
          fn_ready = i >= 1 && rn_mbar_probe(&fp_full[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1));
        if (s < 2 * nb) {  // uniform over the warp
          const double2 x = *reinterpret_cast<const double2*>(xs + s * 512);
          rn_dmma(pe0, pe1, x.x, gfr[s][0]);
          rn_dmma(po0, po1, x.y, gfr[s][1]);
        }
      }
    };
    auto publish = [&](int i) {
      *reinterpret_cast<double2*>(Pw + ((i & 1) * NCW + ci) * 64 + 2 * lane) = make_double2(pe0 + po0, pe1 + po1);
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&pw_full[i & 1]);
      if (lane == 0) rn_fu_trace(vw, i, 8 + ci);  // per-warp publication time
    };
    auto g_phase = [&](int i) {
      const int gs = i % 3;
      if (tid == 0) rn_fu_trace(vw, i, 1);
      {
        long long c0_ = 0;
        if (tr_waits) c0_ = clock64();
        if (!(RN_FU_PROBE && fn_ready)) rn_mbar_wait(&fp_full[i & 1], (uint32_t)((i >> 1) & 1));
        asm volatile("" ::: "memory");
        fn_ready = false;
        if (tr_waits) wait_f += clock64() - c0_;
      }
      if (tid == 0) rn_fu_trace(vw, i, 2);
      if (lane == 0) rn_fu_trace(vw, i, 17 + ci);  // per-warp start of the G phase
      const double fa = Fp[(i & 1) * 64 + ra * 8 + g];
      const double fb = Fp[(i & 1) * 64 + rb * 8 + g];
      const unsigned char* xs = ring + gs * RN_FU_GROUP_BYTES + boff * 1024;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        // X of the group whose F phase follows this G phase (group i + 2), probed early
        if (RN_FU_PROBE && b == NB - (RN_FU_PROBE_LEAD + 1) / 2)
          x_ready = i + 2 < NGL && rn_mbar_probe(&full[((i + 2) % 3) * NCW + ci], (uint32_t)(((i + 2) / 3) & 1));
        if (b < nb) {  // uniform over the warp
          const double2 xa = *reinterpret_cast<const double2*>(xs + off2a + b * 1024);
          const double2 xb = *reinterpret_cast<const double2*>(xs + off2b + b * 1024);
          rn_dmma(tacc[2 * b][0], tacc[2 * b][1], xa.x, fa);
          rn_dmma(tacc[2 * b + 1][0], tacc[2 * b + 1][1], xa.y, fa);
          rn_dmma(tacc[2 * b][0], tacc[2 * b][1], xb.x, fb);
          rn_dmma(tacc[2 * b + 1][0], tacc[2 * b + 1][1], xb.y, fb);
        }
      }
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&empty[gs * NCW + ci]);
      if (tid == 0) rn_fu_trace(vw, i, 3);
    };
    if (NGL > 0) {
      f_phase(0);
      publish(0);
    }
    for (int i = 0; i < NGL; ++i) {
      if (i + 1 < NGL) {
        f_phase(i + 1);
        publish(i + 1);
      }
      g_phase(i);
    }
    if (tr_waits && lane == 0) {
      vw.fu_waits[((int64_t)blockIdx.x * NCW + ci) * 2] = wait_x;
      vw.fu_waits[((int64_t)blockIdx.x * NCW + ci) * 2 + 1] = wait_f;
    }
  }
  if (tid == 0) rn_fu_stamp(vw, 2);
  rn_cluster_sync();  // every st.async of this cluster has landed before any of its CTAs may exit
  if (!is_consumer) return;
  if (tid == 0) rn_fu_stamp(vw, 3);

  // ---- tail (consumer warps): publish T partials, then the column-group epilogues ------------------------
  {
    double* tp = vw.Tpart + (cid * vw.pp8 + colbase) * KP;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (b < nb)
          *reinterpret_cast<double2*>(tp + (16 * b + 2 * g + e) * KP + 2 * t) =
              make_double2(tacc[2 * b + e][0], tacc[2 * b + e][1]);
  }
  __threadfence();
  rn_fu_consumer_sync();
  if (ctid == 0) atomicAdd(&vw.misc_ticket[3], 1);

  double* Ts = reinterpret_cast<double*>(ring);  // the ring is idle now: epilogue scratch lives there
  double* Gs = Ts + RN_FU_TG * KP;
  double* FtFs = Gs + RN_FU_TG * K;
  double* Vs = FtFs + NFF;
  double* fin = Vs + KK;
  double* Us = fin + NOUT;
  double* Sn = Us + KK;
  double* red = Sn + KK;
  const int64_t pp = vw.pp;
  const int64_t NG = (pp + RN_FU_TG - 1) / RN_FU_TG;
  if ((int64_t)blockIdx.x >= NG) return;  // no column group for this CTA (it must not wait: the finisher re-arms)
  if (ctid == 0) {
    while (rn_ld_acquire(&vw.misc_ticket[3]) < (int)gridDim.x) __nanosleep(32);
    while (rn_ld_acquire(&vw.misc_ticket[2]) < (int)n_clusters) __nanosleep(32);
  }
  rn_fu_consumer_sync();
  __threadfence();
  if (ctid == 0) rn_fu_stamp(vw, 4);
  bool ff_ready = false;
  const int64_t tstride = vw.pp8 * KP;
  for (int64_t grp = blockIdx.x; grp < NG; grp += gridDim.x) {
    const int64_t j0 = grp * RN_FU_TG;
    const int njb = (int)min((int64_t)(RN_FU_TG / 8), (pp - j0) >> 3);
    rn_fu_consumer_sync();  // previous group's epilogue is done with Ts / Gs
    for (int i = ctid; i < RN_FU_TG * KP; i += NCT)
      Ts[i] = (i < 8 * njb * KP) ? rn_sum_wide(vw.Tpart + j0 * KP + i, tstride, (int)n_clusters) : 0.0;
    if (!ff_ready) {
      if (ctid < NFF) FtFs[ctid] = rn_sum_wide(vw.FFpart + ctid, NFF, (int)n_clusters);
      rn_fu_consumer_sync();
      for (int o = ctid; o < KK; o += NCT) {  // V = crossprod(F) %*% S
        const int a = o % K, c = o / K;
        double s = 0.0;
        for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
        Vs[a + c * K] = s;
      }
      ff_ready = true;
    }
    rn_fu_consumer_sync();
    if (ctid == 0) rn_fu_stamp(vw, 5);
    if (ctid < RN_FU_TG) {
      const int64_t j = j0 + ctid;
      double gn[K];
      if (j < vw.p) {
        double Tj[K];
#pragma unroll
        for (int c = 0; c < K; ++c) Tj[c] = Ts[ctid * KP + c];
        rn_update_g_row<K>(vw, ft, v, j, Tj, Ssm, Vs, muh, gn);
      } else {
#pragma unroll
        for (int c = 0; c < K; ++c) gn[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < K; ++c) Gs[ctid * K + c] = gn[c];
    }
    rn_fu_consumer_sync();
    for (int o = ctid; o < NOUT; o += NCT) {
      double s = 0.0;
      if (o < KK) {
        const int a = o % K, b = o / K;
        for (int i = 0; i < RN_FU_TG; ++i) s = fma(Gs[i * K + a], Gs[i * K + b], s);
      } else if (o < 2 * KK) {
        const int a = (o - KK) % K, b = (o - KK) / K;
        for (int i = 0; i < RN_FU_TG; ++i) s = fma(Ts[i * KP + a], Gs[i * K + b], s);
      } else {
        const int c = o - 2 * KK;
        for (int i = 0; i < RN_FU_TG; ++i) s += Gs[i * K + c];
      }
      vw.GGpart[grp * NOUT + o] = s;
    }
    __threadfence();
    rn_fu_consumer_sync();
    if (ctid == 0) *s_flag = (atomicAdd(&vw.misc_ticket[0], 1) == (int)NG - 1);
    rn_fu_consumer_sync();
    if (ctid == 0) rn_fu_stamp(vw, 6);
    if (!*s_flag) continue;
    __threadfence();
    // ---- last column group done: finish the view -----------------------------------------------------------
    {  // G'G | A | colSums(G): the two halves of the column-group partials are summed side by side (one L2 round trip)
      static_assert(2 * NOUT <= NCT, "two threads per output");
      const int half = (int)((NG + 1) / 2);
      double part = 0.0;
      if (ctid < NOUT) part = rn_sum_wide(vw.GGpart + ctid, NOUT, half);
      else if (ctid < 2 * NOUT) part = rn_sum_wide(vw.GGpart + (int64_t)half * NOUT + (ctid - NOUT), NOUT, (int)NG - half);
      if (ctid >= NOUT && ctid < 2 * NOUT) red[ctid - NOUT] = part;  // scratch: red is the last array carved from the idle ring
      if (ctid == 0) {
        vw.misc_ticket[0] = 0;
        vw.misc_ticket[2] = 0;
        vw.misc_ticket[3] = 0;
      }
      rn_fu_consumer_sync();
      if (ctid < NOUT) fin[ctid] = part + red[ctid];
    }
    rn_fu_consumer_sync();
    if (ctid == 0) rn_fu_stamp(vw, 8);
    rn_view_finish<K, NCT, true>(vw, ft, v, ctid, fin, FtFs, Ssm, Us, Sn, red, fuse_finish);
    if (ctid == 0) rn_fu_stamp(vw, 7);
  }
}
