// One-pass fused update of one view, second generation: same dataflow as rn_fused_step (rn_fused.cuh -- F step AND G
// step with a single read of X per update-iteration, 8-row groups resident on chip, the 8 x 8 partial of P = X G
// exchanged between the CTAs of a cluster through distributed shared memory), re-cut so that the FP64 tensor pipe of
// ALL FOUR sub-partitions of an SM is fed and the kernel ends up bound by HBM instead of by DMMA issue:
//
//   * 12 consumer warps, 3 per SM sub-partition (512 threads).  The first kernel kept sub-partition 3 free of
//     consumers because its single epilogue warp sat on the critical path of every row group: F_new of group i+1 had to
//     be out one G phase after the F phase that produced its partial.  Here the consumers run TWO groups ahead (F phase
//     of group i+2, then G phase of group i), so the exchange + F update of a group have two full periods, and there
//     are TWO epilogue warps (even / odd groups) so that one group per period is within reach of warps that share
//     their sub-partition's FP64 pipe with DMMA streams.
//   * a CTA holds at most 42 blocks of 16 columns (42 KB per row group) instead of 63: four ring stages fit (two groups
//     between their F and G phase, one in each phase, the reloads in the released slots), clusters of
//     C = ceil(p / 16 / 42) CTAs; the columns are dealt evenly (blocks per CTA differ by at most one, no padding to a
//     fixed width: the X8 copy has exactly pp columns) and inside a CTA a warp owns 3 or 4 blocks, the extra ones going
//     to sub-partitions 0 and 2 first (1 and 3 carry the epilogue warps).
//   * one bulk copy per PAIR of consumer warps and row group (6 per group, ~7 KB each).
//   * F_new rows are stored to HBM by cluster rank (group mod C) instead of rank 0 only.
// Per-group critical path, partial publication, DSMEM exchange (st.async + complete_tx), the auxiliary warp with the
// phi-coupling gather pipeline, the fixed summation orders and the tail are those of rn_fused_step.
#pragma once
#include <type_traits>

#include "rn_fused.cuh"

#define RN_F2_NCW 12                                     // consumer warps 0..11 (3 per sub-partition)
#define RN_F2_THREADS 512                                // + producer (12), epilogue A (13), auxiliary (14), epilogue B (15)
#define RN_F2_NCT (32 * RN_F2_NCW)                       // consumer threads (named barrier 1)
#define RN_F2_NBW 4                                      // most 16-column blocks per consumer warp
#define RN_F2_MAXB 42                                    // most 16-column blocks per CTA
#define RN_F2_NST 4                                      // ring stages (row groups)
#define RN_F2_STAGE_BYTES (RN_F2_MAXB * 1024)
#define RN_F2_RING_BYTES (RN_F2_NST * RN_F2_STAGE_BYTES)  // 168 KB
#define RN_F2_NCOPY 2                                    // bulk copies per row group: half A / half B of the CTA's blocks
#define RN_F2_NSLOT (RN_F2_NST * RN_F2_NCOPY)
#ifndef RN_F2_FFIRST
#define RN_F2_FFIRST 1                                   // 1: F phase of group i+2, then the whole G phase of group i
#endif
#ifndef RN_F2_SPLIT
#define RN_F2_SPLIT 1                                    // 1: two copies / barrier pairs per stage (halves A, B), 0: one
#endif
#ifndef RN_F2_PROBE
#define RN_F2_PROBE 0                                    // 1: early non-blocking probes of the next phase's barrier
#endif
#ifndef RN_F2_STAGGER
#define RN_F2_STAGGER 0                                  // 1: one consumer warp per sub-partition runs G phase first, then F
#endif
#ifndef RN_F2_SKEW
#define RN_F2_SKEW 0                                     // 1: fewer blocks on the sub-partitions of the epilogue warps
#endif
#ifndef RN_F2_EPRE
#define RN_F2_EPRE 0                                     // 1: epilogue warps form a group's denominator reciprocal one group early
#endif
#ifndef RN_F2_PF
#define RN_F2_PF 4                                       // L2 prefetch distance in row groups (ahead of the ring copy)
#endif
#define RN_F2_NPW 3                                      // slots of warp partials / of F_new (groups in flight + 1)
#define RN_F2_NPEX 4                                     // slots of the cluster exchange (two epilogue warps)
// Pw | Pex | Fp | Dn | Fo | Msm | Ssm | lamh | muh | partner table | SrcIdx | Fg | Pcn
#define RN_F2_AUX_DOUBLES                                                                                        \
  (RN_F2_NPW * RN_F2_NCW * 64 + RN_F2_NPEX * RN_FU_MAXC * 64 + RN_F2_NPW * 64 + 4 * 64 + 4 * 64 + 64 + 64 + 8 + 8 + 48 + \
   6 * 8 * 8 / 2 + 3 * RN_FU_MAXPART * 64 + 4 * 64)
#define RN_F2_NBAR (2 * RN_F2_NSLOT + RN_F2_NPW + RN_F2_NPEX + RN_F2_NPW + 4 + 4)
static inline size_t rn_fused2_smem() {
  return (size_t)RN_F2_RING_BYTES + (size_t)RN_F2_AUX_DOUBLES * 8 + (size_t)RN_F2_NBAR * 8 + 16;
}

// Blocks of consumer warp w (nb) and the blocks before it (off) when `total` blocks are dealt to the 12 warps: total / 12
// each, the total % 12 extra ones to the warps listed first in `prio`.  Half A and half B use different lists so that a
// CTA with 42 blocks ends up with 11 / 10 / 11 / 10 blocks on sub-partitions 0..3 (1 and 3 carry the epilogue warps).
__device__ __forceinline__ void rn_f2_share(int total, int w, bool half_b, int& nb, int& off) {
  const int base = total / RN_F2_NCW, extra = total % RN_F2_NCW;
  // position of every warp in the priority list (4 bits each): A = 0,2,4,6,8,10,1,3,5,7,9,11; B = 11,9,7,0,2,4,6,1,3,8,10,5
  const unsigned long long pos_a = 0xB5A493827160ull;  // nibble w = position of warp w in list A
#if RN_F2_SKEW
  // list B = 0,2,4,6,8,10,7,11,9,1,3,5: a CTA with 42 blocks carries 12 / 9 / 12 / 9 (the epilogue warps' sub-partitions
  // keep more of their FP64 pipe)
  const unsigned long long pos_b = 0x758463B2A190ull;
#else
  const unsigned long long pos_b = 0x0A1926B58473ull;  // nibble w = position of warp w in list B
#endif
  const unsigned long long pos = half_b ? pos_b : pos_a;
  nb = base + ((int)((pos >> (4 * w)) & 15) < extra ? 1 : 0);
  off = base * w;
  for (int u = 0; u < w; ++u) off += ((int)((pos >> (4 * u)) & 15) < extra) ? 1 : 0;
}
// developer trace of CTA 0 (per-group stamps, printed with RESNMTF_FU_TIMELINE=1): compiled in with -DRN_F2_TRACE only
#ifdef RN_F2_TRACE
#define RN_F2_TR(cond, i, which) \
  do {                           \
    if (cond) rn_fu_trace(vw, i, which); \
  } while (0)
#else
#define RN_F2_TR(cond, i, which) \
  do {                           \
  } while (0)
#endif
__device__ __forceinline__ void rn_f2_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RN_F2_NCT) : "memory"); }

template <int K>
__global__ void __launch_bounds__(RN_F2_THREADS, 1) rn_fused2_step(const RnView vw, const RnFit ft, const int v,
                                                                   const int fuse_finish) {
  constexpr int KP = 8, KK = K * K, NFF = KK + K, NOUT = 2 * KK + K;
  constexpr int NBW = RN_F2_NBW, NCW = RN_F2_NCW, NCT = RN_F2_NCT, NST = RN_F2_NST;
  constexpr int NPW = RN_F2_NPW, NPEX = RN_F2_NPEX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const bool is_consumer = warp < NCW;
  const int ci = warp;        // consumer index (consumer warps only)
  const int ctid = tid;       // consumer thread index (consumer warps only)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) rn_fu_stamp(vw, 0);

  extern __shared__ __align__(128) unsigned char rn_smem[];
  unsigned char* ring = rn_smem;
  double* Pw = reinterpret_cast<double*>(rn_smem + RN_F2_RING_BYTES);  // [NPW][NCW][64]   warp partials of P
  double* Pex = Pw + NPW * NCW * 64;                                    // [NPEX][MAXC][64] CTA partials (peers write)
  double* Fp = Pex + NPEX * RN_FU_MAXC * 64;                            // [NPW][8 rows][8] F_new of a group
  double* Dn = Fp + NPW * 64;                                           // [4][8][8]        denominators of update_f (aux warp)
  double* Fo = Dn + 4 * 64;                                             // [4][8][8]        old F rows (cp.async)
  double* Msm = Fo + 4 * 64;                                            // [8][8]           M = S W (zero padded)
  double* Ssm = Msm + 64;                                               // [K*K] (64 reserved)
  double* Wsm = Dn;  // only needed to form M during the set-up, before the auxiliary warp first writes Dn
  double* lamh = Ssm + 64;
  double* muh = lamh + 8;
  double* cpl_ph = muh + 8;                                                // phi partner table
  double* cpl_nw = cpl_ph + 8;
  const double** cpl_F = reinterpret_cast<const double**>(cpl_nw + 8);
  const int32_t** cpl_map = reinterpret_cast<const int32_t**>(cpl_nw + 16);
  int* cpl_kp = reinterpret_cast<int*>(cpl_nw + 24);                       // [8] kp, then [8] = number of partners
  int32_t* SrcIdx = reinterpret_cast<int32_t*>(cpl_nw + 40);               // [6 groups][8 partners][8 rows]
  double* Fg = cpl_nw + 40 + 6 * 8 * 8 / 2;                                // [3 groups][MAXPART][8 rows][8]
  double* Pcn = Fg + 3 * RN_FU_MAXPART * 64;                               // [4 groups][8 rows][8] coupling sums / n
  uint64_t* full = reinterpret_cast<uint64_t*>(Pcn + 4 * 64);
  uint64_t* empty = full + RN_F2_NSLOT;
  uint64_t* pw_full = empty + RN_F2_NSLOT;
  uint64_t* pex_full = pw_full + NPW;
  uint64_t* fp_full = pex_full + NPEX;
  uint64_t* aux_full = fp_full + NPW;    // [4] old F rows + coupling sum of a row group are in shared memory
  uint64_t* aux_empty = aux_full + 4;    // [4] the epilogue warp of that group is done with them
  int* s_flag = reinterpret_cast<int*>(aux_empty + 4);

  const uint32_t rank = rn_cluster_rank(), csize = rn_cluster_size();
  const int64_t n_clusters = gridDim.x / csize, cid = blockIdx.x / csize;
  const int64_t NGT = (vw.n + 7) >> 3;  // row groups that hold data
  const RnSplit gsp(NGT, n_clusters);
  const int64_t g0 = gsp.begin(cid);
  const int NGL = (int)(gsp.begin(cid + 1) - g0);
  const int64_t qrow = vw.pp8 >> 1;
  const RnSplit bsp(vw.pp8 >> 4, csize);  // 16-column blocks of the view over the CTAs of the cluster
  const int b0 = (int)bsp.begin(rank);
  const int cb = (int)bsp.begin(rank + 1) - b0;

  if (tid == 0) {
    for (int i = 0; i < RN_F2_NSLOT; ++i) {
      rn_mbar_init(&full[i], 1);
      rn_mbar_init(&empty[i], NCW);
    }
    for (int i = 0; i < NPW; ++i) {
      rn_mbar_init(&pw_full[i], NCW);
      rn_mbar_init(&fp_full[i], 1);
    }
    for (int i = 0; i < NPEX; ++i) rn_mbar_init(&pex_full[i], 1);
    for (int i = 0; i < 4; ++i) {
      rn_mbar_init(&aux_full[i], 1);
      rn_mbar_init(&aux_empty[i], 1);
    }
    rn_mbar_init_fence();
  }
  __syncthreads();

  // Copy slots.  The CTA's blocks are split into a half A (the first (cb+1)/2) and a half B, each dealt to the 12 consumer
  // warps and each with its own ring barriers, because a warp runs  G phase A of group i -> F phase of group i+2 -> G phase
  // B of group i:  half A of a stage is released (and reloaded with group i+4) two periods before the F phase that needs
  // it, half B 1.5 periods, and F_new of a group still has 1.5 periods after its partial was published.  Measured
  // before: whole-stage slots released after a whole G phase leave the reload exactly one period, less than the HBM
  // latency under load (1.14 us per group instead of 0.85); one copy per pair of warps and half (12 per group) makes the
  // producer warp's serial loop -- ~120 ns per copy -- the bottleneck (1.54 us per group).  Two copies per group.
  const int nA = (cb + 1) >> 1, nB = cb - nA;
  int w_na = 0, w_nb = 0, w_aoff = 0, w_boff = 0;  // this consumer warp: blocks in half A / B, their first blocks
  if (is_consumer) {
    rn_f2_share(nA, ci, false, w_na, w_aoff);
    rn_f2_share(nB, ci, true, w_nb, w_boff);
    w_boff += nA;
  }
  // bulk copies of local row group i into its ring stage, each as soon as all consumer warps have released that half
  // (executed by the whole producer warp)
  // HBM -> L2 prefetch of the CTA's share of local row group i (no shared memory involved): issued RN_F2_PF groups
  // before the copy into the ring, which then finds the data in L2 (~250 cycles instead of an HBM round trip under load)
  auto prefetch_l2 = [&](int i) {
    if (i < NGL && lane == 0) {
      const double* src = vw.X8 + (((g0 + i) * qrow + (int64_t)b0 * 8) << 4);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)cb * 1024u) : "memory");
    }
  };
  auto produce = [&](int i) {
    const double* src = vw.X8 + (((g0 + i) * qrow + (int64_t)b0 * 8) << 4);
    const int st = i % NST;
    prefetch_l2(i + RN_F2_PF);
    const uint32_t ph = (uint32_t)((i / NST) & 1);
#if RN_F2_SPLIT
    if (nA) {
      rn_mbar_wait(&empty[st * 2], ph ^ 1u);
      if (lane == 0) {
        rn_mbar_expect_tx(&full[st * 2], (uint32_t)nA * 1024u);
        rn_bulk_g2s(ring + st * RN_F2_STAGE_BYTES, src, (uint32_t)nA * 1024u, &full[st * 2]);
      }
      __syncwarp();
    }
    if (nB) {
      rn_mbar_wait(&empty[st * 2 + 1], ph ^ 1u);
      if (lane == 0) {
        rn_mbar_expect_tx(&full[st * 2 + 1], (uint32_t)nB * 1024u);
        rn_bulk_g2s(ring + st * RN_F2_STAGE_BYTES + nA * 1024, src + nA * 128, (uint32_t)nB * 1024u, &full[st * 2 + 1]);
      }
      __syncwarp();
    }
#else
    rn_mbar_wait(&empty[st * 2], ph ^ 1u);
    if (lane == 0) {
      rn_mbar_expect_tx(&full[st * 2], (uint32_t)cb * 1024u);
      rn_bulk_g2s(ring + st * RN_F2_STAGE_BYTES, src, (uint32_t)cb * 1024u, &full[st * 2]);
    }
    __syncwarp();
#endif
  };
  // X is never written by a kernel: the ring is filled before the previous launch has finished
  const int npre = NGL < NST ? NGL : NST;
  if (warp == 12) {
    for (int i = 0; i < npre; ++i) produce(i);  // (also prefetches groups RN_F2_PF .. RN_F2_PF + npre - 1)
    for (int i = npre; i < RN_F2_PF; ++i) prefetch_l2(i);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads what the previous launch wrote
  if (ft.ctrl->done) {  // uniform over the grid; the copies in flight must land before the CTA may exit
    if (warp == 12)
      for (int i = 0; i < npre; ++i) {
        if (nA) rn_mbar_wait(&full[i * 2], 0u);
        if (RN_F2_SPLIT && nB) rn_mbar_wait(&full[i * 2 + 1], 0u);
      }
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

  double tacc[2 * NBW][2];  // consumer warps: T accumulators of the warp's columns (tile 2b+e: columns 16b+2g+e)
#pragma unroll
  for (int s = 0; s < 2 * NBW; ++s) tacc[s][0] = tacc[s][1] = 0.0;
  // tile / fragment slot b of a consumer warp: b = 0, 1 half A (blocks w_aoff + b), b = 2, 3 half B (w_boff + b - 2)
  const int64_t colA = 16 * (int64_t)(b0 + w_aoff), colB = 16 * (int64_t)(b0 + w_boff);

  if (warp == 12) {
    // ---- producer warp -------------------------------------------------------------------------------
    for (int i = npre; i < NGL; ++i) produce(i);
  } else if (warp == 14) {
    // ---- auxiliary warp: lane (g,t) <-> row g, factor columns 2t, 2t+1; runs two row groups ahead of the epilogue --
    const int V = ft.n_views;
    const int kp = vw.kp;
    const int c0 = 2 * t, c1 = 2 * t + 1;
    double phisum = 0.0;
    for (int w = 0; w < V; ++w) phisum += ft.phi[w + v * V];
    const double nv = (double)vw.n_glob;
    const double inv_nv = 1.0 / nv;
    // B operand of D = F M (B[b][c] = M[b, c]): lane (g,t) supplies B[t][g] and B[t+4][g]
    const double bm0 = Msm[t * 8 + g], bm1 = Msm[(t + 4) * 8 + g];
    const double lam0 = lamh[c0], lam1 = lamh[c1];
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
    for (int i = 0; i < NGL; ++i) {
      // the slots written below held group i-2 (old F rows) and i-4 (coupling sum): their epilogue warp is done with them
      if (i >= 2) rn_mbar_wait(&aux_empty[(i - 2) & 3], (uint32_t)(((i - 2) >> 2) & 1));
      rn_cp_async_wait1();  // batch i: old F rows + partner rows of this group, map entries of group i+2
      __syncwarp();
      prefetch_f(i + 2);
      if (coupled) {
        prefetch_gather(i + 2);
        prefetch_idx(i + 4);
      }
      rn_cp_async_commit();
      {  // denominator of update_f for the 8 rows (R/update_steps.r:149-163): (F S) W evaluated as F (S W), + lambda / 2,
         // + colSums(phi) F on the coupled branch -- nothing here depends on P, so it is off the epilogue's critical path
        const double* fo = Fo + (i & 3) * 64 + g * 8;
        const double2 fmine = *reinterpret_cast<const double2*>(fo + c0);
        double D0 = 0.0, D1 = 0.0;
        rn_dmma(D0, D1, fo[t], bm0);
        rn_dmma(D0, D1, fo[t + 4], bm1);
        if (coupled) {
          D0 += phisum * fmine.x;
          D1 += phisum * fmine.y;
        }
        *reinterpret_cast<double2*>(Dn + (i & 3) * 64 + g * 8 + c0) = make_double2(D0 + lam0, D1 + lam1);
      }
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
        // pc / nv (R/utils.r:77) as a product with the reciprocal plus one residual correction (<= 1 ulp): three dependent
        // FP64 instructions instead of the division routine's dozen -- on this warp's sub-partition, shared with three
        // DMMA streams, every dependent FP64 instruction costs ~130 cycles and the warp has one period per row group
        double q0 = pc0 * inv_nv, q1 = pc1 * inv_nv;
        q0 = fma(fma(-q0, nv, pc0), inv_nv, q0);
        q1 = fma(fma(-q1, nv, pc1), inv_nv, q1);
        *reinterpret_cast<double2*>(Pcn + (i & 3) * 64 + g * 8 + c0) = make_double2(q0, q1);
      }
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&aux_full[i & 3]);
    }
  } else if (warp == 13 || warp == 15) {
    // ---- epilogue warps (A: even local groups, B: odd ones): lane (g,t) owns row g, factor columns 2t and 2t+1 ------
    // The consumers' B operand is G t(S), so the exchanged sum IS the numerator N = X G t(S) of update_f in this warp's
    // layout; the denominators come from the auxiliary warp.  These warps share their sub-partition's FP64 pipe with
    // three DMMA streams -- every dependent FP64 instruction costs ~130 cycles there (measured) -- so the chain between
    // "last warp partial in" and "F_new out" is kept to the two summation trees and ONE multiplication: the reciprocal of
    // the denominator (hardware seed + two Newton steps, <= 1 ulp) times the old F value is formed before the partials
    // arrive.  Operands outside the safe range (zero / huge / Inf / NaN denominators, huge numerators) take the exact
    // IEEE division so that Inf / NaN behave as in R (NaN ratio -> 1 on the uncoupled branch, update_steps.r:152-155).
    const int ew = (warp == 15) ? 1 : 0;
    const int V = ft.n_views;
    const int kp = vw.kp;
    const int c0 = 2 * t, c1 = 2 * t + 1;
    const bool v0 = c0 < K, v1 = c1 < K;
    double phisum = 0.0;
    for (int w = 0; w < V; ++w) phisum += ft.phi[w + v * V];
    const bool coupled = phisum != 0.0;
    const uint32_t my_pex = rn_smem_u32(Pex + rank * 64 + 2 * lane);
    const int sig0 = rn_sigma(c0), sig1 = rn_sigma(c1);
    auto recip = [](double den, bool& safe) {  // 1 / den to <= 1 ulp for den in the safe range
      const double ad = fabs(den);
      safe = ad > 1.0e-280 && ad < 1.0e280;
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
      double e = fma(-den, r, 1.0);
      r = fma(r, e, r);
      e = fma(-den, r, 1.0);
      return fma(r, e, r);
    };
    auto slow = [&](double num, double den, double f) {  // the reference's operation sequence, exact division
      double ratio = num / den;
      if (!coupled && isnan(ratio)) ratio = 1.0;
      return fabs(f * ratio);
    };
    // Everything of a group that does not depend on its partials -- old F rows, denominator, coupling sum, and the
    // reciprocal (six dependent FP64 instructions) -- is formed for this warp's NEXT group (i + 2) in the shadow of the
    // cluster exchange of group i, after the partial has been sent: off the chain "warp partials in -> F_new out" and
    // off the warp's serial time per group (RN_F2_EPRE; measured before: 0.4 us of every 2.2 us group).
    double2 fmine, den, pcn = make_double2(0.0, 0.0);
    double w0 = 0.0, w1 = 0.0;
    bool safe0 = false, safe1 = false;
    auto prework = [&](int i) {
      RN_F2_TR(lane == 0, i, 8);
      // old F rows, denominators (+ coupling sum) of this group are in shared memory -- in particular before any peer
      // can be released to overwrite those rows in HBM (this warp's partial of the group is sent after this point)
      rn_mbar_wait(&aux_full[i & 3], (uint32_t)((i >> 2) & 1));
      RN_F2_TR(lane == 0, i, 9);
      fmine = *reinterpret_cast<const double2*>(Fo + (i & 3) * 64 + g * 8 + c0);
      den = *reinterpret_cast<const double2*>(Dn + (i & 3) * 64 + g * 8 + c0);
      if (coupled) pcn = *reinterpret_cast<const double2*>(Pcn + (i & 3) * 64 + g * 8 + c0);
      w0 = fmine.x * recip(v0 ? den.x : 1.0, safe0);
      w1 = fmine.y * recip(v1 ? den.y : 1.0, safe1);
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&aux_empty[i & 3]);
    };
#if RN_F2_EPRE
    if (ew < NGL) prework(ew);
#endif
    for (int i = ew; i < NGL; i += 2) {
      const int s3 = i % NPW, s4 = i & (NPEX - 1);
      const uint32_t ph3 = (uint32_t)((i / NPW) & 1), ph4 = (uint32_t)((i >> 2) & 1);
      if (lane == 0) rn_mbar_expect_tx(&pex_full[s4], csize * 512u);
#if !RN_F2_EPRE
      prework(i);
#endif
      const double2 fmine_c = fmine, den_c = den, pcn_c = pcn;  // this group's values (prework(i + 2) overwrites them)
      const double w0_c = w0, w1_c = w1;
      const bool safe0_c = safe0, safe1_c = safe1;
      rn_mbar_wait(&pw_full[s3], ph3);
      RN_F2_TR(lane == 0, i, 10);
      const double* pw = Pw + (s3 * NCW) * 64 + 2 * lane;
      double2 q[NCW];
#pragma unroll
      for (int w = 0; w < NCW; ++w) q[w] = *reinterpret_cast<const double2*>(pw + w * 64);
      // fixed tree: ((0+1)+(2+3)) + ((4+5)+(6+7)) + ((8+9)+(10+11))
#pragma unroll
      for (int w = 0; w < NCW; w += 2) {
        q[w].x += q[w + 1].x;
        q[w].y += q[w + 1].y;
      }
#pragma unroll
      for (int w = 0; w < NCW; w += 4) {
        q[w].x += q[w + 2].x;
        q[w].y += q[w + 2].y;
      }
      double2 acc;
      acc.x = (q[0].x + q[4].x) + q[8].x;
      acc.y = (q[0].y + q[4].y) + q[8].y;
      for (uint32_t rr = 0; rr < csize; ++rr)
        rn_st_async2(rn_mapa(my_pex + s4 * (RN_FU_MAXC * 64 * 8), rr), acc.x, acc.y,
                     rn_mapa(rn_smem_u32(&pex_full[s4]), rr));
      RN_F2_TR(lane == 0, i, 11);
#if RN_F2_EPRE
      if (i + 2 < NGL) prework(i + 2);
#endif
      rn_mbar_wait_cluster(&pex_full[s4], ph4);
      RN_F2_TR(lane == 0, i, 12);
      // CTA partials in rank order, as a fixed tree over 8 slots (absent ranks count as +0)
      double2 c[RN_FU_MAXC];
#pragma unroll
      for (int rr = 0; rr < RN_FU_MAXC; ++rr)
        c[rr] = ((uint32_t)rr < csize) ? *reinterpret_cast<const double2*>(Pex + (s4 * RN_FU_MAXC + rr) * 64 + 2 * lane)
                                      : make_double2(0.0, 0.0);
#pragma unroll
      for (int rr = 0; rr < RN_FU_MAXC; rr += 2) {
        c[rr].x += c[rr + 1].x;
        c[rr].y += c[rr + 1].y;
      }
      double N0 = (c[0].x + c[2].x) + (c[4].x + c[6].x);
      double N1 = (c[0].y + c[2].y) + (c[4].y + c[6].y);
      if (coupled) {  // star_prod_relevant term (utils.r:63-78), formed by the auxiliary warp
        N0 += pcn_c.x;
        N1 += pcn_c.y;
      }
      const int64_t grp = g0 + i;
      const int64_t r = grp * 8 + g;
      double o0 = 0.0, o1 = 0.0;
      if (r < vw.n) {
        if (v0) o0 = (safe0_c && fabs(N0) < 1.0e280) ? fabs(N0 * w0_c) : slow(N0, den_c.x, fmine_c.x);
        if (v1) o1 = (safe1_c && fabs(N1) < 1.0e280) ? fabs(N1 * w1_c) : slow(N1, den_c.y, fmine_c.y);
      }
      *reinterpret_cast<double2*>(Fp + s3 * 64 + g * 8 + c0) = make_double2(o0, o1);
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&fp_full[s3]);
      RN_F2_TR(lane == 0, i, 13);
      // off the critical path: F_new to HBM (the CTAs of the cluster take turns)
      if ((uint32_t)(i % (int)csize) == rank && r < vw.n) {
        // rn_fidx(r, c, kp) with r = 8 grp + g: 8 row groups per 64-row panel of F
        double* fpan = vw.F + (grp >> 3) * kp * RN_ROW_TILE + (g & 1);
        const int piece = (((int)(grp & 7)) << 2) | (g >> 1);
        if (v0) fpan[c0 * RN_ROW_TILE + 2 * (piece ^ sig0)] = o0;
        if (v1) fpan[c1 * RN_ROW_TILE + 2 * (piece ^ sig1)] = o1;
      }
      RN_F2_TR(lane == 0, i, 14);
    }
  } else {
    // ---- consumer warps --------------------------------------------------------------------------------
    // B operand of the F phase: fragments of G t(S) for the warp's columns, so that the exchanged partial sums are the
    // numerator X G t(S) of update_f directly (R/update_steps.r:146; evaluated as X (G t(S)))
    double gfr[2 * NBW][2];
#pragma unroll
    for (int s = 0; s < 2 * NBW; ++s)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool live = (s < 4) ? (s < 2 * w_na) : (s - 4 < 2 * w_nb);
        const int64_t col = ((s < 4) ? colA + 8 * s : colB + 8 * (s - 4)) + 2 * t + e;
        double gs = 0.0;
        if (live && col < vw.pp && g < K) {
          const double* grow = vw.G + col * KP;
#pragma unroll
          for (int b = 0; b < K; ++b) gs = fma(grow[b], Ssm[g + b * K], gs);
        }
        gfr[s][e] = gs;
      }
    // F'F and colSums(F) of this cluster's rows: consumer warp 9 of rank 0 (a warp with 3 blocks) adds them from the
    // F_new fragments it loads for the G phase anyway (A[a][r] = B[r][a] = F_new[r][a]: the same register)
    const bool ff_warp = rank == 0 && ci == 9;
    double ff0 = 0.0, ff1 = 0.0, cs0 = 0.0, cs1 = 0.0;
    const int ra = (t & 1) + 4 * (t >> 1);  // G-phase K slot t <-> rows {0,1,4,5} (first MMA), {2,3,6,7} (second)
    const int rb = ra + 2;
    const uint32_t off1 = (uint32_t)(t * 128 + ((g ^ (2 * t)) * 16));
    const uint32_t off2a = (uint32_t)(g * 128 + ((ra ^ (2 * (g & 3))) * 16));
    const uint32_t off2b = (uint32_t)(g * 128 + ((rb ^ (2 * (g & 3))) * 16));
    const bool liveA = nA > 0, liveB = RN_F2_SPLIT && nB > 0;  // a CTA with a single block has no half B
    const unsigned char* ringA = ring + w_aoff * 1024;
    const unsigned char* ringB = ring + w_boff * 1024;

    // The main loop, compiled for the block counts (NA, NB) of this warp's two halves: with compile-time trip counts the
    // fragment loads of a phase are issued ahead of its MMAs (with run-time predicates every DMMA pair waited for its own
    // LDS: ncu short_scoreboard on every DMMA).
    auto run = [&](auto na_c, auto nb_c) {
      constexpr int NA = decltype(na_c)::value, NB = decltype(nb_c)::value;
      double pe0 = 0.0, pe1 = 0.0, po0 = 0.0, po1 = 0.0;  // two accumulator chains: even / odd column of a pair
      double fa = 0.0, fb = 0.0;                           // F_new fragments of the group in its G phase
      bool x_ready = false, fn_ready = false;              // outcome of the early probes (RN_F2_PROBE)
      auto f_phase = [&](int i) {
        const int st = i % NST;
        const uint32_t ph = (uint32_t)((i / NST) & 1);
        RN_F2_TR(tid == 0, i, 0);
        if (liveA && !(RN_F2_PROBE && x_ready)) rn_mbar_wait(&full[st * 2], ph);
        asm volatile("" ::: "memory");  // nothing below is read before the (probed or awaited) phase completion
        x_ready = false;
        RN_F2_TR(tid == 0, i, 1);
        const unsigned char* xa = ringA + st * RN_F2_STAGE_BYTES + off1;
        const unsigned char* xb = ringB + st * RN_F2_STAGE_BYTES + off1;
        double2 x[2 * NA > 0 ? 2 * NA : 1];
#pragma unroll
        for (int s = 0; s < 2 * NA; ++s) x[s] = *reinterpret_cast<const double2*>(xa + s * 512);
        pe0 = pe1 = po0 = po1 = 0.0;
#pragma unroll
        for (int s = 0; s < 2 * NA; ++s) {
          rn_dmma(pe0, pe1, x[s].x, gfr[s][0]);
          rn_dmma(po0, po1, x[s].y, gfr[s][1]);
        }
        if (RN_F2_PROBE && RN_F2_FFIRST && i >= 2)  // F_new of the group whose G phase follows, probed under the MMAs
          fn_ready = rn_mbar_probe(&fp_full[(i - 2) % NPW], (uint32_t)(((i - 2) / NPW) & 1));
        if (NB > 0) {
          if (liveB) rn_mbar_wait(&full[st * 2 + 1], ph);
          double2 y[2 * NB > 0 ? 2 * NB : 1];
#pragma unroll
          for (int s = 0; s < 2 * NB; ++s) y[s] = *reinterpret_cast<const double2*>(xb + s * 512);
#pragma unroll
          for (int s = 0; s < 2 * NB; ++s) {
            rn_dmma(pe0, pe1, y[s].x, gfr[4 + s][0]);
            rn_dmma(po0, po1, y[s].y, gfr[4 + s][1]);
          }
        }
      };
      auto publish = [&](int i) {
        *reinterpret_cast<double2*>(Pw + ((i % NPW) * NCW + ci) * 64 + 2 * lane) = make_double2(pe0 + po0, pe1 + po1);
        __syncwarp();
        if (lane == 0) rn_mbar_arrive(&pw_full[i % NPW]);
        RN_F2_TR(tid == 0, i, 2);
        RN_F2_TR(lane == 0, i, 16 + ci);
      };
      auto g_phase_a = [&](int i) {
        const int st = i % NST;
        const unsigned char* xs = ringA + st * RN_F2_STAGE_BYTES;
        double2 xa[NA > 0 ? NA : 1], xb[NA > 0 ? NA : 1];  // X does not depend on F_new: loaded before the wait
#pragma unroll
        for (int b = 0; b < NA; ++b) {
          xa[b] = *reinterpret_cast<const double2*>(xs + off2a + b * 1024);
          xb[b] = *reinterpret_cast<const double2*>(xs + off2b + b * 1024);
        }
        RN_F2_TR(tid == 0, i, 3);
        if (!(RN_F2_PROBE && fn_ready)) rn_mbar_wait(&fp_full[i % NPW], (uint32_t)((i / NPW) & 1));
        asm volatile("" ::: "memory");
        fn_ready = false;
        RN_F2_TR(tid == 0, i, 4);
        fa = Fp[(i % NPW) * 64 + ra * 8 + g];
        fb = Fp[(i % NPW) * 64 + rb * 8 + g];
        if (ff_warp) {
          rn_dmma(ff0, ff1, fa, fa);
          rn_dmma(cs0, cs1, 1.0, fa);
          rn_dmma(ff0, ff1, fb, fb);
          rn_dmma(cs0, cs1, 1.0, fb);
        }
#pragma unroll
        for (int b = 0; b < NA; ++b) {
          rn_dmma(tacc[2 * b][0], tacc[2 * b][1], xa[b].x, fa);
          rn_dmma(tacc[2 * b + 1][0], tacc[2 * b + 1][1], xa[b].y, fa);
          rn_dmma(tacc[2 * b][0], tacc[2 * b][1], xb[b].x, fb);
          rn_dmma(tacc[2 * b + 1][0], tacc[2 * b + 1][1], xb[b].y, fb);
        }
        __syncwarp();
        if (RN_F2_SPLIT && liveA && lane == 0) rn_mbar_arrive(&empty[st * 2]);
        if (RN_F2_PROBE && RN_F2_FFIRST && liveA && i + 3 < NGL)  // X of the group whose F phase follows this G phase
          x_ready = rn_mbar_probe(&full[((i + 3) % NST) * 2], (uint32_t)(((i + 3) / NST) & 1));
      };
      auto g_phase_b = [&](int i) {
        const int st = i % NST;
        const unsigned char* xs = ringB + st * RN_F2_STAGE_BYTES;
        double2 xa[NB > 0 ? NB : 1], xb[NB > 0 ? NB : 1];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          xa[b] = *reinterpret_cast<const double2*>(xs + off2a + b * 1024);
          xb[b] = *reinterpret_cast<const double2*>(xs + off2b + b * 1024);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          rn_dmma(tacc[4 + 2 * b][0], tacc[4 + 2 * b][1], xa[b].x, fa);
          rn_dmma(tacc[4 + 2 * b + 1][0], tacc[4 + 2 * b + 1][1], xa[b].y, fa);
          rn_dmma(tacc[4 + 2 * b][0], tacc[4 + 2 * b][1], xb[b].x, fb);
          rn_dmma(tacc[4 + 2 * b + 1][0], tacc[4 + 2 * b + 1][1], xb[b].y, fb);
        }
        __syncwarp();
        if (RN_F2_SPLIT) {
          if (liveB && lane == 0) rn_mbar_arrive(&empty[st * 2 + 1]);
        } else {
          if (liveA && lane == 0) rn_mbar_arrive(&empty[st * 2]);
        }
        RN_F2_TR(tid == 0, i, 5);
      };
      // F phase two groups ahead, placed between the two halves of a G phase (see the copy slots above)
      for (int i = 0; i < 2 && i < NGL; ++i) {
        f_phase(i);
        publish(i);
      }
#if RN_F2_STAGGER
      // The warps of a sub-partition do not all run the same phase at the same time: the middle one of its three consumer
      // warps (ci 4..7) does the whole G phase of group i first and the F phase of group i+2 after it, the other two the
      // F phase first.  All warps wait for the same two events per group (X there, F_new there); staggered, the
      // sub-partition's FP64 pipe has another warp's MMAs to run while one of them sits in such a wait.
      if ((ci >> 2) == 1) {
        for (int i = 0; i < NGL; ++i) {
          g_phase_a(i);
          g_phase_b(i);
          if (i + 2 < NGL) {
            f_phase(i + 2);
            publish(i + 2);
          }
        }
        return;
      }
#endif
      for (int i = 0; i < NGL; ++i) {
#if RN_F2_FFIRST
        if (i + 2 < NGL) {
          f_phase(i + 2);
          publish(i + 2);
        }
        g_phase_a(i);
#else
        g_phase_a(i);
        if (i + 2 < NGL) {
          f_phase(i + 2);
          publish(i + 2);
        }
#endif
        g_phase_b(i);
      }
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    if (w_na == 2 && w_nb == 2) run(I2{}, I2{});
    else if (w_na == 2 && w_nb == 1) run(I2{}, I1{});
    else if (w_na == 1 && w_nb == 2) run(I1{}, I2{});
    else if (w_na == 1 && w_nb == 1) run(I1{}, I1{});
    else if (w_na == 1 && w_nb == 0) run(I1{}, I0{});
    else if (w_na == 0 && w_nb == 1) run(I0{}, I1{});
    else if (w_na == 2 && w_nb == 0) run(I2{}, I0{});
    else if (w_na == 0 && w_nb == 2) run(I0{}, I2{});
    else run(I0{}, I0{});
    if (ff_warp) {
      const int c0 = 2 * t, c1 = 2 * t + 1;
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
  }
  if (tid == 0) rn_fu_stamp(vw, 2);
  rn_cluster_sync();  // every st.async of this cluster has landed before any of its CTAs may exit
  if (!is_consumer) return;
  if (tid == 0) rn_fu_stamp(vw, 3);

  // ---- tail (consumer warps): publish T partials, then the column-group epilogues ------------------------
  {
    double* tpa = vw.Tpart + (cid * vw.pp8 + colA) * KP;
    double* tpb = vw.Tpart + (cid * vw.pp8 + colB) * KP;
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (b < w_na)
          *reinterpret_cast<double2*>(tpa + (16 * b + 2 * g + e) * KP + 2 * t) =
              make_double2(tacc[2 * b + e][0], tacc[2 * b + e][1]);
        if (b < w_nb)
          *reinterpret_cast<double2*>(tpb + (16 * b + 2 * g + e) * KP + 2 * t) =
              make_double2(tacc[4 + 2 * b + e][0], tacc[4 + 2 * b + e][1]);
      }
  }
  __threadfence();
  rn_f2_consumer_sync();
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
  rn_f2_consumer_sync();
  __threadfence();
  if (ctid == 0) rn_fu_stamp(vw, 4);
  bool ff_ready = false;
  const int64_t tstride = vw.pp8 * KP;
  for (int64_t grp = blockIdx.x; grp < NG; grp += gridDim.x) {
    const int64_t j0 = grp * RN_FU_TG;
    const int njb = (int)min((int64_t)(RN_FU_TG / 8), (pp - j0) >> 3);
    rn_f2_consumer_sync();  // previous group's epilogue is done with Ts / Gs
    for (int i = ctid; i < RN_FU_TG * KP; i += NCT)
      Ts[i] = (i < 8 * njb * KP) ? rn_sum_wide(vw.Tpart + j0 * KP + i, tstride, (int)n_clusters) : 0.0;
    if (!ff_ready) {
      if (ctid < NFF) FtFs[ctid] = rn_sum_wide(vw.FFpart + ctid, NFF, (int)n_clusters);
      rn_f2_consumer_sync();
      for (int o = ctid; o < KK; o += NCT) {  // V = crossprod(F) %*% S
        const int a = o % K, c = o / K;
        double s = 0.0;
        for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
        Vs[a + c * K] = s;
      }
      ff_ready = true;
    }
    rn_f2_consumer_sync();
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
    rn_f2_consumer_sync();
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
    rn_f2_consumer_sync();
    if (ctid == 0) *s_flag = (atomicAdd(&vw.misc_ticket[0], 1) == (int)NG - 1);
    rn_f2_consumer_sync();
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
      rn_f2_consumer_sync();
      if (ctid < NOUT) fin[ctid] = part + red[ctid];
    }
    rn_f2_consumer_sync();
    if (ctid == 0) rn_fu_stamp(vw, 8);
    rn_view_finish<K, NCT, true>(vw, ft, v, ctid, fin, FtFs, Ssm, Us, Sn, red, fuse_finish);
    if (ctid == 0) rn_fu_stamp(vw, 7);
  }
}
