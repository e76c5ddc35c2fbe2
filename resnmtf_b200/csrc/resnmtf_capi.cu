// Host side of the C ABI declared in include/resnmtf_b200.h: owns the device-resident X, F, S, G,
// lambda, mu of every view of a fit, the per-iteration CUDA graph and the convergence loop of
// R/main.r:50-109.  No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <thread>

#include "rn_host.h"
#include "rn_kernels.cuh"
#include "rn_fused.cuh"
#include "rn_fused2.cuh"
#include "rn_post.cuh"

// ------------------------------------------------------------------------------------------------
// kernel dispatch on k
// ------------------------------------------------------------------------------------------------
#define RN_K_CASES_LE8(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#define RN_K_CASES_GT8(X) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#include "rn_small.cuh"

typedef void (*FStepSkFn)(const RnView, const RnFit, const int);
typedef void (*GStepSkFn)(const RnView, const RnFit, const int, const int);

static FStepSkFn f_step_sk_fn(int K) {
  switch (K) {
#define X(KC) case KC: return rn_f_step_sk<KC>;
    RN_K_CASES_LE8(X)
#undef X
  }
  return nullptr;
}
static GStepSkFn g_step_sk_fn(int K) {
  switch (K) {
#define X(KC) case KC: return rn_g_step_sk<KC>;
    RN_K_CASES_LE8(X)
#undef X
  }
  return nullptr;
}

FStepSkFn rn_f_step_tma_gt8(int K);  // k = 9..16: rn_tma_gt8.cu
GStepSkFn rn_g_step_tma_gt8(int K);

static FStepSkFn f_step_tma_fn(int K) {
  if (K > 8) return rn_f_step_tma_gt8(K);
  switch (K) {
#define X(KC) case KC: return rn_f_step_tma<KC>;
    RN_K_CASES_LE8(X)
#undef X
  }
  return nullptr;
}
static GStepSkFn g_step_tma_fn(int K) {
  if (K > 8) return rn_g_step_tma_gt8(K);
  switch (K) {
#define X(KC) case KC: return rn_g_step_tma<KC>;
    RN_K_CASES_LE8(X)
#undef X
  }
  return nullptr;
}

static GStepSkFn fused_step_fn(int K, int kind) {
  if (kind == 2) {
    switch (K) {
#define X(KC) case KC: return rn_fused2_step<KC>;
      RN_K_CASES_LE8(X)
#undef X
    }
    return nullptr;
  }
  switch (K) {
#define X(KC) case KC: return rn_fused_step<KC>;
    RN_K_CASES_LE8(X)
#undef X
  }
  return nullptr;
}
static inline size_t fused_smem(int kind) { return kind == 2 ? rn_fused2_smem() : rn_fused_smem(); }
static inline int fused_threads(int kind) { return kind == 2 ? RN_F2_THREADS : RN_FU_THREADS; }

// One-pass fused update of a view: cluster launch (fu_csize CTAs per cluster, fu_clusters clusters).
static void launch_fused(const ViewHost& vh, const RnFit& ft, int v, int fuse, cudaStream_t st, bool allow_pdl = true) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(vh.d.fu_clusters * vh.d.fu_csize));
  cfg.blockDim = dim3((unsigned)fused_threads(vh.d.fu_kind));
  cfg.dynamicSmemBytes = fused_smem(vh.d.fu_kind);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)vh.d.fu_csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // Programmatic dependent launch (RESNMTF_NO_PDL=1 switches it off): this launch may be placed on the SMs while the
  // previous kernel of the stream drains; the kernel waits (griddepcontrol.wait) before it reads anything a kernel
  // wrote.  Captured into the iteration graph as a programmatic edge.
  static const bool pdl = rn_env_int("RESNMTF_NO_PDL", 0) == 0;
  if (pdl && allow_pdl) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaLaunchKernelEx(&cfg, fused_step_fn(vh.d.k, vh.d.fu_kind), vh.d, ft, v, fuse);
}

// Cluster size of the fused path for a view, 0 when it does not qualify: k <= 8, not row-sharded, p within
// 8 x 1008 columns and not padded by more than RESNMTF_FUSED_MAX_PAD percent (default 135: the one-pass kernel
// costs ~0.68 of the two-pass pair per padded column).
static int fused_csize(const RnView& d, bool sharded) {
  if (d.k > 8 || sharded) return 0;
  const double max_pad = rn_env_int("RESNMTF_FUSED_MAX_PAD", 135) / 100.0;
  for (int c = 1; c <= RN_FU_MAXC; ++c)
    if ((int64_t)c * RN_FU_CCOLS >= d.p) return ((double)c * RN_FU_CCOLS <= max_pad * (double)d.p) ? c : 0;
  return 0;
}

// Most clusters of `csz` CTAs of the fused kernel the device keeps resident at once (0: the cluster does not fit).
static int fused_max_clusters(int K, int csz, int sms, int kind) {
  GStepSkFn fn = fused_step_fn(K, kind);
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(kind)) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(csz * sms));
  cfg.blockDim = dim3((unsigned)fused_threads(kind));
  cfg.dynamicSmemBytes = fused_smem(kind);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csz;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// Launch of a streaming kernel with (optionally) a persisting-L2 access-policy window over the head of X.
template <typename... Args>
static void launch_windowed(void (*fn)(Args...), int grid, int block, size_t smem, cudaStream_t st, const void* win_ptr,
                            size_t win_bytes, Args... args) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (win_bytes > 0) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(win_ptr);
    attr[0].val.accessPolicyWindow.num_bytes = win_bytes;
    attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaLaunchKernelEx(&cfg, fn, args...);
}

// FP64 tensor-core kernels: the register-path pair (RESNMTF_IMPL_DMMA) serves k <= 8, the TMA pair k <= 16 (two 8-wide
// tiles in the factor dimension for k = 9..16; RESNMTF_TMA_GT8=0 sends those k back to the CUDA-core kernels)
static inline bool use_mma(const ViewHost& vh, int impl) {
  if (impl == RESNMTF_IMPL_DMMA) return vh.d.k <= 8;
  if (impl != RESNMTF_IMPL_TMA && impl != RESNMTF_IMPL_FUSED) return false;
  static const bool gt8 = rn_env_int("RESNMTF_TMA_GT8", 1) != 0;
  return vh.d.k <= 8 || gt8;
}

static void launch_f_step(const ViewHost& vh, const RnFit& ft, int v, int impl, cudaStream_t st) {
  const int K = vh.d.k;
  if (use_mma(vh, impl)) {
    if (impl != RESNMTF_IMPL_DMMA)
      launch_windowed(f_step_tma_fn(K), vh.d.f_ctas, RN_TMA_THREADS, rn_f_tma_smem(K), st, vh.d.X, vh.l2_window,
                      vh.d, ft, v);
    else
      f_step_sk_fn(K)<<<vh.d.f_ctas, 256, 0, st>>>(vh.d, ft, v);
    return;
  }
  dim3 grid(vh.d.row_tiles, vh.d.cs);
  switch (K) {
#define X(KC) case KC: rn_f_step_dfma<KC, 8><<<grid, 256, 0, st>>>(vh.d, ft, v); break;
    RN_K_CASES_LE8(X)
#undef X
#define X(KC) case KC: rn_f_step_dfma<KC, 16><<<grid, 256, 0, st>>>(vh.d, ft, v); break;
    RN_K_CASES_GT8(X)
#undef X
  }
}

static void launch_g_epilogue(const ViewHost& vh, const RnFit& ft, int v, int fuse, cudaStream_t st) {
  const int K = vh.d.k;
  switch (K) {
#define X(KC) case KC: rn_g_epilogue<KC, 8><<<vh.d.gepi_ctas, RN_GEPI_THREADS(KC), 0, st>>>(vh.d, ft, v, fuse); break;
    RN_K_CASES_LE8(X)
#undef X
#define X(KC) case KC: rn_g_epilogue<KC, 16><<<vh.d.gepi_ctas, RN_GEPI_THREADS(KC), 0, st>>>(vh.d, ft, v, fuse); break;
    RN_K_CASES_GT8(X)
#undef X
  }
}

// number of stream-K CTAs that stream part of column group 0 (the F'F contributors) -- host twin of the
// `nffc` the G-step kernels compute from RnSplit
static int ff_contributors(const RnView& d) {
  const int64_t NS = d.row_tiles, U = NS * d.col_groups, C = d.g_ctas;
  const int64_t q = U / C, rem = U % C, u = NS - 1, cut = rem * (q + 1);
  return (int)(u < cut ? u / (q + 1) : rem + (u - cut) / q) + 1;
}

// G step of one view: returns the number of kernels launched.  `ctx` has joined a communicator for a
// row-sharded view: the stream kernel then stops after T, [T | F'F | colSums(F)] is all-reduced over the ranks
// and the stand-alone epilogue finishes the view on every rank redundantly (no broadcast needed).
static int launch_g_step(const ViewHost& vh, const RnFit& ft, int v, int impl, int fuse, cudaStream_t st,
                         resnmtf_ctx* ctx, int* comm_rc) {
  const int K = vh.d.k;
  const bool sharded = ctx && ctx->comm;
  if (use_mma(vh, impl) && !sharded) {
    if (impl != RESNMTF_IMPL_DMMA)
      launch_windowed(g_step_tma_fn(K), vh.d.g_ctas, RN_TMA_THREADS, rn_g_tma_smem(K), st, vh.d.X, vh.l2_window,
                      vh.d, ft, v, fuse);
    else
      g_step_sk_fn(K)<<<vh.d.g_ctas, 128, 0, st>>>(vh.d, ft, v, fuse);
    return 1;
  }
  if (sharded) {
    int n = 0, count;
    if (use_mma(vh, impl)) {
      if (impl != RESNMTF_IMPL_DMMA)
        launch_windowed(g_step_tma_fn(K), vh.d.g_ctas, RN_TMA_THREADS, rn_g_tma_smem(K), st, vh.d.X, vh.l2_window,
                        vh.d, ft, v, -1);
      else
        g_step_sk_fn(K)<<<vh.d.g_ctas, 128, 0, st>>>(vh.d, ft, v, -1);
      count = ff_contributors(vh.d);
      n = 1;
    } else {
      dim3 grid(vh.d.col_groups, vh.d.rs);
      switch (K) {
#define X(KC) case KC: rn_gram_f<KC><<<vh.d.nff, 256, 0, st>>>(vh.d, ft); rn_g_stream_dfma<KC, 8><<<grid, 256, 0, st>>>(vh.d, ft); break;
        RN_K_CASES_LE8(X)
#undef X
#define X(KC) case KC: rn_gram_f<KC><<<vh.d.nff, 256, 0, st>>>(vh.d, ft); rn_g_stream_dfma<KC, 16><<<grid, 256, 0, st>>>(vh.d, ft); break;
        RN_K_CASES_GT8(X)
#undef X
      }
      count = vh.d.nff;
      n = 2;
    }
    rn_pack_ff<<<1, 128, 0, st>>>(vh.d, ft, count);
    const size_t words = (size_t)vh.d.pp * vh.d.kp + (size_t)K * K + K;
    const int rc = rn_allreduce(ctx, vh.d.T, words);
    if (rc && comm_rc) *comm_rc = rc;
    ViewHost tmp = vh;  // the epilogue reads the reduced F'F | colSums(F) as a single "partial"
    tmp.d.FFpart = vh.d.T + (size_t)vh.d.pp * vh.d.kp;
    tmp.d.nff = 1;
    launch_g_epilogue(tmp, ft, v, fuse, st);
    return n + 3;
  }
  dim3 grid(vh.d.col_groups, vh.d.rs);
  switch (K) {
#define X(KC) case KC: rn_gram_f<KC><<<vh.d.nff, 256, 0, st>>>(vh.d, ft); rn_g_stream_dfma<KC, 8><<<grid, 256, 0, st>>>(vh.d, ft); break;
    RN_K_CASES_LE8(X)
#undef X
#define X(KC) case KC: rn_gram_f<KC><<<vh.d.nff, 256, 0, st>>>(vh.d, ft); rn_g_stream_dfma<KC, 16><<<grid, 256, 0, st>>>(vh.d, ft); break;
    RN_K_CASES_GT8(X)
#undef X
  }
  launch_g_epilogue(vh, ft, v, fuse, st);
  return 3;
}

static void launch_residual_kernel(const ViewHost& vh, const RnFit& ft, int force, cudaStream_t st) {
  dim3 grid(vh.d.row_tiles, vh.d.resid_cs);
  const int K = vh.d.k;
  switch (K) {
#define X(KC) case KC: rn_residual<KC, 8><<<grid, 256, 0, st>>>(vh.d, ft, force); break;
    RN_K_CASES_LE8(X)
#undef X
#define X(KC) case KC: rn_residual<KC, 16><<<grid, 256, 0, st>>>(vh.d, ft, force); break;
    RN_K_CASES_GT8(X)
#undef X
  }
}

// direct error of one view; row-sharded: local sum of squares -> all-reduce -> scale.  Returns launches.
static int launch_residual(const ViewHost& vh, const RnFit& ft, int force, cudaStream_t st, resnmtf_ctx* ctx,
                           int* comm_rc) {
  launch_residual_kernel(vh, ft, force, st);
  if (!(ctx && ctx->comm)) return 1;
  const int rc = rn_allreduce(ctx, vh.d.scal + 3, 1);
  if (rc && comm_rc) *comm_rc = rc;
  rn_residual_scale<<<1, 1, 0, st>>>(vh.d, ft, force);
  return 3;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
std::string& rn_err_slot() {
  static thread_local std::string g_err = "";
  return g_err;
}
extern "C" const char* resnmtf_last_error(void) { return rn_err_slot().c_str(); }
extern "C" const char* resnmtf_version(void) { return "resnmtf_b200 0.1.0 (sm_100a)"; }

extern "C" int resnmtf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int resnmtf_ctx_create(int device, resnmtf_ctx** out) {
  RN_CHECK(out != nullptr, RESNMTF_E_INVALID, "resnmtf_ctx_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return rn_fail(RESNMTF_E_CUDA,
                   "resnmtf_ctx_create: no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0) RN_CUDA(cudaGetDevice(&device));
  RN_CHECK(device < n, RESNMTF_E_INVALID, "resnmtf_ctx_create: device index out of range");
  RN_CUDA(cudaSetDevice(device));
  resnmtf_ctx* c = new (std::nothrow) resnmtf_ctx();
  RN_CHECK(c != nullptr, RESNMTF_E_NOMEM, "resnmtf_ctx_create: out of host memory");
  c->device = device;
  cudaDeviceProp prop;
  RN_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  // Persisting L2 (experiment knob, OFF by default): pin the head of every view's X in L2 across the two
  // passes of an iteration with an access-policy window on the streaming kernels.  Measured on B200: no gain
  // for the TMA bulk loads, and the 79 MB set-aside alone slows the G step by 6 % (116 -> 123 us at C2, k=8).
  if (rn_env_int("RESNMTF_L2_PERSIST", 0) && prop.persistingL2CacheMaxSize > 0 &&
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) == cudaSuccess) {
    c->l2_persist_bytes = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)prop.accessPolicyMaxWindowSize);
  } else {
    cudaGetLastError();
  }
  RN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  RN_CUDA(cudaEventCreate(&c->ev0));
  RN_CUDA(cudaEventCreate(&c->ev1));
  if (rn_env_int("RESNMTF_NO_POOL", 0) == 0) {
    int pools = 0;
    if (cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && pools) {
      cudaMemPoolProps props;
      std::memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      uint64_t keep = UINT64_MAX;
      if (cudaMemPoolCreate(&c->pool, &props) == cudaSuccess) {
        if (cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess) {
          c->pooled = true;
        } else {
          cudaMemPoolDestroy(c->pool);
          c->pool = nullptr;
        }
      }
    }
    cudaGetLastError();
  }
  *out = c;
  return RESNMTF_OK;
}

void rn_ctx_release(resnmtf_ctx* ctx) {
  if (!ctx || ctx->refs.fetch_sub(1) != 1) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) rn_nccl().CommDestroy(ctx->comm);
  for (int b = 0; b < 2; ++b) {
    if (ctx->stage[b]) cudaFree(ctx->stage[b]);
    if (ctx->copied[b]) cudaEventDestroy(ctx->copied[b]);
    if (ctx->tiled[b]) cudaEventDestroy(ctx->tiled[b]);
  }
  if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
  for (int t = 0; t < RN_UP_T; ++t) {
    for (int b = 0; b < 2; ++b) {
      if (ctx->up_pin[t][b]) cudaFreeHost(ctx->up_pin[t][b]);
      if (ctx->up_ev[t][b]) cudaEventDestroy(ctx->up_ev[t][b]);
    }
    if (ctx->up_st[t]) cudaStreamDestroy(ctx->up_st[t]);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);  // every block of the private pool goes back to the device
  delete ctx;
}

extern "C" int resnmtf_ctx_destroy(resnmtf_ctx* ctx) {
  rn_ctx_release(ctx);  // fits / data handles still alive keep the context until the last of them is destroyed
  return RESNMTF_OK;
}

extern "C" void* resnmtf_ctx_stream(resnmtf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int resnmtf_ctx_device(resnmtf_ctx* ctx) { return ctx ? ctx->device : -1; }

extern "C" int resnmtf_ctx_synchronize(resnmtf_ctx* ctx) {
  RN_CHECK(ctx != nullptr, RESNMTF_E_INVALID, "resnmtf_ctx_synchronize: ctx is NULL");
  RN_CUDA(cudaSetDevice(ctx->device));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------
// fit: creation, data, factors, restrictions
// ------------------------------------------------------------------------------------------------
// Peer access between the GPUs of a placed fit: direct loads / stores over NVLink into the other context's memory,
// including what comes out of its private stream-ordered pool (pool memory needs its own access grant).
static int enable_peers(const std::vector<resnmtf_ctx*>& ctxs) {
  for (resnmtf_ctx* a : ctxs)
    for (resnmtf_ctx* b : ctxs) {
      if (a == b || a->device == b->device) continue;
      int can = 0;
      RN_CUDA(cudaDeviceCanAccessPeer(&can, b->device, a->device));
      RN_CHECK(can, RESNMTF_E_CUDA, "resnmtf_fit_create_placed: the GPUs of the fit cannot access each other's memory");
      RN_CUDA(cudaSetDevice(b->device));  // b reads / writes a's memory
      cudaError_t e = cudaDeviceEnablePeerAccess(a->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
      } else {
        RN_CUDA(e);
      }
      if (a->pooled) {
        cudaMemAccessDesc desc;
        std::memset(&desc, 0, sizeof(desc));
        desc.location.type = cudaMemLocationTypeDevice;
        desc.location.id = b->device;
        desc.flags = cudaMemAccessFlagsProtReadWrite;
        RN_CUDA(cudaMemPoolSetAccess(a->pool, &desc, 1));
      }
    }
  return RESNMTF_OK;
}

static int fit_create_common(resnmtf_ctx* ctx, resnmtf_ctx* const* view_ctx, int n_views, const int64_t* n,
                             const int64_t* p, const int32_t* k, resnmtf_fit** out) {
  RN_CHECK(ctx && n && p && k && out, RESNMTF_E_INVALID, "resnmtf_fit_create: NULL argument");
  RN_CHECK(n_views >= 1, RESNMTF_E_INVALID, "resnmtf_fit_create: n_views must be >= 1");
  for (int v = 0; v < n_views; ++v) {
    RN_CHECK(n[v] >= 1 && p[v] >= 1, RESNMTF_E_INVALID, "resnmtf_fit_create: empty view");
    RN_CHECK(k[v] >= 1 && k[v] <= RESNMTF_MAX_K, RESNMTF_E_UNSUPPORTED,
             "resnmtf_fit_create: k must be in 1..16");
    RN_CHECK(n[v] < (int64_t)1 << 31 && p[v] < (int64_t)1 << 31, RESNMTF_E_UNSUPPORTED,
             "resnmtf_fit_create: a view dimension exceeds the int32 gather-map range");
  }
  RN_CUDA(cudaSetDevice(ctx->device));
  resnmtf_fit* f = new (std::nothrow) resnmtf_fit();
  RN_CHECK(f != nullptr, RESNMTF_E_NOMEM, "resnmtf_fit_create: out of host memory");
  f->ctx = f->cur = ctx;
  ctx->refs.fetch_add(1);
  f->V = n_views;
  f->views.resize(n_views);
  int rc = RESNMTF_OK;
  auto fail = [&](int code) {
    resnmtf_fit_destroy(f);
    return code;
  };
  const int V = n_views;
  if (view_ctx) {  // placed fit: the distinct contexts (home first), every one referenced by the fit
    f->vctx.assign(view_ctx, view_ctx + V);
    f->vdev.assign(V, 0);
    std::vector<resnmtf_ctx*> distinct{ctx};
    for (int v = 0; v < V; ++v) {
      auto it = std::find(distinct.begin(), distinct.end(), view_ctx[v]);
      if (it == distinct.end()) {
        distinct.push_back(view_ctx[v]);
        view_ctx[v]->refs.fetch_add(1);
        it = distinct.end() - 1;
      }
      f->vdev[v] = (int)(it - distinct.begin());
    }
    f->devs.resize(distinct.size());
    for (size_t i = 0; i < distinct.size(); ++i) f->devs[i].ctx = distinct[i];
    if ((rc = enable_peers(distinct))) return fail(rc);
    cudaSetDevice(ctx->device);
    f->vev.assign(V, nullptr);
    for (int v = 0; v < V; ++v) {
      RnViewScope scope(f, v);
      if (cudaEventCreateWithFlags(&f->vev[v], cudaEventDisableTiming) != cudaSuccess) {
        rn_fail(RESNMTF_E_CUDA, "resnmtf_fit_create_placed: cudaEventCreate failed");
        return fail(RESNMTF_E_CUDA);
      }
    }
  }
  for (int v = 0; v < V && rc == RESNMTF_OK; ++v) {
    RnViewScope scope(f, v);
    ViewHost& vh = f->views[v];
    std::memset(&vh.d, 0, sizeof(RnView));
    vh.d.n = n[v];
    vh.d.p = p[v];
    vh.d.k = k[v];
    vh.d.kp = k[v] <= 8 ? 8 : 16;
    vh.d.n_glob = n[v];
    vh.d.ldx = rn_round_up(n[v], RN_ROW_TILE);
    vh.d.pp = rn_round_up(p[v], 32);
    vh.d.row_tiles = (int)(vh.d.ldx / RN_ROW_TILE);
    vh.d.sharded = ctx->comm ? 1 : 0;
    vh.rowmaps.assign(V, nullptr);
    vh.colmaps.assign(V, nullptr);
    const int K = k[v], KP = vh.d.kp;
    // X is allocated by set_data (or borrowed from a shared resnmtf_data handle by attach_data)
    if ((rc = rn_alloc(f, &vh.d.F, (size_t)vh.d.ldx * KP))) break;
    if ((rc = rn_alloc(f, &vh.d.G, (size_t)vh.d.pp * KP))) break;
    if ((rc = rn_alloc(f, &vh.d.T, (size_t)vh.d.pp * KP + K * K + K))) break;
    double* small = nullptr;
    if ((rc = rn_alloc(f, &small, (size_t)4 * K * K + 4 * K + 8))) break;
    vh.d.S = small;
    vh.d.FtF = small + K * K;
    vh.d.GtG = small + 2 * K * K;
    vh.d.A = small + 3 * K * K;
    vh.d.lam = small + 4 * K * K;
    vh.d.mu = vh.d.lam + K;
    vh.d.csF = vh.d.mu + K;
    vh.d.csG = vh.d.csF + K;
    vh.d.scal = vh.d.csG + K;
    if ((rc = rn_alloc(f, &vh.d.misc_ticket, 4))) break;
    if ((rc = rn_alloc(f, &vh.d.flags, 4))) break;
    if ((rc = rn_alloc(f, &vh.xpart, 1024))) break;
    if ((rc = rn_alloc(f, &vh.xticket, 1))) break;
  }
  if (rc) return fail(rc);
  for (size_t i = 1; i < f->devs.size(); ++i) {  // metadata copies on the other GPUs of a placed fit
    RnDevMeta& m = f->devs[i];
    resnmtf_ctx* saved = f->cur;
    f->cur = m.ctx;
    cudaSetDevice(m.ctx->device);
    if (!rc) rc = rn_alloc(f, &m.d_views, (size_t)V);
    if (!rc) rc = rn_alloc(f, &m.d_phi, (size_t)V * V);
    if (!rc) rc = rn_alloc(f, &m.d_xi, (size_t)V * V);
    if (!rc) rc = rn_alloc(f, &m.d_psi, (size_t)V * V);
    if (!rc) rc = rn_alloc(f, &m.d_rowmap, (size_t)V * V);
    if (!rc) rc = rn_alloc(f, &m.d_colmap, (size_t)V * V);
    if (!rc) rc = rn_alloc(f, &m.d_rowmode, (size_t)V * V);
    if (!rc) rc = rn_alloc(f, &m.d_colmode, (size_t)V * V);
    f->cur = saved;
    cudaSetDevice(saved->device);
    if (rc) return fail(rc);
  }
  if ((rc = rn_alloc(f, &f->d_views, (size_t)V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_ctrl, 1))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_phi, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_xi, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_psi, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_rowmap, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_colmap, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_rowmode, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_colmode, (size_t)V * V))) return fail(rc);
  if ((rc = rn_alloc(f, &f->d_hist, (size_t)f->hist_cap))) return fail(rc);
  f->h_rowmap.assign((size_t)V * V, nullptr);
  f->h_colmap.assign((size_t)V * V, nullptr);
  f->h_rowmode.assign((size_t)V * V, RN_MODE_NULL);
  f->h_colmode.assign((size_t)V * V, RN_MODE_NULL);
  f->h_phi.assign((size_t)V * V, 0.0);
  f->h_xi.assign((size_t)V * V, 0.0);
  f->h_psi.assign((size_t)V * V, 0.0);
  std::memset(&f->d, 0, sizeof(RnFit));
  std::memset(&f->h_ctrl, 0, sizeof(RnCtrl));
  if (ctx->comm) {  // row-sharded views: n is local, the coupling weights need the global row count
    int64_t* d_n = nullptr;
    if ((rc = rn_alloc(f, &d_n, (size_t)V))) return fail(rc);
    std::vector<int64_t> hn(n, n + V);
    cudaError_t e = cudaMemcpyAsync(d_n, hn.data(), V * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && (rc = rn_allreduce(ctx, d_n, (size_t)V, true))) return fail(rc);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hn.data(), d_n, V * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      rn_fail(RESNMTF_E_CUDA, std::string("resnmtf_fit_create: ") + cudaGetErrorString(e));
      return fail(RESNMTF_E_CUDA);
    }
    for (int v = 0; v < V; ++v) f->views[v].d.n_glob = hn[v];
  }
  if (!f->devs.empty()) {  // devs[0] mirrors the home context's tables
    RnDevMeta& m = f->devs[0];
    m.d_views = f->d_views;
    m.d_phi = f->d_phi;
    m.d_xi = f->d_xi;
    m.d_psi = f->d_psi;
    m.d_rowmap = f->d_rowmap;
    m.d_colmap = f->d_colmap;
    m.d_rowmode = f->d_rowmode;
    m.d_colmode = f->d_colmode;
  }
  *out = f;
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_create(resnmtf_ctx* ctx, int n_views, const int64_t* n, const int64_t* p,
                                  const int32_t* k, resnmtf_fit** out) {
  return fit_create_common(ctx, nullptr, n_views, n, p, k, out);
}

extern "C" int resnmtf_fit_create_placed(resnmtf_ctx* const* view_ctx, int n_views, const int64_t* n, const int64_t* p,
                                         const int32_t* k, resnmtf_fit** out) {
  RN_CHECK(view_ctx && n_views >= 1, RESNMTF_E_INVALID, "resnmtf_fit_create_placed: NULL argument");
  bool one = true;
  for (int v = 0; v < n_views; ++v) {
    RN_CHECK(view_ctx[v] != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_create_placed: NULL context");
    RN_CHECK(view_ctx[v]->comm == nullptr, RESNMTF_E_INVALID,
             "resnmtf_fit_create_placed: a row-sharded context cannot hold a placed view");
    one = one && view_ctx[v] == view_ctx[0];
  }
  // all views on one context: an ordinary fit (graphs, chained launches)
  return fit_create_common(view_ctx[0], one ? nullptr : view_ctx, n_views, n, p, k, out);
}

extern "C" int resnmtf_fit_destroy(resnmtf_fit* fit) {
  if (!fit) return RESNMTF_OK;
  for (size_t i = 0; i < fit->devs.size(); ++i) {
    cudaSetDevice(fit->devs[i].ctx->device);
    cudaStreamSynchronize(fit->devs[i].ctx->stream);
  }
  cudaSetDevice(fit->ctx->device);
  cudaStreamSynchronize(fit->ctx->stream);
  if (fit->graph_exec) cudaGraphExecDestroy(fit->graph_exec);
  if (fit->graph) cudaGraphDestroy(fit->graph);
  for (auto& a : fit->allocs) {
    if (a.first != fit->ctx) cudaSetDevice(a.first->device);
    rn_dev_free(a.first, a.second);
    if (a.first != fit->ctx) cudaSetDevice(fit->ctx->device);
  }
  for (cudaEvent_t e : fit->vev)
    if (e) cudaEventDestroy(e);
  for (ViewHost& vh : fit->views) rn_data_release(vh.shared);
  cudaSetDevice(fit->ctx->device);
  resnmtf_ctx* ctx = fit->ctx;
  std::vector<resnmtf_ctx*> others;
  for (size_t i = 1; i < fit->devs.size(); ++i) others.push_back(fit->devs[i].ctx);
  delete fit;
  for (resnmtf_ctx* o : others) rn_ctx_release(o);
  rn_ctx_release(ctx);
  return RESNMTF_OK;
}

// Copies a column-major matrix (host or device) into the swizzled panel layout at g.X (already zeroed) and
// leaves ||X||_F^2 (all-reduced when the context is row-sharded) in g.scal[0].  `g` needs n, p, pp, ldx,
// row_tiles, X, scal.  Host sources go through two <= 128 MB device staging buffers so that the H2D copy of
// chunk c+1 overlaps the re-tiling of chunk c.
// Pinned pieces, events and streams of the threaded upload of pageable host memory (created on first use).
static cudaError_t rn_upload_setup(resnmtf_ctx* ctx) {
  if (ctx->up_ready) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  for (int t = 0; t < RN_UP_T && e == cudaSuccess; ++t) {
    if (!ctx->up_st[t]) e = cudaStreamCreateWithFlags(&ctx->up_st[t], cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      if (!ctx->up_pin[t][b]) e = cudaHostAlloc((void**)&ctx->up_pin[t][b], RN_UP_BYTES, cudaHostAllocDefault);
      if (e == cudaSuccess && !ctx->up_ev[t][b]) e = cudaEventCreateWithFlags(&ctx->up_ev[t][b], cudaEventDisableTiming);
    }
  }
  ctx->up_ready = e == cudaSuccess;
  return e;
}

// nc columns of n doubles from PAGEABLE host memory (leading dimension ld) into the compact device buffer dst: column
// groups of <= RN_UP_BYTES are dealt round-robin to RN_UP_T host threads; each copies its group into one of its two
// pinned pieces and sends it on its own stream while it fills the other.  Returns when everything has landed.
static cudaError_t rn_upload_pageable(resnmtf_ctx* ctx, double* dst, const double* x, int64_t ld, int64_t n, int64_t nc) {
  const int64_t gcols = std::max<int64_t>(1, (int64_t)(RN_UP_BYTES / ((size_t)n * sizeof(double))));
  const int64_t groups = (nc + gcols - 1) / gcols;
  cudaError_t errs[RN_UP_T];
  std::vector<std::thread> threads;
  const int nt = (int)std::min<int64_t>(RN_UP_T, groups);
  for (int t = 0; t < nt; ++t) {
    errs[t] = cudaSuccess;
    threads.emplace_back([=, &errs]() {
      cudaError_t e = cudaSetDevice(ctx->device);
      int used = 0;
      for (int64_t gi = t; gi < groups && e == cudaSuccess; gi += nt, ++used) {
        const int b = used & 1;
        if (used >= 2) e = cudaEventSynchronize(ctx->up_ev[t][b]);  // the piece's previous copy has left it
        if (e != cudaSuccess) break;
        const int64_t c0 = gi * gcols, cn = std::min(gcols, nc - c0);
        double* pin = ctx->up_pin[t][b];
        if (ld == n) {
          std::memcpy(pin, x + c0 * ld, (size_t)cn * n * sizeof(double));
        } else {
          for (int64_t c = 0; c < cn; ++c) std::memcpy(pin + c * n, x + (c0 + c) * ld, (size_t)n * sizeof(double));
        }
        e = cudaMemcpyAsync(dst + c0 * n, pin, (size_t)cn * n * sizeof(double), cudaMemcpyHostToDevice, ctx->up_st[t]);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->up_ev[t][b], ctx->up_st[t]);
      }
      cudaError_t e2 = cudaStreamSynchronize(ctx->up_st[t]);
      errs[t] = e != cudaSuccess ? e : e2;
    });
  }
  for (auto& th : threads) th.join();
  for (int t = 0; t < nt; ++t)
    if (errs[t] != cudaSuccess) return errs[t];
  return cudaSuccess;
}

static int upload_panels(resnmtf_ctx* ctx, const RnView& g, const double* x, int64_t ld, bool on_device,
                         double* xpart, int32_t* xticket, const char* who) {
  cudaStream_t st = ctx->stream;
  const int64_t n = g.n, p = g.p;
  if (on_device) {
    const int blocks = (int)std::min<int64_t>(((int64_t)g.row_tiles * p * 32 + 255) / 256, 1 << 20);
    rn_to_panels<<<blocks, 256, 0, st>>>(g, x, ld, 0, p);
    RN_CUDA(cudaGetLastError());
  } else {
    const int64_t max_cols = std::max<int64_t>(1, ((int64_t)128 << 20) / (n * (int64_t)sizeof(double)));
    const int64_t chunk = std::min<int64_t>(p, max_cols);
    const int nbuf = chunk < p ? 2 : 1;
    // staging buffers, copy stream and events live in the context (created on first use, grown when needed)
    cudaError_t e = cudaSuccess;
    const size_t need = (size_t)chunk * n * sizeof(double);
    if (!ctx->copy_st) e = cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      if (!ctx->copied[b]) e = cudaEventCreateWithFlags(&ctx->copied[b], cudaEventDisableTiming);
      if (e == cudaSuccess && !ctx->tiled[b]) e = cudaEventCreateWithFlags(&ctx->tiled[b], cudaEventDisableTiming);
    }
    if (e == cudaSuccess && (ctx->stage_bytes < need || (nbuf == 2 && !ctx->stage[1]))) {
      e = cudaStreamSynchronize(st);
      for (int b = 0; b < 2; ++b) {
        if (ctx->stage[b]) cudaFree(ctx->stage[b]);
        ctx->stage[b] = nullptr;
      }
      ctx->stage_bytes = 0;
      for (int b = 0; b < nbuf && e == cudaSuccess; ++b) e = cudaMalloc(&ctx->stage[b], need);
      if (e == cudaSuccess) ctx->stage_bytes = need;
    }
    // pageable source (what R hands over) and columns that fit a pinned piece: the threaded path
    bool pageable = false;
    if (e == cudaSuccess && rn_env_int("RESNMTF_UPLOAD_THREADS", 1) != 0 && (size_t)n * sizeof(double) <= RN_UP_BYTES) {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, x) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
      else cudaGetLastError();
      if (pageable && rn_upload_setup(ctx) != cudaSuccess) {
        cudaGetLastError();
        pageable = false;
      }
    }
    double** stage = ctx->stage;
    cudaStream_t copy_st = ctx->copy_st;
    cudaEvent_t* copied = ctx->copied;
    cudaEvent_t* tiled = ctx->tiled;
    int64_t ci = 0;
    for (int64_t c0 = 0; c0 < p && e == cudaSuccess; c0 += chunk, ++ci) {
      const int b = (int)(ci % nbuf);
      const int64_t nc = std::min(chunk, p - c0);
      if (pageable) {
        // the buffer's previous chunk is re-tiled; then RN_UP_T threads bring this chunk over (they return when their
        // last copy has landed, so the re-tiling kernel below needs no event)
        if (ci >= nbuf) e = cudaEventSynchronize(tiled[b]);
        if (e == cudaSuccess) e = rn_upload_pageable(ctx, stage[b], x + c0 * ld, ld, n, nc);
      } else {
        if (ci >= nbuf) e = cudaStreamWaitEvent(copy_st, tiled[b], 0);  // the buffer's previous chunk is re-tiled
        if (e == cudaSuccess)
          e = cudaMemcpy2DAsync(stage[b], (size_t)n * sizeof(double), x + c0 * ld, (size_t)ld * sizeof(double),
                                (size_t)n * sizeof(double), (size_t)nc, cudaMemcpyHostToDevice, copy_st);
        if (e == cudaSuccess) e = cudaEventRecord(copied[b], copy_st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, copied[b], 0);
      }
      if (e == cudaSuccess) {
        const int blocks = (int)std::min<int64_t>(((int64_t)g.row_tiles * nc * 32 + 255) / 256, 1 << 20);
        rn_to_panels<<<blocks, 256, 0, st>>>(g, stage[b], n, c0, nc);
        e = cudaGetLastError();
      }
      if (e == cudaSuccess) e = cudaEventRecord(tiled[b], st);
    }
    cudaError_t e2 = copy_st ? cudaStreamSynchronize(copy_st) : cudaSuccess;
    cudaError_t e3 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = e3;
    if (e != cudaSuccess) return rn_fail(RESNMTF_E_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
  }
  rn_xnorm2<<<1024, 256, 0, st>>>(g, xpart, xticket);
  RN_CUDA(cudaGetLastError());
  int rc = rn_allreduce(ctx, g.scal, 1);  // data_norms of the whole view when row-sharded
  if (rc) return rc;
  RN_CUDA(cudaStreamSynchronize(st));  // the caller's buffer is only borrowed for the call
  return RESNMTF_OK;
}

static int set_data_common(resnmtf_fit* fit, int v, const double* x, int64_t ld, bool on_device,
                           const char* who) {
  RN_CHECK(fit && x, RESNMTF_E_INVALID, std::string(who) + ": NULL argument");
  RN_CHECK(v >= 0 && v < fit->V, RESNMTF_E_INVALID, std::string(who) + ": view index out of range");
  ViewHost& vh = fit->views[v];
  RN_CHECK(ld >= vh.d.n, RESNMTF_E_INVALID, std::string(who) + ": ld < n");
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  RnViewScope scope(fit, v);  // a placed view lives on its own GPU
  if (vh.shared) {  // the view was attached to a shared handle: give it its own buffer again
    rn_data_release(vh.shared);
    vh.shared = nullptr;
    vh.d.X = nullptr;
    if (vh.d.X8) fit->plan_dirty = true;
    vh.d.X8 = nullptr;
  }
  if (!vh.d.X) {
    int rc = rn_alloc(fit, &vh.d.X, (size_t)vh.d.ldx * vh.d.pp);  // zeroed: padding rows / columns stay zero
    if (rc) return rc;
    fit->meta_dirty = true;  // X is a by-value kernel parameter: the iteration graph must be rebuilt
  }
  int rc = upload_panels(fit->cur, vh.d, x, ld, on_device, vh.xpart, vh.xticket, who);
  if (rc) return rc;
  vh.has_data = true;
  if (vh.d.X8) {  // the fused path's copy of X is stale: the next plan rebuilds it
    vh.d.X8 = nullptr;
    fit->plan_dirty = true;
  }
  return RESNMTF_OK;
}

// ---- shared device data -------------------------------------------------------------------------
static int data_create_common(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x, int64_t ld, bool on_device,
                              resnmtf_data** out) {
  RN_CHECK(ctx && x && out, RESNMTF_E_INVALID, "resnmtf_data_create: NULL argument");
  RN_CHECK(n >= 1 && p >= 1 && ld >= n, RESNMTF_E_INVALID, "resnmtf_data_create: bad shape");
  RN_CUDA(cudaSetDevice(ctx->device));
  resnmtf_data* d = new (std::nothrow) resnmtf_data();
  RN_CHECK(d != nullptr, RESNMTF_E_NOMEM, "resnmtf_data_create: out of host memory");
  d->ctx = ctx;
  ctx->refs.fetch_add(1);
  d->n = n;
  d->p = p;
  d->ldx = rn_round_up(n, RN_ROW_TILE);
  d->pp = rn_round_up(p, 32);
  double* scratch = nullptr;  // [0..7] scal, then 1024 partials, then the ticket
  const size_t xbytes = (size_t)d->ldx * d->pp * sizeof(double);
  cudaError_t e = rn_dev_alloc(ctx, (void**)&d->X, xbytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(d->X, 0, xbytes, ctx->stream);
  if (e == cudaSuccess) e = rn_dev_alloc(ctx, (void**)&scratch, (8 + 1024 + 2) * sizeof(double));
  if (e == cudaSuccess) e = cudaMemsetAsync(scratch, 0, (8 + 1024 + 2) * sizeof(double), ctx->stream);
  int rc = RESNMTF_OK;
  if (e != cudaSuccess) {
    rc = rn_fail(e == cudaErrorMemoryAllocation ? RESNMTF_E_NOMEM : RESNMTF_E_CUDA,
                 std::string("resnmtf_data_create: ") + cudaGetErrorString(e));
  } else {
    RnView g;
    std::memset(&g, 0, sizeof(g));
    g.n = n;
    g.p = p;
    g.ldx = d->ldx;
    g.pp = d->pp;
    g.row_tiles = (int)(d->ldx / RN_ROW_TILE);
    g.X = d->X;
    g.scal = scratch;
    rc = upload_panels(ctx, g, x, ld, on_device, scratch + 8, reinterpret_cast<int32_t*>(scratch + 8 + 1024),
                       "resnmtf_data_create");
    if (rc == RESNMTF_OK && cudaMemcpy(&d->xnorm2, scratch, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = rn_fail(RESNMTF_E_CUDA, "resnmtf_data_create: reading back ||X||^2 failed");
  }
  rn_dev_free(ctx, scratch);
  if (rc) {
    rn_data_release(d);
    return rc;
  }
  *out = d;
  return RESNMTF_OK;
}

// An empty (zeroed) view in the panel layout for the kernels of rn_native.cu to fill, and its completion:
// data_norms = ||X||_F^2 (R/main.r:48), all-reduced when the context is row-sharded.
int rn_data_alloc(resnmtf_ctx* ctx, int64_t n, int64_t p, resnmtf_data** out) {
  RN_CHECK(ctx && out, RESNMTF_E_INVALID, "rn_data_alloc: NULL argument");
  RN_CHECK(n >= 1 && p >= 1, RESNMTF_E_INVALID, "rn_data_alloc: bad shape");
  RN_CUDA(cudaSetDevice(ctx->device));
  resnmtf_data* d = new (std::nothrow) resnmtf_data();
  RN_CHECK(d != nullptr, RESNMTF_E_NOMEM, "rn_data_alloc: out of host memory");
  d->ctx = ctx;
  ctx->refs.fetch_add(1);
  d->n = n;
  d->p = p;
  d->ldx = rn_round_up(n, RN_ROW_TILE);
  d->pp = rn_round_up(p, 32);
  const size_t xbytes = (size_t)d->ldx * d->pp * sizeof(double);
  cudaError_t e = rn_dev_alloc(ctx, (void**)&d->X, xbytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(d->X, 0, xbytes, ctx->stream);
  if (e != cudaSuccess) {
    rn_data_release(d);
    return rn_fail(e == cudaErrorMemoryAllocation ? RESNMTF_E_NOMEM : RESNMTF_E_CUDA,
                   std::string("rn_data_alloc: ") + cudaGetErrorString(e));
  }
  *out = d;
  return RESNMTF_OK;
}

int rn_data_seal(resnmtf_data* d) {
  resnmtf_ctx* ctx = d->ctx;
  RN_CUDA(cudaSetDevice(ctx->device));
  double* scratch = nullptr;  // [0..7] scal, then 1024 partials, then the ticket
  RN_CUDA(rn_dev_alloc(ctx, (void**)&scratch, (8 + 1024 + 2) * sizeof(double)));
  RN_CUDA(cudaMemsetAsync(scratch, 0, (8 + 1024 + 2) * sizeof(double), ctx->stream));
  RnView g;
  std::memset(&g, 0, sizeof(g));
  g.n = d->n;
  g.p = d->p;
  g.ldx = d->ldx;
  g.pp = d->pp;
  g.row_tiles = (int)(d->ldx / RN_ROW_TILE);
  g.X = d->X;
  g.scal = scratch;
  rn_xnorm2<<<1024, 256, 0, ctx->stream>>>(g, scratch + 8, reinterpret_cast<int32_t*>(scratch + 8 + 1024));
  cudaError_t e = cudaGetLastError();
  int rc = e == cudaSuccess ? rn_allreduce(ctx, g.scal, 1) : RESNMTF_OK;
  if (e == cudaSuccess && rc == RESNMTF_OK)
    e = cudaMemcpyAsync(&d->xnorm2, scratch, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  rn_dev_free(ctx, scratch);
  if (rc) return rc;
  if (e != cudaSuccess) return rn_fail(RESNMTF_E_CUDA, std::string("rn_data_seal: ") + cudaGetErrorString(e));
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_create(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x, int64_t ld,
                                   resnmtf_data** out) {
  return data_create_common(ctx, n, p, x, ld, false, out);
}
extern "C" int resnmtf_data_create_device(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x_dev, int64_t ld,
                                          resnmtf_data** out) {
  return data_create_common(ctx, n, p, x_dev, ld, true, out);
}

extern "C" int resnmtf_data_destroy(resnmtf_data* data) {
  if (data) {
    cudaSetDevice(data->ctx->device);
    rn_data_release(data);
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_attach_data(resnmtf_fit* fit, int v, resnmtf_data* data) {
  RN_CHECK(fit && data, RESNMTF_E_INVALID, "resnmtf_fit_attach_data: NULL argument");
  RN_CHECK(v >= 0 && v < fit->V, RESNMTF_E_INVALID, "resnmtf_fit_attach_data: view index out of range");
  RN_CHECK(data->ctx == rn_vctx(fit, v), RESNMTF_E_INVALID, "resnmtf_fit_attach_data: data lives on another context");
  ViewHost& vh = fit->views[v];
  RN_CHECK(data->n == vh.d.n && data->p == vh.d.p, RESNMTF_E_INVALID, "resnmtf_fit_attach_data: shape mismatch");
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  RnViewScope scope(fit, v);
  RN_CUDA(cudaStreamSynchronize(fit->cur->stream));
  if (vh.shared) rn_data_release(vh.shared);
  else if (vh.d.X) {
    int rc = rn_free(fit, vh.d.X);
    if (rc) return rc;
  }
  vh.shared = data;
  data->refs.fetch_add(1);
  vh.d.X = data->X;
  if (vh.d.X8) fit->plan_dirty = true;
  vh.d.X8 = nullptr;
  RN_CUDA(cudaMemcpy(vh.d.scal, &data->xnorm2, sizeof(double), cudaMemcpyHostToDevice));
  vh.has_data = true;
  fit->meta_dirty = true;
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_set_data(resnmtf_fit* fit, int v, const double* x, int64_t ld) {
  return set_data_common(fit, v, x, ld, false, "resnmtf_fit_set_data");
}
extern "C" int resnmtf_fit_set_data_device(resnmtf_fit* fit, int v, const double* x, int64_t ld) {
  return set_data_common(fit, v, x, ld, true, "resnmtf_fit_set_data_device");
}

extern "C" int resnmtf_fit_set_factors(resnmtf_fit* fit, int v, const double* f, const double* s,
                                       const double* g, const double* lambda, const double* mu) {
  RN_CHECK(fit && f && s && g, RESNMTF_E_INVALID, "resnmtf_fit_set_factors: NULL argument");
  RN_CHECK(v >= 0 && v < fit->V, RESNMTF_E_INVALID, "resnmtf_fit_set_factors: view index out of range");
  ViewHost& vh = fit->views[v];
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  RnViewScope scope(fit, v);
  cudaStream_t st = fit->cur->stream;
  const int K = vh.d.k, KP = vh.d.kp;
  const int64_t n = vh.d.n, p = vh.d.p;
  std::vector<double> fp((size_t)vh.d.ldx * KP, 0.0);  // F in the device's swizzled 64-row panel layout
  for (int c = 0; c < K; ++c)
    for (int64_t r = 0; r < n; ++r) fp[(size_t)rn_fidx_host(r, c, KP)] = f[(size_t)c * n + r];
  RN_CUDA(cudaMemcpyAsync(vh.d.F, fp.data(), fp.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  std::vector<double> gt((size_t)vh.d.pp * KP, 0.0);  // G is row-major [pp][kp] on the device
  for (int c = 0; c < K; ++c)
    for (int64_t j = 0; j < p; ++j) gt[(size_t)j * KP + c] = g[(size_t)c * p + j];
  RN_CUDA(cudaMemcpyAsync(vh.d.G, gt.data(), gt.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(vh.d.S, s, (size_t)K * K * sizeof(double), cudaMemcpyHostToDevice, st));
  if (lambda) RN_CUDA(cudaMemcpyAsync(vh.d.lam, lambda, K * sizeof(double), cudaMemcpyHostToDevice, st));
  if (mu) RN_CUDA(cudaMemcpyAsync(vh.d.mu, mu, K * sizeof(double), cudaMemcpyHostToDevice, st));
  rn_factor_sums<<<2 * vh.d.k + vh.d.k * vh.d.k, 1024, 0, st>>>(vh.d);
  {
    int rc2 = rn_allreduce(fit->ctx, vh.d.csF, (size_t)K);  // F rows are sharded, G is replicated
    if (rc2) return rc2;
  }
  rn_default_lm<<<1, 32, 0, st>>>(vh.d, lambda ? 0 : 1, mu ? 0 : 1);
  RN_CUDA(cudaGetLastError());
  RN_CUDA(cudaStreamSynchronize(st));
  vh.has_factors = true;
  // a new set of factors starts a new loop: reset the history and the stop-rule state
  fit->errors.clear();
  std::memset(&fit->h_ctrl, 0, sizeof(RnCtrl));
  fit->counters.iterations = 0;
  if (fit->auto_direct) {
    fit->auto_direct = false;
    fit->meta_dirty = true;
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_set_restrictions(resnmtf_fit* fit, const double* phi, const double* xi,
                                            const double* psi) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_set_restrictions: fit is NULL");
  const size_t VV = (size_t)fit->V * fit->V;
  auto put = [&](std::vector<double>& dst, const double* src, const char* name) -> int {
    for (size_t i = 0; i < VV; ++i) {
      const double x = src ? src[i] : 0.0;
      RN_CHECK(!(x < 0.0), RESNMTF_E_INVALID, std::string(name) + " must be a non-negative matrix");
      // init_rest_mats() (R/update_steps.r:19-22) zeroes the diagonal before it symmetrises; a view coupled to itself
      // would read its own factor rows while they are overwritten in place
      RN_CHECK(!(x != 0.0 && i % ((size_t)fit->V + 1) == 0), RESNMTF_E_INVALID,
               std::string(name) + " must have a zero diagonal (init_rest_mats, R/update_steps.r:12-24)");
      dst[i] = x;
    }
    return RESNMTF_OK;
  };
  int rc;
  if ((rc = put(fit->h_phi, phi, "phi"))) return rc;
  if ((rc = put(fit->h_xi, xi, "xi"))) return rc;
  if ((rc = put(fit->h_psi, psi, "psi"))) return rc;
  // xi couples S matrices element-wise: the coupled views must share k (star_prod, R/utils.r:39-47)
  for (int v = 0; v < fit->V; ++v)
    for (int w = 0; w < fit->V; ++w)
      if (fit->h_xi[w + (size_t)v * fit->V] != 0.0 && w != v)
        RN_CHECK(fit->views[v].d.k == fit->views[w].d.k, RESNMTF_E_INVALID,
                 "xi couples views with different k (non-conformable arrays in the reference)");
  fit->meta_dirty = true;
  fit->plan_dirty = true;  // which views qualify for the fused kernel depends on their phi partners
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_set_shared_map(resnmtf_fit* fit, int kind, int v, int w, const int32_t* idx_v,
                                          const int32_t* idx_w, int64_t len) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_set_shared_map: fit is NULL");
  RN_CHECK(kind == RESNMTF_MAP_ROW || kind == RESNMTF_MAP_COL, RESNMTF_E_INVALID,
           "resnmtf_fit_set_shared_map: bad kind");
  RN_CHECK(v >= 0 && v < fit->V && w >= 0 && w < fit->V && v != w, RESNMTF_E_INVALID,
           "resnmtf_fit_set_shared_map: bad view pair");
  RN_CHECK(len >= 0 && (len == 0 || (idx_v && idx_w)), RESNMTF_E_INVALID,
           "resnmtf_fit_set_shared_map: NULL index array");
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  RnViewScope scope(fit, v);  // the map is read by view v's kernels: it lives on v's GPU
  const bool row = kind == RESNMTF_MAP_ROW;
  ViewHost& vh = fit->views[v];
  const ViewHost& wh = fit->views[w];
  const int64_t dim_v = row ? vh.d.n : vh.d.p, dim_w = row ? wh.d.n : wh.d.p;
  const size_t slot = (size_t)w + (size_t)v * fit->V;
  std::vector<int8_t>& modes = row ? fit->h_rowmode : fit->h_colmode;
  std::vector<const int32_t*>& ptrs = row ? fit->h_rowmap : fit->h_colmap;
  std::vector<int32_t*>& own = row ? vh.rowmaps : vh.colmaps;
  if (len == 0) {
    modes[slot] = RN_MODE_NA;
    ptrs[slot] = nullptr;
    fit->meta_dirty = true;
    fit->plan_dirty = true;
    return RESNMTF_OK;
  }
  std::vector<int32_t> map((size_t)dim_v, -1);
  for (int64_t i = 0; i < len; ++i) {
    RN_CHECK(idx_v[i] >= 0 && idx_v[i] < dim_v && idx_w[i] >= 0 && idx_w[i] < dim_w, RESNMTF_E_INVALID,
             "resnmtf_fit_set_shared_map: index out of range");
    map[(size_t)idx_v[i]] = idx_w[i];
  }
  if (!own[w]) {
    int rc = rn_alloc(fit, &own[w], (size_t)dim_v, false);
    if (rc) return rc;
  }
  RN_CUDA(cudaMemcpyAsync(own[w], map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                          fit->cur->stream));
  RN_CUDA(cudaStreamSynchronize(fit->cur->stream));
  modes[slot] = RN_MODE_MAP;
  ptrs[slot] = own[w];
  fit->meta_dirty = true;
  fit->plan_dirty = true;
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_set_options(resnmtf_fit* fit, int err_mode, int impl) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_set_options: fit is NULL");
  RN_CHECK(err_mode >= 0 && err_mode <= 2 && impl >= 0 && impl <= RESNMTF_IMPL_SMALL, RESNMTF_E_INVALID,
           "resnmtf_fit_set_options: unknown option value");
  if (fit->err_mode != err_mode) {
    fit->meta_dirty = true;
    fit->auto_direct = false;
  }
  if (fit->impl_req != impl) fit->plan_dirty = true;
  fit->err_mode = err_mode;
  fit->impl_req = impl;
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------
// plan: grids, workspaces, device metadata, per-iteration graph
// ------------------------------------------------------------------------------------------------

static int build_plan(resnmtf_fit* fit) {
  const int sms = fit->ctx->sm_count;
  int impl = fit->impl_req;
  if (impl == RESNMTF_IMPL_AUTO) impl = rn_env_int("RESNMTF_IMPL", RESNMTF_IMPL_AUTO);
  // Fits whose every view is tiny run the whole loop as one persistent CTA (rn_small.cuh): picked by AUTO and by
  // RESNMTF_IMPL_SMALL when every view qualifies (k <= 8, at most 1024 padded rows and columns, at most 1 MB of X; one
  // GPU, not row-sharded).  The streaming plan below is still built: resnmtf_fit_profile and the AUTO hand-over use it.
  bool small = (impl == RESNMTF_IMPL_AUTO || impl == RESNMTF_IMPL_SMALL) && !rn_placed(fit) && fit->ctx->comm == nullptr &&
               rn_env_int("RESNMTF_SMALL", 1) != 0;
  int64_t small_elems = 0;
  for (int v = 0; v < fit->V && small; ++v) {
    const RnView& d = fit->views[v].d;
    small = d.k <= 8 && d.ldx <= RN_SM_MAXDIM && d.pp <= RN_SM_MAXDIM && d.ldx * d.pp <= RN_SM_MAXELEMS;
    small_elems += d.ldx * d.pp;
  }
  // AUTO takes it only where it wins: measured 25.6 us per sweep against 34.8 us for two launches per view on the
  // README toy (12288 padded entries), but 59.7 against 35.5 us on 2 x 180 x 180 (73728): one SM against a cluster
  if (impl == RESNMTF_IMPL_AUTO && small_elems > RN_SM_AUTO_ELEMS) small = false;
  fit->small = small;
  fit->small_dim = 64;
  for (int v = 0; v < fit->V; ++v)
    fit->small_dim = (int)std::max<int64_t>(fit->small_dim, std::max(fit->views[v].d.ldx, fit->views[v].d.pp));
  if (small)
    RN_CUDA(cudaFuncSetAttribute(rn_small_sweeps, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)rn_small_smem(RN_SM_MAXDIM)));
  if (impl < RESNMTF_IMPL_DFMA || impl > RESNMTF_IMPL_FUSED) impl = RESNMTF_IMPL_FUSED;
  const int impl_fit = impl;
  bool any_fused = false;
  const int f_target = sms * rn_env_int("RESNMTF_F_CTAS_PER_SM", 2);
  const int g_target_dfma = sms * rn_env_int("RESNMTF_G_CTAS_PER_SM_DFMA", 3);
  int rc;
  for (int v = 0; v < fit->V; ++v) {
    RnViewScope scope(fit, v);  // placed fit: occupancy queries, kernel attributes and workspaces on the view's GPU
    ViewHost& vh = fit->views[v];
    RnView& d = vh.d;
    // the one-pass fused kernel serves the views that qualify (fused_csize); the others of the fit run the
    // two-pass TMA kernels
    // Two generations of the one-pass kernel.  rn_fused_step (kind 1: 1008 columns per CTA, 9 consumer warps on three
    // sub-partitions) is the faster one at the bench shape (143 vs 146 us per update-iteration at 20000 x 4000) and is
    // preferred whenever its fixed CTA width fits the view: p <= 8064 and at most RESNMTF_FUSED2_PAD percent (default
    // 108) of padded columns.  rn_fused2_step (kind 2: <= 42 blocks of 16 columns per CTA dealt evenly, no padding, 12
    // consumer warps on all four sub-partitions; p <= 8 x 672) takes the widths kind 1 would pad heavily and the narrow
    // views (p < 747) that kind 1 does not accept at all.  RESNMTF_FUSED_KIND=1 / 2 forces one of them (A/B runs, parity
    // tests of both).
    const int K = d.k, KP = d.kp;
    const bool sharded = fit->ctx->comm != nullptr;
    int partners = 0;  // the fused kernels prefetch the phi gathers of at most RN_FU_MAXPART partner views
    for (int w = 0; w < fit->V; ++w) {
      const size_t slot = (size_t)w + (size_t)v * fit->V;
      if (w != v && fit->h_phi[slot] != 0.0 && fit->h_rowmode[slot] != RN_MODE_NA) ++partners;
    }
    const bool phi_ok = partners <= RN_FU_MAXPART && !(partners > 0 && rn_env_int("RESNMTF_FUSED_PHI", 1) == 0);
    const int kind_req = rn_env_int("RESNMTF_FUSED_KIND", 0);
    const int min_pct = rn_env_int("RESNMTF_FUSED_MIN_SM_PCT", 70);
    int csz = 0, kind = 0, max_clusters = 0;
    if (impl_fit == RESNMTF_IMPL_FUSED && phi_ok && K <= 8 && !sharded) {
      // the persistent cluster grid must be resident at once; a cluster shape that strands too many SMs loses to the
      // two-pass kernels, which use all of them
      int c1 = 0, mc1 = 0, c2 = 0, mc2 = 0;
      if (kind_req != 2) {
        c1 = fused_csize(d, sharded);
        mc1 = c1 ? fused_max_clusters(K, c1, sms, 1) : 0;
        if (!(c1 && mc1 >= 1 && mc1 * c1 * 100 >= sms * min_pct)) c1 = 0;
      }
      if (kind_req != 1) {
        const int64_t blocks = d.pp / 16;
        c2 = (int)((blocks + RN_F2_MAXB - 1) / RN_F2_MAXB);
        if (c2 < 1 || c2 > RN_FU_MAXC) c2 = 0;
        mc2 = c2 ? fused_max_clusters(K, c2, sms, 2) : 0;
        if (!(c2 && mc2 >= 1 && mc2 * c2 * 100 >= sms * min_pct)) c2 = 0;
      }
      const bool tight1 = c1 && (int64_t)c1 * RN_FU_CCOLS * 100 <= (int64_t)d.p * rn_env_int("RESNMTF_FUSED2_PAD", 108);
      if (c1 && (tight1 || !c2)) {
        csz = c1;
        kind = 1;
        max_clusters = mc1;
      } else if (c2) {
        csz = c2;
        kind = 2;
        max_clusters = mc2;
      }
    }
    const int impl = (impl_fit == RESNMTF_IMPL_FUSED && !csz) ? RESNMTF_IMPL_TMA : impl_fit;
    vh.impl = impl;
    const bool mma = use_mma(vh, impl);
    size_t n_ppart, n_tpart, n_ffpart, n_ggpart;
    d.fu_csize = d.fu_clusters = d.fu_kind = 0;
    if (csz) {
      any_fused = true;
      const int64_t groups = (d.n + 7) / 8;
      int nc = (int)std::min<int64_t>(max_clusters, groups);
      nc = std::max(1, std::min(nc, rn_env_int("RESNMTF_FU_CLUSTERS", nc)));
      d.fu_csize = csz;
      d.fu_clusters = nc;
      d.fu_kind = kind;
      const int64_t pp8_new = kind == 2 ? d.pp : (int64_t)csz * RN_FU_CCOLS;
      if (d.X8 && d.pp8 != pp8_new) d.X8 = nullptr;  // the copy at hand has the other kernel's width
      d.pp8 = pp8_new;
      d.col_groups = (int)((d.pp + RN_FU_TG - 1) / RN_FU_TG);
      d.cs = d.rs = 1;
      d.nff = 0;
      d.f_ctas = d.g_ctas = 0;
      d.gepi_ctas = (int)((d.p + RN_GEPI_THREADS(K) - 1) / RN_GEPI_THREADS(K));
      n_ppart = 0;
      n_tpart = (size_t)nc * d.pp8 * KP;
      n_ffpart = (size_t)nc * (K * K + K);
      n_ggpart = (size_t)std::max(d.col_groups, d.gepi_ctas) * (2 * K * K + K);
      if (rn_env_int("RESNMTF_FU_TIMELINE", 0) && !d.fu_timeline) {
        if ((rc = rn_alloc(fit, &d.fu_timeline, (size_t)sms * 12))) return rc;
        if ((rc = rn_alloc(fit, &d.fu_trace, (size_t)((d.n + 7) / 8 + 8) * 32))) return rc;
        if ((rc = rn_alloc(fit, &d.fu_waits, (size_t)sms * RN_FU_NCW * 2))) return rc;
      }
      if (!d.X8) {  // second copy of X in the 8-row-group layout (shared by every fit attached to the same data)
        const size_t x8_count = (size_t)d.ldx * d.pp8;
        bool convert = true;
        if (vh.shared) {
          if (vh.shared->X8 && vh.shared->pp8 == d.pp8) {
            convert = false;
            d.X8 = vh.shared->X8;
          } else if (!vh.shared->X8) {
            RN_CUDA(rn_dev_alloc(fit->cur, (void**)&vh.shared->X8, x8_count * sizeof(double)));
            vh.shared->pp8 = d.pp8;
            d.X8 = vh.shared->X8;
          } else {  // the handle's copy has another width and other fits may be running on it: this fit keeps its own
            if (!vh.x8_own && (rc = rn_alloc(fit, &vh.x8_own, x8_count, false))) return rc;
            d.X8 = vh.x8_own;
          }
        } else {
          if (!vh.x8_own && (rc = rn_alloc(fit, &vh.x8_own, x8_count, false))) return rc;
          d.X8 = vh.x8_own;
        }
        if (convert) {
          const int blocks = (int)std::min<int64_t>(((int64_t)x8_count / 2 + 255) / 256, 1 << 20);
          rn_panels_to_x8<<<blocks, 256, 0, fit->cur->stream>>>(d);
          RN_CUDA(cudaGetLastError());
        }
      }
    } else if (mma) {
      // persistent stream-K grids: exactly the CTAs that are resident at once (never more than units)
      int occ_f = 1, occ_g = 1;
      if (impl == RESNMTF_IMPL_TMA) {  // one CTA per SM: the shared-memory ring takes ~200 KB
        RN_CUDA(cudaFuncSetAttribute(f_step_tma_fn(K), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)rn_f_tma_smem(K)));
        RN_CUDA(cudaFuncSetAttribute(g_step_tma_fn(K), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)rn_g_tma_smem(K)));
        RN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, f_step_tma_fn(K), RN_TMA_THREADS,
                                                              rn_f_tma_smem(K)));
        RN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, g_step_tma_fn(K), RN_TMA_THREADS,
                                                              rn_g_tma_smem(K)));
        RN_CHECK(occ_f >= 1 && occ_g >= 1, RESNMTF_E_CUDA, "TMA kernels do not fit on this device");
      } else {
        RN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, f_step_sk_fn(K), 256, 0));
        RN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, g_step_sk_fn(K), 128, 0));
      }
      occ_f = std::max(1, rn_env_int("RESNMTF_F_OCC", occ_f));
      occ_g = std::max(1, rn_env_int("RESNMTF_G_OCC", occ_g));
      d.col_groups = (int)((d.pp + RN_COL_GROUP - 1) / RN_COL_GROUP);
      const int64_t f_units = (int64_t)d.row_tiles * (d.pp / 32);
      const int64_t g_units = (int64_t)d.row_tiles * d.col_groups;
      d.f_ctas = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * occ_f, f_units));
      d.g_ctas = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * occ_g, g_units));
      d.f_ctas = std::max(1, std::min<int>(rn_env_int("RESNMTF_F_CTAS", d.f_ctas), (int)f_units));
      d.g_ctas = std::max(1, std::min<int>(rn_env_int("RESNMTF_G_CTAS", d.g_ctas), (int)g_units));
      d.cs = d.rs = 1;
      d.nff = 0;
      d.gepi_ctas = (int)((d.p + RN_GEPI_THREADS(K) - 1) / RN_GEPI_THREADS(K));  // row-sharded path only
      n_ppart = (size_t)d.f_ctas * 2 * RN_ROW_TILE * KP;
      n_tpart = (size_t)d.g_ctas * 2 * RN_COL_GROUP * KP;
      n_ffpart = (size_t)d.g_ctas * (K * K + K);
      n_ggpart = (size_t)std::max(d.col_groups, d.gepi_ctas) * (2 * K * K + K);
    } else {
      // F step: one CTA per 64-row tile; split the columns only when there are too few tiles to fill
      // the machine (each split keeps >= 64 data columns)
      int cs = 1;
      if (d.row_tiles < f_target) cs = (f_target + d.row_tiles - 1) / d.row_tiles;
      cs = std::max(1, std::min<int>(cs, (int)std::max<int64_t>(1, d.pp / 64)));
      cs = rn_env_int("RESNMTF_F_CS", cs);
      d.cs = std::max(1, std::min<int>(cs, (int)(d.pp / 8)));
      // G stream: one CTA per 32-column group; split the row steps to fill the machine
      d.col_groups = (int)((d.pp + RN_COL_GROUP_DFMA - 1) / RN_COL_GROUP_DFMA);
      int rs = std::max(1, (g_target_dfma + d.col_groups / 2) / d.col_groups);
      rs = rn_env_int("RESNMTF_G_RS", rs);
      d.rs = std::max(1, std::min(rs, d.row_tiles));
      d.nff = std::max(1, std::min(sms, (int)((d.ldx + 255) / 256)));
      d.gepi_ctas = (int)((d.p + RN_GEPI_THREADS(K) - 1) / RN_GEPI_THREADS(K));
      d.f_ctas = d.g_ctas = 0;
      n_ppart = d.cs > 1 ? (size_t)d.cs * d.row_tiles * RN_ROW_TILE * KP : 0;
      n_tpart = d.rs > 1 ? (size_t)d.rs * d.pp * KP : 0;
      n_ffpart = (size_t)d.nff * (K * K + K);
      n_ggpart = (size_t)d.gepi_ctas * (2 * K * K + K);
    }
    d.resid_cs = std::max(1, std::min<int>((2 * sms + d.row_tiles - 1) / d.row_tiles, (int)(d.pp / 64)));
    d.resid_cs = std::max(1, d.resid_cs);
    // workspaces (old ones are released first when the plan is rebuilt)
    if ((rc = rn_free(fit, d.Ppart))) return rc;
    if ((rc = rn_free(fit, d.Tpart))) return rc;
    if ((rc = rn_free(fit, d.FFpart))) return rc;
    if ((rc = rn_free(fit, d.GGpart))) return rc;
    if ((rc = rn_free(fit, d.Rpart))) return rc;
    if ((rc = rn_free(fit, d.tile_ticket))) return rc;
    if ((rc = rn_free(fit, d.group_ticket))) return rc;
    d.Ppart = d.Tpart = d.FFpart = d.GGpart = d.Rpart = nullptr;
    d.tile_ticket = d.group_ticket = nullptr;
    if (n_ppart && (rc = rn_alloc(fit, &d.Ppart, n_ppart))) return rc;
    if (n_tpart && (rc = rn_alloc(fit, &d.Tpart, n_tpart))) return rc;
    if ((rc = rn_alloc(fit, &d.FFpart, n_ffpart))) return rc;
    if ((rc = rn_alloc(fit, &d.GGpart, n_ggpart))) return rc;
    if ((rc = rn_alloc(fit, &d.Rpart, (size_t)d.row_tiles * d.resid_cs))) return rc;
    if ((rc = rn_alloc(fit, &d.tile_ticket, (size_t)d.row_tiles))) return rc;
    if ((rc = rn_alloc(fit, &d.group_ticket, (size_t)d.col_groups))) return rc;
    RN_CUDA(cudaMemsetAsync(d.misc_ticket, 0, 4 * sizeof(int32_t), fit->cur->stream));
  }
  {  // persisting-L2 budget: 90 % of the carve-out, split over the views in proportion to their size
    double total = 0.0;
    for (int v = 0; v < fit->V; ++v) total += (double)fit->views[v].d.ldx * fit->views[v].d.pp;
    const double budget = 0.9 * (double)fit->ctx->l2_persist_bytes * rn_env_int("RESNMTF_L2_PERSIST_PCT", 100) / 100.0;
    for (int v = 0; v < fit->V; ++v) {
      ViewHost& vh = fit->views[v];
      const double xbytes = (double)vh.d.ldx * vh.d.pp * sizeof(double);
      double w = std::min(xbytes, budget * ((double)vh.d.ldx * vh.d.pp / total));
      vh.l2_window = (size_t)(w / 32768.0) * 32768;  // whole 32 KB units
    }
  }
  fit->impl = fit->small ? RESNMTF_IMPL_SMALL : (impl_fit == RESNMTF_IMPL_FUSED && !any_fused) ? RESNMTF_IMPL_TMA : impl_fit;
  fit->plan_dirty = false;
  fit->meta_dirty = true;
  if (fit->graph_exec) {
    cudaGraphExecDestroy(fit->graph_exec);
    fit->graph_exec = nullptr;
  }
  if (fit->graph) {
    cudaGraphDestroy(fit->graph);
    fit->graph = nullptr;
  }
  return RESNMTF_OK;
}

static int sync_meta(resnmtf_fit* fit) {
  cudaStream_t st = fit->ctx->stream;
  const int V = fit->V;
  const size_t VV = (size_t)V * V;
  std::vector<RnView> hv(V);
  for (int v = 0; v < V; ++v) hv[v] = fit->views[v].d;
  RN_CUDA(cudaMemcpyAsync(fit->d_views, hv.data(), V * sizeof(RnView), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_phi, fit->h_phi.data(), VV * sizeof(double), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_xi, fit->h_xi.data(), VV * sizeof(double), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_psi, fit->h_psi.data(), VV * sizeof(double), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_rowmap, fit->h_rowmap.data(), VV * sizeof(int32_t*), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_colmap, fit->h_colmap.data(), VV * sizeof(int32_t*), cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_rowmode, fit->h_rowmode.data(), VV, cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaMemcpyAsync(fit->d_colmode, fit->h_colmode.data(), VV, cudaMemcpyHostToDevice, st));
  RN_CUDA(cudaStreamSynchronize(st));
  RnFit& d = fit->d;
  d.n_views = V;
  d.err_mode = (fit->err_mode == RESNMTF_ERR_AUTO && fit->auto_direct) ? RESNMTF_ERR_DIRECT : fit->err_mode;
  d.views = fit->d_views;
  d.ctrl = fit->d_ctrl;
  d.phi = fit->d_phi;
  d.xi = fit->d_xi;
  d.psi = fit->d_psi;
  d.rowmap = fit->d_rowmap;
  d.colmap = fit->d_colmap;
  d.rowmode = fit->d_rowmode;
  d.colmode = fit->d_colmode;
  d.hist = fit->d_hist;
  d.hist_cap = fit->hist_cap;
  double ps = 0.0, xs = 0.0;  // whole-matrix sums: the branch tests of update_g / update_s
  for (size_t i = 0; i < VV; ++i) {
    ps += fit->h_psi[i];
    xs += fit->h_xi[i];
  }
  d.psi_total = ps;
  d.xi_total = xs;
  // placed fit: every other GPU gets its own copy of the tables (the pointers inside reach peer memory); the control
  // block and the error history stay on the home GPU
  for (size_t i = 1; i < fit->devs.size(); ++i) {
    RnDevMeta& m = fit->devs[i];
    RN_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t ms = m.ctx->stream;
    RN_CUDA(cudaMemcpyAsync(m.d_views, hv.data(), V * sizeof(RnView), cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_phi, fit->h_phi.data(), VV * sizeof(double), cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_xi, fit->h_xi.data(), VV * sizeof(double), cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_psi, fit->h_psi.data(), VV * sizeof(double), cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_rowmap, fit->h_rowmap.data(), VV * sizeof(int32_t*), cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_colmap, fit->h_colmap.data(), VV * sizeof(int32_t*), cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_rowmode, fit->h_rowmode.data(), VV, cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaMemcpyAsync(m.d_colmode, fit->h_colmode.data(), VV, cudaMemcpyHostToDevice, ms));
    RN_CUDA(cudaStreamSynchronize(ms));
    m.d = d;
    m.d.views = m.d_views;
    m.d.phi = m.d_phi;
    m.d.xi = m.d_xi;
    m.d.psi = m.d_psi;
    m.d.rowmap = m.d_rowmap;
    m.d.colmap = m.d_colmap;
    m.d.rowmode = m.d_rowmode;
    m.d.colmode = m.d_colmode;
  }
  if (!fit->devs.empty()) {
    fit->devs[0].d = d;
    RN_CUDA(cudaSetDevice(fit->ctx->device));
  }
  fit->meta_dirty = false;
  // kernel parameters are baked into the graph: rebuild it
  if (fit->graph_exec) {
    cudaGraphExecDestroy(fit->graph_exec);
    fit->graph_exec = nullptr;
  }
  if (fit->graph) {
    cudaGraphDestroy(fit->graph);
    fit->graph = nullptr;
  }
  return RESNMTF_OK;
}

// One update-iteration of a PLACED fit (views on several GPUs of this process; SURVEY 8e "several coupled views").
// update_matrices() (R/update_steps.r:282-314) is a Gauss-Seidel sweep: view v sees the NEW factors of the views before
// it and the OLD ones of the views after it.  Across GPUs that order is a chain of events: the kernels of view v are
// enqueued on the stream of v's GPU behind (a) the event of every view w < v it is coupled to by phi, psi or xi -- w's
// kernels of THIS sweep have finished, so the rows v's kernels read over NVLink are the new ones -- and (b) the event of
// the last view of the PREVIOUS sweep, whose kernel did the sweep's bookkeeping (error history, stop flag): no view may
// start a sweep before the stop rule of the one before is decided, and by then every partner w > v has finished reading
// v's old factors.  Views that are not coupled to each other run their sweeps concurrently on their GPUs.  The partner
// rows themselves are read by the update kernels straight from peer memory (the gather maps live on the reader's GPU).
static int64_t enqueue_iteration_placed(resnmtf_fit* fit, bool direct) {
  int64_t launches = 0;
  const int V = fit->V;
  const int last = V - 1;
  auto coupled = [&](int v, int w) {
    const size_t a = (size_t)w + (size_t)v * V, b = (size_t)v + (size_t)w * V;
    return fit->h_phi[a] != 0.0 || fit->h_phi[b] != 0.0 || fit->h_psi[a] != 0.0 || fit->h_psi[b] != 0.0 ||
           fit->h_xi[a] != 0.0 || fit->h_xi[b] != 0.0;
  };
  for (int v = 0; v < V; ++v) {
    RnViewScope scope(fit, v);
    const ViewHost& vh = fit->views[v];
    const RnFit& ft = fit->devs[fit->vdev[v]].d;
    cudaStream_t st = fit->cur->stream;
    cudaStreamWaitEvent(st, fit->vev[last], 0);  // previous sweep's bookkeeping (a no-op before the first record)
    for (int w = 0; w < v; ++w)
      if (coupled(v, w) || (!direct && v == last)) cudaStreamWaitEvent(st, fit->vev[w], 0);
    const int fuse = (!direct && v == last) ? 1 : 0;  // the last view's kernel also reads every view's error
    if (vh.d.fu_csize) {
      launch_fused(vh, ft, v, fuse, st, false);
      launches += 1;
    } else {
      launch_f_step(vh, ft, v, vh.impl, st);
      launches += 1 + launch_g_step(vh, ft, v, vh.impl, fuse, st, nullptr, &fit->comm_rc);
    }
    if (direct) launches += launch_residual(vh, ft, 0, st, nullptr, &fit->comm_rc);
    if (!(direct && v == last)) cudaEventRecord(fit->vev[v], st);
  }
  if (direct) {  // bookkeeping as its own launch on the last view's GPU, behind every view's residual
    RnViewScope scope(fit, last);
    cudaStream_t st = fit->cur->stream;
    for (int w = 0; w < last; ++w) cudaStreamWaitEvent(st, fit->vev[w], 0);
    rn_finish<<<1, 1, 0, st>>>(fit->devs[fit->vdev[last]].d, 0);
    cudaEventRecord(fit->vev[last], st);
    launches += 1;
  }
  return launches;
}

// Enqueues one update-iteration (all views, Gauss-Seidel order) on `st`.  When ev is non-null an event
// is recorded before every kernel class for the profiling entry point.
// Error mode ALGEBRAIC / AUTO: the last view's G step also does the iteration bookkeeping (fused finish).
// Error mode DIRECT: every view is followed by the residual pass, then rn_finish.
static int64_t enqueue_iteration(resnmtf_fit* fit, cudaStream_t st, std::vector<cudaEvent_t>* ev,
                                 std::vector<int>* ev_class) {
  int64_t launches = 0;
  auto mark = [&](int cls) {
    if (!ev) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev->push_back(e);
    ev_class->push_back(cls);
  };
  const bool direct = fit->d.err_mode == RESNMTF_ERR_DIRECT;
  if (rn_placed(fit)) return enqueue_iteration_placed(fit, direct);
  for (int v = 0; v < fit->V; ++v) {
    const ViewHost& vh = fit->views[v];
    const int fuse = (!direct && v == fit->V - 1) ? 1 : 0;
    if (vh.d.fu_csize) {  // one pass over X: F step and G step in one launch
      mark(2);
      launch_fused(vh, fit->d, v, fuse, st, fit->pdl_ok);
      launches += 1;
    } else {
      mark(0);
      launch_f_step(vh, fit->d, v, vh.impl, st);
      launches += 1;
      mark(1);
      launches += launch_g_step(vh, fit->d, v, vh.impl, fuse, st, fit->ctx, &fit->comm_rc);
    }
    if (direct) {
      mark(3);
      launches += launch_residual(vh, fit->d, 0, st, fit->ctx, &fit->comm_rc);
    }
  }
  if (direct) {
    mark(4);
    rn_finish<<<1, 1, 0, st>>>(fit->d, 0);
    launches += 1;
  }
  mark(-1);
  return launches;
}

static int prepare(resnmtf_fit* fit) {
  for (int v = 0; v < fit->V; ++v) {
    RN_CHECK(fit->views[v].has_data, RESNMTF_E_STATE, "resnmtf: set_data was not called for every view");
    RN_CHECK(fit->views[v].has_factors, RESNMTF_E_STATE, "resnmtf: set_factors was not called for every view");
  }
  // phi / psi overwrite rows of one view's factor with another's (R/utils.r:72): the coupled views
  // must have the same k unless the pair shares nothing (NA pairs are skipped)
  for (int v = 0; v < fit->V; ++v)
    for (int w = 0; w < fit->V; ++w) {
      if (w == v || fit->views[v].d.k == fit->views[w].d.k) continue;
      const size_t slot = (size_t)w + (size_t)v * fit->V;
      RN_CHECK(!(fit->h_phi[slot] != 0.0 && fit->h_rowmode[slot] != RN_MODE_NA), RESNMTF_E_INVALID,
               "phi couples views with different k (non-conformable in the reference)");
      RN_CHECK(!(fit->h_psi[slot] != 0.0 && fit->h_colmode[slot] != RN_MODE_NA), RESNMTF_E_INVALID,
               "psi couples views with different k (non-conformable in the reference)");
    }
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  int rc;
  if (fit->plan_dirty && (rc = build_plan(fit))) return rc;
  if (fit->meta_dirty && (rc = sync_meta(fit))) return rc;
  // A fit whose every view runs the one-pass kernel launches straight into the stream: consecutive launches then
  // chain through programmatic dependent launch (launch_fused), which a chain of single-iteration graph launches
  // cannot do; one kernel per view and iteration leaves nothing for a graph to save.
  // Programmatic dependent launch places the next kernel's clusters on SMs one by one as the CTAs of the running
  // kernel exit.  Clusters of 1, 2, 4, 6 or 8 CTAs are whole TPCs (SM pairs) and always find their full count again;
  // clusters of an ODD size do not: measured on the C3 structure (50000 x 5000, 5-CTA clusters) only about half of
  // the 26 clusters were resident when the grid started, the rest ran as a second wave after the first had finished
  // (2200 us per update-iteration against 1644 us without the attribute).  Such fits launch without it.
  fit->pdl_ok = true;
  for (int v = 0; v < fit->V; ++v) {
    const int cs = fit->views[v].d.fu_csize;
    if (cs > 1 && (cs & 1)) fit->pdl_ok = false;
  }
  bool chained = rn_env_int("RESNMTF_NO_PDL", 0) == 0 && fit->pdl_ok && fit->d.err_mode != RESNMTF_ERR_DIRECT;
  for (int v = 0; v < fit->V; ++v) chained = chained && fit->views[v].d.fu_csize > 0;
  // a placed fit enqueues on one stream per GPU with events in between: plain launches, no graph
  if (!fit->graph_exec && !chained && !rn_placed(fit) && !fit->small && rn_env_int("RESNMTF_NO_GRAPH", 0) == 0) {
    cudaStream_t st = fit->ctx->stream;
    RN_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    fit->launches_per_iter = enqueue_iteration(fit, st, nullptr, nullptr);
    cudaError_t le = cudaGetLastError();
    cudaError_t ce = cudaStreamEndCapture(st, &fit->graph);
    if (le != cudaSuccess) return rn_fail(RESNMTF_E_CUDA, std::string("kernel launch during capture: ") + cudaGetErrorString(le));
    RN_CUDA(ce);
    RN_CUDA(cudaGraphInstantiate(&fit->graph_exec, fit->graph, 0));
  }
  return RESNMTF_OK;
}

// every stream of a placed fit has drained (the home stream alone orders nothing on the other GPUs)
static int sync_placed(resnmtf_fit* fit) {
  for (size_t i = 1; i < fit->devs.size(); ++i) {
    RN_CUDA(cudaSetDevice(fit->devs[i].ctx->device));
    RN_CUDA(cudaStreamSynchronize(fit->devs[i].ctx->stream));
  }
  if (fit->devs.size() > 1) RN_CUDA(cudaSetDevice(fit->ctx->device));
  return RESNMTF_OK;
}

static int push_ctrl(resnmtf_fit* fit) {
  RN_CUDA(cudaMemcpyAsync(fit->d_ctrl, &fit->h_ctrl, sizeof(RnCtrl), cudaMemcpyHostToDevice, fit->ctx->stream));
  if (rn_placed(fit)) RN_CUDA(cudaStreamSynchronize(fit->ctx->stream));  // kernels on the other GPUs read it too
  return RESNMTF_OK;
}

// Reads the control block and drains the error history accumulated since the last call.
static int pull_ctrl(resnmtf_fit* fit) {
  cudaStream_t st = fit->ctx->stream;
  if (int rc = sync_placed(fit)) return rc;
  RN_CUDA(cudaMemcpyAsync(&fit->h_ctrl, fit->d_ctrl, sizeof(RnCtrl), cudaMemcpyDeviceToHost, st));
  RN_CUDA(cudaStreamSynchronize(st));
  const int64_t cnt = std::min<int64_t>(fit->h_ctrl.hist_count, fit->hist_cap);
  if (cnt > 0) {
    const size_t old = fit->errors.size();
    fit->errors.resize(old + (size_t)cnt);
    RN_CUDA(cudaMemcpy(fit->errors.data() + old, fit->d_hist, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost));
  }
  fit->h_ctrl.hist_count = 0;
  if (cnt > 0) {  // nothing is in flight here: hand the drained history buffer back to the device
    RN_CUDA(cudaMemcpyAsync(fit->d_ctrl, &fit->h_ctrl, sizeof(RnCtrl), cudaMemcpyHostToDevice, st));
  }
  return RESNMTF_OK;
}

static int run_batch(resnmtf_fit* fit, int64_t iters) {
  cudaStream_t st = fit->ctx->stream;
  if (fit->small) {  // the whole batch of sweeps is ONE launch of one persistent CTA
    rn_small_sweeps<<<1, RN_SM_THREADS, rn_small_smem(fit->small_dim), st>>>(fit->d, iters, fit->small_dim);
    RN_CUDA(cudaGetLastError());
    fit->launches_per_iter = 0;
    fit->counters.kernel_launches += 1;
    return RESNMTF_OK;
  }
  for (int64_t i = 0; i < iters; ++i) {
    if (fit->graph_exec) {
      RN_CUDA(cudaGraphLaunch(fit->graph_exec, st));
    } else {
      fit->launches_per_iter = enqueue_iteration(fit, st, nullptr, nullptr);
      RN_CUDA(cudaGetLastError());
    }
  }
  fit->counters.kernel_launches += iters * fit->launches_per_iter;
  if (fit->comm_rc) return fit->comm_rc;
  return RESNMTF_OK;
}

static double alg_bytes_per_iter(const resnmtf_fit* fit) {
  // B_alg = sum_v 8 * [2 n p + 4 n k + 4 p k + C_v]
  double b = 0.0;
  const int V = fit->V;
  for (int v = 0; v < V; ++v) {
    const RnView& d = fit->views[v].d;
    double cv = 0.0;
    for (int w = 0; w < V; ++w) {
      if (w == v) continue;
      if (fit->h_phi[w + (size_t)v * V] != 0.0 && fit->h_rowmode[w + (size_t)v * V] != RN_MODE_NA) cv += (double)d.n * d.k;
      if (fit->h_psi[w + (size_t)v * V] != 0.0 && fit->h_colmode[w + (size_t)v * V] != RN_MODE_NA) cv += (double)d.p * d.k;
    }
    // the one-pass fused kernel reads X once: B_min = B_alg - 8 n p
    b += 8.0 * ((d.fu_csize ? 1.0 : 2.0) * d.n * d.p + 4.0 * d.n * d.k + 4.0 * d.p * d.k + cv);
  }
  return b;
}

// AUTO error mode hand-over: the device paused (done == 3) at the end of a sweep whose algebraic error
// fell below RN_AUTO_DIRECT_BELOW (1e-4).  Re-evaluate that sweep's error with the direct residual pass, do its bookkeeping,
// and continue in DIRECT mode for the rest of the fit.
static int handle_pause(resnmtf_fit* fit) {
  cudaStream_t st = fit->ctx->stream;
  fit->h_ctrl.done = 0;
  fit->h_ctrl.want_direct = 0;
  int rc;
  if ((rc = push_ctrl(fit))) return rc;
  int nl = 1;
  if (rn_placed(fit)) {  // every view's residual on its own GPU, then the bookkeeping on the home GPU
    for (int v = 0; v < fit->V; ++v) {
      RnViewScope scope(fit, v);
      nl += launch_residual(fit->views[v], fit->devs[fit->vdev[v]].d, 1, fit->cur->stream, nullptr, &fit->comm_rc);
      RN_CUDA(cudaGetLastError());
    }
    if ((rc = sync_placed(fit))) return rc;
  } else {
    for (int v = 0; v < fit->V; ++v) nl += launch_residual(fit->views[v], fit->d, 1, st, fit->ctx, &fit->comm_rc);
  }
  rn_finish<<<1, 1, 0, st>>>(fit->d, 0);
  RN_CUDA(cudaGetLastError());
  fit->counters.kernel_launches += nl;
  fit->auto_direct = true;
  fit->meta_dirty = true;
  if ((rc = pull_ctrl(fit))) return rc;
  return prepare(fit);
}

// Developer timeline of the last fused launch (RESNMTF_FU_TIMELINE=1): per-CTA stamps, the per-row-group trace of CTA 0
// and the cycles every consumer warp spent waiting for X and for F_new, to stderr.  A no-op without the switch.
static int rn_print_fused_timeline(resnmtf_fit* fit) {
  for (int v = 0; v < fit->V; ++v) {
    const RnView& d = fit->views[v].d;
    if (!d.fu_timeline || !d.fu_csize) continue;
    const int grid = d.fu_clusters * d.fu_csize;
    std::vector<long long> tl((size_t)grid * 12);
    RN_CUDA(cudaMemcpy(tl.data(), d.fu_timeline, tl.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long t0 = tl[0];
    for (int b = 0; b < grid; ++b) t0 = std::min(t0, tl[(size_t)b * 12]);
    static const char* nm[9] = {"entry", "setup done", "main loop done", "cluster sync", "all T published",
                                "Ts + F'F ready", "group epilogue done", "view finished", "G'G|A partials summed"};
    std::fprintf(stderr, "[resnmtf fused timeline] view %d, grid %d (ns after the first CTA's entry: min / max)\n", v, grid);
    for (int sidx = 0; sidx < 9; ++sidx) {
      long long lo = -1, hi = -1;
      for (int b = 0; b < grid; ++b) {
        const long long x = tl[(size_t)b * 12 + sidx];
        if (x < t0) continue;  // stamp not written by this CTA in the last launch
        if (lo < 0 || x - t0 < lo) lo = x - t0;
        if (x - t0 > hi) hi = x - t0;
      }
      std::fprintf(stderr, "  %-22s %8lld %8lld\n", nm[sidx], lo, hi);
    }
    if (d.fu_trace && d.fu_kind == 1) {  // per-row-group trace of CTA 0 (cluster 0): averages over its groups, ns
      const int64_t groups = (d.n + 7) / 8;
      const int ngl = (int)(groups / d.fu_clusters + (groups % d.fu_clusters ? 1 : 0));
      std::vector<long long> tr((size_t)ngl * 32);
      RN_CUDA(cudaMemcpy(tr.data(), d.fu_trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      // stamps of group i: F phase 0 (asks for X) / 7 (X there); G phase 1 (F-phase MMAs of group i+1 issued, asks for
      // F_new of i) / 2 (F_new there) / 3 (done); epilogue 4 (warp partials in) / 5 (cluster partials in) / 6 (F_new out)
      double wait_x = 0, f_ph = 0, wait_f = 0, g_ph = 0, period = 0, ep_exch = 0, ep_math = 0, early = 0, pub2in = 0;
      int cnt = 0;
      for (int i = 4; i + 2 < ngl; ++i) {  // steady state
        const long long* a = &tr[(size_t)i * 32];
        const long long* nx = &tr[(size_t)(i + 1) * 32];
        wait_x += (double)(nx[7] - nx[0]);
        f_ph += (double)(a[1] - nx[7]);
        wait_f += (double)(a[2] - a[1]);
        g_ph += (double)(a[3] - a[2]);
        period += (double)(nx[3] - a[3]);
        ep_exch += (double)(a[5] - a[4]);
        ep_math += (double)(a[6] - a[5]);
        early += (double)(a[1] - a[6]);    // > 0: F_new of group i was out before consumer warp 0 asked for it
        pub2in += (double)(nx[4] - a[2]);  // consumer warp 0 publishes group i+1 just after stamp 2 of group i
        ++cnt;
      }
      if (cnt > 0) {  // per consumer warp: G-phase start and publication of the next group, relative to warp 0's stamp 2
        std::fprintf(stderr, "  per warp (ns after consumer warp 0 got F_new): G phase starts | publishes next group\n   ");
        for (int w = 0; w < 9; ++w) {
          double gs_ = 0, pb = 0;
          for (int i = 4; i + 2 < ngl; ++i) {
            gs_ += (double)(tr[(size_t)i * 32 + 17 + w] - tr[(size_t)i * 32 + 2]);
            pb += (double)(tr[(size_t)(i + 1) * 32 + 8 + w] - tr[(size_t)i * 32 + 2]);
          }
          std::fprintf(stderr, " w%d %.0f|%.0f", w, gs_ / cnt, pb / cnt);
        }
        std::fprintf(stderr, "\n");
      }
      if (cnt > 0)
        std::fprintf(stderr,
                     "  per row group (CTA 0, %d groups, ns): period %.0f | consumer warp 0: wait X %.0f, F phase %.0f, "
                     "wait F_new %.0f, G phase %.0f | epilogue: own publish -> all warp partials in %.0f, -> cluster "
                     "partials in %.0f, -> F_new out %.0f; F_new out %.0f before it is asked for\n",
                     cnt, period / cnt, wait_x / cnt, f_ph / cnt, wait_f / cnt, g_ph / cnt, pub2in / cnt, ep_exch / cnt,
                     ep_math / cnt, early / cnt);
    }
    if (d.fu_trace && d.fu_kind == 2) {  // rn_fused2_step: per-group stamps of CTA 0, ns relative to warp 0's publication
      const int64_t groups = (d.n + 7) / 8;
      const int ngl = (int)(groups / d.fu_clusters + (groups % d.fu_clusters ? 1 : 0));
      std::vector<long long> tr((size_t)ngl * 32);
      RN_CUDA(cudaMemcpy(tr.data(), d.fu_trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      static const char* nm2[32] = {"w0 F phase asks X", "w0 X(A) there", "w0 published", "w0 G phase asks F_new",
                                    "w0 F_new there", "w0 G phase done", "", "", "E starts", "E old F rows there",
                                    "E warp partials in", "E partial sent", "E cluster partials in", "E F_new out",
                                    "E done", "", "pub w0", "pub w1", "pub w2", "pub w3", "pub w4", "pub w5", "pub w6",
                                    "pub w7", "pub w8", "pub w9", "pub w10", "pub w11", "", "", "", ""};
      double mean[32] = {0}, period = 0;
      int cnt = 0;
      for (int i = 6; i + 4 < ngl; ++i) {
        const long long* a = &tr[(size_t)i * 32];
        for (int q = 0; q < 32; ++q) mean[q] += (double)(a[q] - a[2]);
        period += (double)(tr[(size_t)(i + 1) * 32 + 2] - a[2]);
        ++cnt;
      }
      if (cnt > 0) {
        std::fprintf(stderr, "  per row group (CTA 0, %d groups): period %.0f ns; stamps relative to warp 0's publication:\n", cnt,
                     period / cnt);
        for (int q = 0; q < 28; ++q)
          if (nm2[q][0]) std::fprintf(stderr, "    %-24s %8.0f\n", nm2[q], mean[q] / cnt);
      }
    }
    if (d.fu_waits && d.fu_kind == 1) {  // cycles per row group a consumer warp waited for X (ring) and for F_new (epilogue), mean over CTAs
      std::vector<long long> wt((size_t)grid * RN_FU_NCW * 2);
      RN_CUDA(cudaMemcpy(wt.data(), d.fu_waits, wt.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      const double groups_per_cluster = (double)((d.n + 7) / 8) / d.fu_clusters;
      std::fprintf(stderr, "  cycles per row group waited for X | for F_new, per consumer warp (mean over the %d CTAs):\n   ", grid);
      for (int w = 0; w < RN_FU_NCW; ++w) {
        double sx = 0, sf = 0;
        for (int b = 0; b < grid; ++b) {
          sx += (double)wt[((size_t)b * RN_FU_NCW + w) * 2];
          sf += (double)wt[((size_t)b * RN_FU_NCW + w) * 2 + 1];
        }
        std::fprintf(stderr, " w%d %.0f|%.0f", w, sx / grid / groups_per_cluster, sf / grid / groups_per_cluster);
      }
      std::fprintf(stderr, "\n");
    }
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_run(resnmtf_fit* fit, int64_t n_iters, double tol, int64_t max_iters,
                               int64_t* iters_done) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_run: fit is NULL");
  int rc;
  if ((rc = prepare(fit))) return rc;
  cudaStream_t st = fit->ctx->stream;
  const bool conv = n_iters < 0;
  fit->h_ctrl.done = 0;
  fit->h_ctrl.conv_mode = conv ? 1 : 0;
  fit->h_ctrl.tol = tol;
  fit->h_ctrl.hist_count = 0;
  fit->h_ctrl.direct_passes = 0;
  fit->h_ctrl.want_direct = 0;
  if ((rc = push_ctrl(fit))) return rc;
  fit->counters.kernel_launches = 0;
  const int64_t it0 = fit->h_ctrl.iters;
  RN_CUDA(cudaEventRecord(fit->ctx->ev0, st));
  // sweeps enqueued between two reads of the device-side stop flag: a read costs a stream synchronisation (~30 us), a
  // sweep launched after the flag fired returns at once (~3 us)
  const int64_t conv_batch = std::max(1, rn_env_int("RESNMTF_CONV_BATCH", 32));
  for (;;) {
    const int64_t donei = fit->h_ctrl.iters - it0;
    int64_t b;
    if (!conv) {
      b = std::min<int64_t>(n_iters - donei, fit->hist_cap);
    } else {
      b = conv_batch;
      if (max_iters > 0) b = std::min(b, max_iters - donei);
    }
    if (b <= 0) break;
    if ((rc = run_batch(fit, b))) return rc;
    // fixed mode without AUTO hand-over pending needs no read-back until the end
    if ((rc = pull_ctrl(fit))) return rc;
    if (fit->h_ctrl.done == 3) {
      if ((rc = handle_pause(fit))) return rc;
    }
    if (fit->h_ctrl.done == 1 || fit->h_ctrl.done == 2) break;
  }
  RN_CUDA(cudaEventRecord(fit->ctx->ev1, st));
  if ((rc = pull_ctrl(fit))) return rc;
  float ms = 0.f;
  RN_CUDA(cudaEventElapsedTime(&ms, fit->ctx->ev0, fit->ctx->ev1));
  fit->counters.device_ms = ms;
  fit->counters.iterations = fit->h_ctrl.iters;
  fit->counters.converged = fit->h_ctrl.done == 1;
  fit->counters.impl = fit->impl;
  fit->counters.direct_error_passes = fit->h_ctrl.direct_passes;
  fit->counters.alg_bytes_per_iter = alg_bytes_per_iter(fit);
  if (iters_done) *iters_done = fit->h_ctrl.iters - it0;
  if (int trc = rn_print_fused_timeline(fit)) return trc;  // developer switch RESNMTF_FU_TIMELINE=1, else a no-op
  if (conv && fit->h_ctrl.done == 2)
    return rn_fail(RESNMTF_E_NAN, "mean error is NaN: missing value where TRUE/FALSE needed (R/main.r:55)");
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_step(resnmtf_fit* fit) { return resnmtf_fit_run(fit, 1, 0.0, 0, nullptr); }

extern "C" int resnmtf_fit_profile(resnmtf_fit* fit, int64_t n_iters, double ms[5], int64_t launches[5]) {
  RN_CHECK(fit && ms && launches, RESNMTF_E_INVALID, "resnmtf_fit_profile: NULL argument");
  RN_CHECK(!rn_placed(fit), RESNMTF_E_UNSUPPORTED, "resnmtf_fit_profile: not available for a placed fit");
  int rc;
  if ((rc = prepare(fit))) return rc;
  cudaStream_t st = fit->ctx->stream;
  fit->h_ctrl.done = 0;
  fit->h_ctrl.conv_mode = 0;
  fit->h_ctrl.hist_count = 0;
  if ((rc = push_ctrl(fit))) return rc;
  for (int i = 0; i < 5; ++i) {
    ms[i] = 0.0;
    launches[i] = 0;
  }
  for (int64_t it = 0; it < n_iters; ++it) {
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls;
    enqueue_iteration(fit, st, &ev, &cls);
    RN_CUDA(cudaGetLastError());
    RN_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i + 1 < ev.size(); ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      if (cls[i] >= 0 && cls[i] < 5) {
        ms[cls[i]] += t;
        launches[cls[i]] += 1;
      }
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if ((rc = pull_ctrl(fit))) return rc;
    if (fit->h_ctrl.done == 3 && (rc = handle_pause(fit))) return rc;
  }
  if ((rc = pull_ctrl(fit))) return rc;
  fit->counters.iterations = fit->h_ctrl.iters;
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
extern "C" int resnmtf_fit_get_factors(resnmtf_fit* fit, int v, double* f, double* s, double* g,
                                       double* lambda, double* mu) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_get_factors: fit is NULL");
  RN_CHECK(v >= 0 && v < fit->V, RESNMTF_E_INVALID, "resnmtf_fit_get_factors: view index out of range");
  const ViewHost& vh = fit->views[v];
  RN_CHECK(vh.has_factors, RESNMTF_E_STATE, "resnmtf_fit_get_factors: factors were never set");
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  RnViewScope scope(fit, v);
  cudaStream_t st = fit->cur->stream;
  const int K = vh.d.k, KP = vh.d.kp;
  const int64_t n = vh.d.n, p = vh.d.p;
  std::vector<double> gt, fp;
  if (f) {
    fp.resize((size_t)vh.d.ldx * KP);
    RN_CUDA(cudaMemcpyAsync(fp.data(), vh.d.F, fp.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  if (g) {
    gt.resize((size_t)vh.d.pp * KP);
    RN_CUDA(cudaMemcpyAsync(gt.data(), vh.d.G, gt.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  if (s) RN_CUDA(cudaMemcpyAsync(s, vh.d.S, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (lambda) RN_CUDA(cudaMemcpyAsync(lambda, vh.d.lam, K * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (mu) RN_CUDA(cudaMemcpyAsync(mu, vh.d.mu, K * sizeof(double), cudaMemcpyDeviceToHost, st));
  RN_CUDA(cudaStreamSynchronize(st));
  if (f)
    for (int c = 0; c < K; ++c)
      for (int64_t r = 0; r < n; ++r) f[(size_t)c * n + r] = fp[(size_t)rn_fidx_host(r, c, KP)];
  if (g)
    for (int c = 0; c < K; ++c)
      for (int64_t j = 0; j < p; ++j) g[(size_t)c * p + j] = gt[(size_t)j * KP + c];
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_normalise(resnmtf_fit* fit) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_normalise: fit is NULL");
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  for (int v = 0; v < fit->V; ++v) {
    RnViewScope scope(fit, v);
    cudaStream_t st = fit->cur->stream;
    const ViewHost& vh = fit->views[v];
    RN_CHECK(vh.has_factors, RESNMTF_E_STATE, "resnmtf_fit_normalise: factors were never set");
    rn_factor_sums<<<2 * vh.d.k + vh.d.k * vh.d.k, 1024, 0, st>>>(vh.d);
    int rc2 = rn_allreduce(fit->ctx, vh.d.csF, (size_t)vh.d.k);
    if (rc2) return rc2;
    const int64_t m = std::max(vh.d.n, vh.d.p);
    rn_normalise<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(vh.d);
    rn_factor_sums<<<2 * vh.d.k + vh.d.k * vh.d.k, 1024, 0, st>>>(vh.d);  // keep G'G / colsums consistent with the scaled factors
    if ((rc2 = rn_allreduce(fit->ctx, vh.d.csF, (size_t)vh.d.k))) return rc2;
    RN_CUDA(cudaGetLastError());
    RN_CUDA(cudaStreamSynchronize(st));
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_get_errors(resnmtf_fit* fit, double* out, int64_t cap, int64_t* count) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_get_errors: fit is NULL");
  const int64_t n = (int64_t)fit->errors.size();
  if (count) *count = n;
  if (out && cap > 0) std::memcpy(out, fit->errors.data(), (size_t)std::min(n, cap) * sizeof(double));
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_get_view_errors(resnmtf_fit* fit, double* err, double* data_norms) {
  RN_CHECK(fit != nullptr, RESNMTF_E_INVALID, "resnmtf_fit_get_view_errors: fit is NULL");
  RN_CUDA(cudaSetDevice(fit->ctx->device));
  for (int v = 0; v < fit->V; ++v) {
    double sc[2];
    RN_CUDA(cudaMemcpy(sc, fit->views[v].d.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    if (data_norms) data_norms[v] = sc[0];
    if (err) err[v] = sc[1];
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_fit_get_counters(resnmtf_fit* fit, resnmtf_counters* out) {
  RN_CHECK(fit && out, RESNMTF_E_INVALID, "resnmtf_fit_get_counters: NULL argument");
  fit->counters.alg_bytes_per_iter = alg_bytes_per_iter(fit);
  *out = fit->counters;
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------
// row-sharded path (NCCL) -- see DESIGN.md "Multi-GPU"
// ------------------------------------------------------------------------------------------------
// ---- post-fit reductions -----------------------------------------------------------
extern "C" int resnmtf_jsd_pairs(resnmtf_ctx* ctx, const double* vecs, int64_t n, int32_t m, int64_t ld,
                                 const double* bw, const double* vmax, const int32_t* pair_a,
                                 const int32_t* pair_b, int64_t n_pairs, double* out) {
  RN_CHECK(ctx && vecs && bw && vmax && out, RESNMTF_E_INVALID, "resnmtf_jsd_pairs: NULL argument");
  RN_CHECK(n >= 1 && m >= 1 && ld >= n && n_pairs >= 0, RESNMTF_E_INVALID, "resnmtf_jsd_pairs: bad shape");
  if (n_pairs == 0) return RESNMTF_OK;
  RN_CHECK(pair_a && pair_b, RESNMTF_E_INVALID, "resnmtf_jsd_pairs: NULL pair list");
  for (int64_t i = 0; i < n_pairs; ++i)
    RN_CHECK(pair_a[i] >= 0 && pair_a[i] < m && pair_b[i] >= 0 && pair_b[i] < m, RESNMTF_E_INVALID,
             "resnmtf_jsd_pairs: column index out of range");
  RN_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  double *d_vecs = nullptr, *d_par = nullptr, *d_out = nullptr;
  int32_t* d_pairs = nullptr;
  cudaError_t e = rn_dev_alloc(ctx, (void**)&d_vecs, (size_t)n * m * sizeof(double));
  if (e == cudaSuccess) e = rn_dev_alloc(ctx, (void**)&d_par, (size_t)2 * m * sizeof(double));
  if (e == cudaSuccess) e = rn_dev_alloc(ctx, (void**)&d_pairs, (size_t)2 * n_pairs * sizeof(int32_t));
  if (e == cudaSuccess) e = rn_dev_alloc(ctx, (void**)&d_out, (size_t)n_pairs * sizeof(double));
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(d_vecs, (size_t)n * sizeof(double), vecs, (size_t)ld * sizeof(double),
                          (size_t)n * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_par, bw, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_par + m, vmax, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_pairs, pair_a, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_pairs + n_pairs, pair_b, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    const int64_t grid = std::min<int64_t>(n_pairs, (int64_t)ctx->sm_count * 8);
    e = cudaFuncSetAttribute(rn_jsd_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RnKdeSmem));
  }
  if (e == cudaSuccess) {
    const int64_t grid = std::min<int64_t>(n_pairs, (int64_t)ctx->sm_count * 8);
    rn_jsd_pairs<<<(unsigned)grid, RN_KDE_N, sizeof(RnKdeSmem), st>>>(d_vecs, n, n, d_par, d_par + m, d_pairs, d_pairs + n_pairs,
                                                       n_pairs, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n_pairs * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  rn_dev_free(ctx, d_vecs);
  rn_dev_free(ctx, d_par);
  rn_dev_free(ctx, d_pairs);
  rn_dev_free(ctx, d_out);
  if (e != cudaSuccess)
    return rn_fail(e == cudaErrorMemoryAllocation ? RESNMTF_E_NOMEM : RESNMTF_E_CUDA,
                   std::string("resnmtf_jsd_pairs: ") + cudaGetErrorString(e));
  return RESNMTF_OK;
}

extern "C" int resnmtf_comm_id_size(void) { return (int)sizeof(rn_ncclUniqueId); }

extern "C" int resnmtf_comm_id_create(void* id_out) {
  RN_CHECK(id_out != nullptr, RESNMTF_E_INVALID, "resnmtf_comm_id_create: NULL argument");
  RN_CHECK(rn_nccl().ok, RESNMTF_E_COMM, "libnccl.so.2 could not be loaded");
  rn_ncclUniqueId id;
  RN_NCCL(rn_nccl().GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return RESNMTF_OK;
}

extern "C" int resnmtf_ctx_join(resnmtf_ctx* ctx, const void* id, int rank, int n_ranks) {
  RN_CHECK(ctx && id, RESNMTF_E_INVALID, "resnmtf_ctx_join: NULL argument");
  RN_CHECK(n_ranks >= 1 && rank >= 0 && rank < n_ranks, RESNMTF_E_INVALID, "resnmtf_ctx_join: bad rank");
  RN_CHECK(ctx->comm == nullptr, RESNMTF_E_STATE, "resnmtf_ctx_join: the context already joined a communicator");
  RN_CHECK(rn_nccl().ok, RESNMTF_E_COMM, "libnccl.so.2 could not be loaded");
  RN_CUDA(cudaSetDevice(ctx->device));
  rn_ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  RN_NCCL(rn_nccl().CommInitRank(&ctx->comm, n_ranks, uid, rank));
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  return RESNMTF_OK;
}
