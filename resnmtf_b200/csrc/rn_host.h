// Host-side declarations shared by the translation units of libresnmtf_b200.so: handles behind the C ABI
// (include/resnmtf_b200.h), error plumbing, device allocation from the context's private pool, lazy NCCL binding.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/resnmtf_b200.h"
#include "rn_types.h"

#include <dlfcn.h>

// NCCL is only needed by the row-sharded path, so it is resolved lazily with dlopen (no link-time
// dependency): the handful of types / enum values used here are ABI-stable across NCCL 2.x.
typedef struct rn_nccl_comm* rn_ncclComm_t;
typedef struct { char internal[128]; } rn_ncclUniqueId;
enum { RN_NCCL_SUCCESS = 0, RN_NCCL_SUM = 0, RN_NCCL_INT64 = 4, RN_NCCL_FLOAT64 = 8 };
struct RnNccl {
  void* handle = nullptr;
  int (*GetUniqueId)(rn_ncclUniqueId*) = nullptr;
  int (*CommInitRank)(rn_ncclComm_t*, int, rn_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(rn_ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, rn_ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
inline RnNccl& rn_nccl() {
  static RnNccl api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // Order: (1) an NCCL this process has ALREADY loaded (a host framework's bundled copy: loading a second, older
    // libnccl.so.2 beside it -- or before it -- would make the framework bind to the wrong one by SONAME);
    // (2) the path in RESNMTF_NCCL_LIB (the Python binding points it at the pip-bundled library before first use);
    // (3) the system library.  RTLD_LOCAL: this library's NCCL symbols are looked up with dlsym, nobody else's are
    // affected.
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
      if (api.handle) break;
    }
    if (!api.handle) {
      const char* forced = std::getenv("RESNMTF_NCCL_LIB");
      if (forced && *forced) api.handle = dlopen(forced, RTLD_NOW | RTLD_LOCAL);
    }
    for (const char* nm : names) {
      if (api.handle) break;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    }
    if (api.handle) {
      api.GetUniqueId = (int (*)(rn_ncclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(rn_ncclComm_t*, int, rn_ncclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (int (*)(rn_ncclComm_t))dlsym(api.handle, "ncclCommDestroy");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, rn_ncclComm_t, cudaStream_t))dlsym(
          api.handle, "ncclAllReduce");
      api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    }
  }
  return api;
}

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
// thread-local message behind resnmtf_last_error() (defined in resnmtf_capi.cu)
std::string& rn_err_slot();
inline int rn_fail(int code, const std::string& msg) {
  rn_err_slot() = msg;
  return code;
}

#define RN_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return rn_fail(e__ == cudaErrorMemoryAllocation ? RESNMTF_E_NOMEM : RESNMTF_E_CUDA,     \
                     std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
  } while (0)

#define RN_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    int r__ = (expr);                                                                                   \
    if (r__ != RN_NCCL_SUCCESS)                                                                         \
      return rn_fail(RESNMTF_E_COMM, std::string(#expr) + ": " + rn_nccl().GetErrorString(r__));        \
  } while (0)

#define RN_CHECK(cond, code, msg) \
  do {                            \
    if (!(cond)) return rn_fail(code, msg); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
#define RN_UP_T 6
#define RN_UP_BYTES ((size_t)4 << 20)
struct resnmtf_ctx {
  int device = 0;
  int sm_count = 148;
  size_t l2_persist_bytes = 0;  // persisting-L2 carve-out granted to this device (0: unavailable / disabled)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int rank = 0, n_ranks = 1;
  rn_ncclComm_t comm = nullptr;  // non-null: every view of every fit on this context is ROW-SHARDED over the communicator's
                                 // ranks -- also for a communicator of ONE rank, which runs the whole sharded code path
                                 // (pack -> all-reduce -> stand-alone G epilogue) on a single GPU
  // Device memory comes from the stream-ordered pool of the device with an unlimited release threshold: after the
  // first fit the create / destroy path of a fit never reaches the driver's allocator (measured: cudaMalloc +
  // cudaFree cost 3 ms per fit on one box and 15 ms on another -- more than the upload of the factors).
  bool pooled = false;
  cudaMemPool_t pool = nullptr;  // PRIVATE to this context (never the device's default pool, which other users of the
                                 // process -- torch's cudaMallocAsync backend, other contexts -- share)
  // The context outlives every fit / data handle created on it: they hold a reference, resnmtf_ctx_destroy only drops
  // the creator's, and the last one out tears the context down.
  std::atomic<int> refs{1};
  // host -> device upload pipeline (two staging buffers, copy stream, events), created on first use and kept
  double* stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  cudaStream_t copy_st = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, tiled[2] = {nullptr, nullptr};
  // PAGEABLE host sources (what R hands over): RN_UP_T host threads copy column groups into their own pinned buffers
  // (two of RN_UP_BYTES each) and send them on their own streams -- the driver's single-threaded staging of a pageable
  // cudaMemcpy runs at ~10 GB/s, six threads saturate the link
  double* up_pin[RN_UP_T][2] = {};
  cudaEvent_t up_ev[RN_UP_T][2] = {};
  cudaStream_t up_st[RN_UP_T] = {};
  bool up_ready = false;
};

inline cudaError_t rn_dev_alloc(resnmtf_ctx* ctx, void** p, size_t bytes) {
  if (ctx->pooled) return cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream);
  return cudaMalloc(p, bytes);
}
inline cudaError_t rn_dev_free(resnmtf_ctx* ctx, void* p) {
  if (!p) return cudaSuccess;
  if (ctx->pooled) return cudaFreeAsync(p, ctx->stream);
  return cudaFree(p);
}

// A view's X in the device layout, shareable between fits (the k-sweep of apply_resnmtf fits the same
// data for every k, R/main.r:279-287): reference-counted, freed when the last holder lets go.
struct resnmtf_data {
  resnmtf_ctx* ctx = nullptr;
  int64_t n = 0, p = 0, ldx = 0, pp = 0;
  double* X = nullptr;
  double* X8 = nullptr;  // 8-row-group copy for the one-pass fused kernel (built on demand by the first plan)
  int64_t pp8 = 0;
  double xnorm2 = 0.0;
  std::atomic<int> refs{1};
  // |U| (n x svd_kc), d, |V| (p x svd_kc) of the view, svd_kc = min(16, n, p): computed by the first fit that asks for
  // the SVD initialisation (init_mats_inner, R/update_steps.r:92-95) and shared by every fit of this data -- the
  // k-sweep slices the same triplets for every k.  Host copies, column-major.
  std::vector<double> svd_u, svd_d, svd_v;
  int svd_kc = 0;
  std::mutex svd_mu;  // the cache is filled by the first fit that asks and copied by resnmtf_data_copy, possibly from
                      // different worker threads of a pool: held only while the vectors are published or copied
  std::mutex svd_compute_mu;  // held for the whole computation (one at a time per handle); copies do not wait for it
};

void rn_ctx_release(resnmtf_ctx* ctx);  // resnmtf_capi.cu

inline void rn_data_release(resnmtf_data* d) {
  if (d && d->refs.fetch_sub(1) == 1) {
    resnmtf_ctx* ctx = d->ctx;
    cudaSetDevice(ctx->device);
    rn_dev_free(ctx, d->X);
    rn_dev_free(ctx, d->X8);
    delete d;
    rn_ctx_release(ctx);
  }
}

struct ViewHost {
  RnView d;  // device pointers + geometry (passed by value to the kernels)
  resnmtf_data* shared = nullptr;  // non-null: X belongs to a shared data handle
  size_t l2_window = 0;            // bytes at the head of X pinned in L2 (persisting access-policy window)
  int impl = RESNMTF_IMPL_TMA;     // kernel family this view runs (a fit may mix FUSED and TMA views)
  double* x8_own = nullptr;        // X8 owned by the fit (views without a shared data handle)
  bool has_data = false, has_factors = false;
  std::vector<int32_t*> rowmaps, colmaps;  // [V] device maps of this view into view w (or null)
  double* xpart = nullptr;                 // ||X||^2 partials
  int32_t* xticket = nullptr;
};

// Per-GPU copy of a fit's metadata (PLACED fits only: the views of one fit live on several GPUs of one process,
// resnmtf_fit_create_placed).  Every device reads its own copy of the view table, the restriction matrices and the map
// tables; the pointers inside refer to peer memory where a partner view lives on another GPU.
struct RnDevMeta {
  resnmtf_ctx* ctx = nullptr;
  RnFit d;
  RnView* d_views = nullptr;
  double *d_phi = nullptr, *d_xi = nullptr, *d_psi = nullptr;
  const int32_t** d_rowmap = nullptr;
  const int32_t** d_colmap = nullptr;
  int8_t *d_rowmode = nullptr, *d_colmode = nullptr;
};

struct resnmtf_fit {
  resnmtf_ctx* ctx = nullptr;   // home context: control block, error history, metadata of a single-GPU fit
  resnmtf_ctx* cur = nullptr;   // context device allocations / frees go to right now (RnViewScope; normally == ctx)
  int V = 0;
  std::vector<ViewHost> views;
  std::vector<std::pair<resnmtf_ctx*, void*>> allocs;
  // placed fit (coupled views on different GPUs): context of every view, the distinct contexts with their
  // metadata copies (devs[0] is the home context), the view -> devs index, and one event per view recorded after the
  // view's last kernel of a sweep (what the Gauss-Seidel order of update_matrices() chains on across GPUs)
  std::vector<resnmtf_ctx*> vctx;
  std::vector<RnDevMeta> devs;
  std::vector<int> vdev;
  std::vector<cudaEvent_t> vev;
  RnFit d;               // passed by value to the kernels
  RnView* d_views = nullptr;
  RnCtrl* d_ctrl = nullptr;
  double *d_phi = nullptr, *d_xi = nullptr, *d_psi = nullptr;
  const int32_t** d_rowmap = nullptr;
  const int32_t** d_colmap = nullptr;
  int8_t *d_rowmode = nullptr, *d_colmode = nullptr;
  std::vector<const int32_t*> h_rowmap, h_colmap;
  std::vector<int8_t> h_rowmode, h_colmode;
  std::vector<double> h_phi, h_xi, h_psi;
  double* d_hist = nullptr;
  int64_t hist_cap = 4096;
  std::vector<double> errors;  // All_Error
  int err_mode = RESNMTF_ERR_AUTO;
  int impl_req = RESNMTF_IMPL_AUTO;
  int impl = RESNMTF_IMPL_TMA;
  bool meta_dirty = true;   // device copies of views / maps / restrictions need a refresh
  bool plan_dirty = true;   // grids / workspaces / graph need a rebuild
  bool auto_direct = false; // AUTO error mode has handed over to the direct residual pass
  bool pdl_ok = true;       // the fused launches of this fit may carry the programmatic-dependent-launch attribute
  int small_dim = 64;       // largest padded row / column count of the views (shared-memory copies of F and G)
  bool small = false;       // every view fits one SM's caches: the whole loop runs as one persistent CTA (rn_small.cuh)
  int comm_rc = 0;          // first NCCL failure seen while enqueueing (row-sharded path)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int64_t launches_per_iter = 0;
  resnmtf_counters counters{};
  RnCtrl h_ctrl{};
};

inline bool rn_placed(const resnmtf_fit* f) { return !f->vctx.empty(); }
inline resnmtf_ctx* rn_vctx(const resnmtf_fit* f, int v) { return f->vctx.empty() ? f->ctx : f->vctx[v]; }

// Makes the context of view v (v < 0: the home context) the fit's current one -- CUDA device, allocation target, stream
// of fit->cur -- for the lifetime of the scope.  A no-op in effect for single-GPU fits.
struct RnViewScope {
  resnmtf_fit* f;
  resnmtf_ctx* saved;
  RnViewScope(resnmtf_fit* fit, int v) : f(fit), saved(fit->cur) {
    f->cur = v < 0 ? f->ctx : rn_vctx(f, v);
    if (f->cur != saved) cudaSetDevice(f->cur->device);
  }
  ~RnViewScope() {
    if (f->cur != saved) cudaSetDevice(saved->device);
    f->cur = saved;
  }
};

// device memory on the fit's CURRENT context (fit->cur), released with the fit
template <typename T>
inline int rn_alloc(resnmtf_fit* f, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  RN_CUDA(rn_dev_alloc(f->cur, &p, bytes));
  f->allocs.emplace_back(f->cur, p);
  if (zero) RN_CUDA(cudaMemsetAsync(p, 0, bytes, f->cur->stream));
  *out = static_cast<T*>(p);
  return RESNMTF_OK;
}

inline int rn_free(resnmtf_fit* f, void* p) {
  if (!p) return RESNMTF_OK;
  resnmtf_ctx* owner = f->cur;
  for (auto it = f->allocs.begin(); it != f->allocs.end(); ++it)
    if (it->second == p) {
      owner = it->first;
      f->allocs.erase(it);
      break;
    }
  if (owner != f->cur) {
    cudaSetDevice(owner->device);
    cudaError_t e = rn_dev_free(owner, p);
    cudaSetDevice(f->cur->device);
    RN_CUDA(e);
    return RESNMTF_OK;
  }
  RN_CUDA(rn_dev_free(owner, p));
  return RESNMTF_OK;
}

static inline int64_t rn_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

inline int rn_env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return (s && *s) ? std::atoi(s) : dflt;
}

// in-place sum over the ranks of a row-sharded context, on the context's stream (graph-capturable)
inline int rn_allreduce(resnmtf_ctx* ctx, void* buf, size_t count, bool is_int64 = false) {
  if (!ctx->comm) return RESNMTF_OK;
  RN_NCCL(rn_nccl().AllReduce(buf, buf, count, is_int64 ? RN_NCCL_INT64 : RN_NCCL_FLOAT64, RN_NCCL_SUM, ctx->comm,
                              ctx->stream));
  return RESNMTF_OK;
}

// host twin of rn_fidx (rn_kernels.cuh): position of F[r, c] in the swizzled 64-row panel layout
static inline int64_t rn_fidx_host(int64_t r, int c, int kp) {
  const int sigma = ((c & 1) << 2) | (c & 2);
  return ((r >> 6) * kp + c) * RN_ROW_TILE + 2 * ((int)((r & 63) >> 1) ^ sigma) + (r & 1);
}

// resnmtf_capi.cu: an empty (zeroed) view in the panel layout, and its completion (||X||_F^2 into the handle)
int rn_data_alloc(resnmtf_ctx* ctx, int64_t n, int64_t p, resnmtf_data** out);
int rn_data_seal(resnmtf_data* d);
// rn_native.cu: fills the handle's cache of top singular triplets (no-op when it is there)
int rn_data_svd(resnmtf_data* data);
void rn_data_svd_adopt(resnmtf_data* dst, resnmtf_data* src);  // dst takes src's cached triplets (same view on another GPU)
