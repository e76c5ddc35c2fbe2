// The fan-out of apply_resnmtf() behind the C ABI (SURVEY 8a row a13, 8e "independent fits"): one default call is 66
// convergence loops -- per k of the sweep one fit and num_repeats shuffled refits (R/main.r:270-299,
// R/obtain_bicl.r:31-42), then per stability resample the same again (R/stability_analysis.r:302-338) -- which the
// reference runs nested and serially (its %dopar% over k is unreachable).  Here each of them is a UNIT: "take data set
// `key`, optionally sub-sample it, optionally shuffle it, initialise (explicit factors, or the SVD initialisation with
// the caller's noise draw), run the loop, normalise, hand the factors back".  resnmtf_batch_run() deals the units of
// one phase, longest first, to one native worker thread per GPU of the pool; every step of a unit runs on that GPU
// through the entry points of this library.  No collective on the data path: the views are uploaded once and copied
// GPU to GPU on first use.  What a unit returns does not depend on where or when it ran.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

#include "rn_host.h"

struct resnmtf_pool {
  std::vector<resnmtf_ctx*> ctx;  // one context per GPU of the pool
  struct DataSet {
    int n_views = 0;
    std::vector<std::vector<resnmtf_data*>> on_gpu;  // [gpu][view]; empty until that GPU first needs the set
    int home = 0;                                    // GPU that holds the original handles
    // on_gpu[g] is filled under gpu_mu[g] only (by GPU g's worker or a resnmtf_pool_get for that GPU): the GPUs of a pool
    // fetch their copies of a set at the same time instead of one after the other under the pool's mutex
    std::unique_ptr<std::mutex[]> gpu_mu;
  };
  std::map<int, DataSet> sets;
  std::mutex mu;  // guards `sets` (handles are created under it; a set is only dropped between batches)
};

extern "C" int resnmtf_pool_create(const int* devices, int n_devices, resnmtf_pool** out) {
  RN_CHECK(out != nullptr, RESNMTF_E_INVALID, "resnmtf_pool_create: out is NULL");
  const int visible = resnmtf_device_count();
  RN_CHECK(visible >= 1, RESNMTF_E_CUDA, "resnmtf_pool_create: no CUDA device available (this library has no CPU fallback)");
  std::vector<int> devs;
  if (devices && n_devices > 0) {
    devs.assign(devices, devices + n_devices);
  } else {
    const int n = n_devices > 0 ? std::min(n_devices, visible) : visible;
    for (int d = 0; d < n; ++d) devs.push_back(d);
  }
  resnmtf_pool* p = new (std::nothrow) resnmtf_pool();
  RN_CHECK(p != nullptr, RESNMTF_E_NOMEM, "resnmtf_pool_create: out of host memory");
  for (int d : devs) {
    resnmtf_ctx* c = nullptr;
    int rc = resnmtf_ctx_create(d, &c);
    if (rc) {
      for (resnmtf_ctx* x : p->ctx) resnmtf_ctx_destroy(x);
      delete p;
      return rc;
    }
    p->ctx.push_back(c);
  }
  // device-to-device copies of the views go over NVLink when the GPUs can reach each other directly
  for (size_t a = 0; a < devs.size(); ++a)
    for (size_t b = 0; b < devs.size(); ++b) {
      if (a == b) continue;
      if (devs[a] == devs[b]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) == cudaSuccess && can) {
        cudaSetDevice(devs[a]);
        cudaDeviceEnablePeerAccess(devs[b], 0);
        // The views live in the contexts' private stream-ordered pools, and pool memory needs its own access grant: without
        // it cudaMemcpyPeerAsync stages the copy through the host (measured on 8 GPUs: seven 640 MB copies from the home GPU
        // at the start of a batch stretched the first unit of every GPU by 60-130 ms and the home GPU's SVD from 20 to
        // 126 ms).  GPU a may read and write what GPU b's pool hands out.
        if (p->ctx[b]->pooled) {
          cudaMemAccessDesc desc;
          std::memset(&desc, 0, sizeof(desc));
          desc.location.type = cudaMemLocationTypeDevice;
          desc.location.id = devs[a];
          desc.flags = cudaMemAccessFlagsProtReadWrite;
          cudaMemPoolSetAccess(p->ctx[b]->pool, &desc, 1);
        }
      }
      cudaGetLastError();
    }
  *out = p;
  return RESNMTF_OK;
}

extern "C" int resnmtf_pool_size(resnmtf_pool* pool) { return pool ? (int)pool->ctx.size() : 0; }

extern "C" int resnmtf_unit_size(void) { return (int)sizeof(resnmtf_unit); }

extern "C" resnmtf_ctx* resnmtf_pool_ctx(resnmtf_pool* pool, int gpu) {
  return (pool && gpu >= 0 && gpu < (int)pool->ctx.size()) ? pool->ctx[gpu] : nullptr;
}

static void drop_set(resnmtf_pool::DataSet& s) {
  for (auto& g : s.on_gpu)
    for (resnmtf_data* d : g) resnmtf_data_destroy(d);
  s.on_gpu.clear();
}

extern "C" int resnmtf_pool_drop(resnmtf_pool* pool, int key) {
  RN_CHECK(pool != nullptr, RESNMTF_E_INVALID, "resnmtf_pool_drop: pool is NULL");
  std::lock_guard<std::mutex> lk(pool->mu);
  auto it = pool->sets.find(key);
  if (it != pool->sets.end()) {
    drop_set(it->second);
    pool->sets.erase(it);
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_pool_destroy(resnmtf_pool* pool) {
  if (!pool) return RESNMTF_OK;
  for (auto& kv : pool->sets) drop_set(kv.second);
  for (resnmtf_ctx* c : pool->ctx) resnmtf_ctx_destroy(c);
  delete pool;
  return RESNMTF_OK;
}

extern "C" int resnmtf_pool_put(resnmtf_pool* pool, int key, int n_views, resnmtf_data* const* views) {
  RN_CHECK(pool && views && n_views >= 1, RESNMTF_E_INVALID, "resnmtf_pool_put: bad argument");
  int home = -1;
  for (int g = 0; g < (int)pool->ctx.size(); ++g)
    if (views[0] && views[0]->ctx == pool->ctx[g]) home = g;
  RN_CHECK(home >= 0, RESNMTF_E_INVALID, "resnmtf_pool_put: the handles must live on a context of the pool");
  for (int v = 0; v < n_views; ++v)
    RN_CHECK(views[v] && views[v]->ctx == pool->ctx[home], RESNMTF_E_INVALID,
             "resnmtf_pool_put: all views of a set must live on the same context");
  std::lock_guard<std::mutex> lk(pool->mu);
  auto it = pool->sets.find(key);
  if (it != pool->sets.end()) drop_set(it->second);
  resnmtf_pool::DataSet& s = pool->sets[key];
  s.n_views = n_views;
  s.home = home;
  s.on_gpu.assign(pool->ctx.size(), {});
  s.gpu_mu.reset(new std::mutex[pool->ctx.size()]);
  for (int v = 0; v < n_views; ++v) {
    views[v]->refs.fetch_add(1);  // the pool holds its own reference
    s.on_gpu[home].push_back(views[v]);
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_pool_put_host(resnmtf_pool* pool, int key, int n_views, const int64_t* n, const int64_t* p,
                                     const double* const* x, const int64_t* ld, int prep, int32_t* was_negative) {
  RN_CHECK(pool && n && p && x && n_views >= 1, RESNMTF_E_INVALID, "resnmtf_pool_put_host: bad argument");
  std::vector<resnmtf_data*> hs((size_t)n_views, nullptr);
  int rc = RESNMTF_OK;
  int32_t neg_any = 0;
  for (int v = 0; v < n_views && !rc; ++v) {
    int32_t neg = 0;
    const int64_t l = ld ? ld[v] : n[v];
    rc = prep ? resnmtf_data_create_prepped(pool->ctx[0], n[v], p[v], x[v], l, &neg, &hs[v])
              : resnmtf_data_create(pool->ctx[0], n[v], p[v], x[v], l, &hs[v]);
    neg_any |= neg;
  }
  if (!rc) rc = resnmtf_pool_put(pool, key, n_views, hs.data());
  for (resnmtf_data* h : hs) resnmtf_data_destroy(h);  // the pool keeps its own reference
  if (was_negative) *was_negative = neg_any;
  return rc;
}

// the views of set `key` on GPU g (copied from the home GPU, with their cached SVD triplets, on first use)
static int set_on_gpu(resnmtf_pool* pool, int key, int g, std::vector<resnmtf_data*>* out) {
  resnmtf_pool::DataSet* sp = nullptr;
  {
    std::lock_guard<std::mutex> lk(pool->mu);  // the map only; a set is dropped between batches, never under one
    auto it = pool->sets.find(key);
    RN_CHECK(it != pool->sets.end(), RESNMTF_E_INVALID, "resnmtf_batch_run: unknown data set key");
    sp = &it->second;
  }
  resnmtf_pool::DataSet& s = *sp;
  std::lock_guard<std::mutex> lk(s.gpu_mu[g]);
  if (s.on_gpu[g].empty()) {
    std::vector<resnmtf_data*> mine;
    for (resnmtf_data* src : s.on_gpu[s.home]) {
      resnmtf_data* c = nullptr;
      int rc = resnmtf_data_copy(src, pool->ctx[g], &c);
      if (rc) {
        for (resnmtf_data* d : mine) resnmtf_data_destroy(d);
        return rc;
      }
      mine.push_back(c);
    }
    s.on_gpu[g] = mine;
  }
  *out = s.on_gpu[g];
  return RESNMTF_OK;
}

extern "C" int resnmtf_pool_get(resnmtf_pool* pool, int key, int gpu, int view, resnmtf_data** out) {
  RN_CHECK(pool && out, RESNMTF_E_INVALID, "resnmtf_pool_get: NULL argument");
  RN_CHECK(gpu >= 0 && gpu < (int)pool->ctx.size(), RESNMTF_E_INVALID, "resnmtf_pool_get: bad GPU index");
  std::vector<resnmtf_data*> vs;
  int rc = set_on_gpu(pool, key, gpu, &vs);
  if (rc) return rc;
  RN_CHECK(view >= 0 && view < (int)vs.size(), RESNMTF_E_INVALID, "resnmtf_pool_get: bad view index");
  *out = vs[view];  // borrowed: owned by the pool
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// one unit on one GPU
// ------------------------------------------------------------------------------------------------------------------

// init_mats_inner() (R/update_steps.r:93-105) from the view's top singular triplets and the caller's noise draw:
// F = |U_k|, G = |V_k|, S = |diag(d_k)| + noise, columns of S scaled by colSums(F) colSums(G) (quirk Q7), F and G
// scaled to unit column sums, lambda / mu = their column sums.
static int svd_init(resnmtf_data* d, int k, const double* noise, std::vector<double>& F, std::vector<double>& S,
                    std::vector<double>& G, std::vector<double>& lam, std::vector<double>& mu) {
  int rc = rn_data_svd(d);
  if (rc) return rc;
  RN_CHECK(k <= d->svd_kc, RESNMTF_E_INVALID, "SVD initialisation: k exceeds min(n, p, 16)");
  const int64_t n = d->n, p = d->p;
  F.assign(d->svd_u.begin(), d->svd_u.begin() + (size_t)n * k);
  G.assign(d->svd_v.begin(), d->svd_v.begin() + (size_t)p * k);
  S.assign((size_t)k * k, 0.0);
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i)
      S[(size_t)i + (size_t)j * k] = (i == j ? std::fabs(d->svd_d[i]) : 0.0) + (noise ? noise[(size_t)i + (size_t)j * k] : 0.0);
  lam.assign((size_t)k, 0.0);
  mu.assign((size_t)k, 0.0);
  for (int j = 0; j < k; ++j) {
    double csf = 0.0, csg = 0.0;
    for (int64_t i = 0; i < n; ++i) csf += F[(size_t)i + (size_t)j * n];
    for (int64_t i = 0; i < p; ++i) csg += G[(size_t)i + (size_t)j * p];
    for (int i = 0; i < k; ++i) S[(size_t)i + (size_t)j * k] *= csf * csg;
    double sf = 0.0, sg = 0.0;
    for (int64_t i = 0; i < n; ++i) sf += (F[(size_t)i + (size_t)j * n] /= csf);
    for (int64_t i = 0; i < p; ++i) sg += (G[(size_t)i + (size_t)j * p] /= csg);
    lam[j] = sf;
    mu[j] = sg;
  }
  return RESNMTF_OK;
}

// The SVD triplets of resident data that several units of a batch initialise from are computed ONCE, on the set's home
// GPU, by that GPU's worker before it takes its first unit -- while the other GPUs already run the units that derive
// their own data (shuffled refits, sub-samples).  A unit that needs the triplets waits for them here and adopts them when
// its GPU's copy of the view was made before they existed.
struct HomeSvd {
  std::mutex mu;
  std::condition_variable cv;
  bool done = true;  // false while the home worker still owes the computation
  int rc = RESNMTF_OK;
  std::string message;
  int key = -1;
  std::vector<resnmtf_data*> handles;  // the home GPU's handles of set `key`
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
    if (rc) return rn_fail(rc, message);
    return RESNMTF_OK;
  }
};

static int run_unit(resnmtf_pool* pool, int g, resnmtf_unit* u, HomeSvd* hs) {
  resnmtf_ctx* ctx = pool->ctx[g];
  std::vector<resnmtf_data*> base;
  int rc = set_on_gpu(pool, u->data_key, g, &base);
  if (rc) return rc;
  if (hs && u->derive == 0 && !(u->init_f && u->init_s && u->init_g) && u->data_key == hs->key) {
    if ((rc = hs->wait())) return rc;
    for (size_t v = 0; v < base.size() && v < hs->handles.size(); ++v) rn_data_svd_adopt(base[v], hs->handles[v]);
  }
  const int V = (int)base.size();
  RN_CHECK(u->k != nullptr, RESNMTF_E_INVALID, "resnmtf_batch_run: unit without k");
  std::vector<resnmtf_data*> views((size_t)V, nullptr), owned;
  auto cleanup = [&]() {
    for (resnmtf_data* d : owned) resnmtf_data_destroy(d);
  };
  for (int v = 0; v < V; ++v) {
    resnmtf_data* d = base[v];
    if (u->derive & RESNMTF_DERIVE_SUBSAMPLE) {
      if (!(u->rows && u->cols && u->n_rows && u->n_cols)) {
        cleanup();
        return rn_fail(RESNMTF_E_INVALID, "resnmtf_batch_run: sub-sample unit without row / column indices");
      }
      resnmtf_data* s = nullptr;
      if ((rc = resnmtf_data_subsample(d, u->rows[v], u->n_rows[v], u->cols[v], u->n_cols[v], &s))) {
        cleanup();
        return rc;
      }
      owned.push_back(s);
      d = s;
    }
    if (u->derive & RESNMTF_DERIVE_SHUFFLE) {
      resnmtf_data* s = nullptr;
      if ((rc = resnmtf_data_shuffle(d, u->seed + 0x9e3779b97f4a7c15ULL * (uint64_t)(v + 1), u->renormalise, nullptr, &s))) {
        cleanup();
        return rc;
      }
      owned.push_back(s);
      d = s;
    }
    views[v] = d;
  }
  std::vector<int64_t> n((size_t)V), p((size_t)V);
  for (int v = 0; v < V; ++v) {
    n[v] = views[v]->n;
    p[v] = views[v]->p;
  }
  resnmtf_fit* fit = nullptr;
  if ((rc = resnmtf_fit_create(ctx, V, n.data(), p.data(), u->k, &fit))) {
    cleanup();
    return rc;
  }
  auto fail = [&](int code) {
    resnmtf_fit_destroy(fit);
    cleanup();
    return code;
  };
  if ((rc = resnmtf_fit_set_options(fit, u->err_mode, u->impl))) return fail(rc);
  std::vector<double> F, S, G, lam, mu;
  for (int v = 0; v < V; ++v) {
    if ((rc = resnmtf_fit_attach_data(fit, v, views[v]))) return fail(rc);
    if (u->init_f && u->init_s && u->init_g) {
      if ((rc = resnmtf_fit_set_factors(fit, v, u->init_f[v], u->init_s[v], u->init_g[v], nullptr, nullptr))) return fail(rc);
    } else {
      if ((rc = svd_init(views[v], u->k[v], u->noise ? u->noise[v] : nullptr, F, S, G, lam, mu))) return fail(rc);
      if ((rc = resnmtf_fit_set_factors(fit, v, F.data(), S.data(), G.data(), lam.data(), mu.data()))) return fail(rc);
    }
  }
  if (u->phi || u->xi || u->psi)
    if ((rc = resnmtf_fit_set_restrictions(fit, u->phi, u->xi, u->psi))) return fail(rc);
  for (int i = 0; i < u->n_maps; ++i) {
    const resnmtf_map& m = u->maps[i];
    if ((rc = resnmtf_fit_set_shared_map(fit, m.kind, m.v, m.w, m.idx_v, m.idx_w, m.len))) return fail(rc);
  }
  int64_t done = 0;
  rc = resnmtf_fit_run(fit, u->n_iters, u->tol, u->max_iters, &done);
  u->iters = done;
  if (rc) return fail(rc);
  int64_t cnt = 0;
  resnmtf_fit_get_errors(fit, u->errors, u->errors ? u->errors_cap : 0, &cnt);
  u->n_errors = cnt;
  for (int v = 0; v < V; ++v)  // lambda / mu as the loop left them (R/main.r:137-138)
    if ((u->out_lambda && u->out_lambda[v]) || (u->out_mu && u->out_mu[v]))
      if ((rc = resnmtf_fit_get_factors(fit, v, nullptr, nullptr, nullptr, u->out_lambda ? u->out_lambda[v] : nullptr,
                                        u->out_mu ? u->out_mu[v] : nullptr)))
        return fail(rc);
  if ((rc = resnmtf_fit_normalise(fit))) return fail(rc);  // normalisation_check (R/main.r:110)
  for (int v = 0; v < V; ++v)
    if ((rc = resnmtf_fit_get_factors(fit, v, u->out_f ? u->out_f[v] : nullptr, u->out_s ? u->out_s[v] : nullptr,
                                      u->out_g ? u->out_g[v] : nullptr, nullptr, nullptr)))
      return fail(rc);
  resnmtf_fit_destroy(fit);
  cleanup();
  return RESNMTF_OK;
}

static double unit_cost(resnmtf_pool* pool, const resnmtf_unit& u) {
  std::lock_guard<std::mutex> lk(pool->mu);
  auto it = pool->sets.find(u.data_key);
  if (it == pool->sets.end()) return 0.0;
  const resnmtf_pool::DataSet& s = it->second;
  double c = 0.0;
  for (int v = 0; v < s.n_views; ++v) {
    const resnmtf_data* d = s.on_gpu[s.home][v];
    double n = (double)d->n, p = (double)d->p;
    if ((u.derive & RESNMTF_DERIVE_SUBSAMPLE) && u.n_rows && u.n_cols) {
      n = (double)u.n_rows[v];
      p = (double)u.n_cols[v];
    }
    c += n * p * (u.k ? u.k[v] : 1);
  }
  // a derived view pays its own Gram matrix and top-k solve; the fits of the resident data share one
  return c * ((u.derive != 0 && !u.init_f) ? 1.5 : 1.0);
}

extern "C" int resnmtf_batch_run(resnmtf_pool* pool, resnmtf_unit* units, int n_units) {
  RN_CHECK(pool && (units || n_units == 0) && n_units >= 0, RESNMTF_E_INVALID, "resnmtf_batch_run: bad argument");
  if (n_units == 0) return RESNMTF_OK;
  // SVD triplets of resident data that several units initialise from are computed ONCE, on the home GPU (the copies
  // of the views carry them, or adopt them later).  With several workers and one such data set, the home GPU's worker
  // computes them as its first job while the other GPUs start on the units that derive their own data (HomeSvd);
  // otherwise they are computed here, before the views fan out.
  const int n_workers = std::min<int>((int)pool->ctx.size(), n_units);
  HomeSvd hs;
  int hs_gpu = -1;
  {
    bool several_keys = false;
    for (int i = 0; i < n_units; ++i) {
      resnmtf_unit& u = units[i];
      if (u.derive != 0 || (u.init_f && u.init_s && u.init_g)) continue;
      if (hs.key >= 0 && hs.key != u.data_key) several_keys = true;
      if (hs.key >= 0) continue;
      std::lock_guard<std::mutex> lk(pool->mu);
      auto it = pool->sets.find(u.data_key);
      RN_CHECK(it != pool->sets.end(), RESNMTF_E_INVALID, "resnmtf_batch_run: unknown data set key");
      hs.key = u.data_key;
      hs_gpu = it->second.home;
      hs.handles = it->second.on_gpu[hs_gpu];
    }
    const bool overlap = hs.key >= 0 && !several_keys && n_workers > 1 && hs_gpu < n_workers &&
                         rn_env_int("RESNMTF_POOL_SVD_OVERLAP", 1) != 0;
    if (hs.key >= 0 && !overlap) {
      for (int i = 0; i < n_units; ++i) {
        resnmtf_unit& u = units[i];
        if (u.derive != 0 || (u.init_f && u.init_s && u.init_g)) continue;
        std::vector<resnmtf_data*> home;
        {
          std::lock_guard<std::mutex> lk(pool->mu);
          auto it = pool->sets.find(u.data_key);
          RN_CHECK(it != pool->sets.end(), RESNMTF_E_INVALID, "resnmtf_batch_run: unknown data set key");
          home = it->second.on_gpu[it->second.home];
        }
        for (resnmtf_data* d : home) {
          int rc = rn_data_svd(d);
          if (rc) return rc;
        }
      }
      hs.key = -1;  // nothing owed: units find the triplets in their copies
      hs_gpu = -1;
    } else if (overlap) {
      hs.done = false;
    } else {
      hs_gpu = -1;
    }
  }
  std::vector<int> order((size_t)n_units);
  std::vector<double> cost((size_t)n_units);
  for (int i = 0; i < n_units; ++i) {
    order[i] = i;
    cost[i] = unit_cost(pool, units[i]);
    units[i].status = RESNMTF_E_STATE;
    units[i].gpu = -1;
    units[i].message[0] = 0;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  // Two queues, each longest first.  The fits of the resident data (nothing derived) come before the derived units:
  // their duration is the uncertain one -- structured data needs several times the sweeps of its shuffled twin (measured
  // on the C2 view: 90 ms for the k = 3 fit against 45 ms for a k = 3 shuffled refit, although the cost model ranks it
  // last) -- and a long unit started last is the tail of the batch.  While the home GPU's worker still owes the SVD
  // triplets they wait for, the other workers take derived units instead of waiting.
  std::vector<int> q_main, q_rest;
  for (int i : order) (units[i].derive == 0 ? q_main : q_rest).push_back(i);
  if (rn_env_int("RESNMTF_POOL_MAIN_FIRST", 1) == 0) {
    q_rest = order;
    q_main.clear();
  }
  std::mutex q_mu;
  size_t mi = 0, ri = 0;
  std::atomic<bool> svd_ready{hs_gpu < 0};
  auto take = [&]() {
    std::lock_guard<std::mutex> lk(q_mu);
    if (mi < q_main.size() && svd_ready.load()) return q_main[mi++];
    if (ri < q_rest.size()) return q_rest[ri++];
    if (mi < q_main.size()) return q_main[mi++];  // only fits left: wait for the triplets inside the unit
    return -1;
  };
  std::atomic<int> worst{RESNMTF_OK};
  // RESNMTF_POOL_TRACE=1: one line per unit on stderr (GPU, start and end in ms since the batch began, what it was)
  const bool trace = rn_env_int("RESNMTF_POOL_TRACE", 0) != 0;
  const auto batch_t0 = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(t - batch_t0).count();
  };
  auto worker = [&](int g) {
    if (g == hs_gpu) {  // first job of the home GPU's worker: the shared SVD triplets
      int rc = RESNMTF_OK;
      for (resnmtf_data* d : hs.handles)
        if ((rc = rn_data_svd(d))) break;
      {
        std::lock_guard<std::mutex> lk(hs.mu);
        hs.rc = rc;
        if (rc) hs.message = resnmtf_last_error();
        hs.done = true;
      }
      svd_ready.store(true);
      hs.cv.notify_all();
      if (trace)
        std::fprintf(stderr, "  [resnmtf pool] gpu %d  %8.1f .. %8.1f ms  SVD triplets of the resident view\n", g, 0.0,
                     ms_since(std::chrono::steady_clock::now()));
    }
    for (;;) {
      const int ui = take();
      if (ui < 0) return;
      resnmtf_unit* u = &units[ui];
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = run_unit(pool, g, u, hs_gpu >= 0 ? &hs : nullptr);
      u->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      u->status = rc;
      u->gpu = g;
      if (trace)
        std::fprintf(stderr, "  [resnmtf pool] gpu %d  %8.1f .. %8.1f ms  unit %3d  k=%d%s%s  %lld sweeps\n", g, ms_since(t0),
                     ms_since(std::chrono::steady_clock::now()), ui, u->k ? u->k[0] : 0,
                     (u->derive & RESNMTF_DERIVE_SUBSAMPLE) ? " sub-sample" : "",
                     (u->derive & RESNMTF_DERIVE_SHUFFLE) ? " shuffled" : "", (long long)u->iters);
      if (rc) {
        std::snprintf(u->message, sizeof(u->message), "%s", resnmtf_last_error());
        int expected = RESNMTF_OK;
        worst.compare_exchange_strong(expected, rc);
      }
    }
  };
  if (n_workers <= 1) {
    worker(0);
  } else {
    std::vector<std::thread> threads;
    for (int g = 0; g < n_workers; ++g) threads.emplace_back(worker, g);
    for (std::thread& t : threads) t.join();
  }
  const int rc = worst.load();
  if (rc)
    for (int i = 0; i < n_units; ++i)
      if (units[i].status == rc) return rn_fail(rc, std::string("resnmtf_batch_run: unit ") + std::to_string(i) + ": " + units[i].message);
  return RESNMTF_OK;
}
