// Device-side descriptors shared by the kernels and the host C-ABI (resnmtf_capi.cu).
// Data layout in HBM (see DESIGN.md "Data layout"):
//   X   [row_tiles][pp][64]  "panel" layout: 64-row panels, column-major inside a panel, so that both the
//                  X.G pass (a panel's columns left to right) and the X'.F pass (a column group down the
//                  panels) read long contiguous runs.  Rows are padded to a multiple of 64 (ldx), columns
//                  to a multiple of 32 (pp); all padding is zero so the streaming kernels need no bounds
//                  checks.  Element (r, j) lives at ((r/64)*pp + j)*64 + r%64.
//   F   [row_tiles][kp][64]  the same 64-row panels (and the same piece swizzle) as X, kp = 8 (k<=8) or 16;
//                  padding rows/columns zero.  The 64 rows of F a row step needs are one kp*512 B run.
//   G,T [pp][kp]   row-major (one 64 B / 128 B row per data column) so that a column's k values are one
//                  uniform / vector load in the X.G pass and one gather in the psi coupling.
//   S, F'F, G'G, A  k x k column-major compact, lambda/mu/colsums length k.
#pragma once
// error mode AUTO: per-view error below which the algebraic form of calculate_error() (R/utils.r:157-166) hands over to the
// direct residual pass (rn_view_finish, rn_kernels.cuh)
#define RN_AUTO_DIRECT_BELOW 1.0e-4
#include <stdint.h>

#define RN_MAXK 16
#define RN_ROW_TILE 64   // rows per F-step tile and per G-stream row step
#define RN_COL_GROUP 64  // data columns per G-stream CTA (tensor-core path)
#define RN_COL_GROUP_DFMA 32
#define RN_GEPI_THREADS(k) ((k) <= 8 ? 256 : 128)

#define RN_MODE_NULL 0  // pair never set: R's `indices = NULL` (nothing overwritten)
#define RN_MODE_NA 1    // pair shares no names: skipped
#define RN_MODE_MAP 2   // int32 gather map, -1 = name not shared

struct RnCtrl {
  int32_t done;       // 0 running, 1 converged, 2 mean error is NaN (convergence mode only),
                      // 3 paused: AUTO error mode wants the direct residual (host takes over)
  int32_t conv_mode;  // 1: apply the stop rule of R/main.r:55
  int64_t iters;      // sweeps since set_factors
  int64_t hist_count; // entries in hist[] since the host last drained it
  double prev_err;    // err_temp of R/main.r:54,80
  double tol;
  double last_diff;
  int64_t direct_passes;
  int32_t want_direct;  // sticky: some view's algebraic error fell below RN_AUTO_DIRECT_BELOW (AUTO mode)
  int32_t pad_;
};

struct RnView {
  int64_t n, p, ldx, pp;
  int64_t n_glob;  // rows of the whole view (== n unless the view is row-sharded over ranks)
  int32_t k, kp;
  double* X;
  double* F;
  double* G;
  double* T;
  double* S;    // k*k
  double* FtF;  // k*k
  double* GtG;  // k*k
  double* A;    // k*k   A = (F'X)G = T'G
  double* lam;  // k
  double* mu;   // k
  double* csF;  // k
  double* csG;  // k
  double* scal; // [0] ||X||^2  [1] err  [2] err algebraic  [3] err direct  [4] global F'F etc scratch
  double* Ppart;   // stream-K: [f_ctas][2][64][kp]; CUDA-core path: [cs][row_tiles][64][kp]
  double* Tpart;   // stream-K: [g_ctas][2][64][kp]; CUDA-core path: [rs][pp][kp]
  double* FFpart;  // [g_ctas | nff][k*k+k]   partial F'F | colSums(F)
  double* GGpart;  // [col_groups | gepi_ctas][2*k*k+k]  partial G'G | A | colSums(G)
  double* Rpart;   // [resid ctas]    partial residual sums
  int32_t* tile_ticket;   // [row tiles]
  int32_t* group_ticket;  // [col groups]
  int32_t* misc_ticket;   // [0] G-epilogue  [1] residual  [2] F'F partials published  [3] fused: T partials published
  int32_t* flags;         // [0] need_direct
  int32_t row_tiles, cs;       // F-step grid (CUDA-core path)
  int32_t col_groups, rs;      // G-stream grid (CUDA-core path); col_groups is also the stream-K group count
  int32_t f_ctas, g_ctas;      // persistent grids of the stream-K kernels
  int32_t nff;                 // number of FFpart rows
  int32_t gepi_ctas;           // G-epilogue grid
  int32_t resid_cs;            // residual grid y
  int32_t sharded;             // 1: rows are sharded over ranks (epilogue scalars come from all-reduce)
  // one-pass fused path (rn_fused.cuh): second copy of X in the 8-row-group layout, pp8 = fu_csize * 1008 (kind 1) or pp (kind 2) columns;
  // Tpart is then [fu_clusters][pp8][kp], FFpart [fu_clusters][k*k+k], GGpart [ceil(pp/64)][2*k*k+k]
  double* X8;
  int64_t pp8;
  int32_t fu_csize, fu_clusters;  // CTAs per cluster (1..8; 0: view not on the fused path), clusters in the grid
  int32_t fu_kind, fu_pad_;       // 1: rn_fused_step (1008 columns per CTA), 2: rn_fused2_step (<= 672, rn_fused2.cuh)
  long long* fu_trace;            // optional [row groups of cluster 0][8] per-group stamps of CTA 0 (same switch)
  long long* fu_timeline;         // optional [grid][12] %globaltimer stamps of the last launch (RESNMTF_FU_TIMELINE=1)
  long long* fu_waits;            // optional [grid][9 consumer warps][2] cycles waited for X / for F_new (same switch)
};

struct RnFit {
  int32_t n_views;
  int32_t err_mode;
  const RnView* views;   // device array [n_views]
  RnCtrl* ctrl;
  const double* phi;     // [V*V] column-major, symmetrised
  const double* xi;
  const double* psi;
  const int32_t* const* rowmap;  // [V*V]: rowmap[w + v*V] = map of view v's rows into view w
  const int32_t* const* colmap;
  const int8_t* rowmode;         // [V*V]
  const int8_t* colmode;
  double* hist;          // mean error per sweep (drained by the host)
  int64_t hist_cap;
  double psi_total, xi_total;    // whole-matrix sums (branch tests of R/update_steps.r:190,226)
  int64_t n_total_rows_scale;    // unused (kept for ABI stability of the struct)
};
