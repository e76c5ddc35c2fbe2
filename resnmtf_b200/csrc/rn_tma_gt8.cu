// Instantiations of the two-pass TMA kernels for k = 9..16 (two 8-wide MMA tiles in the factor dimension,
// rn_kernels.cuh), in their own translation unit so that they compile beside resnmtf_capi.cu instead of inside it.
// The reference's k-extension loop (R/main.r:306-320) reaches k = 9 whenever the selected k is k_max = 8.
// rn_kernels.cuh also defines the library's non-template kernels, which resnmtf_capi.cu owns: templates only here
#define RN_KERNEL_TEMPLATES_ONLY
#include "rn_kernels.cuh"

typedef void (*FStepSkFn)(const RnView, const RnFit, const int);
typedef void (*GStepSkFn)(const RnView, const RnFit, const int, const int);

#define RN_K_CASES_GT8(X) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

FStepSkFn rn_f_step_tma_gt8(int K) {
  switch (K) {
#define X(KC) case KC: return rn_f_step_tma<KC>;
    RN_K_CASES_GT8(X)
#undef X
  }
  return nullptr;
}

GStepSkFn rn_g_step_tma_gt8(int K) {
  switch (K) {
#define X(KC) case KC: return rn_g_step_tma<KC>;
    RN_K_CASES_GT8(X)
#undef X
  }
  return nullptr;
}
