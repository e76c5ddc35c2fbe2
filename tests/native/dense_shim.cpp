// Test shim: the host-side dense helpers of the library (resnmtf_b200/csrc/rn_dense.h) behind a C ABI, so that the
// CPU test suite can check them against LAPACK without a GPU (compiled by tests/test_host_dense.py with g++).
#include "../../resnmtf_b200/csrc/rn_dense.h"

extern "C" int t_sym_eig(int n, const double* a, double* w, double* v) { return rn_sym_eig(n, a, w, v) ? 0 : 1; }
extern "C" int t_cholesky_upper(int n, const double* g, double* r) { return rn_cholesky_upper(n, g, r) ? 0 : 1; }
extern "C" void t_upper_inverse(int n, const double* r, double* ri) { rn_upper_inverse(n, r, ri); }
