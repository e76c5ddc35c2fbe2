// Small dense symmetric problems on the host (order <= a few hundred): the Rayleigh-Ritz step and the Cholesky factor of
// the block orthonormalisation in the subspace iteration of rn_native.cu.  Plain C++; column-major.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

// Eigenvalues (ascending in w) and eigenvectors (columns of V, n x n column-major) of the symmetric matrix A (n x n,
// column-major, both triangles given): Householder tridiagonalisation followed by the implicit QL iteration (the
// classical tred2 / tql2 pair).  Returns false when QL does not converge.
inline bool rn_sym_eig(int n, const double* A, double* w, double* V) {
  std::vector<double> e((size_t)n, 0.0);
  auto v = [&](int i, int j) -> double& { return V[(size_t)i + (size_t)j * n]; };
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) v(i, j) = A[(size_t)i + (size_t)j * n];
  double* d = w;
  for (int j = 0; j < n; ++j) d[j] = v(n - 1, j);
  // Householder reduction to tridiagonal form
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) {
        d[j] = v(i - 1, j);
        v(i, j) = 0.0;
        v(j, i) = 0.0;
      }
    } else {
      for (int k = 0; k < i; ++k) {
        d[k] /= scale;
        h += d[k] * d[k];
      }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        v(j, i) = f;
        g = e[j] + v(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) {
          g += v(k, j) * d[k];
          e[k] += v(k, j) * f;
        }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) {
        e[j] /= h;
        f += e[j] * d[j];
      }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; ++k) v(k, j) -= (f * e[k] + g * d[k]);
        d[j] = v(i - 1, j);
        v(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  // accumulate the transformations
  for (int i = 0; i < n - 1; ++i) {
    v(n - 1, i) = v(i, i);
    v(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = v(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += v(k, i + 1) * v(k, j);
        for (int k = 0; k <= i; ++k) v(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) v(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) {
    d[j] = v(n - 1, j);
    v(n - 1, j) = 0.0;
  }
  v(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
  // implicit QL
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::pow(2.0, -52.0);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) {
      if (std::fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 200) return false;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = v(k, i + 1);
            v(k, i + 1) = s * v(k, i) + c * h;
            v(k, i) = c * v(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1);
    }
    d[l] += f;
    e[l] = 0.0;
  }
  // ascending order
  for (int i = 0; i < n - 1; ++i) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; ++j)
      if (d[j] < p) {
        k = j;
        p = d[j];
      }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      for (int j = 0; j < n; ++j) std::swap(v(j, i), v(j, k));
    }
  }
  return true;
}

// Upper-triangular R with R'R = G (n x n, column-major, symmetric positive definite); false on a non-positive pivot.
inline bool rn_cholesky_upper(int n, const double* G, double* R) {
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i <= j; ++i) {
      double s = G[(size_t)i + (size_t)j * n];
      for (int k = 0; k < i; ++k) s -= R[(size_t)k + (size_t)i * n] * R[(size_t)k + (size_t)j * n];
      if (i < j) {
        R[(size_t)i + (size_t)j * n] = s / R[(size_t)i + (size_t)i * n];
      } else {
        if (!(s > 0.0)) return false;
        R[(size_t)j + (size_t)j * n] = std::sqrt(s);
      }
    }
    for (int i = j + 1; i < n; ++i) R[(size_t)i + (size_t)j * n] = 0.0;
  }
  return true;
}

// Inverse of an upper-triangular matrix (column-major), upper-triangular again.
inline void rn_upper_inverse(int n, const double* R, double* Ri) {
  for (size_t i = 0; i < (size_t)n * n; ++i) Ri[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    Ri[(size_t)j + (size_t)j * n] = 1.0 / R[(size_t)j + (size_t)j * n];
    for (int i = j - 1; i >= 0; --i) {
      double s = 0.0;
      for (int k = i + 1; k <= j; ++k) s += R[(size_t)i + (size_t)k * n] * Ri[(size_t)k + (size_t)j * n];
      Ri[(size_t)i + (size_t)j * n] = -s / R[(size_t)i + (size_t)i * n];
    }
  }
}
