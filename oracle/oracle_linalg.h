// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Minimal dense linear algebra standing in for the Ravelin calls on Moby's hot
// path.  Ravelin (github.com/PositronicsLab/Ravelin) is an un-vendored,
// un-pinned dependency of the reference (CMakeLists.txt:62,
// CMakeModules/FindRavelin.cmake:10-23); its source is not under
// /root/reference, so these helpers restate its *published* semantics:
//   MatrixNd::norm_inf()      -> max |a_ij|               (call sites LCP.cpp:58,571,761)
//   LinAlgd::solve_fast(A,b)  -> LAPACK dgesv: LU with partial pivoting (first max),
//                                SingularException on an exactly-zero pivot
//                                (call sites LCP.cpp:120,670,838)
//   LinAlgd::factor_chol      -> dpotrf, false when not positive definite
//                                (ImpactConstraintHandler.cpp:366,1733)
//   LinAlgd::inverse_SPD      -> Cholesky-based inverse (ImpactConstraintHandler.cpp:1607)
// PARITY UNPINNED for these routines: no golden vectors exist for them in the
// reference tree, and the bit pattern of LAPACK results depends on the BLAS.
//
// Arithmetic order is fixed and documented (explicit std::fma, compiled with
// -ffp-contract=off) so that an implementation that follows the same order can
// match bit for bit.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace oracle {

// Column-major dense matrix, like Ravelin::MatrixNd.
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  void resize(int r_, int c_) { r = r_; c = c_; a.assign((size_t)r_ * c_, 0.0); }
  double& operator()(int i, int j) { return a[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * r + i]; }
};
typedef std::vector<double> Vec;

// MatrixNd::norm_inf(): largest absolute entry.
inline double norm_inf(const double* M, int rows, int cols) {
  double nrm = 0.0;
  for (size_t i = 0; i < (size_t)rows * cols; i++) nrm = std::fmax(nrm, std::fabs(M[i]));
  return nrm;
}

// LinAlgd::solve_fast: solves A x = b in place (A destroyed, b <- x).
// Right-looking LU with partial pivoting: pivot = first row of maximum |a_ij|,
// multipliers formed with the reciprocal of the pivot (as LAPACK dgetf2 does),
// trailing update with fma; forward substitution fused into the elimination;
// column-oriented back substitution, columns descending, fma.
// Returns false on an exactly zero pivot (SingularException).
inline bool solve_fast(double* A, int n, double* b, int* piv = nullptr) {   // piv (optional): pivot row chosen at each column, as LAPACK's ipiv (0-based)
  for (int j = 0; j < n; j++) {
    int p = j;
    double best = std::fabs(A[(size_t)j * n + j]);
    for (int i = j + 1; i < n; i++) {
      double v = std::fabs(A[(size_t)j * n + i]);
      if (v > best) { best = v; p = i; }
    }
    if (A[(size_t)j * n + p] == 0.0) return false;
    if (piv) piv[j] = p;
    if (p != j) {
      for (int c = 0; c < n; c++) { double t = A[(size_t)c * n + j]; A[(size_t)c * n + j] = A[(size_t)c * n + p]; A[(size_t)c * n + p] = t; }
      double t = b[j]; b[j] = b[p]; b[p] = t;
    }
    const double rinv = 1.0 / A[(size_t)j * n + j];
    for (int i = j + 1; i < n; i++) {
      const double l = A[(size_t)j * n + i] * rinv;
      A[(size_t)j * n + i] = l;
      for (int c = j + 1; c < n; c++) A[(size_t)c * n + i] = std::fma(-l, A[(size_t)c * n + j], A[(size_t)c * n + i]);
      b[i] = std::fma(-l, b[j], b[i]);
    }
  }
  for (int c = n - 1; c >= 0; c--) {
    b[c] = b[c] / A[(size_t)c * n + c];
    for (int i = 0; i < c; i++) b[i] = std::fma(-A[(size_t)c * n + i], b[c], b[i]);
  }
  return true;
}

// LinAlgd::factor_chol: lower Cholesky in place (column-major, lower triangle), false if not PD.
inline bool factor_chol(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d = std::fma(-A[(size_t)k * n + j], A[(size_t)k * n + j], d);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)j * n + i];
      for (int k = 0; k < j; k++) s = std::fma(-A[(size_t)k * n + i], A[(size_t)k * n + j], s);
      A[(size_t)j * n + i] = s / d;
    }
  }
  return true;
}

// solve_chol_fast: solves (L L^T) x = b given the factor from factor_chol.
inline void solve_chol(const double* L, int n, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s = std::fma(-L[(size_t)k * n + i], b[k], s);
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s = std::fma(-L[(size_t)i * n + k], b[k], s);
    b[i] = s / L[(size_t)i * n + i];
  }
}

// LinAlgd::inverse_SPD: A <- A^-1 (full symmetric matrix), false if not PD.
inline bool inverse_SPD(double* A, int n) {
  std::vector<double> L(A, A + (size_t)n * n);
  if (!factor_chol(L.data(), n)) return false;
  std::vector<double> e(n);
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) e[i] = (i == j) ? 1.0 : 0.0;
    solve_chol(L.data(), n, e.data());
    for (int i = 0; i < n; i++) A[(size_t)j * n + i] = e[i];
  }
  return true;
}

}  // namespace oracle
