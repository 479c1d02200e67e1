// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of Moby's LCP solvers (src/LCP.cpp, include/Moby/LCP.h:21-27).
// PARITY UNPINNED: the reference ships no unit test, known-answer test or golden
// (M,q,z) vector for these solvers, and it cannot be built
// here (Ravelin/Boost/qhull/libxml2 absent).  The restatement is validated by
// brute-force enumeration, by the solution conditions the reference's own
// wrappers check (LCP.cpp:381-390), and indirectly by regress/*.dat.
#pragma once
#include <vector>
#include "oracle_linalg.h"

namespace oracle {

enum TieRule {
  TIE_LOWEST_INDEX = 0,  // documented deterministic rule (SURVEY.md 8a' H2)
  TIE_GLIBC_RAND = 1     // literal reference: unseeded rand() (LCP.cpp:199-209, :611-621)
};

enum LcpStatus {
  LCP_OK = 0, LCP_TRIVIAL = 1, LCP_RAY = 2, LCP_MAXITER = 3, LCP_SINGULAR = 4, LCP_EMPTY_RATIO = 5,
  LCP_UNVERIFIED = 6, LCP_REGULARIZED = 16
};

struct LCP {
  TieRule tie = TIE_LOWEST_INDEX;
  unsigned pivots = 0;        // LCP::pivots
  int status = LCP_OK;        // detail for the last solve
  bool keep_log = false;
  std::vector<int> log;       // Lemke: leaving variable per pivot (first entry: initial leaving);
                              // lcp_fast: moved index, | 0x40000000 when moved nonbasic -> basic
  // statistics
  unsigned long long n_fast_calls = 0, n_lemke_calls = 0, n_pivots_total = 0;

  // LCP.cpp:41-196.  z: warm start iff z.size()==n (LCP.cpp:65); result on success.
  bool lcp_fast(int n, const double* M, const double* q, Vec& z, double zero_tol = -1.0);
  // LCP.cpp:545-1003.
  bool lcp_lemke(int n, const double* M, const double* q, Vec& z, double piv_tol = -1.0, double zero_tol = -1.0);
  // LCP.cpp:212-350.
  bool lcp_fast_regularized(int n, const double* M, const double* q, Vec& z, int min_exp = -20, unsigned step_exp = 4,
                            int max_exp = 20, double piv_tol = -1.0, double zero_tol = -1.0);
  // LCP.cpp:353-487.
  bool lcp_lemke_regularized(int n, const double* M, const double* q, Vec& z, int min_exp = -20, unsigned step_exp = 1,
                             int max_exp = 1, double piv_tol = -1.0, double zero_tol = -1.0);
  // LCP.cpp:199-209.
  unsigned rand_min(const Vec& v, double zero_tol);
};

}  // namespace oracle
