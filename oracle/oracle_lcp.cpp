// ORACLE -- TEST INFRASTRUCTURE ONLY.  See oracle_lcp.h.  PARITY UNPINNED (no reference KATs exist).
// Restates src/LCP.cpp of the reference statement by statement; line numbers cite that file.
#include "oracle_lcp.h"
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <limits>

namespace oracle {

static const double EPS = std::numeric_limits<double>::epsilon();
static const double NEAR_ZERO = std::sqrt(std::numeric_limits<double>::epsilon());  // Constants.h:21

// LCP.cpp:199-209.  Candidates: the first minimum, then every other i with v[i] < v[min] + tol in index order.
// Reference picks candidates[rand() % size]; the documented deterministic rule picks the lowest index.
unsigned LCP::rand_min(const Vec& v, double zero_tol) {
  unsigned minv = (unsigned)(std::min_element(v.begin(), v.end()) - v.begin());
  if (tie == TIE_GLIBC_RAND) {
    std::vector<unsigned> minima;
    minima.push_back(minv);
    for (unsigned i = 0; i < v.size(); i++)
      if (i != minv && v[i] < v[minv] + zero_tol) minima.push_back(i);
    return minima[rand() % minima.size()];
  }
  for (unsigned i = 0; i < minv; i++)
    if (v[i] < v[minv] + zero_tol) return i;
  return minv;
}

static void insertion_sort(std::vector<unsigned>& v) { std::sort(v.begin(), v.end()); }  // include/Moby/insertion_sort: plain ascending sort

// LCP.cpp:41-196
bool LCP::lcp_fast(int n, const double* M, const double* q, Vec& z, double zero_tol) {
  const unsigned N = (unsigned)n;
  const unsigned UINF = std::numeric_limits<unsigned>::max();
  n_fast_calls++;
  pivots = 0;
  status = LCP_OK;
  if (N == 0) { z.clear(); return true; }                                   // :49-54
  if (zero_tol < 0.0) zero_tol = n * norm_inf(M, n, n) * EPS;               // :57-58
  std::vector<unsigned> nonbas, bas;
  if (z.size() == N) {                                                      // :65-85 warm start
    for (unsigned i = 0; i < N; i++) {
      if (std::fabs(z[i]) < zero_tol) bas.push_back(i); else nonbas.push_back(i);
    }
  } else {                                                                  // :86-103
    unsigned minw = (unsigned)(std::min_element(q, q + N) - q);
    if (q[minw] > -zero_tol) { z.assign(N, 0.0); status = LCP_TRIVIAL; return true; }
    nonbas.push_back(minw);
    for (unsigned i = 0; i < N; i++) if (i != minw) bas.push_back(i);
  }
  const unsigned MAX_PIV = 2 * N;                                           // :107
  Vec Msub, zz, w;
  for (pivots = 0; pivots < MAX_PIV; pivots++) {
    const unsigned k = (unsigned)nonbas.size(), nb = (unsigned)bas.size();
    Msub.assign((size_t)k * k, 0.0);                                        // :111 select_square
    for (unsigned c = 0; c < k; c++) for (unsigned r = 0; r < k; r++) Msub[(size_t)c * k + r] = M[(size_t)nonbas[c] * N + nonbas[r]];
    zz.resize(k);
    for (unsigned i = 0; i < k; i++) zz[i] = -q[nonbas[i]];                 // :113-115
    if (!solve_fast(Msub.data(), (int)k, zz.data())) { status = LCP_SINGULAR; n_pivots_total += pivots; return false; }  // :118-126
    w.resize(nb);                                                           // :129  w = Mmix*z + qbas
    for (unsigned i = 0; i < nb; i++) {
      double s = 0.0;
      for (unsigned c = 0; c < k; c++) s = std::fma(M[(size_t)nonbas[c] * N + bas[i]], zz[c], s);
      w[i] = s + q[bas[i]];
    }
    unsigned minw = (nb > 0) ? rand_min(w, zero_tol) : UINF;                // :130
    if (minw == UINF || w[minw] > -zero_tol) {                              // :135
      unsigned minz = (k > 0) ? rand_min(zz, zero_tol) : UINF;              // :138
      if (minz < UINF && zz[minz] < -zero_tol) {                            // :141-150
        unsigned idx = nonbas[minz];
        nonbas.erase(nonbas.begin() + minz);
        bas.push_back(idx);
        insertion_sort(bas);
        if (keep_log) log.push_back((int)idx | 0x40000000);
      } else {                                                              // :151-162
        z.assign(N, 0.0);
        for (unsigned j = 0; j < nonbas.size(); j++) z[nonbas[j]] = zz[j];
        n_pivots_total += pivots;
        return true;
      }
    } else {                                                                // :164-189
      unsigned idx = bas[minw];
      bas.erase(bas.begin() + minw);
      nonbas.push_back(idx);
      insertion_sort(nonbas);
      if (keep_log) log.push_back((int)idx);
      unsigned minz = (k > 0) ? rand_min(zz, zero_tol) : UINF;              // :176 (positions in the OLD ordering)
      if (minz < UINF && zz[minz] < -zero_tol) {                            // :179-188 (erases position in the NEW list)
        unsigned idx2 = nonbas[minz];
        nonbas.erase(nonbas.begin() + minz);
        bas.push_back(idx2);
        insertion_sort(bas);
        if (keep_log) log.push_back((int)idx2 | 0x40000000);
      }
    }
  }
  status = LCP_MAXITER;
  n_pivots_total += pivots;
  return false;                                                             // :192-195
}

// LCP.cpp:545-1003
bool LCP::lcp_lemke(int n_, const double* M, const double* q, Vec& z, double piv_tol, double zero_tol) {
  const unsigned n = (unsigned)n_;
  const unsigned MAXITER = std::min((unsigned)1000, 50 * n);                // :548
  n_lemke_calls++;
  pivots = 0;
  status = LCP_OK;
  if (n == 0) { z.clear(); return true; }                                   // :557-561
  std::fill(z.begin(), z.end(), 0.0);                                       // :564 warm start disabled
  const size_t z0_size = z.size();                                          // :567 _z0 = z (all zeros)
  if (zero_tol <= 0.0) zero_tol = EPS * norm_inf(M, n_, n_) * n;            // :570-571
  if (*std::min_element(q, q + n) > -zero_tol) { z.assign(n, 0.0); status = LCP_TRIVIAL; return true; }  // :578-584
  z.assign(2 * n, 0.0);                                                     // :596
  const unsigned t = 2 * n;
  unsigned entering = t, leaving = 0, lvindex;
  std::vector<unsigned> bas, nonbas;
  if (z0_size != n) {                                                       // :611-621 (rand() consumed for an unused restart basis)
    for (unsigned i = 0; i < n; i++) nonbas.push_back(i);
    if (tie == TIE_GLIBC_RAND) for (unsigned i = 0; i < n; i++) (void)rand();
  } else {                                                                  // :622-643 (_z0 is all zeros: nothing basic)
    for (unsigned i = 0; i < n; i++) nonbas.push_back(i);
  }
  // :691-699 standard initial basis
  Vec Bl((size_t)n * n, 0.0), x(q, q + n), Al, Be(n), dl(n), u(n);
  for (unsigned i = 0; i < n; i++) Bl[(size_t)i * n + i] = -1.0;
  // :737-758 initial basis is a solution? (cannot happen after the trivial test unless zero_tol is huge)
  bool anyneg = false;
  for (unsigned i = 0; i < n; i++) if (x[i] < 0.0) { anyneg = true; break; }
  if (!anyneg) { z.assign(n, 0.0); return true; }
  const double PIV_TOL = (piv_tol > 0.0) ? piv_tol : EPS * n * std::max(1.0, norm_inf(M, n_, n_));  // :761
  // :764-771 initial leaving variable: first minimum of x
  lvindex = (unsigned)(std::min_element(x.begin(), x.end()) - x.begin());
  double tval = -x[lvindex];
  for (unsigned i = 0; i < nonbas.size(); i++) bas.push_back(nonbas[i] + n);
  leaving = bas[lvindex];
  if (keep_log) log.push_back((int)leaving);
  // :776-785 pivot in the artificial variable
  bas[lvindex] = t;
  for (unsigned i = 0; i < n; i++) u[i] = (x[i] < 0.0) ? 1.0 : 0.0;
  for (unsigned i = 0; i < n; i++) {                                        // Be = -(Bl*u)
    double s = 0.0;
    for (unsigned c = 0; c < n; c++) s = std::fma(Bl[(size_t)c * n + i], u[c], s);
    Be[i] = -s;
  }
  for (unsigned i = 0; i < n; i++) x[i] += u[i] * tval;
  x[lvindex] = tval;
  for (unsigned i = 0; i < n; i++) Bl[(size_t)lvindex * n + i] = Be[i];

  std::vector<unsigned> j;
  for (pivots = 0; pivots < MAXITER; pivots++) {                            // :789
    if (leaving == t) {                                                     // :800-822
      for (unsigned i = 0; i < bas.size(); i++) z[bas[i]] = x[i];
      z.resize(n);
      n_pivots_total += pivots;
      return true;
    } else if (leaving < n) {                                               // :823-828
      entering = n + leaving;
      std::fill(Be.begin(), Be.end(), 0.0);
      Be[leaving] = -1.0;
    } else {                                                                // :829-833
      entering = leaving - n;
      for (unsigned i = 0; i < n; i++) Be[i] = M[(size_t)entering * n + i];
    }
    dl = Be;                                                                // :834-838
    Al = Bl;
    if (!solve_fast(Al.data(), n_, dl.data())) { status = LCP_SINGULAR; n_pivots_total += pivots; return false; }  // :840-850
    j.clear();                                                              // :886-889
    for (unsigned i = 0; i < n; i++) if (dl[i] > PIV_TOL) j.push_back(i);
    if (j.empty()) { status = LCP_RAY; n_pivots_total += pivots; return false; }   // :892-903
    double theta = std::numeric_limits<double>::max();                      // :920-924
    for (unsigned k = 0; k < j.size(); k++) theta = std::min(theta, (x[j[k]] + zero_tol) / dl[j[k]]);
    {                                                                       // :930-935
      std::vector<unsigned> keep;
      for (unsigned k = 0; k < j.size(); k++) if (x[j[k]] / dl[j[k]] <= theta) keep.push_back(j[k]);
      j.swap(keep);
    }
    if (j.empty()) { z.resize(n); status = LCP_EMPTY_RATIO; n_pivots_total += pivots; return false; }  // :946-958
    bool t_in = false;                                                      // :961-975
    for (unsigned k = 0; k < j.size(); k++) if (bas[j[k]] == t) { t_in = true; break; }
    if (t_in) lvindex = (unsigned)(std::find(bas.begin(), bas.end(), t) - bas.begin());
    else lvindex = j[0];
    leaving = bas[lvindex];                                                 // :978-980
    if (keep_log) log.push_back((int)leaving);
    const double ratio = x[lvindex] / dl[lvindex];                          // :983-988
    for (unsigned i = 0; i < n; i++) dl[i] *= ratio;
    for (unsigned i = 0; i < n; i++) x[i] -= dl[i];
    x[lvindex] = ratio;
    for (unsigned i = 0; i < n; i++) Bl[(size_t)lvindex * n + i] = Be[i];
    bas[lvindex] = entering;
  }
  z.resize(n);                                                              // :992-1002
  status = LCP_MAXITER;
  n_pivots_total += pivots;
  return false;
}

// Solution checks shared by both wrappers (LCP.cpp:240-256 with >=, :303-319 with >).
static bool verify(int n, const double* Mchk, const double* q, const Vec& z, double ZERO_TOL, bool strict) {
  if (z.size() != (size_t)n) return false;
  double minz = *std::min_element(z.begin(), z.end());
  if (strict ? !(minz > -ZERO_TOL) : !(minz >= -ZERO_TOL)) return false;
  Vec wx(n);
  for (int i = 0; i < n; i++) {                                             // M.mult(z, wx) += q
    double s = 0.0;
    for (int c = 0; c < n; c++) s = std::fma(Mchk[(size_t)c * n + i], z[c], s);
    wx[i] = s + q[i];
  }
  double minw = *std::min_element(wx.begin(), wx.end());
  if (strict ? !(minw > -ZERO_TOL) : !(minw >= -ZERO_TOL)) return false;
  double mn = DBL_MAX, mx = -DBL_MAX;
  for (int i = 0; i < n; i++) { double p = z[i] * wx[i]; mn = std::min(mn, p); mx = std::max(mx, p); }
  if (strict ? !(mn > -ZERO_TOL) : !(mn >= -ZERO_TOL)) return false;
  return mx < ZERO_TOL;
}

// LCP.cpp:212-350
bool LCP::lcp_fast_regularized(int n, const double* M, const double* q, Vec& z, int min_exp, unsigned step_exp,
                               int max_exp, double /*piv_tol*/, double zero_tol) {
  if (n == 0) { z.clear(); return true; }                                   // :218-222
  const double ZERO_TOL = (zero_tol > 0.0) ? zero_tol : n * norm_inf(M, n, n) * NEAR_ZERO;  // :228
  unsigned total_piv = 0;
  bool result = lcp_fast(n, M, q, z, zero_tol);                             // :236
  if (result && verify(n, M, q, z, ZERO_TOL, false)) return true;           // :237-256
  total_piv += pivots;                                                      // :278
  Vec MM((size_t)n * n);
  int attempt = 0;
  for (int rf = min_exp; rf < max_exp; rf += (int)step_exp, attempt++) {    // :281-340
    const double lambda = std::pow(10.0, (double)rf);
    std::copy(M, M + (size_t)n * n, MM.begin());
    for (int i = 0; i < n; i++) MM[(size_t)i * n + i] += lambda;
    result = lcp_fast(n, MM.data(), q, z, zero_tol);
    total_piv += pivots;
    if (result && verify(n, MM.data(), q, z, ZERO_TOL, true)) { pivots = total_piv; status = LCP_REGULARIZED + attempt; return true; }
  }
  pivots = total_piv;                                                       // :346
  status = LCP_UNVERIFIED;
  return false;
}

// LCP.cpp:353-487
bool LCP::lcp_lemke_regularized(int n, const double* M, const double* q, Vec& z, int min_exp, unsigned step_exp,
                                int max_exp, double piv_tol, double zero_tol) {
  if (n == 0) { z.clear(); return true; }
  const double ZERO_TOL = (zero_tol > 0.0) ? zero_tol : n * norm_inf(M, n, n) * NEAR_ZERO;  // :369
  unsigned total_piv = 0;
  bool result = lcp_lemke(n, M, q, z, piv_tol, zero_tol);                   // :377
  if (result && verify(n, M, q, z, ZERO_TOL, false)) return true;
  total_piv += pivots;
  Vec MM((size_t)n * n);
  int attempt = 0;
  for (int rf = min_exp; rf < max_exp; rf += (int)step_exp, attempt++) {    // :419-477
    const double lambda = std::pow(10.0, (double)rf);
    std::copy(M, M + (size_t)n * n, MM.begin());
    for (int i = 0; i < n; i++) MM[(size_t)i * n + i] += lambda;
    result = lcp_lemke(n, MM.data(), q, z, piv_tol, zero_tol);
    total_piv += pivots;
    if (result && verify(n, MM.data(), q, z, ZERO_TOL, true)) { pivots = total_piv; status = LCP_REGULARIZED + attempt; return true; }
  }
  pivots = total_piv;
  status = LCP_UNVERIFIED;
  return false;
}

}  // namespace oracle
