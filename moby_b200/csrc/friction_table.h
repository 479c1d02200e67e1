// Host-side table of the friction-cone direction cosines / sines, evaluated with the host libm exactly as the
// reference evaluates them (ImpactConstraintHandlerQP.cpp:464-468; ImpactConstraintHandlerLCP.cpp:259-275), so the
// device never calls its own cos/sin.  Layout: [4][NKMAX+1][NKMAX/2] = QP cos, QP sin, AP cos, AP sin, followed by the
// spoke directions of the rimless wheel (example/rimless-wheel/coldet-plugin.cpp:107): [WHEEL_NS_MAX+1][WHEEL_NS_MAX][2] =
// cos, sin of theta = M_PI * i * 2.0 / N for N spokes, spoke i.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#ifndef B2M_NKMAX
#define B2M_NKMAX 64
#endif
#ifndef B2M_WHEEL_NS_MAX
#define B2M_WHEEL_NS_MAX 16
#endif
#define B2M_WTAB_OFF ((size_t)4 * (B2M_NKMAX + 1) * (B2M_NKMAX / 2))

inline std::vector<double> b2m_friction_table() {
  const int H = B2M_NKMAX / 2, S = (B2M_NKMAX + 1) * H;
  std::vector<double> t((size_t)4 * S + (size_t)2 * (B2M_WHEEL_NS_MAX + 1) * B2M_WHEEL_NS_MAX, 0.0);
  for (unsigned N = 1; N <= B2M_WHEEL_NS_MAX; N++)
    for (unsigned i = 0; i < N; i++) {
      const double theta = M_PI * i * 2.0 / N;
      t[B2M_WTAB_OFF + 2 * ((size_t)N * B2M_WHEEL_NS_MAX + i)] = std::cos(theta);
      t[B2M_WTAB_OFF + 2 * ((size_t)N * B2M_WHEEL_NS_MAX + i) + 1] = std::sin(theta);
    }
  for (int NK = 4; NK <= B2M_NKMAX; NK++) {
    const int half = NK / 2;
    for (int j = 0; j < half && j < H; j++) {
      const double theta = (double)j / (half - 1) * M_PI_2;
      t[(size_t)NK * H + j] = std::cos(theta);
      t[(size_t)S + (size_t)NK * H + j] = std::sin(theta);
    }
    const int nk4 = (NK + 4) / 4;
    for (int k = 0; k < nk4 && k < H; k++) {
      t[(size_t)2 * S + (size_t)NK * H + k] = std::cos((M_PI * k) / (2.0 * nk4));
      t[(size_t)3 * S + (size_t)NK * H + k] = std::sin((M_PI * k) / (2.0 * nk4));
    }
  }
  return t;
}

// Upper bounds of contacts and LCP dimension for one env of a scene (shape/enabled/NK of every body pair).
inline void b2m_env_bounds(int nb, const int* shape, const int* enabled, const int* NK, int model, int& cmax, int& nmax, int& npairs) {
  cmax = 0; nmax = 0; npairs = 0;
  for (int i = 0; i < nb; i++)
    for (int j = i + 1; j < nb; j++) {
      if (!(enabled[i] || enabled[j])) continue;
      if (shape[i] == 0 || shape[j] == 0) continue;
      const int nk = NK[i * nb + j];
      if (nk == 0) continue;
      npairs++;
      int cnt = 1;                                     // sphere-anything: one contact
      const bool bi = shape[i] == 2, bj = shape[j] == 2, pi = shape[i] == 3, pj = shape[j] == 3;
      if ((bi && pj) || (pi && bj)) cnt = 4;           // a non-degenerate box touches a plane with at most 4 vertices
      if (bi && bj) cnt = 8;
      if (pi && pj) cnt = 0;
      if (shape[i] == 4 || shape[j] == 4) cnt = (pi || pj) ? 4 : 0;
      if (shape[i] >= 5 || shape[j] >= 5) cnt = ((shape[i] == 5 && shape[j] == 6) || (shape[i] == 6 && shape[j] == 5)) ? 6 : 0;   // pin joint as six contacts   // rimless wheel: spoke tips against a plane only (two tips, two sides when W > 0)
      cmax += cnt;
      nmax += cnt * (model == 1 ? 5 + (nk > 4 ? (nk + 4) / 4 : 1) : 6 + nk / 2);
    }
}

// LCP classes of the impact phase: ascending bounds on the LCP dimension (and the contact count they allow); an env
// is solved with the working set of the first class its contacts fit.  The last class is the scene's own bound.
// Returns the number of classes (<= max_classes).
inline int b2m_class_table(int nmax, int cmax, int model, int max_classes, int* class_nmax, int* class_cmax) {
  static const int bounds[] = {8, 16, 24, 32, 40, 48, 64, 80, 96, 128, 160};
  const int minper = (model == 1) ? 6 : 8;            // fewest LCP rows one contact brings (NK = 4)
  int k = 0;
  for (int b : bounds)
    if (b < nmax && k < max_classes - 1) { class_nmax[k] = b; class_cmax[k] = std::max(1, std::min(cmax, b / minper)); k++; }
  class_nmax[k] = nmax; class_cmax[k] = cmax; k++;
  return k;
}

// Working-set bounds of a whole batch descriptor (shared by b200moby_create and the host-compiled test harness).
// Returns nullptr, or a message when a friction-cone-edges value is invalid.
#include "../../include/b200moby.h"
inline const char* b2m_scene_bounds(const b200moby_scene_desc* d, int& cmax, int& nmax, int& npmax) {
  const int ne = d->n_envs, nb = d->n_bodies;
  cmax = 0; nmax = 0; npmax = 0;
  std::vector<int> sh(nb), en(nb), nk(nb * nb);
  int per_contact = 0;
  for (int e = 0; e < ne; e++) {
    for (int b = 0; b < nb; b++) {
      sh[b] = d->shape[(size_t)b * ne + e]; en[b] = d->enabled[(size_t)b * ne + e];
      if (sh[b] == 4) {
        const double ns = d->dims[((size_t)b * 3 + 2) * ne + e];
        if (!(ns >= 1 && ns <= B2M_WHEEL_NS_MAX) || ns != (double)(int)ns) return "rimless wheel: dims = (R, W, N_SPOKES) with 1 <= N_SPOKES <= 16";
      }
    }
    for (int i = 0; i < nb; i++) for (int j = i + 1; j < nb; j++) {
      const int k = d->NK[((size_t)i * nb + j) * ne + e];
      if (k != 0 && (k < 4 || k > B2M_NKMAX || (k & 1))) return "friction-cone-edges must be even and in [4,64] (ContactParameters.cpp:129-136)";
      nk[i * nb + j] = k;
      if (k) per_contact = std::max(per_contact, d->impact_model == 1 ? 5 + (k > 4 ? (k + 4) / 4 : 1) : 6 + k / 2);
    }
    int c, n, np;
    b2m_env_bounds(nb, sh.data(), en.data(), nk.data(), d->impact_model, c, n, np);
    cmax = std::max(cmax, c); nmax = std::max(nmax, n); npmax = std::max(npmax, np);
  }
  if (d->max_contacts > 0 && d->max_contacts < cmax) { cmax = d->max_contacts; nmax = std::min(nmax, cmax * per_contact); }
  if (d->max_lcp_n > 0 && d->max_lcp_n < nmax) nmax = d->max_lcp_n;
  cmax = std::max(cmax, 1); nmax = std::max(nmax, 1); npmax = std::max(npmax, 1);
  return nullptr;
}
