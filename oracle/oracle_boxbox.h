// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Box-box narrowphase under rule H5 of SURVEY.md 8(a'): the reference's polyhedron-polyhedron leaf
// (include/Moby/CCD.inl:86-494) runs V-Clip (src/Polyhedron.cpp:1238), a Seidel LP and qhull, and in the resting
// ("kissing", CCD.inl:140-316) case returns the vertices of the intersection of the two faces' 2-D hulls in an order
// that depends on qhull's output.  qhull is not available and its vertex order cannot be restated, so oracle and product
// adopt one analytic rule (a deliberate, documented deviation; contact SETS of resting boxes match the reference's
// polygon-intersection vertices, contact ORDER is canonical):
//   1. signed distance = max separation over the 15 separating axes (face normals of A, of B, edge x edge) -- the
//      penetration depth when negative; for separated boxes whose closest features are not face-vertex / interior
//      edge-edge, the Euclidean distance by exhaustion (vertices against boxes, edge pairs), as V-Clip returns it;
//      a face axis is kept unless an edge axis separates by more than 1e-9 more;
//   2. face axis: clip the incident face of the other box (Sutherland-Hodgman) against the side planes of the reference
//      face, in the order -u, +u, -v, +v with (u, v) the two axes following the face axis cyclically; each surviving
//      vertex within TOL of the reference plane is a contact, violation = its signed distance to that plane;
//   3. edge axis: one contact at the midpoint of the closest points of the two supporting edges;
//   4. normal from geom2 toward geom1, as create_contact expects (CollisionDetection.cpp:57-93).
// PARITY UNPINNED against the reference for this leaf (no golden vectors exist for box-box contacts; test/VClipTest.cpp
// checks distances only): pinned instead by geometric properties in tests/test_boxbox.py.
#pragma once
#include <array>
#include <cmath>
#include <vector>

namespace oracle {
namespace bb {

struct Vec3 { double x, y, z; };
static inline Vec3 add(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline Vec3 mul(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dt(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline Vec3 cr(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline double len(Vec3 a) { return std::sqrt(dt(a, a)); }

struct Box {
  Vec3 c;           // centre
  Vec3 ax[3];       // unit axes (columns of the rotation)
  double ext[3];    // full edge lengths
};

struct Axis { double sep; int code; Vec3 n; };   // n from A toward B; code 0-2 A faces, 3-5 B faces, 6+3i+j edge pairs

static inline Axis best_axis(const Box& A, const Box& B) {
  const Vec3 p = sub(B.c, A.c);
  double R[3][3], Q[3][3], pa[3], pb[3];
  for (int i = 0; i < 3; i++) {
    pa[i] = dt(p, A.ax[i]); pb[i] = dt(p, B.ax[i]);
    for (int j = 0; j < 3; j++) { R[i][j] = dt(A.ax[i], B.ax[j]); Q[i][j] = std::fabs(R[i][j]); }
  }
  Axis face{-1.7976931348623157e308, -1, {0, 0, 0}}, edge{-1.7976931348623157e308, -1, {0, 0, 0}};
  for (int i = 0; i < 3; i++) {
    const double s = std::fabs(pa[i]) - (A.ext[i] * 0.5 + B.ext[0] * 0.5 * Q[i][0] + B.ext[1] * 0.5 * Q[i][1] + B.ext[2] * 0.5 * Q[i][2]);
    if (s > face.sep) face = {s, i, pa[i] < 0.0 ? mul(A.ax[i], -1.0) : A.ax[i]};
  }
  for (int j = 0; j < 3; j++) {
    const double s = std::fabs(pb[j]) - (B.ext[j] * 0.5 + A.ext[0] * 0.5 * Q[0][j] + A.ext[1] * 0.5 * Q[1][j] + A.ext[2] * 0.5 * Q[2][j]);
    if (s > face.sep) face = {s, 3 + j, pb[j] < 0.0 ? mul(B.ax[j], -1.0) : B.ax[j]};
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const Vec3 L = cr(A.ax[i], B.ax[j]);
      const double l = len(L);
      if (l < 1e-12) continue;
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double e = dt(p, L);
      const double s = (std::fabs(e) - (A.ext[i1] * 0.5 * Q[i2][j] + A.ext[i2] * 0.5 * Q[i1][j] + B.ext[j1] * 0.5 * Q[i][j2] + B.ext[j2] * 0.5 * Q[i][j1])) / l;
      if (s > edge.sep) edge = {s, 6 + 3 * i + j, mul(L, (e < 0.0 ? -1.0 : 1.0) / l)};
    }
  if (edge.code >= 0 && edge.sep > face.sep + 1e-9) return edge;
  return face;
}

static inline Vec3 support(const Box& X, Vec3 d) {
  Vec3 v = X.c;
  for (int k = 0; k < 3; k++) v = add(v, mul(X.ax[k], (dt(d, X.ax[k]) < 0.0 ? -0.5 : 0.5) * X.ext[k]));
  return v;
}

static inline void edge_points(const Box& A, const Box& B, int i, int j, Vec3 n, Vec3& pa, Vec3& pb) {
  const Vec3 ua = A.ax[i], ub = B.ax[j];
  Vec3 ca = support(A, n), cb = support(B, mul(n, -1.0));
  ca = sub(ca, mul(ua, dt(sub(ca, A.c), ua)));
  cb = sub(cb, mul(ub, dt(sub(cb, B.c), ub)));
  const Vec3 w = sub(cb, ca);
  const double uaub = dt(ua, ub), q1 = dt(ua, w), q2 = -dt(ub, w);
  double d = 1.0 - uaub * uaub, s = 0.0, t = 0.0;
  if (d > 1e-12) { d = 1.0 / d; s = (q1 + uaub * q2) * d; t = (uaub * q1 + q2) * d; }
  const double ha = A.ext[i] * 0.5, hb = B.ext[j] * 0.5;
  s = std::fmin(std::fmax(s, -ha), ha); t = std::fmin(std::fmax(t, -hb), hb);
  pa = add(ca, mul(ua, s)); pb = add(cb, mul(ub, t));
}

// closest point of box X to the world point p (clamp in the box frame)
static inline Vec3 closest_on_box(const Box& X, Vec3 p) {
  Vec3 v = X.c;
  const Vec3 r = sub(p, X.c);
  for (int k = 0; k < 3; k++) { const double h = 0.5 * X.ext[k]; v = add(v, mul(X.ax[k], std::fmin(std::fmax(dt(r, X.ax[k]), -h), h))); }
  return v;
}
static inline Vec3 corner(const Box& X, int i) {
  Vec3 v = X.c;
  for (int k = 0; k < 3; k++) v = add(v, mul(X.ax[k], (((i >> (2 - k)) & 1) ? -0.5 : 0.5) * X.ext[k]));
  return v;
}
// edge e (0..11) of box X: axis e / 4, the four sign combinations of the other two axes; endpoints p0, p0 + d
static inline void edge_of(const Box& X, int e, Vec3& p0, Vec3& d) {
  const int k = e / 4, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
  const double s1 = (e & 1) ? -0.5 : 0.5, s2 = (e & 2) ? -0.5 : 0.5;
  p0 = add(add(X.c, mul(X.ax[k1], s1 * X.ext[k1])), add(mul(X.ax[k2], s2 * X.ext[k2]), mul(X.ax[k], -0.5 * X.ext[k])));
  d = mul(X.ax[k], X.ext[k]);
}
// closest points of two segments p1 + s d1, p2 + t d2 (s, t in [0,1])
static inline void segment_points(Vec3 p1, Vec3 d1, Vec3 p2, Vec3 d2, Vec3& c1, Vec3& c2) {
  const Vec3 r = sub(p1, p2);
  const double a = dt(d1, d1), e = dt(d2, d2), f = dt(d2, r), c = dt(d1, r), b = dt(d1, d2);
  const double denom = a * e - b * b;
  double s = (denom > 1e-300) ? std::fmin(std::fmax((b * f - c * e) / denom, 0.0), 1.0) : 0.0;
  double t = (b * s + f) / e;
  if (t < 0.0) { t = 0.0; s = std::fmin(std::fmax(-c / a, 0.0), 1.0); }
  else if (t > 1.0) { t = 1.0; s = std::fmin(std::fmax((b - c) / a, 0.0), 1.0); }
  c1 = add(p1, mul(d1, s)); c2 = add(p2, mul(d2, t));
}
// Euclidean distance of two SEPARATED boxes by exhaustion: every vertex of one box against the other box, every edge
// pair -- what Polyhedron::vclip (src/Polyhedron.cpp:1238) returns for them and what test/VClipTest.cpp:24-107 checks
static inline void separated_dist(const Box& A, const Box& B, double& dist, Vec3& pA, Vec3& pB) {
  double best = 1.7976931348623157e308;
  for (int i = 0; i < 8; i++) { const Vec3 v = corner(A, i), c = closest_on_box(B, v); const double d = len(sub(v, c)); if (d < best) { best = d; pA = v; pB = c; } }
  for (int i = 0; i < 8; i++) { const Vec3 v = corner(B, i), c = closest_on_box(A, v); const double d = len(sub(v, c)); if (d < best) { best = d; pA = c; pB = v; } }
  for (int ea = 0; ea < 12; ea++) {
    Vec3 p1, d1; edge_of(A, ea, p1, d1);
    for (int eb = 0; eb < 12; eb++) {
      Vec3 p2, d2, c1, c2; edge_of(B, eb, p2, d2);
      segment_points(p1, d1, p2, d2, c1, c2);
      const double d = len(sub(c1, c2));
      if (d < best) { best = d; pA = c1; pB = c2; }
    }
  }
  dist = best;
}

// Signed distance and closest points.  Touching / penetrating (largest separation over the 15 axes <= 0): that separation
// (the penetration depth of the Minkowski difference, test/VClipTest.cpp:177-247).  Separated: the separation IS the
// Euclidean distance when the closest features are a face and a vertex that projects into the face, or two edges at
// interior points; otherwise (vertex-vertex, vertex-edge, clamped edge-edge) the exhaustive search above.
static inline void signed_dist(const Box& A, const Box& B, double& dist, Vec3& pA, Vec3& pB) {
  const Axis ax = best_axis(A, B);
  dist = ax.sep;
  bool exact = true;
  if (ax.code < 6) {
    const bool refA = ax.code < 3;
    const Box& Rb = refA ? A : B;
    const int k = refA ? ax.code : ax.code - 3;
    if (refA) { pB = support(B, mul(ax.n, -1.0)); pA = sub(pB, mul(ax.n, ax.sep)); }
    else { pA = support(A, ax.n); pB = add(pA, mul(ax.n, ax.sep)); }
    const Vec3 onface = sub(refA ? pA : pB, Rb.c);
    for (int j = 1; j <= 2; j++) { const int kk = (k + j) % 3; if (std::fabs(dt(onface, Rb.ax[kk])) > 0.5 * Rb.ext[kk]) exact = false; }
  } else {
    const int i = (ax.code - 6) / 3, j = (ax.code - 6) % 3;
    edge_points(A, B, i, j, ax.n, pA, pB);
    // interior closest points: the connecting segment is (anti)parallel to the axis and as long as the separation
    const Vec3 w = sub(pB, pA);
    if (std::fabs(dt(w, ax.n) - ax.sep) > 1e-12 * std::fmax(1.0, std::fabs(ax.sep)) || std::fabs(len(w) - std::fabs(ax.sep)) > 1e-12 * std::fmax(1.0, std::fabs(ax.sep))) exact = false;
  }
  if (ax.sep > 0.0 && !exact) separated_dist(A, B, dist, pA, pB);
}

struct Point { Vec3 p; double violation; };

static inline std::vector<Point> contacts(const Box& A, const Box& B, double TOL, Vec3& normal) {
  std::vector<Point> out;
  const Axis ax = best_axis(A, B);
  if (ax.sep > TOL) return out;
  normal = mul(ax.n, -1.0);
  if (ax.code >= 6) {
    Vec3 pa, pb;
    edge_points(A, B, (ax.code - 6) / 3, (ax.code - 6) % 3, ax.n, pa, pb);
    out.push_back({mul(add(pa, pb), 0.5), ax.sep});
    return out;
  }
  const bool refA = ax.code < 3;
  const Box& Rb = refA ? A : B; const Box& Ib = refA ? B : A;
  const int k = refA ? ax.code : ax.code - 3;
  const Vec3 nr = refA ? ax.n : mul(ax.n, -1.0);
  int kin = 0; double bestd = -1.0;
  for (int j = 0; j < 3; j++) { const double d = std::fabs(dt(nr, Ib.ax[j])); if (d > bestd) { bestd = d; kin = j; } }
  const double sgn = dt(nr, Ib.ax[kin]) > 0.0 ? -1.0 : 1.0;
  const int j1 = (kin + 1) % 3, j2 = (kin + 2) % 3;
  const Vec3 fc = add(Ib.c, mul(Ib.ax[kin], sgn * 0.5 * Ib.ext[kin]));
  const Vec3 e1 = mul(Ib.ax[j1], 0.5 * Ib.ext[j1]), e2 = mul(Ib.ax[j2], 0.5 * Ib.ext[j2]);
  std::vector<Vec3> poly = {add(add(fc, e1), e2), add(sub(fc, e1), e2), sub(sub(fc, e1), e2), sub(add(fc, e1), e2)};
  const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
  for (int side = 0; side < 4 && !poly.empty(); side++) {
    const int kk = side < 2 ? k1 : k2;
    const Vec3 u = mul(Rb.ax[kk], (side & 1) ? 1.0 : -1.0);
    const double h = 0.5 * Rb.ext[kk];
    std::vector<Vec3> next;
    for (size_t v = 0; v < poly.size(); v++) {
      const Vec3 P0 = poly[v], P1 = poly[(v + 1) % poly.size()];
      const double d0 = dt(sub(P0, Rb.c), u) - h, d1 = dt(sub(P1, Rb.c), u) - h;
      if (d0 <= 0.0) next.push_back(P0);
      if ((d0 <= 0.0) != (d1 <= 0.0)) next.push_back(add(P0, mul(sub(P1, P0), d0 / (d0 - d1))));
    }
    if (next.size() > 8) next.resize(8);
    poly.swap(next);
  }
  const Vec3 fR = add(Rb.c, mul(nr, 0.5 * Rb.ext[k]));
  for (const Vec3& v : poly) {
    const double d = dt(sub(v, fR), nr);
    if (d <= TOL) out.push_back({v, d});
  }
  return out;
}

}  // namespace bb
}  // namespace oracle
