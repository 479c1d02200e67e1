// Analytic box-box narrowphase: separating-axis signed distance + closest points, face clipping / edge-edge contacts.
//
// Stands in for Moby's polyhedron-polyhedron leaf (include/Moby/CCD.inl:86-494: V-Clip distance, Seidel LP interior
// point, qhull half-space intersection, qhull 2-D hulls + O'Rourke polygon clipping for the "kissing" case) whose
// outputs depend on qhull's vertex order (rule H5).  The rule adopted here, by oracle and kernels alike:
//   * distance = largest separation over the 15 separating axes (3 + 3 face normals, 9 edge x edge); negative =
//     penetration depth.  For separated boxes that is the Euclidean distance whenever the closest features are
//     face-vertex, face-edge, face-face or edge-edge with interior closest points (every resting configuration); for
//     vertex-vertex / vertex-edge near misses the Euclidean distance is found by exhaustion (vertices against boxes, edge
//     pairs), so separated boxes get what V-Clip returns (test/VClipTest.cpp:24-107).
//   * a face axis wins unless an edge axis separates by more than BB_EDGE_SLACK more;
//   * face axis: the incident face (most anti-parallel face of the other box) is clipped against the side planes of
//     the reference face (Sutherland-Hodgman, incident-face vertex order, clip order -u,+u,-v,+v); every clipped vertex
//     within TOL of the reference plane is a contact at that vertex, violation = its signed distance from the plane;
//   * edge axis: one contact at the midpoint of the closest points of the two supporting edges;
//   * the normal points from geom2 (B) toward geom1 (A), like every other leaf (CollisionDetection.cpp:57-93).
#pragma once
namespace b2m {

#define BB_EDGE_SLACK 1e-9
#define BB_PARALLEL 1e-12

struct BoxBoxAxis {
  double s;        // separation along the winning axis (signed distance)
  int code;        // 0..2 face of A, 3..5 face of B, 6..14 edge a_i x b_j (6 + 3 i + j)
  V3 n;            // unit axis, pointing from A toward B
};

B2M_HD B2M_INL V3 box_axis(const double* R, int i) { return V3(R[i], R[3 + i], R[6 + i]); }   // column i of the row-major rotation

B2M_HD B2M_NOINL inline void boxbox_axis(const BodyRef& A, const BodyRef& B, BoxBoxAxis& r) {
  const double hA[3] = {A.dims[0] * 0.5, A.dims[1] * 0.5, A.dims[2] * 0.5}, hB[3] = {B.dims[0] * 0.5, B.dims[1] * 0.5, B.dims[2] * 0.5};
  const V3 p = ld3(B.x) - ld3(A.x);
  V3 a[3], b[3];
  for (int i = 0; i < 3; i++) { a[i] = box_axis(A.R, i); b[i] = box_axis(B.R, i); }
  double Rm[3][3], Q[3][3], pa[3], pb[3];
  for (int i = 0; i < 3; i++) { pa[i] = dot(p, a[i]); pb[i] = dot(p, b[i]); for (int j = 0; j < 3; j++) { Rm[i][j] = dot(a[i], b[j]); Q[i][j] = fabs(Rm[i][j]); } }
  double best = -B2M_INF; int code = -1; V3 n;
  for (int i = 0; i < 3; i++) {                                   // faces of A
    const double s = fabs(pa[i]) - (hA[i] + hB[0] * Q[i][0] + hB[1] * Q[i][1] + hB[2] * Q[i][2]);
    if (s > best) { best = s; code = i; n = pa[i] < 0.0 ? -a[i] : a[i]; }
  }
  for (int j = 0; j < 3; j++) {                                   // faces of B
    const double s = fabs(pb[j]) - (hB[j] + hA[0] * Q[0][j] + hA[1] * Q[1][j] + hA[2] * Q[2][j]);
    if (s > best) { best = s; code = 3 + j; n = pb[j] < 0.0 ? -b[j] : b[j]; }
  }
  double ebest = -B2M_INF; int ecode = -1; V3 en;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const V3 L = cross(a[i], b[j]);
      const double l = norm(L);
      if (l < BB_PARALLEL) continue;
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double e = dot(p, L);
      const double s = (fabs(e) - (hA[i1] * Q[i2][j] + hA[i2] * Q[i1][j] + hB[j1] * Q[i][j2] + hB[j2] * Q[i][j1])) / l;
      if (s > ebest) { ebest = s; ecode = 6 + 3 * i + j; en = L * ((e < 0.0 ? -1.0 : 1.0) / l); }
    }
  if (ecode >= 0 && ebest > best + BB_EDGE_SLACK) { best = ebest; code = ecode; n = en; }
  r.s = best; r.code = code; r.n = n;
}

// support vertex of a box in world direction d (largest d . x), ties toward the + side of each axis
B2M_HD B2M_INL V3 box_support(const BodyRef& X, const V3& d) {
  V3 v = ld3(X.x);
  for (int k = 0; k < 3; k++) { const V3 ax = box_axis(X.R, k); v = v + ax * ((dot(d, ax) < 0.0 ? -0.5 : 0.5) * X.dims[k]); }
  return v;
}

// closest points of the two supporting edges of an edge-edge axis (edge i of A, edge j of B, n from A toward B)
B2M_HD B2M_NOINL inline void boxbox_edge_points(const BodyRef& A, const BodyRef& B, int i, int j, const V3& n, V3& pa, V3& pb) {
  const V3 ua = box_axis(A.R, i), ub = box_axis(B.R, j);
  // a point on each edge: the support vertex moved to the middle of the edge direction
  V3 ca = box_support(A, n), cb = box_support(B, -n);
  ca = ca - ua * dot(ca - ld3(A.x), ua);
  cb = cb - ub * dot(cb - ld3(B.x), ub);
  // closest points of the lines ca + s ua, cb + t ub
  const V3 w = cb - ca;
  const double uaub = dot(ua, ub), q1 = dot(ua, w), q2 = -dot(ub, w);
  double d = 1.0 - uaub * uaub, s = 0.0, t = 0.0;
  if (d > BB_PARALLEL) { d = 1.0 / d; s = (q1 + uaub * q2) * d; t = (uaub * q1 + q2) * d; }
  const double ha = A.dims[i] * 0.5, hb = B.dims[j] * 0.5;
  s = fmin(fmax(s, -ha), ha); t = fmin(fmax(t, -hb), hb);
  pa = ca + ua * s; pb = cb + ub * t;
}

// closest point of box X to the world point p (clamp in the box frame)
B2M_HD B2M_INL V3 box_closest_world(const BodyRef& X, const V3& p) {
  V3 v = ld3(X.x);
  const V3 r = p - ld3(X.x);
  for (int k = 0; k < 3; k++) { const V3 ax = box_axis(X.R, k); const double h = 0.5 * X.dims[k]; v = v + ax * fmin(fmax(dot(r, ax), -h), h); }
  return v;
}
B2M_HD B2M_INL V3 box_corner(const BodyRef& X, int i) {
  V3 v = ld3(X.x);
  for (int k = 0; k < 3; k++) v = v + box_axis(X.R, k) * ((((i >> (2 - k)) & 1) ? -0.5 : 0.5) * X.dims[k]);
  return v;
}
// edge e (0..11) of box X: axis e / 4, the four sign combinations of the other two axes; endpoints p0, p0 + d
B2M_HD B2M_INL void box_edge(const BodyRef& X, int e, V3& p0, V3& d) {
  const int k = e / 4, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
  const double s1 = (e & 1) ? -0.5 : 0.5, s2 = (e & 2) ? -0.5 : 0.5;
  const V3 ak = box_axis(X.R, k);
  p0 = (ld3(X.x) + box_axis(X.R, k1) * (s1 * X.dims[k1])) + (box_axis(X.R, k2) * (s2 * X.dims[k2]) + ak * (-0.5 * X.dims[k]));
  d = ak * X.dims[k];
}
// closest points of two segments p1 + s d1, p2 + t d2 (s, t in [0,1])
B2M_HD B2M_INL void segment_points(const V3& p1, const V3& d1, const V3& p2, const V3& d2, V3& c1, V3& c2) {
  const V3 r = p1 - p2;
  const double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
  const double denom = a * e - b * b;
  double s = (denom > 1e-300) ? fmin(fmax((b * f - c * e) / denom, 0.0), 1.0) : 0.0;
  double t = (b * s + f) / e;
  if (t < 0.0) { t = 0.0; s = fmin(fmax(-c / a, 0.0), 1.0); }
  else if (t > 1.0) { t = 1.0; s = fmin(fmax((b - c) / a, 0.0), 1.0); }
  c1 = p1 + d1 * s; c2 = p2 + d2 * t;
}
// Euclidean distance of two SEPARATED boxes by exhaustion: every vertex of one box against the other box, every edge
// pair -- what Polyhedron::vclip (src/Polyhedron.cpp:1238) returns for them and what test/VClipTest.cpp:24-107 checks
B2M_HD B2M_NOINL inline void boxbox_separated_dist(const BodyRef& A, const BodyRef& B, double& dist, V3& pA, V3& pB) {
  double best = B2M_INF;
  for (int i = 0; i < 8; i++) { const V3 v = box_corner(A, i), c = box_closest_world(B, v); const double d = norm(v - c); if (d < best) { best = d; pA = v; pB = c; } }
  for (int i = 0; i < 8; i++) { const V3 v = box_corner(B, i), c = box_closest_world(A, v); const double d = norm(v - c); if (d < best) { best = d; pA = c; pB = v; } }
  for (int ea = 0; ea < 12; ea++) {
    V3 p1, d1; box_edge(A, ea, p1, d1);
    for (int eb = 0; eb < 12; eb++) {
      V3 p2, d2, c1, c2; box_edge(B, eb, p2, d2);
      segment_points(p1, d1, p2, d2, c1, c2);
      const double d = norm(c1 - c2);
      if (d < best) { best = d; pA = c1; pB = c2; }
    }
  }
  dist = best;
}

// Signed distance and closest points (pA on A, pB on B).  Touching / penetrating (largest separation over the 15 axes
// <= 0): that separation (the penetration depth of the Minkowski difference, test/VClipTest.cpp:177-247).  Separated:
// the separation IS the Euclidean distance when the closest features are a face and a vertex that projects into the
// face, or two edges at interior points; otherwise (vertex-vertex, vertex-edge, clamped edge-edge) the exhaustive search.
B2M_HD B2M_NOINL inline void boxbox_signed_dist(const BodyRef& A, const BodyRef& B, double& dist, V3& pA, V3& pB) {
  BoxBoxAxis ax; boxbox_axis(A, B, ax);
  dist = ax.s;
  bool exact = true;
  if (ax.code < 6) {
    const bool refA = ax.code < 3;
    const BodyRef& Rb = refA ? A : B;
    const int k = refA ? ax.code : ax.code - 3;
    if (refA) { pB = box_support(B, -ax.n); pA = pB - ax.n * ax.s; }
    else { pA = box_support(A, ax.n); pB = pA + ax.n * ax.s; }
    const V3 onface = (refA ? pA : pB) - ld3(Rb.x);
    for (int j = 1; j <= 2; j++) { const int kk = (k + j) % 3; if (fabs(dot(onface, box_axis(Rb.R, kk))) > 0.5 * Rb.dims[kk]) exact = false; }
  } else {
    boxbox_edge_points(A, B, (ax.code - 6) / 3, (ax.code - 6) % 3, ax.n, pA, pB);
    const V3 w = pB - pA;
    if (fabs(dot(w, ax.n) - ax.s) > 1e-12 * fmax(1.0, fabs(ax.s)) || fabs(norm(w) - fabs(ax.s)) > 1e-12 * fmax(1.0, fabs(ax.s))) exact = false;
  }
  if (ax.s > 0.0 && !exact) boxbox_separated_dist(A, B, dist, pA, pB);
}

// Contacts (at most 8).  out[k].n points from B toward A; b1 / b2 are filled by the caller.
B2M_HD B2M_NOINL inline int boxbox_contacts(const BodyRef& A, const BodyRef& B, double TOL, V3* pts, double* depth, V3& normal) {
  BoxBoxAxis ax; boxbox_axis(A, B, ax);
  if (ax.s > TOL) return 0;
  normal = -ax.n;
  if (ax.code >= 6) {
    V3 pa, pb; boxbox_edge_points(A, B, (ax.code - 6) / 3, (ax.code - 6) % 3, ax.n, pa, pb);
    pts[0] = (pa + pb) * 0.5; depth[0] = ax.s;
    return 1;
  }
  // reference box Rb (its face normal nr points toward the incident box Ib)
  const bool refA = ax.code < 3;
  const BodyRef& Rb = refA ? A : B; const BodyRef& Ib = refA ? B : A;
  const int k = refA ? ax.code : ax.code - 3;
  const V3 nr = refA ? ax.n : -ax.n;
  // incident face: axis most parallel to nr, on the side facing the reference box
  int kin = 0; double bestd = -1.0;
  for (int j = 0; j < 3; j++) { const double d = fabs(dot(nr, box_axis(Ib.R, j))); if (d > bestd) { bestd = d; kin = j; } }
  const V3 ai = box_axis(Ib.R, kin);
  const double sgn = dot(nr, ai) > 0.0 ? -1.0 : 1.0;            // incident face outward normal = sgn * ai, anti-parallel to nr
  const int j1 = (kin + 1) % 3, j2 = (kin + 2) % 3;
  const V3 fc = ld3(Ib.x) + ai * (sgn * 0.5 * Ib.dims[kin]);
  const V3 e1 = box_axis(Ib.R, j1) * (0.5 * Ib.dims[j1]), e2 = box_axis(Ib.R, j2) * (0.5 * Ib.dims[j2]);
  V3 poly[8], tmp[8];
  poly[0] = fc + e1 + e2; poly[1] = fc - e1 + e2; poly[2] = fc - e1 - e2; poly[3] = fc + e1 - e2;
  int np = 4;
  // clip against the four side planes of the reference face
  const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
  const V3 cR = ld3(Rb.x);
  for (int side = 0; side < 4; side++) {
    const int kk = side < 2 ? k1 : k2;
    const V3 u = box_axis(Rb.R, kk) * ((side & 1) ? 1.0 : -1.0);
    const double h = 0.5 * Rb.dims[kk];
    int nq = 0;
    for (int v = 0; v < np; v++) {
      const V3 P0 = poly[v], P1 = poly[(v + 1) % np];
      const double d0 = dot(P0 - cR, u) - h, d1 = dot(P1 - cR, u) - h;      // <= 0: inside
      if (d0 <= 0.0) { if (nq < 8) tmp[nq] = P0; nq++; }
      if ((d0 <= 0.0) != (d1 <= 0.0)) { const double t = d0 / (d0 - d1); if (nq < 8) tmp[nq] = P0 + (P1 - P0) * t; nq++; }
    }
    np = nq < 8 ? nq : 8;
    for (int v = 0; v < np; v++) poly[v] = tmp[v];
    if (np == 0) return 0;
  }
  const V3 fR = cR + nr * (0.5 * Rb.dims[k]);
  int cnt = 0;
  for (int v = 0; v < np; v++) {
    const double d = dot(poly[v] - fR, nr);
    if (d <= TOL) { pts[cnt] = poly[v]; depth[cnt] = d; cnt++; }
  }
  return cnt;
}

// CCD::calc_next_CA_Euler_step_polyhedron_polyhedron (CCD.cpp:468-541) for two boxes: per-vertex bound on the time to
// reach the contact plane <n0, x> = offset0.  rvA: velocity of A relative to B in A's frame; rvB the same in B's frame.
B2M_HD B2M_NOINL inline double next_CA_box_box(const BodyRef& A, const BodyRef& B, const V3& rvA_lin, const V3& rvA_ang, const V3& rvB_lin, const V3& rvB_ang,
                                     const V3& n0, double offset0) {
  double max_step = B2M_INF;
  const V3 nA = rotT(A.R, n0), nB = rotT(B.R, -n0);
  const V3 p0 = n0 * offset0;
  const double offsetA = dot(nA, to_local(A, p0)), offsetB = dot(nB, to_local(B, p0));
  const double avA = norm(rvA_ang), avB = norm(rvB_ang);
  const double lvA = -dot(nA, rvA_lin), lvB = dot(nB, rvB_lin);
  for (int i = 0; i < 8; i++) {
    const V3 vtx = box_vertex(A.dims, i);
    const double dist = dot(nA, vtx) - offsetA;
    if (dist < B2M_NEAR_ZERO) continue;
    const double speed = fmax(0.0, lvA + avA * norm(vtx));
    max_step = fmin(max_step, dist / speed);
  }
  for (int i = 0; i < 8; i++) {
    const V3 vtx = box_vertex(B.dims, i);
    const double dist = dot(nB, vtx) - offsetB;
    if (dist < B2M_NEAR_ZERO) continue;
    const double speed = fmax(0.0, lvB + avB * norm(vtx));
    max_step = fmin(max_step, dist / speed);
  }
  return max_step;
}

}  // namespace b2m
