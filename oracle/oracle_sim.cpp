// ORACLE -- TEST INFRASTRUCTURE ONLY.  See oracle_sim.h for scope, citations and the PARITY UNPINNED note.
#include "oracle_sim.h"
#include "oracle_boxbox.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <queue>
#include <set>

namespace oracle {

static const double NEAR_ZERO = std::sqrt(std::numeric_limits<double>::epsilon());  // Constants.h:21
static const double INF = std::numeric_limits<double>::max();

// ---------- small vector helpers ----------
static inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(const V3& a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(const V3& a, const V3& b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(const V3& a) { double n = norm(a); return V3(a.x / n, a.y / n, a.z / n); }
// R (row-major) * v and R^T * v
static inline V3 rot(const double* R, const V3& v) { return V3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z); }
static inline V3 rotT(const double* R, const V3& v) { return V3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z); }

Sim::Sim() { min_step_size = NEAR_ZERO; stab_eps = NEAR_ZERO; }

void Sim::init(int nb) {
  bodies.assign(nb, Body());
  cparams.assign((size_t)nb * nb, ContactParams());
  for (auto& c : cparams) c.NK = 0;
  for (auto& b : bodies) update_pose(b);
}

// Quaternion (x y z w) -> rotation matrix.  Ravelin: Matrix3d(Quatd).
void Sim::update_pose(Body& b) {
  const double x = b.quat[0], y = b.quat[1], z = b.quat[2], w = b.quat[3];
  double* R = b.R;
  R[0] = 1.0 - 2.0 * (y * y + z * z); R[1] = 2.0 * (x * y - w * z);       R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);       R[4] = 1.0 - 2.0 * (x * x + z * z); R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);       R[7] = 2.0 * (y * z + w * x);       R[8] = 1.0 - 2.0 * (x * x + y * y);
}

// RCArticulatedBodyd::update_link_poses / update_link_velocities (RCArticulatedBody.cpp:102,142-143): link bodies follow (jq, jqd)
static void R_to_quat(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) { const double s = std::sqrt(tr + 1.0) * 2.0; q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s; }
  else { const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s; }
}
void Sim::rc_update_links() {
  if (!has_rc) return;
  const Body& base = bodies[rc_first];
  for (int c = 0; c < 3; c++) rc.base_x[c] = base.x[c];
  for (int c = 0; c < 4; c++) rc.base_quat[c] = base.quat[c];
  for (int i = 0; i < rc.n_links; i++) { rc.mass[i] = bodies[rc_first + i].mass; for (int c = 0; c < 3; c++) rc.J[i][c] = bodies[rc_first + i].J[c]; }
  RCKin k;
  rc_kinematics(rc, jq.data(), jqd.data(), k);
  for (int i = 1; i < rc.n_links; i++) {
    Body& b = bodies[rc_first + i];
    b.x = V3(k.x[i][0], k.x[i][1], k.x[i][2]);
    for (int c = 0; c < 9; c++) b.R[c] = k.R[i][c];
    R_to_quat(b.R, b.quat);
    b.vl = V3(k.vl[i][0], k.vl[i][1], k.vl[i][2]);
    b.va = V3(k.va[i][0], k.va[i][1], k.va[i][2]);
  }
}

// velocity of the body-fixed point coincident with p: the linear part of Pose3d::transform(frame at p, v)
static inline V3 point_vel(const Body& b, const V3& p) {
  if (!b.enabled) return V3();
  return b.vl + cross(b.va, p - b.x);
}

// BoxPrimitive::get_vertices order (BoxPrimitive.cpp:358-365), body frame
static inline V3 box_vertex(const Body& b, int i) {
  const double X = b.dims[0] * 0.5, Y = b.dims[1] * 0.5, Z = b.dims[2] * 0.5;
  return V3((i & 4) ? -X : X, (i & 2) ? -Y : Y, (i & 1) ? -Z : Z);
}
static inline V3 to_global(const Body& b, const V3& p) { return b.x + rot(b.R, p); }
static inline V3 to_local(const Body& b, const V3& p) { return rotT(b.R, p - b.x); }

// ---------- broad phase: CollisionDetection.cpp:28-54 all-pairs + ConstraintSimulator.cpp:471-485 filters ----------
void Sim::broad_phase(std::vector<std::pair<int, int> >& pairs) const {
  pairs.clear();
  const int nb = (int)bodies.size();
  for (int i = 0; i < nb; i++)
    for (int j = i + 1; j < nb; j++) {
      if (!(bodies[i].enabled || bodies[j].enabled)) continue;
      if (bodies[i].shape == SHAPE_NONE || bodies[j].shape == SHAPE_NONE) continue;
      if (cparams[(size_t)i * nb + j].NK == 0) continue;  // <DisabledPair>
      pairs.push_back(std::make_pair(i, j));
    }
}

// BoxPrimitive::calc_closest_point (BoxPrimitive.cpp:788-836): point in the box frame
static double box_closest_point(const Body& box, const V3& point, V3& closest) {
  const double ext[3] = {box.dims[0] * 0.5, box.dims[1] * 0.5, box.dims[2] * 0.5};
  closest = point;
  bool inside = true;
  double sqrDist = 0.0, intDist = -INF, delta;
  for (int i = 0; i < 3; i++) {
    if (point[i] < -ext[i]) { delta = point[i] + ext[i]; closest[i] = -ext[i]; sqrDist += delta * delta; inside = false; }
    else if (point[i] > ext[i]) { delta = point[i] - ext[i]; closest[i] = ext[i]; sqrDist += delta * delta; inside = false; }
    else if (inside) { double d = -std::min(std::fabs(ext[i] - point[i]), std::fabs(point[i] + ext[i])); intDist = std::max(intDist, d); }
  }
  return inside ? intDist : std::sqrt(sqrDist);
}

static bb::Box bb_box(const Body& b) {
  bb::Box X;
  X.c = {b.x.x, b.x.y, b.x.z};
  for (int k = 0; k < 3; k++) { X.ax[k] = {b.R[k], b.R[3 + k], b.R[6 + k]}; X.ext[k] = b.dims[k]; }
  return X;
}
static inline V3 from_bb(const bb::Vec3& v) { return V3(v.x, v.y, v.z); }

// signed distance + closest points (global) with (A,B) in the order given; returns false when the pair type is unsupported
static bool signed_dist_ordered(const Body& A, const Body& B, double& dist, V3& pA, V3& pB) {
  if (A.shape == SHAPE_PLANE && B.shape == SHAPE_BOX) {            // PlanePrimitive.cpp:342-380
    double min_dist = INF; V3 pb_best, pthis;
    for (int i = 0; i < 8; i++) {
      V3 vg = to_global(B, box_vertex(B, i));
      V3 pv = to_local(A, vg);
      if (pv.y < min_dist) { min_dist = pv.y; pb_best = vg; pthis = pv; }
    }
    pthis.y = 0.0;
    dist = min_dist; pA = to_global(A, pthis); pB = pb_best;
    return true;
  }
  if (A.shape == SHAPE_PLANE && B.shape == SHAPE_SPHERE) {         // PlanePrimitive.cpp:383-411
    V3 c = to_local(A, B.x);
    V3 lowest(c.x, c.y - B.dims[0], c.z);
    V3 pthis(c.x, 0.0, c.z);
    dist = lowest.y; pA = to_global(A, pthis); pB = to_global(A, lowest);
    return true;
  }
  if (A.shape == SHAPE_SPHERE && B.shape == SHAPE_SPHERE) {        // SpherePrimitive.cpp:104-135
    V3 d = B.x - A.x;
    double len = norm(d);
    double dd = len - A.dims[0] - B.dims[0];
    V3 u = d * (1.0 / len);
    double sa = (dd > 0.0) ? A.dims[0] : A.dims[0] + dd, sb = (dd > 0.0) ? B.dims[0] : B.dims[0] + dd;
    dist = dd; pA = A.x + u * sa; pB = B.x - u * sb;
    return true;
  }
  if (A.shape == SHAPE_BOX && B.shape == SHAPE_SPHERE) {           // BoxPrimitive.cpp:257-276
    V3 c = to_local(A, B.x), pbox;
    dist = box_closest_point(A, c, pbox) - B.dims[0];
    V3 pbox_g = to_global(A, pbox);
    V3 v = pbox_g - B.x;
    double vnorm = norm(v);
    pA = pbox_g;
    pB = (vnorm == 0.0) ? B.x : B.x + v * ((B.dims[0] + std::min(dist, 0.0)) / vnorm);
    return true;
  }
  if (A.shape == SHAPE_BOX && B.shape == SHAPE_BOX) {              // rule H5 (oracle_boxbox.h) for BoxPrimitive.cpp:150-181 -> V-Clip
    bb::Vec3 a, b;
    bb::signed_dist(bb_box(A), bb_box(B), dist, a, b);
    pA = from_bb(a); pB = from_bb(b);
    return true;
  }
  return false;
}

// Spoke tip i of the rimless wheel, side s (+1: y = +W/2, -1: y = -W/2), in the wheel frame (coldet-plugin.cpp:110-115)
static V3 wheel_tip(const Body& Wh, int i, int s) {
  const double theta = M_PI * i * 2.0 / (unsigned)Wh.dims[2];
  return V3(std::cos(theta) * Wh.dims[0], s * (Wh.dims[1] * .5), std::sin(theta) * Wh.dims[0]);
}

// BladePlanePlugin::calc_signed_dist_wheel_plane (coldet-plugin.cpp:86-137): lowest spoke tip over the plane; pwheel is
// the tip, pground its projection (both returned in the global frame here).
static double wheel_plane_signed_dist(const Body& Wh, const Body& P, V3& pwheel, V3& pground) {
  double min_dist = INF;
  const int ns = (int)Wh.dims[2];
  for (int i = 0; i < ns; i++)
    for (int s = 1; s >= -1; s -= 2) {                                // p1 then p2 (:121-134); the test is strict, so p2 never wins when W = 0
      const V3 pg = to_global(Wh, wheel_tip(Wh, i, s));
      V3 pp = to_local(P, pg);
      if (pp.y < min_dist) { min_dist = pp.y; pp.y = 0.0; pground = to_global(P, pp); pwheel = pg; }
    }
  return min_dist;
}

static bool signed_dist(const Body& A, const Body& B, double& dist, V3& pA, V3& pB) {
  // coldet-plugin.cpp:324-334: BOTH argument orders hand (pA, pB) to (pwheel, pground) -- with the pair the plugin
  // queues, (ground, wheel) (:70), the point reported for the ground is the wheel's and vice versa.  Literal.
  // contact-constrained-pendulum-coldet-plugin.cpp:60-75,140-150: minus the distance between the link's anchor point and the world
  // body's origin, never positive; both argument orders hand (pA, pB) to (point on l1, point on world)
  if ((A.shape == SHAPE_PIN && B.shape == SHAPE_PINWORLD) || (A.shape == SHAPE_PINWORLD && B.shape == SHAPE_PIN)) {
    const Body& L = (A.shape == SHAPE_PIN) ? A : B; const Body& W = (A.shape == SHAPE_PIN) ? B : A;
    pA = to_global(L, V3(L.dims[0], L.dims[1], L.dims[2])); pB = W.x;
    dist = -norm(to_local(W, pA));
    return true;
  }
  if (A.shape == SHAPE_WHEEL && B.shape == SHAPE_PLANE) { dist = wheel_plane_signed_dist(A, B, pA, pB); return true; }
  if (A.shape == SHAPE_PLANE && B.shape == SHAPE_WHEEL) { dist = wheel_plane_signed_dist(B, A, pA, pB); return true; }
  if (signed_dist_ordered(A, B, dist, pA, pB)) return true;
  if (signed_dist_ordered(B, A, dist, pB, pA)) return true;        // BoxPrimitive.cpp:150-181, SpherePrimitive.cpp:282-303 swap the roles
  return false;
}

// ConstraintSimulator.cpp:450-468
void Sim::calc_pairwise_distances(const std::vector<std::pair<int, int> >& pairs, std::vector<PairDist>& out) const {
  out.clear();
  for (size_t i = 0; i < pairs.size(); i++) {
    PairDist pdi;
    pdi.a = pairs[i].first; pdi.b = pairs[i].second;
    if (!signed_dist(bodies[pdi.a], bodies[pdi.b], pdi.dist, pdi.pa, pdi.pb)) { pdi.dist = INF; }
    out.push_back(pdi);
  }
}

// Ravelin Vector3d::determine_orthonormal_basis (restated: the coordinate axis of the smallest |component| seeds the basis)
static void determine_orthonormal_basis(const V3& v1, V3& v2, V3& v3) {
  const double x = std::fabs(v1.x), y = std::fabs(v1.y), z = std::fabs(v1.z);
  V3 a;
  if (x < y) { if (x < z) a = V3(1, 0, 0); else a = V3(0, 0, 1); }
  else       { if (y < z) a = V3(0, 1, 0); else a = V3(0, 0, 1); }
  v2 = normalize(cross(v1, a));
  v3 = normalize(cross(v1, v2));
}

// CollisionDetection.cpp:57-93 (non-logging behaviour) + UnilateralConstraint.cpp:1387-1430
static Contact create_contact(int a, int b, const V3& point, const V3& normal, double violation) {
  Contact c;
  c.b1 = a; c.b2 = b; c.p = point; c.n = normal; c.dist = violation;
  determine_orthonormal_basis(c.n, c.t1, c.t2);
  return c;
}

// CCD.inl:3-82 dispatch and leaves
void Sim::find_contacts(int ia, int ib, double TOL, std::vector<Contact>& out) const {
  const Body& A = bodies[ia]; const Body& B = bodies[ib];
  // pin joint as contacts (contact-constrained-pendulum-coldet-plugin.cpp:78-110): six contacts at the midpoint of the anchor
  // point and the GLOBAL origin, normals +y -y +z -z +x -x, signed violation min(0, -p_k) for both normals of an axis; TOL ignored
  if ((A.shape == SHAPE_PIN && B.shape == SHAPE_PINWORLD) || (A.shape == SHAPE_PINWORLD && B.shape == SHAPE_PIN)) {
    const int il = (A.shape == SHAPE_PIN) ? ia : ib, iw = (A.shape == SHAPE_PIN) ? ib : ia;
    const Body& L = bodies[il];
    const V3 p = to_global(L, V3(L.dims[0], L.dims[1], L.dims[2]));
    const V3 point = (p + V3(0, 0, 0)) * 0.5;
    const double pk[3] = {p.x, p.y, p.z};
    static const int axis[6] = {1, 1, 2, 2, 0, 0}; static const double sgn[6] = {+1, -1, +1, -1, +1, -1};
    for (int k = 0; k < 6; k++) {
      V3 n(0, 0, 0); n[axis[k]] = sgn[k];
      out.push_back(create_contact(il, iw, point, n, std::min(0.0, -pk[axis[k]])));
    }
    return;
  }
  // rimless wheel / plane (coldet-plugin.cpp:222-310): one candidate per spoke tip (two when W > 0); the plugin ignores
  // the caller's TOL and tests `< sim->contact_dist_thresh` (:225,:270,:283)
  if ((A.shape == SHAPE_WHEEL && B.shape == SHAPE_PLANE) || (A.shape == SHAPE_PLANE && B.shape == SHAPE_WHEEL)) {
    const int iw = (A.shape == SHAPE_WHEEL) ? ia : ib, ip = (A.shape == SHAPE_WHEEL) ? ib : ia;
    const Body& Wh = bodies[iw]; const Body& P = bodies[ip];
    const V3 n = rot(P.R, V3(0, 1, 0));
    const int ns = (int)Wh.dims[2];
    for (int i = 0; i < ns; i++)
      for (int s = 1; s >= -1; s -= 2) {
        if (s < 0 && !(Wh.dims[1] > 0.0)) continue;                   // :283
        const V3 pg = to_global(Wh, wheel_tip(Wh, i, s));
        V3 pp = to_local(P, pg);
        const double h = pp.y;
        if (!(h < contact_dist_thresh)) continue;
        pp.y = 0.0;
        out.push_back(create_contact(iw, ip, (pg + to_global(P, pp)) * 0.5, n, h));
      }
    return;
  }
  // sphere / plane (CCD.inl:805-846): cgA = sphere, cgB = plane
  if ((A.shape == SHAPE_SPHERE && B.shape == SHAPE_PLANE) || (A.shape == SHAPE_PLANE && B.shape == SHAPE_SPHERE)) {
    const int is = (A.shape == SHAPE_SPHERE) ? ia : ib, ip = (A.shape == SHAPE_SPHERE) ? ib : ia;
    const Body& S = bodies[is]; const Body& P = bodies[ip];
    V3 c = to_local(P, S.x);
    double dist = c.y - S.dims[0];
    if (dist > TOL) return;
    V3 p(c.x, 0.5 * (c.y - S.dims[0]), c.z);
    V3 n = rot(P.R, V3(0, 1, 0));
    out.push_back(create_contact(is, ip, to_global(P, p), n, dist));
    return;
  }
  // plane / box (CCD.inl:850-886): cgA = plane, cgB = box; one contact per vertex within TOL, normal = -(plane +Y)
  if ((A.shape == SHAPE_BOX && B.shape == SHAPE_PLANE) || (A.shape == SHAPE_PLANE && B.shape == SHAPE_BOX)) {
    const int ix = (A.shape == SHAPE_BOX) ? ia : ib, ip = (A.shape == SHAPE_BOX) ? ib : ia;
    const Body& X = bodies[ix]; const Body& P = bodies[ip];
    V3 n = rot(P.R, V3(0, 1, 0));
    for (int i = 0; i < 8; i++) {
      V3 vg = to_global(X, box_vertex(X, i));
      double dist = to_local(P, vg).y;                             // PlanePrimitive::calc_dist_and_normal :477-492
      if (dist <= TOL) out.push_back(create_contact(ip, ix, vg, -n, dist));
    }
    return;
  }
  // sphere / sphere (CCD.inl:1165-1206)
  if (A.shape == SHAPE_SPHERE && B.shape == SHAPE_SPHERE) {
    V3 d = A.x - B.x;
    double dist = norm(d) - A.dims[0] - B.dims[0];
    if (dist > TOL) return;
    V3 n = normalize(d);
    V3 closest_A = A.x - n * A.dims[0], closest_B = B.x + n * B.dims[0];
    out.push_back(create_contact(ia, ib, (closest_A + closest_B) * 0.5, n, dist));
    return;
  }
  // box / sphere (CCD.inl:1210-1259 + BoxPrimitive.cpp:184-254): cgA = box, cgB = sphere
  if ((A.shape == SHAPE_BOX && B.shape == SHAPE_SPHERE) || (A.shape == SHAPE_SPHERE && B.shape == SHAPE_BOX)) {
    const int ix = (A.shape == SHAPE_BOX) ? ia : ib, is = (A.shape == SHAPE_BOX) ? ib : ia;
    const Body& X = bodies[ix]; const Body& S = bodies[is];
    const double HX = X.dims[0] * 0.5, HY = X.dims[1] * 0.5, HZ = X.dims[2] * 0.5, Rr = S.dims[0];
    V3 c = to_local(X, S.x);
    // QP::qp_gradproj of 1/2|p-c|^2 over the box == clamp
    V3 pbox(std::min(std::max(c.x, -HX), HX), std::min(std::max(c.y, -HY), HY), std::min(std::max(c.z, -HZ), HZ));
    V3 pbox_g = to_global(X, pbox);
    V3 psph = rotT(S.R, pbox_g - S.x);                              // closest box point in the sphere frame
    double psph_nrm = norm(psph), dist;
    V3 psph_g;
    if (std::fabs(pbox.x) < HX || std::fabs(pbox.y) < HY || std::fabs(pbox.z) < HZ || psph_nrm < Rr) {
      double box_dist = std::min(HX - std::fabs(pbox.x), std::min(HY - std::fabs(pbox.y), HZ - std::fabs(pbox.z)));
      dist = -std::min(box_dist, Rr - psph_nrm);
    } else {
      psph = psph * (Rr / psph_nrm);
      dist = norm(to_local(X, to_global(S, psph)) - pbox);
    }
    if (dist > TOL) return;
    psph_g = to_global(S, psph);
    V3 p, normal;
    if (dist > 0.0) {
      p = (psph_g + pbox_g) * 0.5;
      normal = pbox_g - psph_g;
      double nrm = norm(normal);
      if (nrm > NEAR_ZERO) normal = normal * (1.0 / nrm);
      else normal = normalize(rot(S.R, psph));
    } else {
      p = psph_g;
      normal = normalize(rot(S.R, psph));
    }
    out.push_back(create_contact(ix, is, p, normal, dist));
    return;
  }
  // box / box (CCD.inl:86-494), rule H5
  if (A.shape == SHAPE_BOX && B.shape == SHAPE_BOX) {
    bb::Vec3 n;
    const std::vector<bb::Point> pts = bb::contacts(bb_box(A), bb_box(B), TOL, n);
    for (const bb::Point& c : pts) out.push_back(create_contact(ia, ib, from_bb(c.p), from_bb(n), c.violation));
    return;
  }
}

// ConstraintSimulator.cpp:488-537 + preprocess_constraint :390-417
void Sim::find_unilateral_constraints(const std::vector<PairDist>& pd, std::vector<Contact>& out) const {
  out.clear();
  const int nb = (int)bodies.size();
  for (size_t i = 0; i < pd.size(); i++)
    if (pd[i].dist < contact_dist_thresh) find_contacts(pd[i].a, pd[i].b, contact_dist_thresh, out);
  for (size_t i = 0; i < out.size(); i++) {
    int lo = std::min(out[i].b1, out[i].b2), hi = std::max(out[i].b1, out[i].b2);
    out[i].cp = cparams[(size_t)lo * nb + hi];
  }
}

// UnilateralConstraint.cpp:695-747 / :1360-1384
double Sim::calc_constraint_vel(const Contact& c) const {
  V3 ta = point_vel(bodies[c.b1], c.p), tb = point_vel(bodies[c.b2], c.p);
  return dot(c.n, ta - tb);
}

// CCD::calc_max_dist (CCD.cpp:585-607).  Literal: the body velocity is first transformed to the GLOBAL
// frame, so its linear part is the velocity of the body-fixed point at the global origin.
static double calc_max_dist(const Body& rb, const V3& n, double rmax) {
  if (!rb.enabled) return 0.0;
  V3 xd0 = rb.vl - cross(rb.va, rb.x);
  return dot(n, xd0) + norm(cross(rb.va, n)) * rmax;
}

// bounding radius used by CA (CCD.cpp:1023-1101)
static double calc_rmax(const Body& b) {
  if (b.shape == SHAPE_SPHERE) return b.dims[0];
  if (b.shape == SHAPE_BOX) return std::sqrt((b.dims[0] / 2.0) * (b.dims[0] / 2.0) + (b.dims[1] / 2.0) * (b.dims[1] / 2.0) + (b.dims[2] / 2.0) * (b.dims[2] / 2.0));
  // SHAPE_WHEEL: 0.  The plugin takes the wheel out of the body list before CCD::broad_phase (coldet-plugin.cpp:53-67),
  // the only place _rmax is filled (CCD.cpp:739), so _rmax[wheel_cg] is the map's default-constructed 0.0.
  return 0.0;
}

// CompGeom::collinear (CompGeom.cpp:1923-1931, CompGeom.h:110)
static bool rel_equal(double x, double y, double tol = NEAR_ZERO) { return std::fabs(x - y) <= tol * std::max(std::fabs(x), std::max(std::fabs(y), 1.0)); }
static bool collinear(const V3& a, const V3& b, const V3& c) {
  return rel_equal((c.z - a.z) * (b.y - a.y), (b.z - a.z) * (c.y - a.y)) &&
         rel_equal((b.z - a.z) * (c.x - a.x), (b.x - a.x) * (c.z - a.z)) &&
         rel_equal((b.x - a.x) * (c.y - a.y), (b.y - a.y) * (c.x - a.x));
}

// CCD::calc_next_CA_Euler_step_polyhedron_plane (CCD.cpp:407-460); box = polyhedron, rv = relative velocity at the box pose
static double next_CA_box_plane(const Body& box, const V3& rv_lin_boxframe, const V3& rv_ang_boxframe, const V3& normal, double offset0) {
  double max_step = INF;
  V3 nP = rotT(box.R, normal);
  V3 p0 = normal * offset0;
  const double offset = dot(nP, to_local(box, p0));
  double av_norm = norm(rv_ang_boxframe);
  double lv_dot_n = -dot(nP, rv_lin_boxframe);
  for (int i = 0; i < 8; i++) {
    V3 vtx = box_vertex(box, i);
    double r = norm(vtx);
    double dist = dot(nP, vtx) - offset;
    if (dist < NEAR_ZERO) continue;
    double speed = std::max(0.0, lv_dot_n + av_norm * r);
    max_step = std::min(max_step, dist / speed);
  }
  return max_step;
}

// CCD::calc_next_CA_Euler_step_polyhedron_polyhedron (CCD.cpp:468-541) with both polyhedra boxes
static double next_CA_box_box(const Body& A, const Body& B, const V3& rvA_lin, const V3& rvA_ang, const V3& rvB_lin, const V3& rvB_ang,
                              const V3& n0, double offset0) {
  double max_step = INF;
  const V3 nA = rotT(A.R, n0), nB = rotT(B.R, -n0);                   // :474-475
  const V3 p0 = n0 * offset0;
  const double offsetA = dot(nA, to_local(A, p0)), offsetB = dot(nB, to_local(B, p0));   // :478-480
  const double avA_norm = norm(rvA_ang), avB_norm = norm(rvB_ang);
  const double lvA_dot_n = -dot(nA, rvA_lin), lvB_dot_n = dot(nB, rvB_lin);              // :493-494
  for (int i = 0; i < 8; i++) {                                       // :497-516
    const V3 vertex = box_vertex(A, i);
    const double r = norm(vertex), dist = dot(nA, vertex) - offsetA;
    if (dist < NEAR_ZERO) continue;
    const double speed = std::max(0.0, lvA_dot_n + avA_norm * r);
    max_step = std::min(max_step, dist / speed);
  }
  for (int i = 0; i < 8; i++) {                                       // :519-538
    const V3 vertex = box_vertex(B, i);
    const double r = norm(vertex), dist = dot(nB, vertex) - offsetB;
    if (dist < NEAR_ZERO) continue;
    const double speed = std::max(0.0, lvB_dot_n + avB_norm * r);
    max_step = std::min(max_step, dist / speed);
  }
  return max_step;
}

// CCD.cpp:122-400
double Sim::calc_CA_Euler_step(const PairDist& pdi) const {
  const Body& A = bodies[pdi.a]; const Body& B = bodies[pdi.b];
  if (pdi.dist == INF) return INF;
  // :138-166 sphere special case
  if (A.shape == SHAPE_SPHERE || B.shape == SHAPE_SPHERE) {
    if (!(pdi.dist > NEAR_ZERO)) {
      std::vector<Contact> contacts;
      find_contacts(pdi.a, pdi.b, NEAR_ZERO, contacts);
      if (contacts.size() == 1 && std::fabs(calc_constraint_vel(contacts.front())) < NEAR_ZERO * 10) return INF;
    }
  }
  // :169-235 generic
  if (pdi.dist <= 0.0 && (A.shape == SHAPE_WHEEL || B.shape == SHAPE_WHEEL || A.shape == SHAPE_PIN || B.shape == SHAPE_PIN)) return INF;   // both plugins override calc_next_CA_Euler_step (coldet-plugin.cpp:214-217, contact-constrained-pendulum-coldet-plugin.cpp:152-155)
  if (pdi.dist <= 0.0) {
    // :238-400 bodies in contact
    std::vector<Contact> contacts;
    find_contacts(pdi.a, pdi.b, NEAR_ZERO, contacts);               // CCD.h:45 default TOL
    if (contacts.empty()) return INF;
    const Contact& c = contacts.front();
    double d = dot(c.n, c.p);
    for (size_t i = 0; i < contacts.size(); i++)
      if (calc_constraint_vel(contacts[i]) < -NEAR_ZERO) return 0.0;  // :272-284
    if (contacts.size() >= 3) {                                      // :288-330 (H8: always tests points 0,1,2)
      bool twosimplex = false;
      for (size_t i = 2; i < contacts.size(); i++)
        if (!collinear(contacts[0].p, contacts[1].p, contacts[2].p)) { twosimplex = true; break; }
      if (twosimplex) return INF;
    }
    const Body& gA = bodies[c.b1]; const Body& gB = bodies[c.b2];   // :333-399
    if (gA.shape == SHAPE_BOX && gB.shape == SHAPE_PLANE) {
      V3 rl = rotT(gA.R, gA.enabled ? gA.vl : V3()) - rotT(gA.R, point_vel(gB, gA.x));
      V3 ra = rotT(gA.R, (gA.enabled ? gA.va : V3()) - (gB.enabled ? gB.va : V3()));
      return next_CA_box_plane(gA, rl, ra, c.n, d);
    }
    if (gA.shape == SHAPE_PLANE && gB.shape == SHAPE_BOX) {
      V3 rl = rotT(gB.R, point_vel(gA, gB.x)) - rotT(gB.R, gB.enabled ? gB.vl : V3());
      V3 ra = rotT(gB.R, (gA.enabled ? gA.va : V3()) - (gB.enabled ? gB.va : V3()));
      return next_CA_box_plane(gB, -rl, -ra, -c.n, -d);
    }
    if (gA.shape == SHAPE_BOX && gB.shape == SHAPE_BOX) {            // :350-364 -> :468-541
      const V3 wrel = (gA.enabled ? gA.va : V3()) - (gB.enabled ? gB.va : V3());
      const V3 rvA_lin = rotT(gA.R, (gA.enabled ? gA.vl : V3()) - point_vel(gB, gA.x));
      const V3 rvB_lin = rotT(gB.R, point_vel(gA, gB.x) - (gB.enabled ? gB.vl : V3()));
      return next_CA_box_box(gA, gB, rvA_lin, rotT(gA.R, wrel), rvB_lin, rotT(gB.R, wrel), c.n, d);
    }
    return INF;                                                      // :397-399
  }
  V3 d0 = pdi.pa - pdi.pb;                                           // :193-202
  double d0_norm = norm(d0);
  V3 n0 = d0 * (1.0 / d0_norm);
  double dist_per_tA = calc_max_dist(A, -n0, calc_rmax(A));          // :214-217
  double dist_per_tB = calc_max_dist(B, n0, calc_rmax(B));
  double total = dist_per_tA + dist_per_tB;
  if (total < 0.0) total = 0.0;
  return std::min(INF, pdi.dist / total);                            // :229
}

// Simulator::precalc_fwd_dyn + calc_fwd_dyn (Simulator.cpp:319-350,482-602) for free bodies, then
// v += h*a (TimeSteppingSimulator.cpp:181-192).  Ravelin RigidBodyd::calc_fwd_dyn restated:
// a_lin = f/m, alpha = J^-1 (tau - w x J w) with J = R diag(Jb) R^T at the COM (global-aligned frame).
void Sim::calc_fwd_dyn_and_integrate_velocity(double h) {
  if (has_rc) {   // Simulator.cpp:339-348 (controller at current_time), :544-553 (RCArticulatedBodyd::calc_fwd_dyn), TimeSteppingSimulator.cpp:181-192
    const int nd = rc.ndof();
    Vec tau(nd), qdd(nd);
    for (int k = 0; k < nd; k++) {
      double u = jtau.empty() ? 0.0 : jtau[k];
      if (has_ctrl) {
        const double ph = current_time * ctrl_freq[k];
        u += ctrl_kp[k] * (std::sin(ph) * ctrl_amp[k] - jq[k]) + ctrl_kv[k] * (std::cos(ph) * ctrl_amp[k] - jqd[k]);
      }
      tau[k] = u;
    }
    const double g[3] = {gravity.x, gravity.y, gravity.z};
    if (rc_fdyn == 1) rc_crb_fwd_dyn(rc, jq.data(), jqd.data(), tau.data(), g, qdd.data());
    else rc_aba(rc, jq.data(), jqd.data(), tau.data(), g, qdd.data());
    for (int k = 0; k < nd; k++) jqd[k] = jqd[k] + qdd[k] * h;
    rc_update_links();
  }
  for (size_t i = 0; i < bodies.size(); i++) {
    Body& b = bodies[i];
    if (!b.enabled || is_link((int)i)) continue;
    V3 f = gravity * b.mass + b.fext;                                // GravityForce.cpp:32-48
    V3 tau = b.text;
    V3 wb = rotT(b.R, b.va);
    V3 Jw = rot(b.R, V3(b.J[0] * wb.x, b.J[1] * wb.y, b.J[2] * wb.z));
    V3 rhs = tau - cross(b.va, Jw);
    V3 rb = rotT(b.R, rhs);
    V3 alpha = rot(b.R, V3(rb.x / b.J[0], rb.y / b.J[1], rb.z / b.J[2]));
    V3 a = f * (1.0 / b.mass);
    b.vl = b.vl + a * h;
    b.va = b.va + alpha * h;
  }
}

// TimeSteppingSimulator::do_mini_step (TimeSteppingSimulator.cpp:114-222)
double Sim::do_mini_step(double dt) {
  mini_failed = false;
  const size_t nb = bodies.size();
  std::vector<V3> xsave(nb);
  std::vector<double> qsave(nb * 4);
  for (size_t i = 0; i < nb; i++) { xsave[i] = bodies[i].x; for (int k = 0; k < 4; k++) qsave[i * 4 + k] = bodies[i].quat[k]; }
  const Vec jqsave = jq;
  double h = 0.0;
  std::vector<std::pair<int, int> > pairs;
  std::vector<PairDist> pd;
  while (h < dt) {                                                   // :133-168
    cnt.ca_iterations++;
    broad_phase(pairs);
    calc_pairwise_distances(pairs, pd);
    double CA_step = INF;                                            // :272-331 (no joints here)
    for (size_t i = 0; i < pd.size(); i++) CA_step = std::min(CA_step, calc_CA_Euler_step(pd[i]));
    if (CA_step <= 0.0) break;
    double tc = std::max(min_step_size, CA_step);
    tc = std::min(dt - h, tc);
    if (has_rc) {                                                    // joint coordinates: q = qsave + (h + tc) qd
      for (size_t k = 0; k < jq.size(); k++) jq[k] = jqd[k] * (h + tc) + jqsave[k];
      rc_update_links();
    }
    for (size_t i = 0; i < nb; i++) {                                // :156-164
      Body& b = bodies[i];
      if (!b.enabled || is_link((int)i)) continue;
      const double s = h + tc;
      const double qx = qsave[i * 4 + 0], qy = qsave[i * 4 + 1], qz = qsave[i * 4 + 2], qw = qsave[i * 4 + 3];
      const V3& w = b.va;
      // Ravelin Quatd::deriv(q, w): qd = 1/2 (0,w) * q
      const double dw = 0.5 * (-qx * w.x - qy * w.y - qz * w.z);
      const double dx = 0.5 * (+qw * w.x + qz * w.y - qy * w.z);
      const double dy = 0.5 * (-qz * w.x + qw * w.y + qx * w.z);
      const double dz = 0.5 * (+qy * w.x - qx * w.y + qw * w.z);
      b.x = V3(b.vl.x * s + xsave[i].x, b.vl.y * s + xsave[i].y, b.vl.z * s + xsave[i].z);
      double nx = dx * s + qx, ny = dy * s + qy, nz = dz * s + qz, nw = dw * s + qw;
      const double nrm = std::sqrt(nx * nx + ny * ny + nz * nz + nw * nw);   // set_generalized_coordinates_euler normalises
      b.quat[0] = nx / nrm; b.quat[1] = ny / nrm; b.quat[2] = nz / nrm; b.quat[3] = nw / nrm;
      update_pose(b);
    }
    h += tc;
  }
  calc_fwd_dyn_and_integrate_velocity(h);                            // :173-192
  broad_phase(pairs);                                                // pairs are unchanged (all-pairs table)
  calc_pairwise_distances(pairs, pd);                                // :206
  std::vector<Contact> contacts;
  find_unilateral_constraints(pd, contacts);                         // :209
  cnt.contacts += (long long)contacts.size();
  process_constraints(contacts);                                     // :212 -> ConstraintSimulator.cpp:298-355
  current_time += h;
  cnt.mini_steps++;
  return h;
}

// TimeSteppingSimulator::step + step_si_Euler (TimeSteppingSimulator.cpp:52-111,433-455)
double Sim::step(double dt) {
  double h = 0.0;
  int stalled = 0;
  while (h < dt) {
    const double hh = do_mini_step(dt - h);
    h += hh;
    // the reference would spin (or leave through LCPSolverException) when an impact cannot be resolved; the batch
    // contract gives up the rest of the step after 64 zero-length mini-steps in a row and counts a failure
    stalled = (hh > 0.0) ? 0 : stalled + 1;
    if (stalled >= 64) { cnt.lcp_failures++; break; }
    if (hh == 0.0 && mini_failed) break;   // LCPSolverException in the reference (ImpactConstraintHandlerQP.cpp:224): the env gives up this step
  }
  stabilize();                             // TimeSteppingSimulator.cpp:95-98
  cnt.env_steps++;
  return dt;
}

// ---------- impact handling ----------
namespace {
struct ProblemData {
  int nc = 0, ngc = 0;
  std::vector<int> sb;            // super bodies (enabled), ascending scene index
  std::vector<int> gc;            // gc offset per scene body (-1 if absent)
  std::vector<Contact*> cons;
  // Jacobian rows as two 1x6 blocks per contact and direction: [dir][contact][block][6]
  std::vector<double> Jr;         // 3 * nc * 2 * 6
  Mat X;                          // ngc x ngc
  Mat XT[3];                      // X_CnT, X_CsT, X_CtT : ngc x nc
  Mat D[3][3];                    // Cd1_X_Cd2T, upper triangle used: nn ns nt ss st tt
  Vec Cv[3];                      // Cn_v, Cs_v, Ct_v
  Vec cn, cs, ct;
  double* jrow(int d, int i, int blk) { return &Jr[(((size_t)d * nc + i) * 2 + blk) * 6]; }
  // scenes with an articulated body: dense rows Cn, Cs, Ct over the island's generalized coordinates (nc x ngc each)
  bool dense = false;
  Mat C[3];
};
}  // namespace

// compute_problem_data with an RCArticulatedBody in the island (ImpactConstraintHandler.cpp:1898-2166): super bodies are
// the enabled free bodies and the articulated body (all moving links share its joint coordinates, :1905-1916);
// X = blockdiag(inverse_SPD(generalized inertia)) with the articulated body's inertia from get_generalized_inertia
// (:1599-1611); a contact wrench [d, r x d] on a link is post-multiplied by the link Jacobian (:1869-1878).
static void compute_problem_data_rc(Sim& S, ProblemData& q, const std::vector<Contact*>& cons, const std::vector<int>& island_bodies) {
  const int nb = (int)S.bodies.size();
  const int rep = S.rc_first + 1, nd = S.rc.ndof();
  q.dense = true;
  q.cons = cons;
  q.nc = (int)cons.size();
  q.sb.clear();
  for (size_t i = 0; i < island_bodies.size(); i++) if (S.bodies[island_bodies[i]].enabled) q.sb.push_back(S.super_of(island_bodies[i]));
  std::sort(q.sb.begin(), q.sb.end());                               // :1915 (pointer order -> scene order, H4)
  q.sb.erase(std::unique(q.sb.begin(), q.sb.end()), q.sb.end());
  q.gc.assign(nb, -1);
  q.ngc = 0;
  for (size_t i = 0; i < q.sb.size(); i++) { q.gc[q.sb[i]] = q.ngc; q.ngc += (q.sb[i] == rep) ? nd : 6; }
  q.X = Mat(q.ngc, q.ngc);
  RCKin kin;
  rc_kinematics(S.rc, S.jq.data(), S.jqd.data(), kin);
  Vec v(q.ngc, 0.0);                                                 // get_generalized_velocity :1801-1814
  for (size_t i = 0; i < q.sb.size(); i++) {
    const int g = q.gc[q.sb[i]];
    if (q.sb[i] == rep) {
      Vec H((size_t)nd * nd);
      rc_crb(S.rc, kin, H.data());
      inverse_SPD(H.data(), nd);
      for (int r = 0; r < nd; r++) for (int c = 0; c < nd; c++) q.X(g + r, g + c) = H[(size_t)c * nd + r];
      for (int k = 0; k < nd; k++) v[g + k] = S.jqd[k];
      continue;
    }
    const Body& b = S.bodies[q.sb[i]];
    double Mg[36] = {0};
    for (int k = 0; k < 3; k++) Mg[k * 6 + k] = b.mass;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += b.R[r * 3 + k] * b.J[k] * b.R[c * 3 + k];
        Mg[(3 + c) * 6 + (3 + r)] = s;
      }
    inverse_SPD(Mg, 6);
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) q.X(g + r, g + c) = Mg[c * 6 + r];
    v[g] = b.vl.x; v[g + 1] = b.vl.y; v[g + 2] = b.vl.z; v[g + 3] = b.va.x; v[g + 4] = b.va.y; v[g + 5] = b.va.z;
  }
  // contact Jacobians (:1817-1895)
  std::vector<std::vector<double> > Jl(S.rc.n_links);               // link Jacobians, computed on demand
  for (int d = 0; d < 3; d++) q.C[d] = Mat(q.nc, q.ngc);
  for (int i = 0; i < q.nc; i++) {
    const Contact& c = *cons[i];
    const V3 dirs[3] = {c.n, c.t1, c.t2};
    for (int blk = 0; blk < 2; blk++) {
      const int bi = blk == 0 ? c.b1 : c.b2;
      const Body& b = S.bodies[bi];
      if (!b.enabled) continue;
      for (int d = 0; d < 3; d++) {
        const V3 dd = blk == 0 ? dirs[d] : -dirs[d];
        const V3 rxd = cross(c.p - b.x, dd);
        const double w[6] = {dd.x, dd.y, dd.z, rxd.x, rxd.y, rxd.z};
        if (S.is_link(bi)) {
          const int li = bi - S.rc_first;
          if (Jl[li].empty()) { Jl[li].assign((size_t)6 * nd, 0.0); rc_link_jacobian(S.rc, kin, li, Jl[li].data()); }
          const int g = q.gc[rep];
          for (int k = 0; k < nd; k++) {
            double s = 0.0;
            for (int r = 0; r < 6; r++) s += w[r] * Jl[li][(size_t)r * nd + k];
            q.C[d](i, g + k) += s;
          }
        } else {
          const int g = q.gc[bi];
          for (int r = 0; r < 6; r++) q.C[d](i, g + r) += w[r];
        }
      }
    }
  }
  for (int d = 0; d < 3; d++) {                                      // X_CdT = (Cd X)^T (:2125-2127)
    q.XT[d] = Mat(q.ngc, q.nc);
    for (int i = 0; i < q.nc; i++)
      for (int k = 0; k < q.ngc; k++) {
        double s = 0.0;
        for (int kk = 0; kk < q.ngc; kk++) s = std::fma(q.C[d](i, kk), q.X(kk, k), s);
        q.XT[d](k, i) = s;
      }
  }
  for (int d1 = 0; d1 < 3; d1++)                                     // Delassus blocks (:2133-2149)
    for (int d2 = d1; d2 < 3; d2++) {
      q.D[d1][d2] = Mat(q.nc, q.nc);
      for (int i = 0; i < q.nc; i++)
        for (int j = 0; j < q.nc; j++) {
          double s = 0.0;
          for (int k = 0; k < q.ngc; k++) s = std::fma(q.C[d1](i, k), q.XT[d2](k, j), s);
          q.D[d1][d2](i, j) = s;
        }
    }
  for (int d = 0; d < 3; d++) {                                      // Cd v (:2157-2159)
    q.Cv[d].assign(q.nc, 0.0);
    for (int i = 0; i < q.nc; i++) {
      double s = 0.0;
      for (int k = 0; k < q.ngc; k++) s = std::fma(q.C[d](i, k), v[k], s);
      q.Cv[d][i] = s;
    }
  }
  q.cn.assign(q.nc, 0.0); q.cs.assign(q.nc, 0.0); q.ct.assign(q.nc, 0.0);
}

// ImpactConstraintHandler::compute_problem_data (ImpactConstraintHandler.cpp:1898-2166), free bodies only
static void compute_problem_data(Sim& S, ProblemData& q, const std::vector<Contact*>& cons, const std::vector<int>& island_bodies) {
  if (S.has_rc) { compute_problem_data_rc(S, q, cons, island_bodies); return; }
  const int nb = (int)S.bodies.size();
  q.cons = cons;
  q.nc = (int)cons.size();
  q.sb.clear();
  for (size_t i = 0; i < island_bodies.size(); i++) if (S.bodies[island_bodies[i]].enabled) q.sb.push_back(island_bodies[i]);
  std::sort(q.sb.begin(), q.sb.end());                               // :1915 (pointer order -> scene order, H4)
  q.sb.erase(std::unique(q.sb.begin(), q.sb.end()), q.sb.end());
  q.gc.assign(nb, -1);
  q.ngc = 0;
  for (size_t i = 0; i < q.sb.size(); i++) { q.gc[q.sb[i]] = q.ngc; q.ngc += 6; }
  // compute_X (:1590-1695) with no bilateral constraints: X = blockdiag(inverse_SPD(generalized inertia))
  q.X = Mat(q.ngc, q.ngc);
  for (size_t i = 0; i < q.sb.size(); i++) {
    const Body& b = S.bodies[q.sb[i]];
    double Mg[36] = {0};
    for (int k = 0; k < 3; k++) Mg[k * 6 + k] = b.mass;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += b.R[r * 3 + k] * b.J[k] * b.R[c * 3 + k];
        Mg[(3 + c) * 6 + (3 + r)] = s;
      }
    inverse_SPD(Mg, 6);
    const int g = q.gc[q.sb[i]];
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) q.X(g + r, g + c) = Mg[c * 6 + r];
  }
  // contact Jacobian rows (:1817-1895): [d, r x d], +d on body 1, -d on body 2
  q.Jr.assign((size_t)3 * q.nc * 2 * 6, 0.0);
  for (int i = 0; i < q.nc; i++) {
    const Contact& c = *cons[i];
    const V3 dirs[3] = {c.n, c.t1, c.t2};
    for (int d = 0; d < 3; d++)
      for (int blk = 0; blk < 2; blk++) {
        const int bi = blk == 0 ? c.b1 : c.b2;
        const Body& b = S.bodies[bi];
        if (!b.enabled) continue;
        V3 dd = blk == 0 ? dirs[d] : -dirs[d];
        V3 rxd = cross(c.p - b.x, dd);
        double* row = q.jrow(d, i, blk);
        row[0] = dd.x; row[1] = dd.y; row[2] = dd.z; row[3] = rxd.x; row[4] = rxd.y; row[5] = rxd.z;
      }
  }
  // X_CdT = (Cd * X)^T (:2125-2127)
  for (int d = 0; d < 3; d++) {
    q.XT[d] = Mat(q.ngc, q.nc);
    for (int i = 0; i < q.nc; i++)
      for (int blk = 0; blk < 2; blk++) {
        const int bi = blk == 0 ? cons[i]->b1 : cons[i]->b2;
        if (q.gc[bi] < 0) continue;
        const int g = q.gc[bi];
        const double* row = q.jrow(d, i, blk);
        for (int k = 0; k < 6; k++) {
          double s = 0.0;
          for (int kk = 0; kk < 6; kk++) s = std::fma(row[kk], q.X(g + kk, g + k), s);
          q.XT[d](g + k, i) += s;
        }
      }
  }
  // Delassus blocks (:2133-2149): Cd1 * X_Cd2T
  for (int d1 = 0; d1 < 3; d1++)
    for (int d2 = d1; d2 < 3; d2++) {
      q.D[d1][d2] = Mat(q.nc, q.nc);
      for (int i = 0; i < q.nc; i++)
        for (int j = 0; j < q.nc; j++) {
          double s = 0.0;
          for (int blk = 0; blk < 2; blk++) {
            const int bi = blk == 0 ? cons[i]->b1 : cons[i]->b2;
            if (q.gc[bi] < 0) continue;
            const int g = q.gc[bi];
            const double* row = q.jrow(d1, i, blk);
            for (int k = 0; k < 6; k++) s = std::fma(row[k], q.XT[d2](g + k, j), s);
          }
          q.D[d1][d2](i, j) = s;
        }
    }
  // Cd * v (:2157-2159)
  for (int d = 0; d < 3; d++) {
    q.Cv[d].assign(q.nc, 0.0);
    for (int i = 0; i < q.nc; i++) {
      double s = 0.0;
      for (int blk = 0; blk < 2; blk++) {
        const int bi = blk == 0 ? cons[i]->b1 : cons[i]->b2;
        if (q.gc[bi] < 0) continue;
        const Body& b = S.bodies[bi];
        const double v[6] = {b.vl.x, b.vl.y, b.vl.z, b.va.x, b.va.y, b.va.z};
        const double* row = q.jrow(d, i, blk);
        for (int k = 0; k < 6; k++) s = std::fma(row[k], v[k], s);
      }
      q.Cv[d][i] = s;
    }
  }
  q.cn.assign(q.nc, 0.0); q.cs.assign(q.nc, 0.0); q.ct.assign(q.nc, 0.0);
}

static inline double Dn(const ProblemData& q, int d1, int d2, int i, int j) {   // block (d1,d2) incl. the transposed lower ones
  return (d1 <= d2) ? q.D[d1][d2](i, j) : q.D[d2][d1](j, i);
}

// setup_QP + the LCP wrap of solve_qp_work (ImpactConstraintHandlerQP.cpp:129-148,216,271-497).  nl = 0.
static void build_qp_lcp(const ProblemData& q, int& n, Vec& MM, Vec& qq) {
  const int nc = q.nc;
  const int NVARS = 5 * nc;
  int NK_TOTAL = 0;
  for (int i = 0; i < nc; i++) NK_TOTAL += q.cons[i]->cp.NK / 2;      // ImpactConstraintHandler.cpp:1984-1993
  n = NVARS + nc + NK_TOTAL;
  MM.assign((size_t)n * n, 0.0);
  qq.assign(n, 0.0);
  auto at = [&](int r, int c) -> double& { return MM[(size_t)c * n + r]; };
  // H: variable blocks [cn, cs+, ct+, cs-, ct-] -> (direction, sign)
  const int dir[5] = {0, 1, 2, 1, 2};
  const double sg[5] = {1, 1, 1, -1, -1};
  for (int br = 0; br < 5; br++)
    for (int bc = 0; bc < 5; bc++)
      for (int i = 0; i < nc; i++)
        for (int j = 0; j < nc; j++) {
          double v = Dn(q, dir[br], dir[bc], i, j);
          at(br * nc + i, bc * nc + j) = (sg[br] * sg[bc] < 0) ? -v : v;   // :383-411
        }
  for (int i = 0; i < nc; i++) at(i, i) += q.cons[i]->cp.compliance;    // :437-440
  for (int i = 0; i < nc; i++) {                                        // :429-435
    qq[i] = q.Cv[0][i]; qq[nc + i] = q.Cv[1][i]; qq[2 * nc + i] = q.Cv[2][i];
    qq[3 * nc + i] = -q.Cv[1][i]; qq[4 * nc + i] = -q.Cv[2][i];
  }
  // A rows (the "M" block): normal rows = top of H (:445-447), rhs Cn_v
  for (int i = 0; i < nc; i++) {
    for (int c = 0; c < NVARS; c++) at(NVARS + i, c) = at(i, c);
    qq[NVARS + i] = q.Cv[0][i];                                         // after the two negations (:496, :216)
  }
  int row = NVARS + nc;                                                 // :456-479
  for (int i = 0; i < nc; i++) {
    const ContactParams& cp = q.cons[i]->cp;
    const double vel = std::sqrt(q.Cv[1][i] * q.Cv[1][i] + q.Cv[2][i] * q.Cv[2][i]);
    const int half = cp.NK / 2;
    for (int j = 0; j < half; j++) {
      const double theta = (double)j / (half - 1) * M_PI_2;
      const double ct = std::cos(theta), st = std::sin(theta);
      at(row, i) = cp.mu_c;
      at(row, nc + i) = -ct; at(row, 3 * nc + i) = -ct;
      at(row, 2 * nc + i) = -st; at(row, 4 * nc + i) = -st;
      qq[row] = cp.mu_v * vel;
      row++;
    }
  }
  // MT = -M' (:145-148)
  for (int r = NVARS; r < n; r++)
    for (int c = 0; c < NVARS; c++) at(c, r) = -at(r, c);
}

// apply_ap_model assembly (ImpactConstraintHandlerLCP.cpp:94-310).  nl = 0.
static void build_ap_lcp(const ProblemData& q, int& n, Vec& MM, Vec& qq) {
  const int NC = q.nc;
  const int N_CONST = 5 * NC;
  int NK_DIRS = 0;
  for (int i = 0; i < NC; i++) NK_DIRS += (q.cons[i]->cp.NK > 4) ? (q.cons[i]->cp.NK + 4) / 4 : 1;   // :117-124
  n = N_CONST + NK_DIRS;
  MM.assign((size_t)n * n, 0.0);
  qq.assign(n, 0.0);
  auto at = [&](int r, int c) -> double& { return MM[(size_t)c * n + r]; };
  // UL: blocks ordered [n, s+, s-, t+, t-] (:170-244)
  const int dir[5] = {0, 1, 1, 2, 2};
  const double sg[5] = {1, 1, -1, 1, -1};
  for (int br = 0; br < 5; br++)
    for (int bc = 0; bc < 5; bc++)
      for (int i = 0; i < NC; i++)
        for (int j = 0; j < NC; j++) {
          double v = Dn(q, dir[br], dir[bc], i, j);
          at(br * NC + i, bc * NC + j) = (sg[br] * sg[bc] < 0) ? -v : v;
        }
  for (int i = 0, r = 0; i < NC; i++) {                                 // :247-295
    const ContactParams& cp = q.cons[i]->cp;
    if (cp.NK > 4) {
      const int nk4 = (cp.NK + 4) / 4;
      for (int k = 0; k < nk4; k++) {
        const double cs = std::cos((M_PI * k) / (2.0 * nk4)), sn = std::sin((M_PI * k) / (2.0 * nk4));
        at(N_CONST + r + k, i) = cp.mu_c;
        at(N_CONST + r + k, NC + i) = -cs; at(N_CONST + r + k, 2 * NC + i) = -cs;
        at(N_CONST + r + k, 3 * NC + i) = -sn; at(N_CONST + r + k, 4 * NC + i) = -sn;
        at(NC + i, N_CONST + r + k) = cs; at(2 * NC + i, N_CONST + r + k) = cs;
        at(3 * NC + i, N_CONST + r + k) = sn; at(4 * NC + i, N_CONST + r + k) = sn;
      }
      r += nk4;
    } else {
      at(N_CONST + r, i) = cp.mu_c;
      for (int b = 1; b < 5; b++) { at(N_CONST + r, b * NC + i) = -1.0; at(b * NC + i, N_CONST + r) = 1.0; }
      r += 1;
    }
  }
  for (int i = 0; i < NC; i++) {                                        // :174-175,302-308
    qq[i] = q.Cv[0][i]; qq[NC + i] = q.Cv[1][i]; qq[2 * NC + i] = -q.Cv[1][i];
    qq[3 * NC + i] = q.Cv[2][i]; qq[4 * NC + i] = -q.Cv[2][i];
  }
}

// update_from_stacked (:298-397) without bilateral joints: dv = X_CnT cn + X_CsT cs + X_CtT ct; v += dv
static void apply_to_bodies(Sim& S, ProblemData& q) {
  Vec dv(q.ngc, 0.0);
  for (int g = 0; g < q.ngc; g++) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < q.nc; i++) a = std::fma(q.XT[0](g, i), q.cn[i], a);
    for (int i = 0; i < q.nc; i++) b = std::fma(q.XT[1](g, i), q.cs[i], b);
    for (int i = 0; i < q.nc; i++) c = std::fma(q.XT[2](g, i), q.ct[i], c);
    dv[g] = (a + b) + c;
  }
  for (size_t i = 0; i < q.sb.size(); i++) {                           // update_generalized_velocities :1784-1798
    Body& b = S.bodies[q.sb[i]];
    const int g = q.gc[q.sb[i]];
    if (S.has_rc && q.sb[i] == S.rc_first + 1) {                       // the articulated body: joint velocities, then its links
      for (int k = 0; k < S.rc.ndof(); k++) S.jqd[k] = S.jqd[k] + dv[g + k];
      S.rc_update_links();
      continue;
    }
    b.vl = b.vl + V3(dv[g], dv[g + 1], dv[g + 2]);
    b.va = b.va + V3(dv[g + 3], dv[g + 4], dv[g + 5]);
  }
}

// update_constraint_velocities_from_impulses (:427-464), nl = 0
static void update_constraint_velocities(ProblemData& q) {
  const int nc = q.nc;
  const Vec* imp[3] = {&q.cn, &q.cs, &q.ct};
  for (int d = 0; d < 3; d++)
    for (int k = 0; k < 3; k++) {
      for (int i = 0; i < nc; i++) {
        double s = 0.0;
        for (int j = 0; j < nc; j++) s = std::fma(Dn(q, d, k, i, j), (*imp[k])[j], s);
        q.Cv[d][i] += s;
      }
    }
}

static double calc_min_constraint_velocity(const ProblemData& q) {     // :413-424
  double minv = INF;
  if (!q.Cv[0].empty()) minv = *std::min_element(q.Cv[0].begin(), q.Cv[0].end());
  return minv;
}

// solve_qp_work (ImpactConstraintHandlerQP.cpp:94-263): returns cn/cs/ct in q
static void solve_qp(Sim& S, ProblemData& q) {
  int n; Vec MM, qq;
  build_qp_lcp(q, n, MM, qq);
  Vec z;
  if (S.zlast.size() == (size_t)n) z = S.zlast; else z.assign(n, 0.0);  // :158-162 with rule H1
  S.cnt.lcp_solves++;
  S.cnt.max_lcp_n = std::max<long long>(S.cnt.max_lcp_n, n);
  const unsigned long long f0 = S.lcp.n_fast_calls, l0 = S.lcp.n_lemke_calls, p0 = S.lcp.n_pivots_total;
  if (!S.lcp.lcp_fast_regularized(n, MM.data(), qq.data(), z, -20, 4, -8)) {   // :219
    z.assign(n, 0.0);                                                   // :222
    if (!S.lcp.lcp_lemke_regularized(n, MM.data(), qq.data(), z)) {     // :224
      S.cnt.lcp_failures++; S.mini_failed = true;                       // LCPSolverException: impulses are not applied
      z.assign(n, 0.0);
    }
  }
  S.cnt.lcp_fast_calls += S.lcp.n_fast_calls - f0; S.cnt.lemke_calls += S.lcp.n_lemke_calls - l0; S.cnt.pivots += S.lcp.n_pivots_total - p0;
  S.cnt.pivot_flops += (long long)(S.lcp.n_pivots_total - p0) * 2 * n * (n + 1);
  S.zlast = z;                                                          // :233
  S.last_n = n; S.last_MM = MM; S.last_qq = qq; S.last_z = z;
  const int nc = q.nc;                                                  // update_from_stacked_qp (UnilateralConstraintProblemData.h:218-228)
  for (int i = 0; i < nc; i++) {
    q.cn[i] = z[i];
    q.cs[i] = z[nc + i] - z[3 * nc + i];
    q.ct[i] = z[2 * nc + i] - z[4 * nc + i];
  }
}

// apply_model_to_connected_constraints (ImpactConstraintHandler.cpp:530-626)
static void apply_qp_model(Sim& S, ProblemData& q) {
  solve_qp(S, q);
  apply_to_bodies(S, q);                                                // update_from_stacked(_epd, _z) :569
  update_constraint_velocities(q);                                      // :572
  double minv = calc_min_constraint_velocity(q);
  // apply_restitution(epd, z) (:470-491): only cn entries of z are scaled; cs/ct keep the friction values (H3)
  bool changed = false;
  for (int i = 0; i < q.nc; i++) {
    q.cn[i] *= q.cons[i]->cp.eps;
    if (!changed && q.cn[i] > NEAR_ZERO) changed = true;
  }
  if (changed) {
    apply_to_bodies(S, q);                                              // :581 (friction applied a second time)
    update_constraint_velocities(q);
    double minv_plus = calc_min_constraint_velocity(q);
    if (minv_plus < 0.0 && minv_plus < minv - NEAR_ZERO) {              // :591-601
      solve_qp(S, q);
      apply_to_bodies(S, q);
    }
  }
}

// apply_ap_model_to_connected_constraints (ImpactConstraintHandlerLCP.cpp:36-91)
static void solve_ap(Sim& S, ProblemData& q, Vec& acc_cn, Vec& acc_cs, Vec& acc_ct) {
  int n; Vec MM, qq;
  build_ap_lcp(q, n, MM, qq);
  Vec z;                                                                // :332 fresh vector: size 0 != n
  S.cnt.lcp_solves++;
  S.cnt.max_lcp_n = std::max<long long>(S.cnt.max_lcp_n, n);
  const unsigned long long l0 = S.lcp.n_lemke_calls, p0 = S.lcp.n_pivots_total;
  if (!S.lcp.lcp_lemke_regularized(n, MM.data(), qq.data(), z, -20, 1, -2)) {   // :333
    S.cnt.lcp_failures++; S.mini_failed = true;
    z.assign(n, 0.0);
  }
  S.cnt.lemke_calls += S.lcp.n_lemke_calls - l0; S.cnt.pivots += S.lcp.n_pivots_total - p0;
  S.cnt.pivot_flops += (long long)(S.lcp.n_pivots_total - p0) * 2 * n * (n + 1);
  S.last_n = n; S.last_MM = MM; S.last_qq = qq; S.last_z = z;
  const int NC = q.nc;
  for (int i = 0; i < NC; i++) {                                        // :336-342
    q.cn[i] = z[i];
    q.cs[i] = z[NC + i] - z[2 * NC + i];
    q.ct[i] = z[3 * NC + i] - z[4 * NC + i];
    acc_cn[i] += q.cn[i]; acc_cs[i] += q.cs[i]; acc_ct[i] += q.ct[i];   // propagate_impulse_data :350
  }
}

static void apply_ap_model(Sim& S, ProblemData& q) {
  Vec acn(q.nc, 0.0), acs(q.nc, 0.0), act(q.nc, 0.0);
  solve_ap(S, q, acn, acs, act);
  update_constraint_velocities(q);
  double minv = calc_min_constraint_velocity(q);
  bool changed = false;                                                 // apply_restitution(q) :497-524
  for (int i = 0; i < q.nc; i++) {
    q.cn[i] *= q.cons[i]->cp.eps;
    if (!changed && q.cn[i] > NEAR_ZERO) changed = true;
  }
  if (changed) {
    std::fill(q.cs.begin(), q.cs.end(), 0.0);
    std::fill(q.ct.begin(), q.ct.end(), 0.0);
    update_constraint_velocities(q);
    double minv_plus = calc_min_constraint_velocity(q);
    if (minv_plus < 0.0 && minv_plus < minv - NEAR_ZERO) solve_ap(S, q, acn, acs, act);
    else for (int i = 0; i < q.nc; i++) { acn[i] += q.cn[i]; acs[i] += q.cs[i]; act[i] += q.ct[i]; }   // propagate_impulse_data :80
  }
  q.cn = acn; q.cs = acs; q.ct = act;                                   // apply_impulses(:676-748): accumulated contact_impulse
  apply_to_bodies(S, q);
}

// apply_no_slip_model (ImpactConstraintHandler.cpp:1009-1417), nl = 0, no implicit joints.  The tangential directions
// become equality constraints; the largest set of them that keeps [S;T] X [S;T]^T positive definite is chosen greedily
// by trial Cholesky of the matrix skewed by -NEAR_ZERO (:1089-1145); the normal impulses solve the nc x nc Schur LCP
// MM = Cn X Cn^T - (Cn X W^T) Y (W X Cn^T), qq = Cn v - (Cn X W^T) Y W v with W = [S;T], Y = (W X W^T)^-1 (:1170-1236)
// through lcp_fast, falling back to lcp_lemke_regularized (:1239-1284).  Velocities are updated inside (:1370-1400).
// Rule H1 extended to the member `_v` handed to lcp_fast (:1239): per env, warm start iff its size equals nc; a size
// change takes lcp_fast's cold branch (LCP.cpp:89-103) exactly as the reference's size test would.
static void apply_no_slip_model(Sim& S, ProblemData& q) {
  const int nc = q.nc;
  std::vector<int> Si, Ti;
  Vec Y;
  auto form_Y = [&](double skew) {                                       // :1098-1112
    const int ns = (int)Si.size(), nt = (int)Ti.size(), m = ns + nt;
    Y.assign((size_t)m * m, 0.0);
    for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) Y[(size_t)b * m + a] = Dn(q, 1, 1, Si[a], Si[b]);
    for (int a = 0; a < nt; a++) for (int b = 0; b < nt; b++) Y[(size_t)(ns + b) * m + ns + a] = Dn(q, 2, 2, Ti[a], Ti[b]);
    for (int a = 0; a < ns; a++) for (int b = 0; b < nt; b++) { const double v = Dn(q, 1, 2, Si[a], Ti[b]); Y[(size_t)(ns + b) * m + a] = v; Y[(size_t)a * m + ns + b] = v; }
    for (int j = 0; j < m; j++) Y[(size_t)j * m + j] -= skew;
    return m;
  };
  for (int i = 0; i < nc; i++) {
    Si.push_back(i);
    int m = form_Y(NEAR_ZERO);
    if (!factor_chol(Y.data(), m)) Si.pop_back();                         // :1115-1116
    Ti.push_back(i);
    m = form_Y(NEAR_ZERO);
    if (!factor_chol(Y.data(), m)) Ti.pop_back();                         // :1140-1141
  }
  const int ns = (int)Si.size(), nt = (int)Ti.size(), m = ns + nt;
  form_Y(0.0);                                                            // :1165-1176
  const bool ok = (m == 0) || factor_chol(Y.data(), m);                   // :1179 (asserted in the reference)
  // Q X W^T : nc x m (:1195-1204), column-major
  Vec QXW((size_t)nc * m, 0.0);
  for (int a = 0; a < ns; a++) for (int i = 0; i < nc; i++) QXW[(size_t)a * nc + i] = Dn(q, 0, 1, i, Si[a]);
  for (int b = 0; b < nt; b++) for (int i = 0; i < nc; i++) QXW[(size_t)(ns + b) * nc + i] = Dn(q, 0, 2, i, Ti[b]);
  // workM = Y (W X Q^T): m x nc, one Cholesky solve per column (:1207-1208)
  Vec WM((size_t)m * nc, 0.0);
  for (int i = 0; i < nc; i++) {
    for (int a = 0; a < m; a++) WM[(size_t)i * m + a] = QXW[(size_t)a * nc + i];
    if (ok && m) solve_chol(Y.data(), m, &WM[(size_t)i * m]);
  }
  Vec MM((size_t)nc * nc), qq(nc);
  for (int j = 0; j < nc; j++)
    for (int i = 0; i < nc; i++) {
      double s = 0.0;
      for (int a = 0; a < m; a++) s = std::fma(QXW[(size_t)a * nc + i], WM[(size_t)j * m + a], s);   // :1211
      MM[(size_t)j * nc + i] = Dn(q, 0, 0, i, j) - s;                   // :1212
    }
  Vec YXv(m);
  for (int a = 0; a < ns; a++) YXv[a] = q.Cv[1][Si[a]];                  // :1220-1224
  for (int b = 0; b < nt; b++) YXv[ns + b] = q.Cv[2][Ti[b]];
  if (ok && m) solve_chol(Y.data(), m, YXv.data());                      // :1227-1228
  for (int i = 0; i < nc; i++) {
    double s = 0.0;
    for (int a = 0; a < m; a++) s = std::fma(QXW[(size_t)a * nc + i], YXv[a], s);   // :1231
    qq[i] = q.Cv[0][i] - s;                                              // :1215-1216,1234
  }
  S.cnt.lcp_solves++;
  S.cnt.max_lcp_n = std::max<long long>(S.cnt.max_lcp_n, nc);
  const unsigned long long f0 = S.lcp.n_fast_calls, l0 = S.lcp.n_lemke_calls, p0 = S.lcp.n_pivots_total;
  Vec v = S.vlast;                                                       // the member _v: warm start iff sizes match
  bool solved = ok && S.lcp.lcp_fast(nc, MM.data(), qq.data(), v);       // :1239
  if (ok && !solved) { v.clear(); solved = S.lcp.lcp_lemke_regularized(nc, MM.data(), qq.data(), v); }   // :1279
  if (!solved) { S.cnt.lcp_failures++; S.mini_failed = true; v.assign(nc, 0.0); }             // std::runtime_error in the reference (:1280)
  S.cnt.lcp_fast_calls += S.lcp.n_fast_calls - f0; S.cnt.lemke_calls += S.lcp.n_lemke_calls - l0; S.cnt.pivots += S.lcp.n_pivots_total - p0;
  S.cnt.pivot_flops += (long long)(S.lcp.n_pivots_total - p0) * 2 * nc * (nc + 1);
  S.vlast = v;
  S.last_n = nc; S.last_MM = MM; S.last_qq = qq; S.last_z = v;
  // [cs; ct] = -(Y W v + Y W X Q^T cn) (:1294-1299)
  Vec w(m, 0.0);
  for (int a = 0; a < m; a++) { double s = 0.0; for (int i = 0; i < nc; i++) s = std::fma(QXW[(size_t)a * nc + i], v[i], s); w[a] = s; }
  if (ok && m) solve_chol(Y.data(), m, w.data());
  for (int i = 0; i < nc; i++) { q.cn[i] = v[i]; q.cs[i] = 0.0; q.ct[i] = 0.0; }
  if (solved) {
    for (int a = 0; a < ns; a++) q.cs[Si[a]] = -(YXv[a] + w[a]);
    for (int b = 0; b < nt; b++) q.ct[Ti[b]] = -(YXv[ns + b] + w[ns + b]);
  }
  apply_to_bodies(S, q);                                                 // :1370-1400
}

// apply_no_slip_model_to_connected_constraints (ImpactConstraintHandler.cpp:236-293).  Rule H10: the trailing
// update_from_stacked(_epd, _z) after a second solve (:288) re-applies the QP handler's stale member _z (empty in a
// no-slip-only run: out-of-bounds in the reference); it is skipped here and in the kernels.
static void apply_no_slip_to_connected(Sim& S, ProblemData& q) {
  apply_no_slip_model(S, q);
  update_constraint_velocities(q);                                      // :262
  double minv = calc_min_constraint_velocity(q);
  bool changed = false;                                                 // apply_restitution(q) :497-524
  for (int i = 0; i < q.nc; i++) {
    q.cn[i] *= q.cons[i]->cp.eps;
    if (!changed && q.cn[i] > NEAR_ZERO) changed = true;
  }
  if (changed) {
    std::fill(q.cs.begin(), q.cs.end(), 0.0);
    std::fill(q.ct.begin(), q.ct.end(), 0.0);
    apply_to_bodies(S, q);                                              // update_from_stacked(q) :271
    update_constraint_velocities(q);                                    // :274
    double minv_plus = calc_min_constraint_velocity(q);
    if (minv_plus < 0.0 && minv_plus < minv - NEAR_ZERO) apply_no_slip_model(S, q);   // :281-285
  }
}

void Sim::assemble_island_lcp(const std::vector<Contact*>& cons, const std::vector<int>& island_bodies, int& n, Vec& MM, Vec& qq) {
  ProblemData q;
  compute_problem_data(*this, q, cons, island_bodies);
  if (model == MODEL_AP) build_ap_lcp(q, n, MM, qq); else build_qp_lcp(q, n, MM, qq);
}

// islands: UnilateralConstraint::determine_connected_constraints (UnilateralConstraint.cpp:940-1194), canonical order H4
typedef std::vector<std::pair<std::vector<Contact*>, std::vector<int> > > IslandList;
static void determine_connected_constraints(const Sim& S, std::vector<Contact>& contacts, IslandList& groups) {
  const std::vector<Body>& bodies = S.bodies;
  const int nb = (int)bodies.size();
  std::set<int> nodes;
  std::vector<std::vector<int> > adj(nb);
  for (size_t i = 0; i < contacts.size(); i++) {
    const int b1 = S.super_of(contacts[i].b1), b2 = S.super_of(contacts[i].b2);      // single bodies of one articulated body are one node
    const bool e1 = bodies[contacts[i].b1].enabled, e2 = bodies[contacts[i].b2].enabled;
    if (e1) nodes.insert(b1);
    if (e2) nodes.insert(b2);
    if (e1 && e2) { adj[b1].push_back(b2); adj[b2].push_back(b1); }
  }
  // std::multimap keeps equal keys in insertion order: neighbours are visited in contact order
  std::vector<char> taken(contacts.size(), 0);
  while (!nodes.empty()) {
    int node = *nodes.begin();
    groups.push_back(std::make_pair(std::vector<Contact*>(), std::vector<int>()));
    std::queue<int> nq;
    nq.push(node);
    std::set<int> processed;
    while (!nq.empty()) {
      node = nq.front(); nq.pop();
      nodes.erase(node);
      groups.back().second.push_back(node);
      processed.insert(node);
      for (size_t k = 0; k < adj[node].size(); k++) if (!processed.count(adj[node][k])) nq.push(adj[node][k]);
      for (size_t i = 0; i < contacts.size(); i++)
        if (!taken[i] && ((bodies[contacts[i].b1].enabled && S.super_of(contacts[i].b1) == node) || (bodies[contacts[i].b2].enabled && S.super_of(contacts[i].b2) == node))) { taken[i] = 1; groups.back().first.push_back(&contacts[i]); }
    }
    if (groups.back().first.empty()) groups.pop_back();
  }
}

// ConstraintSimulator::calc_impacting_unilateral_constraint_forces (:298-355) -> ImpactConstraintHandler::apply_model (:96-168)
void Sim::process_constraints(std::vector<Contact>& contacts) {
  last_contacts = contacts;
  if (contacts.empty()) return;
  bool none_impacting = true;
  for (size_t i = 0; i < contacts.size(); i++)
    if (calc_constraint_vel(contacts[i]) < -NEAR_ZERO) { none_impacting = false; break; }   // eNegative, UnilateralConstraint.cpp:1433-1446
  if (none_impacting) return;
  IslandList groups;
  determine_connected_constraints(*this, contacts, groups);
  // remove_inactive_groups (:1197-1225), evaluated before any island is solved
  std::vector<char> active(groups.size(), 0);
  for (size_t g = 0; g < groups.size(); g++)
    for (size_t i = 0; i < groups[g].first.size(); i++)
      if (calc_constraint_vel(*groups[g].first[i]) < -NEAR_ZERO) { active[g] = 1; break; }
  for (size_t g = 0; g < groups.size(); g++) {
    if (!active[g]) continue;
    ProblemData q;
    compute_problem_data(*this, q, groups[g].first, groups[g].second);
    bool all_inf = true;                                                // ImpactConstraintHandler.cpp:122-135
    for (size_t i = 0; i < groups[g].first.size(); i++) if (groups[g].first[i]->cp.mu_c < 1e2) all_inf = false;
    if (all_inf) apply_no_slip_to_connected(*this, q);
    else if (model == MODEL_AP) apply_ap_model(*this, q); else apply_qp_model(*this, q);
  }
  // ImpactToleranceException check over the remaining groups (ImpactConstraintHandler.cpp:153-167): only logged by the caller
  bool still = false;
  for (size_t g = 0; g < groups.size() && !still; g++) {
    if (!active[g]) continue;
    for (size_t i = 0; i < groups[g].first.size(); i++)
      if (calc_constraint_vel(*groups[g].first[i]) < -NEAR_ZERO) { still = true; break; }
  }
  if (still) cnt.impact_tol_events++;
}

// ---------- constraint stabilization: ConstraintStabilization.cpp (no implicit joints, no joint limits) ----------
namespace {
// Euler coordinates of every body of the simulator (ConstraintStabilization::get_body_configurations :1215-1237): seven
// per enabled free body (x y z qx qy qz qw), the joint positions of the articulated body; the same layout for dq.
struct StabQ { std::vector<double> x; Vec j; };     // x: [body][7] (entries of disabled bodies and links unused)
}
static void stab_get(const Sim& S, StabQ& q) {
  const size_t nb = S.bodies.size();
  q.x.assign(nb * 7, 0.0);
  for (size_t b = 0; b < nb; b++) { const Body& B = S.bodies[b]; for (int k = 0; k < 3; k++) q.x[b * 7 + k] = B.x[k]; for (int k = 0; k < 4; k++) q.x[b * 7 + 3 + k] = B.quat[k]; }
  q.j = S.jq;
}
// update_body_configurations(q + t dq) (:1252-1264): set_generalized_coordinates_euler normalises the quaternion
static void stab_set(Sim& S, const StabQ& q, const StabQ& dq, double t) {
  for (size_t b = 0; b < S.bodies.size(); b++) {
    Body& B = S.bodies[b];
    if (!B.enabled || S.is_link((int)b)) continue;
    double c[7];
    for (int k = 0; k < 7; k++) c[k] = dq.x[b * 7 + k] * t + q.x[b * 7 + k];                 // qstar = dq; qstar *= t; qstar += q
    B.x = V3(c[0], c[1], c[2]);
    const double nrm = std::sqrt(c[3] * c[3] + c[4] * c[4] + c[5] * c[5] + c[6] * c[6]);
    for (int k = 0; k < 4; k++) B.quat[k] = c[3 + k] / nrm;
    S.update_pose(B);
  }
  if (S.has_rc) { for (size_t k = 0; k < S.jq.size(); k++) S.jq[k] = dq.j[k] * t + q.j[k]; S.rc_update_links(); }
}
// evaluate_unilateral_constraints (:88-131): the pairwise distances at the current configuration; returns the smallest
static double stab_eval(const Sim& S, const std::vector<std::pair<int, int> >& pairs, std::vector<PairDist>& pd, std::vector<double>& uC) {
  S.calc_pairwise_distances(pairs, pd);
  double vio = INF;
  uC.resize(pd.size());
  for (size_t i = 0; i < pd.size(); i++) { uC[i] = pd[i].dist; vio = std::min(vio, uC[i]); }
  return vio;
}
static double stab_sign(double x, double y) { return (y > 0.0) ? std::fabs(x) : -std::fabs(x); }
// ridders_unilateral (:1322-1380), literally: note that the caller passes x2 = the current t with fh = the value at t = 1
static double stab_ridders(Sim& S, const std::vector<std::pair<int, int> >& pairs, double x1, double x2, double fl, double fh, size_t idx,
                           const StabQ& dq, const StabQ& q) {
  const unsigned MAX_ITERATIONS = 25;
  const double TOL = 1e-4;
  std::vector<PairDist> pd; std::vector<double> uC;
  auto eval = [&](double t) { stab_set(S, q, dq, t); stab_eval(S, pairs, pd, uC); return uC[idx]; };
  double ans = INF, fm, fnew, s, xh, xl, xm, xnew;
  if ((fl > 0.0 && fh < 0.0) || (fl < 0.0 && fh > 0.0)) {
    xl = x1; xh = x2;
    for (unsigned j = 0; j < MAX_ITERATIONS; j++) {
      xm = 0.5 * (xl + xh);
      fm = eval(xm);
      s = std::sqrt(fm * fm - fl * fh);
      if (s == 0.0) return ans;
      xnew = xm + (xm - xl) * ((fl >= fh ? 1.0 : -1.0) * fm / s);
      ans = xnew;
      fnew = eval(ans);
      if (std::fabs(fnew) < TOL && fnew >= 0.0) return xnew;
      if (stab_sign(fm, fnew) != fm) { xl = xm; fl = fm; xh = ans; fh = fnew; }
      else if (stab_sign(fl, fnew) != fl) { xh = ans; fh = fnew; }
      else if (stab_sign(fh, fnew) != fh) { xl = ans; fl = fnew; }
      else return 0.0;                                                     // assert(false) in the reference
    }
  } else {
    if (fl == 0.0) return x1;
    if (fh == 0.0) return x2;
  }
  return 0.0;
}
// update_q (:1055-1212): line search along dq; leaves the bodies at q + t dq and stores that in q
static bool stab_update_q(Sim& S, const std::vector<std::pair<int, int> >& pairs, const StabQ& dq, StabQ& q) {
  const double MIN_T = NEAR_ZERO, BETA = 0.6;
  std::vector<PairDist> pd; std::vector<double> uC, uC_old;
  stab_eval(S, pairs, pd, uC_old);
  stab_set(S, q, dq, 1.0);
  stab_eval(S, pairs, pd, uC);
  std::vector<char> bracket(uC.size());
  for (size_t i = 0; i < uC.size(); i++) bracket[i] = (uC_old[i] < 0.0 && uC[i] > 0.0) || (uC_old[i] > 0.0 && uC[i] < 0.0);
  double t = 1.0;
  const std::vector<double> uC1 = uC;                                      // the values at t = 1 (ridders evaluates into its own arrays)
  for (size_t i = 0; i < bracket.size(); i++) {
    if (!bracket[i]) continue;
    const double root = stab_ridders(S, pairs, 0.0, t, uC_old[i], uC1[i], i, dq, q);
    if (root > 0.0 && root < 1.0) t = std::min(root, t);
  }
  stab_set(S, q, dq, t);
  stab_eval(S, pairs, pd, uC);
  for (;;) {
    bool stop = true;
    for (size_t i = 0; i < bracket.size(); i++) if (!bracket[i] && uC[i] < 0.0 && uC_old[i] > uC[i]) { stop = false; break; }
    if (stop) break;                                                       // no bilateral constraints: violation 0 < bilateral_eps
    t *= BETA;
    if (t < MIN_T) return false;
    stab_set(S, q, dq, t);
    stab_eval(S, pairs, pd, uC);
  }
  for (size_t k = 0; k < q.x.size(); k++) q.x[k] = dq.x[k] * t + q.x[k];   // q = qstar (NOT renormalised: the stored vector, :1209)
  for (size_t k = 0; k < q.j.size(); k++) q.j[k] = dq.j[k] * t + q.j[k];
  return true;
}

void Sim::stabilize() {
  if (stab_max_iterations == 0) return;                                    // :173-174
  const size_t nb = bodies.size();
  std::vector<V3> vl_save(nb), va_save(nb);                                // save_velocities :66-75
  for (size_t b = 0; b < nb; b++) { vl_save[b] = bodies[b].vl; va_save[b] = bodies[b].va; }
  const Vec jqd_save = jqd;
  StabQ q, dq;
  stab_get(*this, q);
  std::vector<std::pair<int, int> > pairs;
  broad_phase(pairs);
  std::vector<PairDist> pd; std::vector<double> uC;
  double max_uvio = stab_eval(*this, pairs, pd, uC);                        // :187
  const long long cap = stab_max_iterations < 0 ? 100 : std::min(100, stab_max_iterations);   // rule H12
  long long iterations = 0;
  while (max_uvio < stab_eps) {                                             // :197 (bilateral violation is 0)
    if (iterations == cap) { if (stab_max_iterations < 0 || stab_max_iterations > 100) cnt.stab_line_search_failures++; break; }
    for (size_t b = 0; b < nb; b++) { bodies[b].vl = V3(); bodies[b].va = V3(); }      // :211-217
    if (has_rc) { for (size_t k = 0; k < jqd.size(); k++) jqd[k] = 0.0; rc_update_links(); }
    // compute_problem_data (:348-478): one contact at the closest points of a separated pair, the narrowphase's
    // contacts (TOL = NEAR_ZERO, CollisionDetection.h:46) of a touching / penetrating one
    std::vector<Contact> contacts;
    calc_pairwise_distances(pairs, pd);
    for (size_t p = 0; p < pd.size(); p++) {                               // add_contact_constraints :304-345
      if (pd[p].dist == INF) continue;
      if (pd[p].dist >= NEAR_ZERO) {
        const V3 normal = normalize(pd[p].pb - pd[p].pa);
        contacts.push_back(create_contact(pd[p].a, pd[p].b, pd[p].pa, normal, pd[p].dist));
      } else find_contacts(pd[p].a, pd[p].b, NEAR_ZERO, contacts);
    }
    IslandList groups;
    determine_connected_constraints(*this, contacts, groups);
    dq.x.assign(nb * 7, 0.0); dq.j.assign(jq.size(), 0.0);
    for (size_t g = 0; g < groups.size(); g++) {                           // determine_dq :932-970
      ProblemData pq;
      compute_problem_data(*this, pq, groups[g].first, groups[g].second);
      const int n = pq.nc;
      Vec MM((size_t)n * n), qq(n), z;
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) MM[(size_t)j * n + i] = pq.D[0][0](i, j);
      for (int i = 0; i < n; i++) qq[i] = pq.cons[i]->dist - std::fabs(stab_eps) - NEAR_ZERO;   // :432-433
      const unsigned long long f0 = stab_lcp.n_fast_calls, l0 = stab_lcp.n_lemke_calls, p0 = stab_lcp.n_pivots_total;
      cnt.stab_lcp_solves++;
      if (!stab_lcp.lcp_fast(n, MM.data(), qq.data(), z))                   // :961 (z is a fresh vector: cold start)
        if (!stab_lcp.lcp_lemke_regularized(n, MM.data(), qq.data(), z)) { z.assign(n, 0.0); cnt.lcp_failures++; }   // rule H12
      cnt.lcp_fast_calls += stab_lcp.n_fast_calls - f0; cnt.lemke_calls += stab_lcp.n_lemke_calls - l0; cnt.pivots += stab_lcp.n_pivots_total - p0;
      for (int i = 0; i < n; i++) { pq.cn[i] = z[i]; pq.cs[i] = 0.0; pq.ct[i] = 0.0; }
      apply_to_bodies(*this, pq);                                          // update_from_stacked :965
      for (size_t k = 0; k < pq.sb.size(); k++) {                          // dq <- generalized velocity in Euler coordinates :968-975
        const int b = pq.sb[k];
        if (has_rc && b == rc_first + 1) { dq.j = jqd; continue; }
        const Body& B = bodies[b];
        const double qx = B.quat[0], qy = B.quat[1], qz = B.quat[2], qw = B.quat[3];
        const V3& w = B.va;
        dq.x[b * 7 + 0] = B.vl.x; dq.x[b * 7 + 1] = B.vl.y; dq.x[b * 7 + 2] = B.vl.z;
        dq.x[b * 7 + 3] = 0.5 * (+qw * w.x + qz * w.y - qy * w.z);         // Quatd::deriv, as in do_mini_step
        dq.x[b * 7 + 4] = 0.5 * (-qz * w.x + qw * w.y + qx * w.z);
        dq.x[b * 7 + 5] = 0.5 * (+qy * w.x - qx * w.y + qw * w.z);
        dq.x[b * 7 + 6] = 0.5 * (-qx * w.x - qy * w.y - qz * w.z);
      }
    }
    if (!stab_update_q(*this, pairs, dq, q)) { cnt.stab_line_search_failures++; break; }   // :231-235
    max_uvio = stab_eval(*this, pairs, pd, uC);                             // :238
    iterations++;
  }
  cnt.stab_iterations += iterations;
  for (size_t b = 0; b < nb; b++) { bodies[b].vl = vl_save[b]; bodies[b].va = va_save[b]; }   // restore_velocities :78-85
  if (has_rc) { jqd = jqd_save; rc_update_links(); }
}

}  // namespace oracle
