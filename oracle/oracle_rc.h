// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Reduced-coordinate articulated body (Moby RCArticulatedBody, include/Moby/RCArticulatedBody.h:43; the
// arithmetic is Ravelin::RCArticulatedBodyd -- un-vendored, un-pinned, NOT under /root/reference).
// What the reference selects:
//   fdyn-algorithm="fsab" -> Featherstone's articulated-body algorithm   (RCArticulatedBody.cpp:178-201,
//   fdyn-algorithm="crb"  -> composite-rigid-body inertia + dense solve    feeder.xml:37, pendulum.xml:21,
//                            (hard-wired for SDF robots such as the UR10)  SDFReader.cpp:931-935,973-978)
// Both are restated here from Featherstone, "Rigid Body Dynamics Algorithms" (2008), tables 7.1 (ABA), 6.2 (CRB)
// and 5.1 (RNEA), in LINK coordinates (each link's frame sits at its centre of mass, as Moby's eLinkCOM frame does)
// with Pluecker transforms between links.  The product kernels use a different formulation (every spatial quantity
// in world coordinates, no transforms), so agreement between the two is an independent check of both.
// PARITY UNPINNED against Ravelin itself: the only pins the reference tree offers for RC bodies are
// regress/fixed-articulated-table.dat and contact-constrained-pendulum.dat, compared at 1e-2 (regression-test:49).
//
// Model: fixed base (link 0, welded to the world), every other link hangs off `parent[i] < i` by a one-DoF
// revolute or prismatic joint; generalized coordinate k = i-1 belongs to link i.  Spatial vectors are
// [angular; linear], 6x6 matrices row-major.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include "oracle_linalg.h"

namespace oracle {

enum { RC_MAX_LINKS = 16 };
enum JointType { JOINT_REVOLUTE = 1, JOINT_PRISMATIC = 2 };

struct RCModel {
  int n_links = 0;                       // base included
  int parent[RC_MAX_LINKS];
  int jtype[RC_MAX_LINKS];
  double axis[RC_MAX_LINKS][3];          // unit joint axis, outboard link frame
  double loc_parent[RC_MAX_LINKS][3];    // joint location, inboard link (COM) frame
  double loc_child[RC_MAX_LINKS][3];     // joint location, outboard link (COM) frame
  double rel_quat[RC_MAX_LINKS][4];      // outboard orientation relative to inboard at q = 0 (x y z w)
  double mass[RC_MAX_LINKS];
  double J[RC_MAX_LINKS][3];             // principal inertia about the COM, link frame
  double base_x[3] = {0, 0, 0};          // pose of the base link in the world
  double base_quat[4] = {0, 0, 0, 1};
  int ndof() const { return n_links - 1; }
};

namespace rc {

typedef double M3[9];
typedef double S6[6];
typedef double M6[36];

inline void quat_R(const double* q, double* R) {   // same formula as Sim::update_pose
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1.0 - 2.0 * (y * y + z * z); R[1] = 2.0 * (x * y - w * z);       R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);       R[4] = 1.0 - 2.0 * (x * x + z * z); R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);       R[7] = 2.0 * (y * z + w * x);       R[8] = 1.0 - 2.0 * (x * x + y * y);
}
inline void axis_angle_R(const double* a, double th, double* R) {   // Rodrigues
  const double c = std::cos(th), s = std::sin(th), t = 1.0 - c;
  R[0] = t * a[0] * a[0] + c;        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c;        R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
}
inline void mm3(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  std::memcpy(C, T, sizeof(T));
}
inline void mv3(const double* A, const double* v, double* o) { double t[3]; for (int i = 0; i < 3; i++) t[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2]; o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; }
inline void mtv3(const double* A, const double* v, double* o) { double t[3]; for (int i = 0; i < 3; i++) t[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2]; o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; }
inline void cross3(const double* a, const double* b, double* o) { double t[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; }

// Pluecker transform parent -> child: E = R_rel^T (3x3), r = child origin in parent coordinates.
struct Xf { double E[9]; double r[3]; };
inline void xf_motion(const Xf& X, const double* m, double* o) {          // o = X m
  double t[3], w[3], v[3];
  cross3(X.r, m, t);                                                      // r x w
  for (int k = 0; k < 3; k++) t[k] = m[3 + k] - t[k];                     // v - r x w
  mv3(X.E, m, w); mv3(X.E, t, v);
  for (int k = 0; k < 3; k++) { o[k] = w[k]; o[3 + k] = v[k]; }
}
inline void xf_force_T(const Xf& X, const double* f, double* o) {         // o = X^T f  (child force -> parent coordinates)
  double n[3], l[3], t[3];
  mtv3(X.E, f, n); mtv3(X.E, f + 3, l);
  cross3(X.r, l, t);
  for (int k = 0; k < 3; k++) { o[k] = n[k] + t[k]; o[3 + k] = l[k]; }
}
inline void xf_matrix(const Xf& X, double* M) {                           // 6x6 motion transform
  double rx[9] = {0, -X.r[2], X.r[1], X.r[2], 0, -X.r[0], -X.r[1], X.r[0], 0}, Erx[9];
  mm3(X.E, rx, Erx);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    M[i * 6 + j] = X.E[i * 3 + j]; M[i * 6 + 3 + j] = 0.0;
    M[(3 + i) * 6 + j] = -Erx[i * 3 + j]; M[(3 + i) * 6 + 3 + j] = X.E[i * 3 + j];
  }
}
inline void xt_I_x(const Xf& X, const double* I, double* out) {           // out = X^T I X
  double M[36], T[36];
  xf_matrix(X, M);
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += I[i * 6 + k] * M[k * 6 + j]; T[i * 6 + j] = s; }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += M[k * 6 + i] * T[k * 6 + j]; out[i * 6 + j] = s; }
}
inline void crm(const double* v, const double* m, double* o) {            // v x m (motion)
  double a[3], b[3], c[3];
  cross3(v, m, a); cross3(v, m + 3, b); cross3(v + 3, m, c);
  for (int k = 0; k < 3; k++) { o[k] = a[k]; o[3 + k] = b[k] + c[k]; }
}
inline void crf(const double* v, const double* f, double* o) {            // v x* f (force)
  double a[3], b[3], c[3];
  cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
  for (int k = 0; k < 3; k++) { o[k] = a[k] + b[k]; o[3 + k] = c[k]; }
}
inline void mv6(const double* I, const double* v, double* o) { double t[6]; for (int i = 0; i < 6; i++) { double s = 0; for (int k = 0; k < 6; k++) s += I[i * 6 + k] * v[k]; t[i] = s; } std::memcpy(o, t, sizeof(t)); }
inline double dot6(const double* a, const double* b) { double s = 0; for (int k = 0; k < 6; k++) s += a[k] * b[k]; return s; }

}  // namespace rc

// Kinematic state of every link at (q, qd): world pose, joint transforms, motion subspaces, link-frame velocities.
struct RCKin {
  double R[RC_MAX_LINKS][9], x[RC_MAX_LINKS][3];    // world pose of each link frame (COM)
  rc::Xf X[RC_MAX_LINKS];                            // parent -> link
  double S[RC_MAX_LINKS][6];                         // motion subspace, link coordinates
  double v[RC_MAX_LINKS][6];                         // spatial velocity, link coordinates
  double vl[RC_MAX_LINKS][3], va[RC_MAX_LINKS][3];   // COM linear / angular velocity, world axes
};

inline void rc_kinematics(const RCModel& m, const double* q, const double* qd, RCKin& k) {
  using namespace rc;
  quat_R(m.base_quat, k.R[0]);
  for (int c = 0; c < 3; c++) k.x[0][c] = m.base_x[c];
  for (int c = 0; c < 6; c++) k.v[0][c] = 0.0;
  for (int c = 0; c < 3; c++) k.vl[0][c] = k.va[0][c] = 0.0;
  for (int i = 1; i < m.n_links; i++) {
    const int p = m.parent[i];
    double R0[9], Rq[9], Rrel[9], t[3], r[3];
    quat_R(m.rel_quat[i], R0);
    if (m.jtype[i] == JOINT_REVOLUTE) {
      axis_angle_R(m.axis[i], q[i - 1], Rq); mm3(R0, Rq, Rrel);
      mv3(Rrel, m.loc_child[i], t);
      for (int c = 0; c < 3; c++) r[c] = m.loc_parent[i][c] - t[c];
      cross3(m.loc_child[i], m.axis[i], t);
      for (int c = 0; c < 3; c++) { k.S[i][c] = m.axis[i][c]; k.S[i][3 + c] = t[c]; }
    } else {
      std::memcpy(Rrel, R0, sizeof(R0));
      for (int c = 0; c < 3; c++) t[c] = m.axis[i][c] * q[i - 1] - m.loc_child[i][c];
      mv3(Rrel, t, t);
      for (int c = 0; c < 3; c++) r[c] = m.loc_parent[i][c] + t[c];
      for (int c = 0; c < 3; c++) { k.S[i][c] = 0.0; k.S[i][3 + c] = m.axis[i][c]; }
    }
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) k.X[i].E[a * 3 + b] = Rrel[b * 3 + a];
    for (int c = 0; c < 3; c++) k.X[i].r[c] = r[c];
    mm3(k.R[p], Rrel, k.R[i]);
    mv3(k.R[p], r, t);
    for (int c = 0; c < 3; c++) k.x[i][c] = k.x[p][c] + t[c];
    double vp[6];
    xf_motion(k.X[i], k.v[p], vp);
    for (int c = 0; c < 6; c++) k.v[i][c] = vp[c] + k.S[i][c] * qd[i - 1];
    mv3(k.R[i], k.v[i], k.va[i]); mv3(k.R[i], k.v[i] + 3, k.vl[i]);
  }
}

inline void rc_link_inertia(const RCModel& m, int i, double* I) {
  for (int c = 0; c < 36; c++) I[c] = 0.0;
  for (int c = 0; c < 3; c++) { I[c * 6 + c] = m.J[i][c]; I[(3 + c) * 6 + 3 + c] = m.mass[i]; }
}

// Featherstone ABA (RBDA table 7.1).  g: gravity, world axes.  tau, qdd: ndof.
inline void rc_aba(const RCModel& m, const double* q, const double* qd, const double* tau, const double* g, double* qdd) {
  using namespace rc;
  const int N = m.n_links;
  RCKin k; rc_kinematics(m, q, qd, k);
  static thread_local double IA[RC_MAX_LINKS][36], pA[RC_MAX_LINKS][6], c[RC_MAX_LINKS][6], U[RC_MAX_LINKS][6], a[RC_MAX_LINKS][6];
  double D[RC_MAX_LINKS], u[RC_MAX_LINKS];
  for (int i = 1; i < N; i++) {
    double vJ[6], Iv[6];
    for (int t = 0; t < 6; t++) vJ[t] = k.S[i][t] * qd[i - 1];
    crm(k.v[i], vJ, c[i]);
    rc_link_inertia(m, i, IA[i]);
    mv6(IA[i], k.v[i], Iv);
    crf(k.v[i], Iv, pA[i]);
  }
  for (int i = N - 1; i >= 1; i--) {
    mv6(IA[i], k.S[i], U[i]);
    D[i] = dot6(k.S[i], U[i]);
    u[i] = tau[i - 1] - dot6(k.S[i], pA[i]);
    const int p = m.parent[i];
    if (p != 0) {
      double Ia[36], pa[6], Iac[6], T[36], tp[6];
      for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) Ia[r * 6 + s] = IA[i][r * 6 + s] - U[i][r] * U[i][s] / D[i];
      mv6(Ia, c[i], Iac);
      for (int r = 0; r < 6; r++) pa[r] = pA[i][r] + Iac[r] + U[i][r] * (u[i] / D[i]);
      xt_I_x(k.X[i], Ia, T);
      for (int r = 0; r < 36; r++) IA[p][r] += T[r];
      xf_force_T(k.X[i], pa, tp);
      for (int r = 0; r < 6; r++) pA[p][r] += tp[r];
    }
  }
  double gb[3];
  mtv3(k.R[0], g, gb);
  for (int t = 0; t < 3; t++) { a[0][t] = 0.0; a[0][3 + t] = -gb[t]; }
  for (int i = 1; i < N; i++) {
    double ap[6];
    xf_motion(k.X[i], a[m.parent[i]], ap);
    for (int t = 0; t < 6; t++) ap[t] += c[i][t];
    qdd[i - 1] = (u[i] - dot6(U[i], ap)) / D[i];
    for (int t = 0; t < 6; t++) a[i][t] = ap[t] + k.S[i][t] * qdd[i - 1];
  }
}

// Joint-space inertia H (ndof x ndof, column-major, symmetric) by the composite-rigid-body algorithm (RBDA table 6.2).
inline void rc_crb(const RCModel& m, const RCKin& k, double* H) {
  using namespace rc;
  const int N = m.n_links, nd = N - 1;
  static thread_local double Ic[RC_MAX_LINKS][36];
  for (int i = 1; i < N; i++) rc_link_inertia(m, i, Ic[i]);
  for (int i = 0; i < nd * nd; i++) H[i] = 0.0;
  for (int i = N - 1; i >= 1; i--) {
    const int p = m.parent[i];
    if (p != 0) { double T[36]; xt_I_x(k.X[i], Ic[i], T); for (int r = 0; r < 36; r++) Ic[p][r] += T[r]; }
    double F[6];
    mv6(Ic[i], k.S[i], F);
    H[(size_t)(i - 1) * nd + (i - 1)] = dot6(k.S[i], F);
    int j = i;
    while (m.parent[j] != 0) {
      double Fp[6];
      xf_force_T(k.X[j], F, Fp);
      for (int t = 0; t < 6; t++) F[t] = Fp[t];
      j = m.parent[j];
      const double h = dot6(F, k.S[j]);
      H[(size_t)(j - 1) * nd + (i - 1)] = h; H[(size_t)(i - 1) * nd + (j - 1)] = h;
    }
  }
}

// Bias forces C(q,qd) - gravity terms by the recursive Newton-Euler algorithm with qdd = 0 (RBDA table 5.1).
inline void rc_bias(const RCModel& m, const RCKin& k, const double* qd, const double* g, double* C) {
  using namespace rc;
  const int N = m.n_links;
  double a[RC_MAX_LINKS][6], f[RC_MAX_LINKS][6], gb[3];
  mtv3(k.R[0], g, gb);
  for (int t = 0; t < 3; t++) { a[0][t] = 0.0; a[0][3 + t] = -gb[t]; }
  for (int i = 1; i < N; i++) {
    double vJ[6], c[6], I[36], Ia[6], Iv[6], vIv[6];
    for (int t = 0; t < 6; t++) vJ[t] = k.S[i][t] * qd[i - 1];
    crm(k.v[i], vJ, c);
    xf_motion(k.X[i], a[m.parent[i]], a[i]);
    for (int t = 0; t < 6; t++) a[i][t] += c[t];
    rc_link_inertia(m, i, I);
    mv6(I, a[i], Ia); mv6(I, k.v[i], Iv); crf(k.v[i], Iv, vIv);
    for (int t = 0; t < 6; t++) f[i][t] = Ia[t] + vIv[t];
  }
  for (int i = N - 1; i >= 1; i--) {
    C[i - 1] = dot6(k.S[i], f[i]);
    const int p = m.parent[i];
    if (p != 0) { double fp[6]; xf_force_T(k.X[i], f[i], fp); for (int t = 0; t < 6; t++) f[p][t] += fp[t]; }
  }
}

// CRB forward dynamics: H qdd = tau - C, Cholesky (what Ravelin's eCRB path does with factor_chol / solve_chol_fast).
inline bool rc_crb_fwd_dyn(const RCModel& m, const double* q, const double* qd, const double* tau, const double* g, double* qdd) {
  const int nd = m.ndof();
  RCKin k; rc_kinematics(m, q, qd, k);
  std::vector<double> H((size_t)nd * nd), C(nd);
  rc_crb(m, k, H.data());
  rc_bias(m, k, qd, g, C.data());
  for (int i = 0; i < nd; i++) qdd[i] = tau[i] - C[i];
  if (!factor_chol(H.data(), nd)) return false;
  solve_chol(H.data(), nd, qdd);
  return true;
}

// Link Jacobian in Moby's convention (RCArticulatedBodyd::calc_jacobian with the link's mixed pose: linear velocity of
// the COM and angular velocity, world axes): J is 6 x ndof, row-major, rows [linear; angular].
inline void rc_link_jacobian(const RCModel& m, const RCKin& k, int link, double* J) {
  using namespace rc;
  const int nd = m.ndof();
  for (int i = 0; i < 6 * nd; i++) J[i] = 0.0;
  for (int j = link; j != 0; j = m.parent[j]) {
    double aw[3], pj[3], t[3];
    mv3(k.R[j], m.axis[j], aw);
    if (m.jtype[j] == JOINT_REVOLUTE) {
      mv3(k.R[j], m.loc_child[j], pj);
      for (int c = 0; c < 3; c++) pj[c] = k.x[link][c] - (k.x[j][c] + pj[c]);   // joint point -> link COM
      cross3(aw, pj, t);
      for (int c = 0; c < 3; c++) { J[c * nd + (j - 1)] = t[c]; J[(3 + c) * nd + (j - 1)] = aw[c]; }
    } else {
      for (int c = 0; c < 3; c++) J[c * nd + (j - 1)] = aw[c];
    }
  }
}

// Total mechanical energy (for the conservation tests): 1/2 qd^T H qd - sum m g . x_com
inline double rc_energy(const RCModel& m, const double* q, const double* qd, const double* g) {
  const int nd = m.ndof();
  RCKin k; rc_kinematics(m, q, qd, k);
  std::vector<double> H((size_t)nd * nd);
  rc_crb(m, k, H.data());
  double ke = 0.0, pe = 0.0;
  for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) ke += 0.5 * qd[i] * H[(size_t)j * nd + i] * qd[j];
  for (int i = 1; i < m.n_links; i++) pe -= m.mass[i] * (g[0] * k.x[i][0] + g[1] * k.x[i][1] + g[2] * k.x[i][2]);
  return ke + pe;
}

}  // namespace oracle
