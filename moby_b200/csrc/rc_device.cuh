// Reduced-coordinate articulated body (fixed base, one-DoF revolute / prismatic joints): kinematics, Featherstone
// articulated-body forward dynamics, composite-rigid-body joint-space inertia and recursive Newton-Euler bias forces.
//
// Replaces what Moby reaches through Simulator::calc_fwd_dyn -> RCArticulatedBodyd::calc_fwd_dyn
// (src/Simulator.cpp:544-553; algorithm choice src/RCArticulatedBody.cpp:178-201, SDFReader.cpp:931-935),
// RCArticulatedBodyd::calc_jacobian (src/ImpactConstraintHandler.cpp:1875) and get_generalized_inertia
// (src/ImpactConstraintHandler.cpp:1605).  The arithmetic itself is Ravelin's and is not in the reference tree.
//
// Formulation: every spatial quantity is expressed in WORLD coordinates about the world origin (motion vectors
// [omega; v_O], force vectors [n_O; f]).  The recursions then need no Pluecker transforms at all -- a link passes its
// articulated inertia to its parent by plain addition -- and one thread runs a whole env with the running quantities
// in registers.  Symmetric 6x6 inertias are kept as 21 doubles (upper triangle, row-major).
// The CPU checker used by the tests follows the classical link-coordinate formulation with transforms instead, so the
// two agree only up to rounding (~1e-13 relative): that difference is the independent check.
#pragma once
#include "common.cuh"

namespace b2m {

#define B2M_MAX_LINKS 16
enum { JT_REVOLUTE = 1, JT_PRISMATIC = 2 };
enum { FDYN_FSAB = 0, FDYN_CRB = 1 };

// Kinematic tree shared by all envs (device global memory, read through the read-only path)
struct RCTree {
  int n_links, first_body, fdyn, has_ctrl;
  int parent[B2M_MAX_LINKS], jtype[B2M_MAX_LINKS];
  double axis[B2M_MAX_LINKS][3], loc_parent[B2M_MAX_LINKS][3], loc_child[B2M_MAX_LINKS][3], R0[B2M_MAX_LINKS][9];
  double kp[B2M_MAX_LINKS], kv[B2M_MAX_LINKS], amp[B2M_MAX_LINKS], freq[B2M_MAX_LINKS];
};

// upper-triangle index of a symmetric 6x6
B2M_HD B2M_INL constexpr int sym_idx(int i, int j) { return i <= j ? i * 6 - i * (i - 1) / 2 + (j - i) : j * 6 - j * (j - 1) / 2 + (i - j); }

B2M_HD B2M_INL void sym_mv(const double* I, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 6; j++) s = fma(I[sym_idx(i, j)], x[j], s);
    y[i] = s;
  }
}
B2M_HD B2M_INL double dot6(const double* a, const double* b) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) s = fma(a[k], b[k], s);
  return s;
}
B2M_HD B2M_INL void cross3(const double* a, const double* b, double* o) {
  const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
// spatial cross products, [angular; linear]
B2M_HD B2M_INL void crm(const double* v, const double* m, double* o) {
  double a[3], b[3], c[3];
  cross3(v, m, a); cross3(v, m + 3, b); cross3(v + 3, m, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
B2M_HD B2M_INL void crf(const double* v, const double* f, double* o) {
  double a[3], b[3], c[3];
  cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
B2M_HD B2M_INL void mat3_mul(const double* A, const double* B, double* C) {
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = T[i];
}
B2M_HD B2M_INL void mat3_vec(const double* A, const double* v, double* o) {
  const double x = A[0] * v[0] + A[1] * v[1] + A[2] * v[2], y = A[3] * v[0] + A[4] * v[1] + A[5] * v[2], z = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
B2M_HD B2M_INL void axis_angle_R(const double* a, double th, double* R) {
  double s, c;
#ifdef __CUDA_ARCH__
  sincos(th, &s, &c);
#else
  s = sin(th); c = cos(th);
#endif
  const double t = 1.0 - c;
  R[0] = t * a[0] * a[0] + c;        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c;        R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
}

// rotation matrix (row-major) -> unit quaternion x y z w, largest-component branch (Shepperd)
B2M_HD inline void R_to_quat(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0) * 2.0;
    q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
    q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
    q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s;
  } else {
    const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
    q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s;
  }
}

// World pose of link i from its parent's (x_p, R_p) and the joint coordinate; also the joint's motion subspace S_i in
// world coordinates about the world origin.
B2M_HD inline void rc_link_fk(const RCTree& T, int i, double qi, const double* xp, const double* Rp, double* x, double* R, double* S) {
  double Rrel[9], t[3], r[3], aw[3];
  if (T.jtype[i] == JT_REVOLUTE) {
    double Rq[9];
    axis_angle_R(T.axis[i], qi, Rq);
    mat3_mul(T.R0[i], Rq, Rrel);
    mat3_vec(Rrel, T.loc_child[i], t);
    r[0] = T.loc_parent[i][0] - t[0]; r[1] = T.loc_parent[i][1] - t[1]; r[2] = T.loc_parent[i][2] - t[2];
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) Rrel[k] = T.R0[i][k];
    t[0] = T.axis[i][0] * qi - T.loc_child[i][0]; t[1] = T.axis[i][1] * qi - T.loc_child[i][1]; t[2] = T.axis[i][2] * qi - T.loc_child[i][2];
    mat3_vec(Rrel, t, t);
    r[0] = T.loc_parent[i][0] + t[0]; r[1] = T.loc_parent[i][1] + t[1]; r[2] = T.loc_parent[i][2] + t[2];
  }
  mat3_mul(Rp, Rrel, R);
  mat3_vec(Rp, r, t);
  x[0] = xp[0] + t[0]; x[1] = xp[1] + t[1]; x[2] = xp[2] + t[2];
  mat3_vec(R, T.axis[i], aw);
  if (T.jtype[i] == JT_REVOLUTE) {
    double pj[3];
    mat3_vec(R, T.loc_child[i], pj);
    pj[0] += x[0]; pj[1] += x[1]; pj[2] += x[2];
    cross3(pj, aw, t);
    S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2]; S[3] = t[0]; S[4] = t[1]; S[5] = t[2];
  } else {
    S[0] = S[1] = S[2] = 0.0; S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
  }
}

// Spatial inertia of a link about the world origin: [[Ic + m cx cx^T, m cx], [m cx^T, m 1]], Ic = R diag(J) R^T
B2M_HD inline void rc_link_inertia(double mass, const double* J, const double* c, const double* R, double* I) {
  double Ic[6];   // xx xy xz yy yz zz
  {
    int k = 0;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = a; b < 3; b++) Ic[k++] = R[a * 3] * J[0] * R[b * 3] + R[a * 3 + 1] * J[1] * R[b * 3 + 1] + R[a * 3 + 2] * J[2] * R[b * 3 + 2];
  }
  const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  I[sym_idx(0, 0)] = Ic[0] + mass * (cc - c[0] * c[0]); I[sym_idx(0, 1)] = Ic[1] - mass * c[0] * c[1]; I[sym_idx(0, 2)] = Ic[2] - mass * c[0] * c[2];
  I[sym_idx(1, 1)] = Ic[3] + mass * (cc - c[1] * c[1]); I[sym_idx(1, 2)] = Ic[4] - mass * c[1] * c[2];
  I[sym_idx(2, 2)] = Ic[5] + mass * (cc - c[2] * c[2]);
  // m cx: [[0,-cz,cy],[cz,0,-cx],[-cy,cx,0]]
  I[sym_idx(0, 3)] = 0.0;          I[sym_idx(0, 4)] = -mass * c[2]; I[sym_idx(0, 5)] = mass * c[1];
  I[sym_idx(1, 3)] = mass * c[2];  I[sym_idx(1, 4)] = 0.0;          I[sym_idx(1, 5)] = -mass * c[0];
  I[sym_idx(2, 3)] = -mass * c[1]; I[sym_idx(2, 4)] = mass * c[0];  I[sym_idx(2, 5)] = 0.0;
  I[sym_idx(3, 3)] = mass; I[sym_idx(3, 4)] = 0.0; I[sym_idx(3, 5)] = 0.0; I[sym_idx(4, 4)] = mass; I[sym_idx(4, 5)] = 0.0; I[sym_idx(5, 5)] = mass;
}

// Per-thread view of one env's articulated body.  Arrays indexed by link live in local memory when the tree is
// indexed dynamically; the running quantities of each recursion stay in registers.
struct RCState {
  double *x, *R;   // link poses (COM frames), world: [link][3], [link][9] row-major
  double *S;       // motion subspaces, world coordinates: [link][6]
  double *v;       // spatial velocities, world coordinates: [link][6]
};
// thread-local backing store for an RCState (the thread-per-env kernels)
struct RCLocal {
  double x[B2M_MAX_LINKS * 3], R[B2M_MAX_LINKS * 9], S[B2M_MAX_LINKS * 6], v[B2M_MAX_LINKS * 6];
  B2M_HD RCState view() { RCState s; s.x = x; s.R = R; s.S = S; s.v = v; return s; }
};

// Link poses, motion subspaces and spatial velocities.  x[0], R[0] must hold the base pose.
B2M_HD inline void rc_kinematics(const RCTree& T, const double* q, const double* qd, RCState& s) {
#pragma unroll
  for (int k = 0; k < 6; k++) s.v[6 * (0) + k] = 0.0;
  for (int i = 1; i < T.n_links; i++) {
    const int p = T.parent[i];
    rc_link_fk(T, i, q[i - 1], s.x + 3 * (p), s.R + 9 * (p), s.x + 3 * (i), s.R + 9 * (i), s.S + 6 * (i));
#pragma unroll
    for (int k = 0; k < 6; k++) s.v[6 * i + k] = fma(s.S[6 * i + k], qd[i - 1], s.v[6 * p + k]);
  }
}
// COM linear velocity and angular velocity of link i, world axes (what narrowphase / conservative advancement read)
B2M_HD B2M_INL void rc_link_velocity(const RCState& s, int i, double* vl, double* va) {
  double t[3];
  cross3(s.v + 6 * (i), s.x + 3 * (i), t);
  va[0] = s.v[6 * (i) + 0]; va[1] = s.v[6 * (i) + 1]; va[2] = s.v[6 * (i) + 2];
  vl[0] = s.v[6 * (i) + 3] + t[0]; vl[1] = s.v[6 * (i) + 4] + t[1]; vl[2] = s.v[6 * (i) + 5] + t[2];
}

// Featherstone's articulated-body algorithm (RBDA table 7.1) in world coordinates.
// mass[i], J[i*3..] are link i's; tau may be null; g = gravity; qdd out [ndof].
B2M_HD inline void rc_aba(const RCTree& T, const RCState& s, const double* mass, const double* J, const double* qd, const double* tau,
                          const double* g, double* qdd) {
  const int N = T.n_links;
  double IA[B2M_MAX_LINKS][21], pA[B2M_MAX_LINKS][6], c[B2M_MAX_LINKS][6], U[B2M_MAX_LINKS][6], Dinv[B2M_MAX_LINKS], u[B2M_MAX_LINKS];
  double a[B2M_MAX_LINKS][6];
  for (int i = 1; i < N; i++) {
    double vJ[6], Iv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) vJ[k] = s.S[6 * (i) + k] * qd[i - 1];
    crm(s.v + 6 * (i), vJ, c[i]);
    rc_link_inertia(mass[i], J + 3 * i, s.x + 3 * (i), s.R + 9 * (i), IA[i]);
    sym_mv(IA[i], s.v + 6 * (i), Iv);
    crf(s.v + 6 * (i), Iv, pA[i]);
  }
  for (int i = N - 1; i >= 1; i--) {
    sym_mv(IA[i], s.S + 6 * (i), U[i]);
    const double D = dot6(s.S + 6 * (i), U[i]);
    Dinv[i] = 1.0 / D;
    u[i] = (tau ? tau[i - 1] : 0.0) - dot6(s.S + 6 * (i), pA[i]);
    const int p = T.parent[i];
    if (p != 0) {
      double Ia[21], Iac[6];
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = r; cc < 6; cc++) Ia[sym_idx(r, cc)] = fma(-U[i][r] * Dinv[i], U[i][cc], IA[i][sym_idx(r, cc)]);
      sym_mv(Ia, c[i], Iac);
      const double ud = u[i] * Dinv[i];
#pragma unroll
      for (int k = 0; k < 21; k++) IA[p][k] += Ia[k];
#pragma unroll
      for (int k = 0; k < 6; k++) pA[p][k] += pA[i][k] + Iac[k] + U[i][k] * ud;
    }
  }
  a[0][0] = a[0][1] = a[0][2] = 0.0; a[0][3] = -g[0]; a[0][4] = -g[1]; a[0][5] = -g[2];
  for (int i = 1; i < N; i++) {
    const int p = T.parent[i];
    double ap[6];
#pragma unroll
    for (int k = 0; k < 6; k++) ap[k] = a[p][k] + c[i][k];
    const double qi = (u[i] - dot6(U[i], ap)) * Dinv[i];
    qdd[i - 1] = qi;
#pragma unroll
    for (int k = 0; k < 6; k++) a[i][k] = fma(s.S[6 * i + k], qi, ap[k]);
  }
}

// Joint-space inertia H (ndof x ndof, column-major, leading dimension ld) by the composite-rigid-body algorithm
// (RBDA table 6.2), world coordinates.
B2M_HD inline void rc_crb(const RCTree& T, const RCState& s, const double* mass, const double* J, double* H, int ld) {
  const int N = T.n_links;
  double Ic[B2M_MAX_LINKS][21];
  for (int i = 0; i < N - 1; i++) for (int j = 0; j < N - 1; j++) H[(size_t)j * ld + i] = 0.0;
  for (int i = 1; i < N; i++) rc_link_inertia(mass[i], J + 3 * i, s.x + 3 * (i), s.R + 9 * (i), Ic[i]);
  for (int i = N - 1; i >= 1; i--) {
    const int p = T.parent[i];
    if (p != 0) {
#pragma unroll
      for (int k = 0; k < 21; k++) Ic[p][k] += Ic[i][k];
    }
    double F[6];
    sym_mv(Ic[i], s.S + 6 * (i), F);
    H[(size_t)(i - 1) * ld + (i - 1)] = dot6(s.S + 6 * (i), F);
    for (int j = T.parent[i]; j != 0; j = T.parent[j]) {
      const double h = dot6(F, s.S + 6 * (j));
      H[(size_t)(j - 1) * ld + (i - 1)] = h; H[(size_t)(i - 1) * ld + (j - 1)] = h;
    }
  }
}

// Bias forces C(q, qd) including gravity: recursive Newton-Euler with qdd = 0 (RBDA table 5.1), world coordinates.
B2M_HD inline void rc_bias(const RCTree& T, const RCState& s, const double* mass, const double* J, const double* qd, const double* g, double* C) {
  const int N = T.n_links;
  double a[B2M_MAX_LINKS][6], f[B2M_MAX_LINKS][6];
  a[0][0] = a[0][1] = a[0][2] = 0.0; a[0][3] = -g[0]; a[0][4] = -g[1]; a[0][5] = -g[2];
  for (int i = 1; i < N; i++) {
    const int p = T.parent[i];
    double vJ[6], c[6], I[21], Ia[6], Iv[6], vIv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) vJ[k] = s.S[6 * (i) + k] * qd[i - 1];
    crm(s.v + 6 * (i), vJ, c);
#pragma unroll
    for (int k = 0; k < 6; k++) a[i][k] = a[p][k] + c[k];
    rc_link_inertia(mass[i], J + 3 * i, s.x + 3 * (i), s.R + 9 * (i), I);
    sym_mv(I, a[i], Ia); sym_mv(I, s.v + 6 * (i), Iv); crf(s.v + 6 * (i), Iv, vIv);
#pragma unroll
    for (int k = 0; k < 6; k++) f[i][k] = Ia[k] + vIv[k];
  }
  for (int i = N - 1; i >= 1; i--) {
    C[i - 1] = dot6(s.S + 6 * (i), f[i]);
    const int p = T.parent[i];
    if (p != 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) f[p][k] += f[i][k];
    }
  }
}

// In-place Cholesky solve of the dense SPD system H x = b (column-major, ld), H destroyed.  Returns false if not PD.
B2M_HD inline bool rc_chol_solve(double* H, int n, int ld, double* b) {
  for (int j = 0; j < n; j++) {
    double d = H[(size_t)j * ld + j];
    for (int k = 0; k < j; k++) d = fma(-H[(size_t)k * ld + j], H[(size_t)k * ld + j], d);
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    H[(size_t)j * ld + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = H[(size_t)j * ld + i];
      for (int k = 0; k < j; k++) s = fma(-H[(size_t)k * ld + i], H[(size_t)k * ld + j], s);
      H[(size_t)j * ld + i] = s / d;
    }
  }
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s = fma(-H[(size_t)k * ld + i], b[k], s); b[i] = s / H[(size_t)i * ld + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s = fma(-H[(size_t)i * ld + k], b[k], s); b[i] = s / H[(size_t)i * ld + i]; }
  return true;
}

// Joint torques of the built-in controller (law of example/ur10/controller.cpp:46-96) plus the feed-forward term
B2M_HD inline void rc_controller(const RCTree& T, const double* q, const double* qd, double t, const double* ff, double* tau) {
  for (int k = 0; k < T.n_links - 1; k++) {
    double u = ff ? ff[k] : 0.0;
    if (T.has_ctrl) {
      const double ph = t * T.freq[k];
      u += T.kp[k] * (sin(ph) * T.amp[k] - q[k]) + T.kv[k] * (cos(ph) * T.amp[k] - qd[k]);
    }
    tau[k] = u;
  }
}

// qdd by the selected algorithm (B200MOBY_FDYN_*).  Hwork: ndof*ndof doubles (CRB only).
B2M_HD inline void rc_fwd_dyn(const RCTree& T, int algo, const RCState& s, const double* mass, const double* J, const double* qd,
                              const double* tau, const double* g, double* qdd, double* Hwork) {
  if (algo == FDYN_CRB) {
    const int nd = T.n_links - 1;
    rc_crb(T, s, mass, J, Hwork, nd);
    rc_bias(T, s, mass, J, qd, g, qdd);
    for (int k = 0; k < nd; k++) qdd[k] = (tau ? tau[k] : 0.0) - qdd[k];
    rc_chol_solve(Hwork, nd, nd, qdd);
  } else {
    rc_aba(T, s, mass, J, qd, tau, g, qdd);
  }
}

}  // namespace b2m
