// Articulated-body stage kernels: one THREAD per env, SoA state so a warp's loads and stores are contiguous, the spatial
// recursions of rc_device.cuh in registers / thread-local memory.
#include "sim_kernel_util.cuh"
using namespace b2m;

namespace {

// pose of the base link and the links' mass properties for env e
__device__ __forceinline__ void rc_load_env(const SimParams& P, const RCTree& T, int e, RCState& s, double* mass, double* J) {   // s: a view
  const size_t ne = P.n_envs;
  const int b0 = T.first_body;
  double qt[4];
  for (int c = 0; c < 3; c++) s.x[c] = P.q[((size_t)b0 * 7 + c) * ne + e];
  for (int c = 0; c < 4; c++) qt[c] = P.q[((size_t)b0 * 7 + 3 + c) * ne + e];
  quat_to_R(qt, s.R);
  for (int i = 0; i < T.n_links; i++) {
    mass[i] = P.mass[(size_t)(b0 + i) * ne + e];
    for (int c = 0; c < 3; c++) J[3 * i + c] = P.inertia[((size_t)(b0 + i) * 3 + c) * ne + e];
  }
}

}  // namespace

// qdd = forward dynamics(q, qd, tau) by ABA or CRB
__global__ void __launch_bounds__(128) rc_fwd_dyn_kernel(SimParams P, int algo, const double* jq, const double* jqd, const double* tau, double* qdd) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n_envs) return;
  const RCTree& T = *P.rc;
  const size_t ne = P.n_envs;
  const int nd = T.n_links - 1;
  RCLocal loc; RCState s = loc.view();
  double mass[B2M_MAX_LINKS], J[3 * B2M_MAX_LINKS], q[B2M_MAX_LINKS], qd[B2M_MAX_LINKS], tq[B2M_MAX_LINKS], out[B2M_MAX_LINKS];
  double H[(B2M_MAX_LINKS - 1) * (B2M_MAX_LINKS - 1)];
  rc_load_env(P, T, e, s, mass, J);
  for (int k = 0; k < nd; k++) { q[k] = jq[(size_t)k * ne + e]; qd[k] = jqd[(size_t)k * ne + e]; tq[k] = tau ? tau[(size_t)k * ne + e] : 0.0; }
  rc_kinematics(T, q, qd, s);
  const double g[3] = {P.gx, P.gy, P.gz};
  rc_fwd_dyn(T, algo, s, mass, J, qd, tq, g, out, H);
  for (int k = 0; k < nd; k++) qdd[(size_t)k * ne + e] = out[k];
}

// H(q), [dof*dof][env] column-major
__global__ void __launch_bounds__(128) rc_inertia_kernel(SimParams P, const double* jq, double* Hout) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n_envs) return;
  const RCTree& T = *P.rc;
  const size_t ne = P.n_envs;
  const int nd = T.n_links - 1;
  RCLocal loc; RCState s = loc.view();
  double mass[B2M_MAX_LINKS], J[3 * B2M_MAX_LINKS], q[B2M_MAX_LINKS], qd[B2M_MAX_LINKS];
  double H[(B2M_MAX_LINKS - 1) * (B2M_MAX_LINKS - 1)];
  rc_load_env(P, T, e, s, mass, J);
  for (int k = 0; k < nd; k++) { q[k] = jq[(size_t)k * ne + e]; qd[k] = 0.0; }
  rc_kinematics(T, q, qd, s);
  rc_crb(T, s, mass, J, H, nd);
  for (int k = 0; k < nd * nd; k++) Hout[(size_t)k * ne + e] = H[k];
}

// link poses and velocities (P.q, P.v rows of the link bodies) from the joint state
__global__ void __launch_bounds__(128) rc_refresh_links_kernel(SimParams P) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n_envs) return;
  const RCTree& T = *P.rc;
  const size_t ne = P.n_envs;
  const int nd = T.n_links - 1, b0 = T.first_body;
  RCLocal loc; RCState s = loc.view();
  double mass[B2M_MAX_LINKS], J[3 * B2M_MAX_LINKS], q[B2M_MAX_LINKS], qd[B2M_MAX_LINKS];
  rc_load_env(P, T, e, s, mass, J);
  for (int k = 0; k < nd; k++) { q[k] = P.jq[(size_t)k * ne + e]; qd[k] = P.jqd[(size_t)k * ne + e]; }
  rc_kinematics(T, q, qd, s);
  for (int i = 1; i < T.n_links; i++) {
    double qt[4], vl[3], va[3];
    R_to_quat(s.R + 9 * i, qt);
    rc_link_velocity(s, i, vl, va);
    for (int c = 0; c < 3; c++) P.q[((size_t)(b0 + i) * 7 + c) * ne + e] = s.x[3 * i + c];
    for (int c = 0; c < 4; c++) P.q[((size_t)(b0 + i) * 7 + 3 + c) * ne + e] = qt[c];
    for (int c = 0; c < 3; c++) { P.v[((size_t)(b0 + i) * 6 + c) * ne + e] = vl[c]; P.v[((size_t)(b0 + i) * 6 + 3 + c) * ne + e] = va[c]; }
  }
}

const void* b2m_k_rc_fwd_dyn() { return (const void*)rc_fwd_dyn_kernel; }
const void* b2m_k_rc_inertia() { return (const void*)rc_inertia_kernel; }
const void* b2m_k_rc_refresh() { return (const void*)rc_refresh_links_kernel; }
