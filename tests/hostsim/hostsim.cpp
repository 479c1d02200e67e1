// TEST INFRASTRUCTURE ONLY -- never part of libb200moby.so and not reachable from the C ABI.
// Compiles the kernels' device code (moby_b200/csrc/*.cuh) for the host with a single-thread "group" so the
// kernel logic can be checked against the oracle in the CPU test suite, where no GPU exists.  Floating-point
// results are the ones the GPU produces (the .cu files are built with -fmad=false, this file with
// -ffp-contract=off, and both use explicit fma at the same places).
#include <cstring>
#include <vector>
#include "../../include/b200moby.h"
#include "../../moby_b200/csrc/friction_table.h"
#include "../../moby_b200/csrc/sim_device.cuh"
#include "../../moby_b200/csrc/rc_host.h"

using namespace b2m;

static int* g_env_stat = nullptr;   // optional [5][n_envs] per-env solver statistics (SimParams::env_stat), set by hostsim_set_env_stat

// SimParams over host arrays, exactly as b200moby_create fills it
static bool setup_params(SimParams& P, RCTree& tree, std::vector<double>& tau0, const std::vector<double>& tab, const b200moby_scene_desc* d, int cmax,
                         int nmax, int npmax, double* q, double* v, double* time, double* zlast, int* zlast_n, unsigned long long* counters,
                         double* jq, double* jqd) {
  const int ne = d->n_envs, nb = d->n_bodies;
  memset(&P, 0, sizeof(P));
  P.n_envs = ne; P.nb = nb; P.cmax = cmax; P.nmax = nmax; P.npmax = npmax; P.model = d->impact_model;
  P.shape = d->shape; P.enabled = d->enabled; P.mass = d->mass; P.dims = d->dims; P.inertia = d->inertia;
  P.mu_c = d->mu_coulomb; P.mu_v = d->mu_viscous; P.eps = d->epsilon; P.compliance = d->compliance; P.NK = d->NK;
  P.fr_tab = tab.data(); P.gx = d->gravity[0]; P.gy = d->gravity[1]; P.gz = d->gravity[2];
  P.contact_dist_thresh = d->contact_dist_thresh; P.min_step_size = d->min_step_size; P.min_step_env = d->min_step_size_env;
  P.stab_max_iterations = d->stabilization_max_iterations; P.stab_eps = B2M_NEAR_ZERO;
  P.q = q; P.v = v; P.time = time; P.zlast = zlast; P.zlast_n = zlast_n; P.counters = counters;
  P.vlast = zlast + (size_t)nmax * ne; P.vlast_n = zlast_n + ne;   // the caller's arrays carry both warm starts
  P.env_stat = g_env_stat;
  if (d->rc && d->rc->n_links > 0) {
    bool unsup;
    if (b2m_rc_tree_from_desc(*d->rc, nb, tree, &unsup)) return false;
    P.rc = &tree; P.rc_links = tree.n_links; P.rc_first = tree.first_body; P.ngc = b2m_dense_ngc(d);
    P.jq = jq; P.jqd = jqd;
    tau0.assign((size_t)(tree.n_links - 1) * ne, 0.0);
    P.jtau = tau0.data();
  }
  return true;
}

// the stabilization phase of one env, as k_stabilize.cu runs it (working set with nmax = cmax + the stabilization extras)
static void stabilize_env_host(const SimParams& P, int e, unsigned long long* lc) {
  if (P.stab_max_iterations == 0) return;
  SimParams Ps = P; Ps.nmax = Ps.cmax;
  const EnvDims D = env_dims(Ps);
  std::vector<double> wd(env_doubles(D) + stab_extra_doubles(D));
  std::vector<int> wi(env_ints(D) + stab_extra_ints(D));
  EnvMem m; StabMem s;
  env_carve(m, wd.data(), wi.data(), D);
  stab_carve(s, wd.data() + env_doubles(D), wi.data() + env_ints(D), D);
  SerialGroup g(nullptr);
  env_load(g, Ps, e, m);
  const EnvStatBase sb = env_stat_base(lc);
  env_stabilize(g, Ps, e, m, s, lc);
  env_stat_commit(g, Ps, e, lc, sb);
  env_store(g, Ps, e, m, ST_POS);
}

extern "C" {

void hostsim_set_env_stat(int* stat) { g_env_stat = stat; }

// q [nb][7][ne], v [nb][6][ne], time [ne], zlast [2 nmax][ne] (QP warm start, then the no-slip one), zlast_n [2][ne], counters [CNT_COUNT] all host, updated in place.
// Returns nmax (call with q == NULL to query sizes only).
int hostsim_run(const b200moby_scene_desc* d, double* q, double* v, double* time, double* zlast, int* zlast_n,
                unsigned long long* counters, double dt, int n_steps, int e0, int e1, double* tapMM, double* tapqq, double* tapz, int* tapn,
                double* jq, double* jqd) {
  const int ne = d->n_envs, nb = d->n_bodies;
  int cmax = 0, nmax = 0, npmax = 0;
  if (b2m_scene_bounds(d, cmax, nmax, npmax)) return -1;
  if (!q) return nmax;
  std::vector<double> tab = b2m_friction_table();
  SimParams P; RCTree tree; std::vector<double> tau0;
  if (!setup_params(P, tree, tau0, tab, d, cmax, nmax, npmax, q, v, time, zlast, zlast_n, counters, jq, jqd)) return -1;
  P.tap_MM = tapMM; P.tap_qq = tapqq; P.tap_z = tapz; P.tap_n = tapn;
  const EnvDims D = env_dims(P);
  std::vector<double> wd(env_doubles(D));
  std::vector<int> wi(env_ints(D));
  EnvMem m; env_carve(m, wd.data(), wi.data(), D);
  SerialGroup g(nullptr);
  unsigned long long lc[CNT_COUNT]; memset(lc, 0, sizeof(lc));
  EnvCtx cx; cx.limit = false; cx.budget = 0;
  for (int e = e0; e < e1; e++) {
    if (P.stab_max_iterations == 0) { env_run(g, P, e, m, dt, n_steps, lc, cx); continue; }
    for (int s = 0; s < n_steps; s++) { env_run(g, P, e, m, dt, 1, lc, cx); stabilize_env_host(P, e, lc); }   // TimeSteppingSimulator.cpp:95-98
  }
  for (int k = 0; k < CNT_COUNT; k++) { if (k == CNT_MAX_N) counters[k] = std::max(counters[k], lc[k]); else counters[k] += lc[k]; }
  return nmax;
}

// The phased step (advance -> impact per LCP class -> ... -> finish) exactly as sim_kernels.cu's launch_step sequences
// it, each "kernel" a serial loop over its queue.  pivot_budget > 0 exercises the straggler path.
int hostsim_run_phased(const b200moby_scene_desc* d, double* q, double* v, double* time, double* zlast, int* zlast_n,
                       unsigned long long* counters, double dt, int n_steps, int rounds, int pivot_budget, double* jq, double* jqd) {
  const int ne = d->n_envs, nb = d->n_bodies;
  int cmax = 0, nmax = 0, npmax = 0;
  if (b2m_scene_bounds(d, cmax, nmax, npmax)) return -1;
  std::vector<double> tab = b2m_friction_table();
  SimParams P; RCTree tree; std::vector<double> tau0;
  if (!setup_params(P, tree, tau0, tab, d, cmax, nmax, npmax, q, v, time, zlast, zlast_n, counters, jq, jqd)) return -1;
  std::vector<double> hacc(ne), hpend(ne);
  std::vector<int> queue((size_t)B2M_ROUNDS_MAX * B2M_SLOTS * ne), qctl(2 * B2M_ROUNDS_MAX * (B2M_SLOTS + 1));
  P.hacc = hacc.data(); P.hpend = hpend.data(); P.queue = queue.data(); P.qctl = qctl.data();
  P.pivot_budget = pivot_budget;
  std::vector<int> cost(ne, 0);                 // longest-job-first queue, with a low threshold so the tests exercise it
  P.cost = cost.data(); P.hard_cost = 8; P.cost_shift = 2;
  P.n_classes = b2m_class_table(nmax, cmax, P.model, B2M_MAX_CLASSES, P.class_nmax, P.class_cmax);
  const EnvDims D = env_dims(P);
  std::vector<double> wd(env_doubles(D));
  std::vector<int> wi(env_ints(D));
  SerialGroup g(nullptr);
  unsigned long long tot[CNT_COUNT]; memset(tot, 0, sizeof(tot));
  auto add = [&](const unsigned long long* lc) { for (int k = 0; k < CNT_COUNT; k++) { if (k == CNT_MAX_N) tot[k] = std::max(tot[k], lc[k]); else tot[k] += lc[k]; } };
  for (int s = 0; s < n_steps; s++) {
    std::fill(qctl.begin(), qctl.end(), 0);
    for (int r = 0; r < rounds; r++) {
      {   // advance
        EnvMem m; env_carve_small(m, wd.data(), wi.data(), D);
        const int count = (r == 0) ? ne : *q_count(P, r - 1, B2M_SLOT_CONT);
        const int* list = (r == 0) ? nullptr : q_list(P, r - 1, B2M_SLOT_CONT);
        for (int i = 0; i < count; i++) { unsigned long long lc[CNT_COUNT] = {0}; env_advance(g, P, list ? list[i] : i, m, dt, r, lc); add(lc); }
      }
      {   // hard queue first, full working set, no budget
        EnvMem m; env_carve(m, wd.data(), wi.data(), D);
        const int count = q_size(P, r, B2M_SLOT_HARD);
        for (int i = 0; i < count; i++) {
          unsigned long long lc[CNT_COUNT] = {0};
          EnvCtx cx; cx.limit = false; cx.budget = 0;
          env_impact(g, P, q_at(P, r, B2M_SLOT_HARD, i), m, dt, r, lc, cx); add(lc);
        }
      }
      for (int c = 0; c < P.n_classes; c++) {   // impact, per class, with the class's working-set size
        SimParams Pc = P; Pc.cmax = P.class_cmax[c]; Pc.nmax = P.class_nmax[c];
        EnvMem m; env_carve(m, wd.data(), wi.data(), env_dims(Pc));
        const int count = *q_count(P, r, c);
        const int* list = q_list(P, r, c);
        for (int i = 0; i < count; i++) {
          unsigned long long lc[CNT_COUNT] = {0};
          EnvCtx cx; cx.limit = pivot_budget > 0; cx.budget = pivot_budget;
          if (env_impact(g, Pc, list[i], m, dt, r, lc, cx)) add(lc);
        }
      }
      if (pivot_budget > 0) {   // stragglers, full working set
        EnvMem m; env_carve(m, wd.data(), wi.data(), D);
        const int count = *q_count(P, r, B2M_SLOT_STRAGGLER);
        const int* list = q_list(P, r, B2M_SLOT_STRAGGLER);
        for (int i = 0; i < count; i++) {
          unsigned long long lc[CNT_COUNT] = {0};
          EnvCtx cx; cx.limit = false; cx.budget = 0;
          env_impact(g, P, list[i], m, dt, r, lc, cx); add(lc);
        }
      }
    }
    {   // finish
      EnvMem m; env_carve(m, wd.data(), wi.data(), D);
      const int count = *q_count(P, rounds - 1, B2M_SLOT_CONT);
      const int* list = q_list(P, rounds - 1, B2M_SLOT_CONT);
      for (int i = 0; i < count; i++) { unsigned long long lc[CNT_COUNT] = {0}; env_finish(g, P, list[i], m, dt, lc); add(lc); }
    }
    for (int e = 0; e < ne; e++) { unsigned long long lc[CNT_COUNT] = {0}; stabilize_env_host(P, e, lc); add(lc); }   // stabilization phase (k_stabilize.cu)
  }
  for (int k = 0; k < CNT_COUNT; k++) { if (k == CNT_MAX_N) counters[k] = std::max(counters[k], tot[k]); else counters[k] += tot[k]; }
  return nmax;
}

// LCP solvers through the same device code (serial group)
int hostsim_lcp(int mode, int n, const double* M, const double* q, double* z, int warm, double piv_tol, double zero_tol,
                int min_exp, int step_exp, int max_exp, int* pivots, int* log, int log_cap, int* log_len) {
  std::vector<double> wd(std::max(lemke_work_doubles(n), fast_work_doubles(n)) + 8);
  std::vector<int> wi(std::max(lemke_work_ints(n), fast_work_ints(n)) + 8);
  SerialGroup g(nullptr);
  int piv = 0, nlog = 0, st;
  if (!warm) for (int i = 0; i < n; i++) z[i] = 0.0;
  std::vector<double> zz(z, z + n);
  if (mode == 0) st = lemke_solve(g, n, M, n, q, 0.0, piv_tol, zero_tol, zz.data(), wd.data(), wi.data(), &piv, log, log_cap, &nlog);
  else if (mode == 1) st = lcp_fast_solve(g, n, M, n, q, 0.0, zero_tol, warm != 0, zz.data(), wd.data(), wi.data(), &piv, log, log_cap, &nlog);
  else if (mode == 2) st = lcp_lemke_regularized(g, n, M, n, q, piv_tol, zero_tol, min_exp, step_exp, max_exp, zz.data(), wd.data(), wi.data(), &piv, nullptr);
  else st = lcp_fast_regularized(g, n, M, n, q, zero_tol, warm != 0, min_exp, step_exp, max_exp, zz.data(), wd.data(), wi.data(), &piv, nullptr);
  const bool ok = (st == LCP_OK || st == LCP_TRIVIAL || st >= LCP_REGULARIZED);
  if (ok || mode == 0 || mode == 2) for (int i = 0; i < n; i++) z[i] = zz[i];
  if (pivots) *pivots = piv;
  if (log_len) *log_len = nlog;
  return st;
}

// box-box signed distance of boxbox_device.cuh for two posed boxes (same argument layout as the checker's oracle_boxbox_dist)
void hostsim_boxbox_dist(const double* A, const double* B, double* out) {
  BodyRef X[2];
  const double* src[2] = {A, B};
  for (int b = 0; b < 2; b++) { X[b].x = src[b]; X[b].R = src[b] + 3; X[b].dims = src[b] + 12; X[b].vl = X[b].va = src[b]; X[b].shape = SH_BOX; X[b].enabled = 1; X[b].wtab = nullptr; }
  double dist; V3 pA, pB;
  boxbox_signed_dist(X[0], X[1], dist, pA, pB);
  out[0] = dist; out[1] = pA.x; out[2] = pA.y; out[3] = pA.z; out[4] = pB.x; out[5] = pB.y; out[6] = pB.z;
}

// ---- articulated body: the device functions of rc_device.cuh on the host ----
// mass [link], J [link][3], base_pose [7]; what: 0 ABA qdd, 1 CRB qdd, 2 H (ndof*ndof col-major) -> out; links (optional):
// x [link][3], quat [link][4], vl [link][3], va [link][3]
int hostsim_rc(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, const double* g, int what,
               const double* q, const double* qd, const double* tau, double* out, double* lx, double* lquat, double* lvl, double* lva) {
  RCTree T; bool unsup;
  if (b2m_rc_tree_from_desc(*r, r->first_body + r->n_links, T, &unsup)) return -1;
  RCLocal loc; RCState s = loc.view();
  double qt[4], nrm = 0;
  for (int c = 0; c < 4; c++) nrm += base_pose[3 + c] * base_pose[3 + c];
  nrm = std::sqrt(nrm);
  for (int c = 0; c < 3; c++) s.x[c] = base_pose[c];
  for (int c = 0; c < 4; c++) qt[c] = base_pose[3 + c] / nrm;
  quat_to_R(qt, s.R);
  rc_kinematics(T, q, qd, s);
  const int nd = T.n_links - 1;
  std::vector<double> H((size_t)nd * nd);
  if (what == 0 || what == 1) rc_fwd_dyn(T, what, s, mass, J, qd, tau, g, out, H.data());
  else if (what == 2) rc_crb(T, s, mass, J, out, nd);
  if (lx) for (int i = 0; i < T.n_links; i++) {
    for (int c = 0; c < 3; c++) lx[3 * i + c] = s.x[3 * i + c];
    R_to_quat(s.R + 9 * i, lquat + 4 * i);
    rc_link_velocity(s, i, lvl + 3 * i, lva + 3 * i);
  }
  return 0;
}

}  // extern "C"
