// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
// extern "C" surface so tests / bench.py's cpu_baseline leg can drive the CPU oracle through ctypes.
// It consumes the same plain-C batch descriptor as the product (include/b200moby.h is a data-format
// header only) so one synthetic scene feeds both sides.
#include <cstring>
#include <thread>
#include <vector>
#include "../include/b200moby.h"
#include "oracle_lcp.h"
#include "oracle_sim.h"
#include "oracle_rc.h"
#include "oracle_boxbox.h"

using namespace oracle;

extern "C" {

// ---- LCP solvers.  M column-major n*n.  z: in (warm start when warm!=0) / out.  log may be NULL. Returns 1 on success.
int oracle_lcp_lemke(int n, const double* M, const double* q, double* z, double piv_tol, double zero_tol, int tie,
                     int* pivots, int* status, int* log, int log_cap, int* log_len) {
  LCP lcp; lcp.tie = (TieRule)tie; lcp.keep_log = (log != nullptr);
  Vec zz(n, 0.0);
  bool ok = lcp.lcp_lemke(n, M, q, zz, piv_tol, zero_tol);
  for (int i = 0; i < n; i++) z[i] = (i < (int)zz.size()) ? zz[i] : 0.0;
  if (pivots) *pivots = (int)lcp.pivots;
  if (status) *status = lcp.status;
  if (log) { int m = std::min<int>(log_cap, (int)lcp.log.size()); for (int i = 0; i < m; i++) log[i] = lcp.log[i]; }
  if (log_len) *log_len = (int)lcp.log.size();
  return ok ? 1 : 0;
}

int oracle_lcp_fast(int n, const double* M, const double* q, double* z, int warm, double zero_tol, int tie, int* pivots,
                    int* status, int* log, int log_cap, int* log_len) {
  LCP lcp; lcp.tie = (TieRule)tie; lcp.keep_log = (log != nullptr);
  Vec zz;
  if (warm) zz.assign(z, z + n);
  bool ok = lcp.lcp_fast(n, M, q, zz, zero_tol);
  if (ok) for (int i = 0; i < n; i++) z[i] = zz[i];
  if (pivots) *pivots = (int)lcp.pivots;
  if (status) *status = lcp.status;
  if (log) { int m = std::min<int>(log_cap, (int)lcp.log.size()); for (int i = 0; i < m; i++) log[i] = lcp.log[i]; }
  if (log_len) *log_len = (int)lcp.log.size();
  return ok ? 1 : 0;
}

int oracle_lcp_fast_regularized(int n, const double* M, const double* q, double* z, int warm, int min_exp, int step_exp,
                                int max_exp, double zero_tol, int tie, int* pivots, int* status) {
  LCP lcp; lcp.tie = (TieRule)tie;
  Vec zz;
  if (warm) zz.assign(z, z + n);
  bool ok = lcp.lcp_fast_regularized(n, M, q, zz, min_exp, (unsigned)step_exp, max_exp, -1.0, zero_tol);
  if (zz.size() == (size_t)n) for (int i = 0; i < n; i++) z[i] = zz[i];
  if (pivots) *pivots = (int)lcp.pivots;
  if (status) *status = lcp.status;
  return ok ? 1 : 0;
}

int oracle_lcp_lemke_regularized(int n, const double* M, const double* q, double* z, int min_exp, int step_exp, int max_exp,
                                 double piv_tol, double zero_tol, int tie, int* pivots, int* status) {
  LCP lcp; lcp.tie = (TieRule)tie;
  Vec zz(n, 0.0);
  bool ok = lcp.lcp_lemke_regularized(n, M, q, zz, min_exp, (unsigned)step_exp, max_exp, piv_tol, zero_tol);
  for (int i = 0; i < n; i++) z[i] = (ok && i < (int)zz.size()) ? zz[i] : 0.0;
  if (pivots) *pivots = (int)lcp.pivots;
  if (status) *status = lcp.status;
  return ok ? 1 : 0;
}

// ---- simulator ----
struct OracleSim { Sim sim; };
// checker counters -> the product's counter struct
static void add_counters(b200moby_counters* t, const Counters& k) {
  t->env_steps += k.env_steps; t->mini_steps += k.mini_steps; t->lcp_solves += k.lcp_solves; t->lcp_fast_calls += k.lcp_fast_calls;
  t->lemke_calls += k.lemke_calls; t->pivots += k.pivots; t->lcp_failures += k.lcp_failures; t->impact_tol_events += k.impact_tol_events;
  t->contacts += k.contacts; t->max_lcp_n = std::max<long long>(t->max_lcp_n, k.max_lcp_n); t->pivot_flops += k.pivot_flops; t->ca_iterations += k.ca_iterations;
  t->stab_iterations += k.stab_iterations; t->stab_lcp_solves += k.stab_lcp_solves; t->stab_line_search_failures += k.stab_line_search_failures;
}

static void rc_model_from_desc(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, RCModel& m);

static void fill_from_desc(Sim& S, const b200moby_scene_desc* d, int e) {
  const int nb = d->n_bodies, ne = d->n_envs;
  S.init(nb);
  for (int b = 0; b < nb; b++) {
    Body& B = S.bodies[b];
    B.shape = d->shape[(size_t)b * ne + e];
    B.enabled = d->enabled[(size_t)b * ne + e] != 0;
    B.mass = d->mass[(size_t)b * ne + e];
    for (int k = 0; k < 3; k++) {
      B.dims[k] = d->dims[((size_t)b * 3 + k) * ne + e];
      B.J[k] = d->inertia[((size_t)b * 3 + k) * ne + e];
    }
  }
  for (int i = 0; i < nb; i++)
    for (int j = i + 1; j < nb; j++) {
      ContactParams& c = S.cparams[(size_t)i * nb + j];
      const size_t o = ((size_t)i * nb + j) * ne + e;
      c.mu_c = d->mu_coulomb[o]; c.mu_v = d->mu_viscous[o]; c.eps = d->epsilon[o]; c.compliance = d->compliance[o]; c.NK = d->NK[o];
    }
  S.gravity = V3(d->gravity[0], d->gravity[1], d->gravity[2]);
  S.contact_dist_thresh = d->contact_dist_thresh;
  S.min_step_size = d->min_step_size_env ? d->min_step_size_env[e] : d->min_step_size;
  S.model = d->impact_model;
  S.stab_max_iterations = d->stabilization_max_iterations;
  if (d->rc && d->rc->n_links > 0) {
    const b200moby_rc_desc* r = d->rc;
    S.has_rc = true; S.rc_first = r->first_body; S.rc_fdyn = r->fdyn_algorithm;
    std::vector<double> mass(r->n_links), J(3 * r->n_links);
    for (int i = 0; i < r->n_links; i++) { mass[i] = S.bodies[r->first_body + i].mass; for (int c = 0; c < 3; c++) J[3 * i + c] = S.bodies[r->first_body + i].J[c]; }
    const double pose0[7] = {0, 0, 0, 0, 0, 0, 1};
    rc_model_from_desc(r, mass.data(), J.data(), pose0, S.rc);
    const int nd = r->n_links - 1;
    S.jq.assign(nd, 0.0); S.jqd.assign(nd, 0.0); S.jtau.assign(nd, 0.0);
    if (r->ctrl_kp) {
      S.has_ctrl = true;
      S.ctrl_kp.assign(r->ctrl_kp, r->ctrl_kp + nd);
      S.ctrl_kv.assign(nd, 0.0); S.ctrl_amp.assign(nd, 0.0); S.ctrl_freq.assign(nd, 0.0);
      if (r->ctrl_kv) S.ctrl_kv.assign(r->ctrl_kv, r->ctrl_kv + nd);
      if (r->ctrl_amp) S.ctrl_amp.assign(r->ctrl_amp, r->ctrl_amp + nd);
      if (r->ctrl_freq) S.ctrl_freq.assign(r->ctrl_freq, r->ctrl_freq + nd);
    }
  }
}

// dense helpers, exposed so the tests can pin them against LAPACK (scipy): LinAlgd::solve_fast / factor_chol / inverse_SPD
int oracle_solve_fast(int n, double* A, double* b, int* piv) { return oracle::solve_fast(A, n, b, piv) ? 1 : 0; }
int oracle_factor_chol(int n, double* A) { return oracle::factor_chol(A, n) ? 1 : 0; }
int oracle_inverse_spd(int n, double* A) { return oracle::inverse_SPD(A, n) ? 1 : 0; }

void* oracle_sim_create(const b200moby_scene_desc* d, int env, int tie) {
  OracleSim* o = new OracleSim;
  fill_from_desc(o->sim, d, env);
  o->sim.lcp.tie = (TieRule)tie;
  return o;
}
void oracle_sim_destroy(void* h) { delete (OracleSim*)h; }

// q: [body][7], v: [body][6] (AoS for one env)
static void set_state(Sim& S, const double* q, const double* v) {
  for (size_t b = 0; b < S.bodies.size(); b++) {
    Body& B = S.bodies[b];
    B.x = V3(q[b * 7 + 0], q[b * 7 + 1], q[b * 7 + 2]);
    double nrm = 0; for (int k = 0; k < 4; k++) nrm += q[b * 7 + 3 + k] * q[b * 7 + 3 + k];
    nrm = std::sqrt(nrm);
    for (int k = 0; k < 4; k++) B.quat[k] = q[b * 7 + 3 + k] / nrm;
    S.update_pose(B);
    B.vl = V3(v[b * 6 + 0], v[b * 6 + 1], v[b * 6 + 2]);
    B.va = V3(v[b * 6 + 3], v[b * 6 + 4], v[b * 6 + 5]);
  }
}
void oracle_sim_set_state(void* h, const double* q, const double* v) { set_state(((OracleSim*)h)->sim, q, v); }
static void get_state(Sim& S, double* q, double* v) {
  for (size_t b = 0; b < S.bodies.size(); b++) {
    const Body& B = S.bodies[b];
    q[b * 7 + 0] = B.x.x; q[b * 7 + 1] = B.x.y; q[b * 7 + 2] = B.x.z;
    for (int k = 0; k < 4; k++) q[b * 7 + 3 + k] = B.quat[k];
    v[b * 6 + 0] = B.vl.x; v[b * 6 + 1] = B.vl.y; v[b * 6 + 2] = B.vl.z;
    v[b * 6 + 3] = B.va.x; v[b * 6 + 4] = B.va.y; v[b * 6 + 5] = B.va.z;
  }
}
void oracle_sim_get_state(void* h, double* q, double* v) { get_state(((OracleSim*)h)->sim, q, v); }
// joint state of the articulated body (after oracle_sim_set_state, which places the base link)
void oracle_sim_set_joint_state(void* h, const double* jq, const double* jqd) {
  Sim& S = ((OracleSim*)h)->sim;
  const int nd = S.rc.ndof();
  S.jq.assign(jq, jq + nd); S.jqd.assign(jqd, jqd + nd);
  S.rc_update_links();
}
void oracle_sim_get_joint_state(void* h, double* jq, double* jqd) {
  Sim& S = ((OracleSim*)h)->sim;
  for (int k = 0; k < S.rc.ndof(); k++) { jq[k] = S.jq[k]; jqd[k] = S.jqd[k]; }
}
void oracle_sim_set_joint_forces(void* h, const double* tau) {
  Sim& S = ((OracleSim*)h)->sim;
  for (int k = 0; k < S.rc.ndof(); k++) S.jtau[k] = tau ? tau[k] : 0.0;
}
void oracle_sim_step(void* h, double dt, int n_steps) {
  Sim& S = ((OracleSim*)h)->sim;
  for (int i = 0; i < n_steps; i++) S.step(dt);
}
double oracle_sim_time(void* h) { return ((OracleSim*)h)->sim.current_time; }
void oracle_sim_counters(void* h, b200moby_counters* c) {
  std::memset(c, 0, sizeof(*c));
  add_counters(c, ((OracleSim*)h)->sim.cnt);
}
// LCP of the most recent impact solve: returns n; copies min(n*n, cap) etc.
int oracle_sim_last_lcp(void* h, double* MM, double* qq, double* z, int ncap) {
  Sim& S = ((OracleSim*)h)->sim;
  const int n = S.last_n;
  if (n <= ncap) {
    if (MM) std::memcpy(MM, S.last_MM.data(), sizeof(double) * (size_t)n * n);
    if (qq) std::memcpy(qq, S.last_qq.data(), sizeof(double) * n);
    if (z) std::memcpy(z, S.last_z.data(), sizeof(double) * n);
  }
  return n;
}
// contacts of the most recent mini-step: point, normal, tan1, tan2 as [cap][3]; pair = b1*nb+b2; returns count
int oracle_sim_last_contacts(void* h, int cap, double* point, double* normal, double* tan1, double* tan2, int* pair, double* dist) {
  Sim& S = ((OracleSim*)h)->sim;
  const int nb = (int)S.bodies.size();
  const int m = (int)S.last_contacts.size();
  for (int i = 0; i < m && i < cap; i++) {
    const Contact& c = S.last_contacts[i];
    for (int k = 0; k < 3; k++) { point[i * 3 + k] = c.p[k]; normal[i * 3 + k] = c.n[k]; tan1[i * 3 + k] = c.t1[k]; tan2[i * 3 + k] = c.t2[k]; }
    pair[i] = c.b1 * nb + c.b2; dist[i] = c.dist;
  }
  return m;
}
// Contacts + first-island LCP at the current state without stepping (assembly parity): returns n (0 if no impact)
int oracle_sim_assemble(void* h, double* MM, double* qq, int ncap, int* n_contacts) {
  Sim& S = ((OracleSim*)h)->sim;
  std::vector<std::pair<int, int> > pairs; std::vector<PairDist> pd; std::vector<Contact> cons;
  S.broad_phase(pairs); S.calc_pairwise_distances(pairs, pd); S.find_unilateral_constraints(pd, cons);
  if (n_contacts) *n_contacts = (int)cons.size();
  if (cons.empty()) return 0;
  std::vector<Contact*> cp; std::vector<int> ib;
  for (auto& c : cons) { cp.push_back(&c); ib.push_back(c.b1); ib.push_back(c.b2); }
  int n; Vec M, q;
  S.assemble_island_lcp(cp, ib, n, M, q);
  if (n <= ncap) { std::memcpy(MM, M.data(), sizeof(double) * (size_t)n * n); std::memcpy(qq, q.data(), sizeof(double) * n); }
  return n;
}

// Batch driver for tests / the CPU baseline: steps envs [e0,e1) of the descriptor `n_steps` times with `threads`
// host threads.  q,v use the product's SoA layout ([body][7][env], [body][6][env]) and are updated in place.
void oracle_batch_step(const b200moby_scene_desc* d, double* q, double* v, int e0, int e1, double dt, int n_steps, int tie,
                       int threads, b200moby_counters* total) {
  const int nb = d->n_bodies, ne = d->n_envs;
  if (threads < 1) threads = 1;
  std::vector<b200moby_counters> cs(threads);
  for (auto& c : cs) std::memset(&c, 0, sizeof(c));
  auto work = [&](int t) {
    std::vector<double> qa(nb * 7), va(nb * 6);
    for (int e = e0 + t; e < e1; e += threads) {
      Sim S; fill_from_desc(S, d, e); S.lcp.tie = (TieRule)tie;
      for (int b = 0; b < nb; b++) { for (int k = 0; k < 7; k++) qa[b * 7 + k] = q[((size_t)b * 7 + k) * ne + e]; for (int k = 0; k < 6; k++) va[b * 6 + k] = v[((size_t)b * 6 + k) * ne + e]; }
      set_state(S, qa.data(), va.data());
      for (int i = 0; i < n_steps; i++) S.step(dt);
      get_state(S, qa.data(), va.data());
      for (int b = 0; b < nb; b++) { for (int k = 0; k < 7; k++) q[((size_t)b * 7 + k) * ne + e] = qa[b * 7 + k]; for (int k = 0; k < 6; k++) v[((size_t)b * 6 + k) * ne + e] = va[b * 6 + k]; }
      add_counters(&cs[t], S.cnt);
    }
  };
  if (threads == 1) work(0);
  else { std::vector<std::thread> th; for (int t = 0; t < threads; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
  if (total) {
    std::memset(total, 0, sizeof(*total));
    for (auto& c : cs) {
      total->env_steps += c.env_steps; total->mini_steps += c.mini_steps; total->lcp_solves += c.lcp_solves; total->lcp_fast_calls += c.lcp_fast_calls;
      total->lemke_calls += c.lemke_calls; total->pivots += c.pivots; total->lcp_failures += c.lcp_failures; total->impact_tol_events += c.impact_tol_events;
      total->contacts += c.contacts; total->max_lcp_n = std::max(total->max_lcp_n, c.max_lcp_n); total->pivot_flops += c.pivot_flops; total->ca_iterations += c.ca_iterations;
      total->stab_iterations += c.stab_iterations; total->stab_lcp_solves += c.stab_lcp_solves; total->stab_line_search_failures += c.stab_line_search_failures;
    }
  }
}

// Persistent batch (keeps every env's simulator, including its warm start, across calls): the CPU baseline of bench.py.
struct OracleBatch { std::vector<Sim> sims; int e0; };

void* oracle_batch_create(const b200moby_scene_desc* d, const double* q, const double* v, int e0, int e1, int tie) {
  OracleBatch* B = new OracleBatch;
  B->e0 = e0;
  const int nb = d->n_bodies, ne = d->n_envs;
  B->sims.resize(e1 - e0);
  std::vector<double> qa(nb * 7), va(nb * 6);
  for (int e = e0; e < e1; e++) {
    Sim& S = B->sims[e - e0];
    fill_from_desc(S, d, e); S.lcp.tie = (TieRule)tie;
    for (int b = 0; b < nb; b++) { for (int k = 0; k < 7; k++) qa[b * 7 + k] = q[((size_t)b * 7 + k) * ne + e]; for (int k = 0; k < 6; k++) va[b * 6 + k] = v[((size_t)b * 6 + k) * ne + e]; }
    set_state(S, qa.data(), va.data());
  }
  return B;
}
// joint state of the articulated body for every env of the batch, SoA [dof][n_envs of the descriptor]
void oracle_batch_set_joint_state(void* h, const double* jq, const double* jqd, int ne) {
  OracleBatch* B = (OracleBatch*)h;
  for (size_t i = 0; i < B->sims.size(); i++) {
    Sim& S = B->sims[i];
    if (!S.has_rc) continue;
    const int e = B->e0 + (int)i, nd = S.rc.ndof();
    for (int k = 0; k < nd; k++) { S.jq[k] = jq[(size_t)k * ne + e]; S.jqd[k] = jqd[(size_t)k * ne + e]; }
    S.rc_update_links();
  }
}
void oracle_batch_get_joint_state(void* h, int i, double* jq, double* jqd) {
  Sim& S = ((OracleBatch*)h)->sims[i];
  for (int k = 0; k < S.rc.ndof(); k++) { jq[k] = S.jq[k]; jqd[k] = S.jqd[k]; }
}
void oracle_batch_destroy(void* h) { delete (OracleBatch*)h; }
void oracle_batch_run(void* h, double dt, int n_steps, int threads, b200moby_counters* total) {
  OracleBatch* B = (OracleBatch*)h;
  if (threads < 1) threads = 1;
  const int n = (int)B->sims.size();
  auto work = [&](int t) { for (int i = t; i < n; i += threads) for (int s = 0; s < n_steps; s++) B->sims[i].step(dt); };
  if (threads == 1) work(0);
  else { std::vector<std::thread> th; for (int t = 0; t < threads; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
  if (total) {
    std::memset(total, 0, sizeof(*total));
    for (auto& S : B->sims) add_counters(total, S.cnt);
  }
}
// state of env (e0 + i) as AoS [body][7], [body][6]
void oracle_batch_get_state(void* h, int i, double* q, double* v) { get_state(((OracleBatch*)h)->sims[i], q, v); }
// state of every env of the batch in the product's SoA layout: q [body][7][n], v [body][6][n] with n = e1 - e0
void oracle_batch_get_state_soa(void* h, double* q, double* v) {
  OracleBatch* B = (OracleBatch*)h;
  const int n = (int)B->sims.size();
  if (!n) return;
  const int nb = (int)B->sims[0].bodies.size();
  std::vector<double> qa(nb * 7), va(nb * 6);
  for (int i = 0; i < n; i++) {
    get_state(B->sims[i], qa.data(), va.data());
    for (int b = 0; b < nb; b++) { for (int k = 0; k < 7; k++) q[((size_t)b * 7 + k) * n + i] = qa[b * 7 + k]; for (int k = 0; k < 6; k++) v[((size_t)b * 6 + k) * n + i] = va[b * 6 + k]; }
  }
}
// per-env solver statistics since creation, stat [5][n]: LCP failures, lcp_lemke calls, lcp_fast calls, LCP solves, pivots
// (the checker's side of b200moby_get_env_stats)
void oracle_batch_env_stats(void* h, int* stat) {
  OracleBatch* B = (OracleBatch*)h;
  const size_t n = B->sims.size();
  for (size_t i = 0; i < n; i++) {
    const Counters& k = B->sims[i].cnt;
    stat[i] = (int)k.lcp_failures; stat[n + i] = (int)k.lemke_calls; stat[2 * n + i] = (int)k.lcp_fast_calls; stat[3 * n + i] = (int)k.lcp_solves; stat[4 * n + i] = (int)k.pivots;
  }
}

// box-box signed distance of oracle_boxbox.h for two posed boxes: box = centre[3], rotation R[9] row-major, edge lengths[3];
// out = dist, pA[3], pB[3]
void oracle_boxbox_dist(const double* A, const double* B, double* out) {
  bb::Box X[2];
  const double* src[2] = {A, B};
  for (int b = 0; b < 2; b++) {
    X[b].c = {src[b][0], src[b][1], src[b][2]};
    const double* R = src[b] + 3;
    for (int k = 0; k < 3; k++) { X[b].ax[k] = {R[k], R[3 + k], R[6 + k]}; X[b].ext[k] = src[b][12 + k]; }
  }
  bb::Vec3 pA{0, 0, 0}, pB{0, 0, 0};
  bb::signed_dist(X[0], X[1], out[0], pA, pB);
  out[1] = pA.x; out[2] = pA.y; out[3] = pA.z; out[4] = pB.x; out[5] = pB.y; out[6] = pB.z;
}

// ---- reduced-coordinate articulated body (oracle_rc.h) ----
// mass [link], J [link][3], base_pose [7] = x y z qx qy qz qw; the tree comes from the product's plain-C descriptor.
static void rc_model_from_desc(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, RCModel& m) {
  m.n_links = r->n_links;
  for (int i = 0; i < r->n_links; i++) {
    m.mass[i] = mass[i];
    for (int c = 0; c < 3; c++) m.J[i][c] = J[3 * i + c];
    if (i == 0) continue;
    m.parent[i] = r->parent[i]; m.jtype[i] = r->joint_type[i];
    double an = 0, qn = 0;
    for (int c = 0; c < 3; c++) an += r->joint_axis[3 * i + c] * r->joint_axis[3 * i + c];
    for (int c = 0; c < 4; c++) qn += r->rel_quat[4 * i + c] * r->rel_quat[4 * i + c];
    an = std::sqrt(an); qn = std::sqrt(qn);
    for (int c = 0; c < 3; c++) { m.axis[i][c] = r->joint_axis[3 * i + c] / an; m.loc_parent[i][c] = r->loc_parent[3 * i + c]; m.loc_child[i][c] = r->loc_child[3 * i + c]; }
    for (int c = 0; c < 4; c++) m.rel_quat[i][c] = r->rel_quat[4 * i + c] / qn;
  }
  double qn = 0; for (int c = 0; c < 4; c++) qn += base_pose[3 + c] * base_pose[3 + c];
  qn = std::sqrt(qn);
  for (int c = 0; c < 3; c++) m.base_x[c] = base_pose[c];
  for (int c = 0; c < 4; c++) m.base_quat[c] = base_pose[3 + c] / qn;
}
// algo: 0 = ABA (fsab), 1 = CRB + Cholesky.  Returns 1 on success.
int oracle_rc_fwd_dyn(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, const double* g,
                      int algo, const double* q, const double* qd, const double* tau, double* qdd) {
  RCModel m; rc_model_from_desc(r, mass, J, base_pose, m);
  std::vector<double> t(m.ndof(), 0.0);
  if (tau) t.assign(tau, tau + m.ndof());
  if (algo == 1) return rc_crb_fwd_dyn(m, q, qd, t.data(), g, qdd) ? 1 : 0;
  rc_aba(m, q, qd, t.data(), g, qdd);
  return 1;
}
// H: ndof x ndof column-major
void oracle_rc_inertia(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, const double* q, double* H) {
  RCModel m; rc_model_from_desc(r, mass, J, base_pose, m);
  std::vector<double> qd(m.ndof(), 0.0);
  RCKin k; rc_kinematics(m, q, qd.data(), k);
  rc_crb(m, k, H);
}
// link world poses x [link][3], R [link][9] (row-major), COM linear / angular velocity [link][3], Jacobians [link][6*ndof]
void oracle_rc_links(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, const double* q,
                     const double* qd, double* x, double* R, double* vl, double* va, double* jac) {
  RCModel m; rc_model_from_desc(r, mass, J, base_pose, m);
  RCKin k; rc_kinematics(m, q, qd, k);
  for (int i = 0; i < m.n_links; i++) {
    for (int c = 0; c < 3; c++) { x[3 * i + c] = k.x[i][c]; vl[3 * i + c] = k.vl[i][c]; va[3 * i + c] = k.va[i][c]; }
    for (int c = 0; c < 9; c++) R[9 * i + c] = k.R[i][c];
    if (jac) rc_link_jacobian(m, k, i, jac + (size_t)i * 6 * m.ndof());
  }
}
double oracle_rc_energy(const b200moby_rc_desc* r, const double* mass, const double* J, const double* base_pose, const double* g,
                        const double* q, const double* qd) {
  RCModel m; rc_model_from_desc(r, mass, J, base_pose, m);
  return rc_energy(m, q, qd, g);
}

}  // extern "C"
