// Constraint stabilization for one env, group-cooperative: ConstraintStabilization::stabilize (Moby
// src/ConstraintStabilization.cpp:167-254), called by TimeSteppingSimulator::step after the Euler step
// (src/TimeSteppingSimulator.cpp:95-98) unless constraint-stabilization-max-iterations = 0.
//
// While some pair of geometries is closer than eps = +sqrt(eps_machine) (rule H7: the sign the code has, :58-59,197):
//   zero the velocities; one contact constraint at the closest points of every separated pair, the narrowphase's contacts
//   (TOL = sqrt(eps), CollisionDetection.h:46) of every touching pair (:304-345); per island a frictionless nc x nc LCP
//   Cn X Cn^T z + (dist - |eps| - sqrt(eps)) >= 0 by lcp_fast (cold), else lcp_lemke_regularized (:932-970); dq = the
//   bodies' velocity after applying z, in Euler coordinates; a line search along dq -- Ridders' root finder on every
//   pairwise distance that changes sign, then backtracking by 0.6 while a non-bracketed distance is negative and got
//   worse (:1055-1212) -- moves the bodies.
// Velocities are restored at the end (:78-85).  No implicit joints and no joint limits.  Same
// arithmetic, same order as the CPU checker's restatement of stabilize(): bit-identical positions.
// Included at the end of sim_device.cuh (inside namespace b2m).
#pragma once

// extra working set of the stabilization phase, carved after an env's full working set
struct StabMem {
  double *vls, *vas, *jqds;      // saved velocities
  double *q, *dq;                // [body][7] Euler coordinates of the free bodies and the step along them
  double *qj, *dqj;              // joint coordinates of the articulated body
  double *uC, *uC0, *uC1;        // pairwise distances: current trial, before the step, at t = 1
  int* bracket;
};
B2M_HD inline size_t stab_extra_doubles(const EnvDims& d) { const size_t nd = d.rcl ? d.rcl - 1 : 0; return (size_t)20 * d.nb + 3 * nd + 3 * (size_t)d.npmax; }
B2M_HD inline size_t stab_extra_ints(const EnvDims& d) { return (size_t)d.npmax; }
B2M_HD inline void stab_carve(StabMem& s, double* d, int* i, const EnvDims& D) {
  const size_t nb = D.nb, nd = D.rcl ? D.rcl - 1 : 0, np = D.npmax;
  s.vls = d; d += 3 * nb; s.vas = d; d += 3 * nb; s.jqds = d; d += nd;
  s.q = d; d += 7 * nb; s.dq = d; d += 7 * nb; s.qj = d; d += nd; s.dqj = d; d += nd;
  s.uC = d; d += np; s.uC0 = d; d += np; s.uC1 = d; d += np;
  s.bracket = i;
}
// The stabilization LCP has one row per contact: its working set is the env's with nmax = cmax.
B2M_HD inline EnvDims stab_dims(const SimParams& P) { EnvDims d = env_dims(P); d.nmax = d.cmax; return d; }

// update_body_configurations(q + t dq) (:1252-1264); set_generalized_coordinates_euler normalises the quaternion
template <class G>
B2M_DEV B2M_NOINL void stab_set(const G& g, const SimParams& P, EnvMem& m, const StabMem& s, double t) {
  for (int b = g.tid; b < P.nb; b += G::size) {
    if (!m.ben[b] || is_link(P, b)) continue;
    double c[7];
    for (int k = 0; k < 7; k++) c[k] = s.dq[7 * b + k] * t + s.q[7 * b + k];
    m.bx[3 * b] = c[0]; m.bx[3 * b + 1] = c[1]; m.bx[3 * b + 2] = c[2];
    const double nrm = sqrt(c[3] * c[3] + c[4] * c[4] + c[5] * c[5] + c[6] * c[6]);
    double* qt = m.bq + 4 * b;
    qt[0] = c[3] / nrm; qt[1] = c[4] / nrm; qt[2] = c[5] / nrm; qt[3] = c[6] / nrm;
    quat_to_R(qt, m.bR + 9 * b);
  }
  if (B2M_RC(P)) {
    for (int k = g.tid; k < P.rc_links - 1; k += G::size) m.jq[k] = s.dqj[k] * t + s.qj[k];
    g.sync();
    rc_refresh(g, P, m);
  }
  g.sync();
}

// evaluate_unilateral_constraints (:88-131): pairwise distances at the current configuration into out[]; returns the smallest
template <class G>
B2M_DEV B2M_NOINL double stab_eval(const G& g, EnvMem& m, double* out) {
  calc_pairwise_distances(g, m);
  const int np = m.scal[S_NPAIRS];
  double vio = B2M_INF;
  for (int p = g.tid; p < np; p += G::size) { const double d = m.pd_dist[p]; out[p] = d; vio = fmin(vio, d); }
  vio = g.min(vio);
  g.sync();
  return vio;
}

B2M_HD B2M_INL double stab_sign(double x, double y) { return (y > 0.0) ? fabs(x) : -fabs(x); }

// ridders_unilateral (:1322-1380), literally (the caller passes x2 = the current t with fh = the value at t = 1)
template <class G>
B2M_DEV B2M_NOINL double stab_ridders(const G& g, const SimParams& P, EnvMem& m, const StabMem& s, double x1, double x2, double fl, double fh, int idx) {
  const int MAX_ITERATIONS = 25;
  const double TOL = 1e-4;
  double ans = B2M_INF, fm, fnew, sq, xh, xl, xm, xnew;
  if ((fl > 0.0 && fh < 0.0) || (fl < 0.0 && fh > 0.0)) {
    xl = x1; xh = x2;
    for (int j = 0; j < MAX_ITERATIONS; j++) {
      xm = 0.5 * (xl + xh);
      stab_set(g, P, m, s, xm); stab_eval(g, m, s.uC); fm = s.uC[idx];
      g.sync();
      sq = sqrt(fm * fm - fl * fh);
      if (sq == 0.0) return ans;
      xnew = xm + (xm - xl) * ((fl >= fh ? 1.0 : -1.0) * fm / sq);
      ans = xnew;
      stab_set(g, P, m, s, ans); stab_eval(g, m, s.uC); fnew = s.uC[idx];
      g.sync();
      if (fabs(fnew) < TOL && fnew >= 0.0) return xnew;
      if (stab_sign(fm, fnew) != fm) { xl = xm; fl = fm; xh = ans; fh = fnew; }
      else if (stab_sign(fl, fnew) != fl) { xh = ans; fh = fnew; }
      else if (stab_sign(fh, fnew) != fh) { xl = ans; fl = fnew; }
      else return 0.0;
    }
  } else {
    if (fl == 0.0) return x1;
    if (fh == 0.0) return x2;
  }
  return 0.0;
}

// update_q (:1055-1212): leaves the bodies at q + t dq and stores that in q; false when t fell below sqrt(eps)
template <class G>
B2M_DEV B2M_NOINL bool stab_update_q(const G& g, const SimParams& P, EnvMem& m, const StabMem& s) {
  const double MIN_T = B2M_NEAR_ZERO, BETA = 0.6;
  const int np = m.scal[S_NPAIRS];
  stab_eval(g, m, s.uC0);
  stab_set(g, P, m, s, 1.0);
  stab_eval(g, m, s.uC1);
  for (int i = g.tid; i < np; i += G::size) s.bracket[i] = ((s.uC0[i] < 0.0 && s.uC1[i] > 0.0) || (s.uC0[i] > 0.0 && s.uC1[i] < 0.0)) ? 1 : 0;
  g.sync();
  double t = 1.0;
  for (int i = 0; i < np; i++) {
    if (!s.bracket[i]) continue;
    const double root = stab_ridders(g, P, m, s, 0.0, t, s.uC0[i], s.uC1[i], i);
    if (root > 0.0 && root < 1.0) t = fmin(root, t);
  }
  stab_set(g, P, m, s, t);
  stab_eval(g, m, s.uC);
  for (;;) {
    bool worse = false;
    for (int i = g.tid; i < np; i += G::size) if (!s.bracket[i] && s.uC[i] < 0.0 && s.uC0[i] > s.uC[i]) worse = true;
    if (!g.any(worse)) break;
    t *= BETA;
    if (t < MIN_T) return false;
    stab_set(g, P, m, s, t);
    stab_eval(g, m, s.uC);
  }
  for (int k = g.tid; k < 7 * P.nb; k += G::size) s.q[k] = s.dq[k] * t + s.q[k];        // q = qstar: the stored vector is not renormalised (:1209)
  if (B2M_RC(P)) for (int k = g.tid; k < P.rc_links - 1; k += G::size) s.qj[k] = s.dqj[k] * t + s.qj[k];
  g.sync();
  return true;
}

#define B2M_STAB_CAP 100   /* rule H12: the reference's default loop is unbounded */

// The env is loaded (full working set with nmax = cmax); positions are stored by the caller afterwards.
template <class G>
B2M_DEV B2M_NOINL void env_stabilize(const G& g, const SimParams& P, int e, EnvMem& m, const StabMem& s, unsigned long long* lc) {
  if (P.stab_max_iterations == 0) return;
  const int nb = P.nb, ne = P.n_envs;
  double vio = stab_eval(g, m, s.uC);                                                   // :187
  if (!(vio < P.stab_eps)) return;
  for (int k = g.tid; k < 3 * nb; k += G::size) { s.vls[k] = m.bvl[k]; s.vas[k] = m.bva[k]; }   // save_velocities :66-75
  for (int b = g.tid; b < nb; b += G::size) {                                            // get_body_configurations :1215-1237
    for (int k = 0; k < 3; k++) s.q[7 * b + k] = m.bx[3 * b + k];
    for (int k = 0; k < 4; k++) s.q[7 * b + 3 + k] = m.bq[4 * b + k];
  }
  if (B2M_RC(P)) for (int k = g.tid; k < P.rc_links - 1; k += G::size) { s.jqds[k] = m.jqd[k]; s.qj[k] = m.jq[k]; }
  g.sync();
  const int cap = (P.stab_max_iterations < 0 || P.stab_max_iterations > B2M_STAB_CAP) ? B2M_STAB_CAP : P.stab_max_iterations;
  int iterations = 0;
  while (vio < P.stab_eps) {                                                             // :197
    if (iterations == cap) { if (cap == B2M_STAB_CAP && P.stab_max_iterations != B2M_STAB_CAP && g.tid == 0) lc[CNT_STAB_LSFAIL]++; break; }
    for (int k = g.tid; k < 3 * nb; k += G::size) { m.bvl[k] = 0.0; m.bva[k] = 0.0; }    // :211-217
    if (B2M_RC(P)) { for (int k = g.tid; k < P.rc_links - 1; k += G::size) m.jqd[k] = 0.0; g.sync(); rc_refresh(g, P, m); }
    g.sync();
    calc_pairwise_distances(g, m);
    if (g.tid == 0) {                                                                    // add_contact_constraints :304-345
      const int np = m.scal[S_NPAIRS];
      int nc = 0; bool overflow = false;
      ContactOut con[8];
      for (int p = 0; p < np && !overflow; p++) {
        const double dist = m.pd_dist[p];
        if (dist == B2M_INF) continue;
        int k;
        if (dist >= B2M_NEAR_ZERO) {
          con[0].p = ld3(m.pd_pa + 3 * p); con[0].n = normalize(ld3(m.pd_pb + 3 * p) - ld3(m.pd_pa + 3 * p));
          con[0].b1 = m.pair_a[p]; con[0].b2 = m.pair_b[p]; con[0].dist = dist;
          k = 1;
        } else k = pair_contacts(m, m.pair_a[p], m.pair_b[p], B2M_NEAR_ZERO, con, 8);
        for (int i = 0; i < k && i < 8; i++) {
          if (nc >= P.cmax) { overflow = true; break; }
          st3(m.cp + 3 * nc, con[i].p); st3(m.cnrm + 3 * nc, con[i].n);
          V3 t1, t2; orthonormal_basis(con[i].n, t1, t2);
          st3(m.ct1 + 3 * nc, t1); st3(m.ct2 + 3 * nc, t2);
          m.cb1[nc] = con[i].b1; m.cb2[nc] = con[i].b2; m.cdist[nc] = con[i].dist;
          nc++;
        }
      }
      m.scal[S_NCON] = nc;
      if (overflow) lc[CNT_OVERFLOW]++;
      build_islands(P, m, nc, false);
    }
    for (int k = g.tid; k < 7 * nb; k += G::size) s.dq[k] = 0.0;
    if (B2M_RC(P)) for (int k = g.tid; k < P.rc_links - 1; k += G::size) s.dqj[k] = 0.0;
    g.sync();
    const int nisl = m.scal[S_NISL];
    for (int isl = 0; isl < nisl; isl++) {                                               // determine_dq :932-970
      if (g.tid == 0) select_island(P, m, isl);
      g.sync();
      compute_problem_data(g, P, m);
      const int n = m.scal[S_NC];
      for (int t = g.tid; t < n * n; t += G::size) { const int j = t / n, i = t - j * n; m.MM[t] = Dn(m, n, 0, 0, i, j); }
      for (int i = g.tid; i < n; i += G::size) { m.qq[i] = m.cdist[m.icon[i]] - fabs(P.stab_eps) - B2M_NEAR_ZERO; m.z[i] = 0.0; }   // :432-433
      g.sync();
      int piv = 0, ex = 0;
      long long fast_calls = 1, lemke_calls = 0, pivots = 0;
      int st = lcp_fast_solve(g, n, m.MM, n, m.qq, 0.0, -1.0, false, m.z, m.work, m.iwork, &piv, nullptr, 0, nullptr, nullptr, &ex);   // :961, cold
      pivots = piv;
      bool solved = (st == LCP_OK || st == LCP_TRIVIAL);
      if (!solved) {
        g.sync();
        long long stats[3] = {0, 0, 0};
        st = lcp_lemke_regularized(g, n, m.MM, n, m.qq, -1.0, -1.0, -20, 1, 1, m.z, m.work, m.iwork, &piv, stats, nullptr);          // :962
        lemke_calls = stats[0]; pivots += stats[1];
        solved = (st != LCP_UNVERIFIED);                                                 // rule H12: z = 0 otherwise
      }
      g.sync();
      if (g.tid == 0) {
        lc[CNT_STAB_SOLVES]++; lc[CNT_FAST_CALLS] += fast_calls; lc[CNT_LEMKE_CALLS] += lemke_calls; lc[CNT_PIVOTS] += pivots;
        if (!solved) lc[CNT_LCP_FAIL]++;
      }
      for (int i = g.tid; i < n; i += G::size) { m.imp[i] = m.z[i]; m.imp[n + i] = 0.0; m.imp[2 * n + i] = 0.0; }
      g.sync();
      apply_to_bodies(g, P, m, m.imp);                                                   // update_from_stacked :965
      for (int b = g.tid; b < nb; b += G::size) {                                        // dq <- velocity in Euler coordinates :968-975
        if (m.bisl[b] != isl || !m.ben[b] || is_link(P, b)) continue;
        const double qx = m.bq[4 * b], qy = m.bq[4 * b + 1], qz = m.bq[4 * b + 2], qw = m.bq[4 * b + 3];
        const V3 w = ld3(m.bva + 3 * b);
        s.dq[7 * b + 0] = m.bvl[3 * b]; s.dq[7 * b + 1] = m.bvl[3 * b + 1]; s.dq[7 * b + 2] = m.bvl[3 * b + 2];
        s.dq[7 * b + 3] = 0.5 * (+qw * w.x + qz * w.y - qy * w.z);
        s.dq[7 * b + 4] = 0.5 * (-qz * w.x + qw * w.y + qx * w.z);
        s.dq[7 * b + 5] = 0.5 * (+qy * w.x - qx * w.y + qw * w.z);
        s.dq[7 * b + 6] = 0.5 * (-qx * w.x - qy * w.y - qz * w.z);
      }
      if (B2M_RC(P) && m.bisl[P.rc_first + 1] == isl) for (int k = g.tid; k < P.rc_links - 1; k += G::size) s.dqj[k] = m.jqd[k];
      g.sync();
    }
    if (!stab_update_q(g, P, m, s)) { if (g.tid == 0) lc[CNT_STAB_LSFAIL]++; break; }   // :231-235
    vio = stab_eval(g, m, s.uC);                                                         // :238
    iterations++;
  }
  if (g.tid == 0) lc[CNT_STAB_ITERS] += iterations;
  for (int k = g.tid; k < 3 * nb; k += G::size) { m.bvl[k] = s.vls[k]; m.bva[k] = s.vas[k]; }   // restore_velocities :78-85
  if (B2M_RC(P)) { for (int k = g.tid; k < P.rc_links - 1; k += G::size) m.jqd[k] = s.jqds[k]; g.sync(); rc_refresh(g, P, m); }
  g.sync();
  (void)ne; (void)e;
}
