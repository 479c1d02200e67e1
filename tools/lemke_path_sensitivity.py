#!/usr/bin/env python
"""How sensitive is LCP::lcp_lemke's pivot path to the LAST BITS of the basis solve?  (CPU only; writes
profiles/r02_lemke_path_sensitivity.json.)

LCP.cpp:834-838 solves B d = Be with a fresh dense LU (Ravelin solve_fast = LAPACK dgesv) every pivot and compares the
entries of d with a tolerance at rounding level (PIV_TOL = eps n max(1,|M|), :761).  On the singular LCPs that resting
face contacts produce, which entries of d pass that test is decided by rounding noise.  This script takes the impact
LCPs of a scene (dumped from the host build of the kernels' device code) and runs Lemke's rules with three basis solves:

  oracle : oracle/ (hand-written LU with partial pivoting, LU per pivot -- the restatement of the reference)
  lapack : the same rules with numpy.linalg.solve (LAPACK dgesv: what the reference really links) -- fresh per pivot
  tableau: the same rules on an incrementally updated inverse (the arithmetic of the kernels' tableau)

and reports on how many problems each pair takes the identical leaving-variable sequence.  `lapack` vs `oracle` is the
floor: two faithful implementations of the reference that differ only in the rounding of the LU."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
EPS = np.finfo(float).eps


def lemke(M, q, mode):
    n = len(q)
    nrm = np.abs(M).max()
    zero_tol, PIV = EPS * nrm * n, EPS * n * max(1.0, nrm)
    if q.min() > -zero_tol:
        return True, [], 0
    B, x, bas, t = -np.eye(n), q.copy(), list(range(n, 2 * n)), 2 * n
    lv = int(np.argmin(x)); tval = -x[lv]; leaving = bas[lv]; bas[lv] = t
    u = (x < 0).astype(float); Be = -(B @ u); x = x + tval * u; x[lv] = tval; B[:, lv] = Be
    log = [leaving]
    Binv = np.linalg.inv(B) if mode == "tableau" else None
    for it in range(min(1000, 50 * n)):
        if leaving == t:
            return True, log, it
        if leaving < n:
            entering = n + leaving; Be = np.zeros(n); Be[leaving] = -1.0
        else:
            entering = leaving - n; Be = M[:, entering].copy()
        if mode == "lapack":
            try:
                d = np.linalg.solve(B, Be)
            except np.linalg.LinAlgError:
                return False, log, it
        else:
            d = Binv @ Be
        J = np.where(d > PIV)[0]
        if len(J) == 0:
            return False, log, it
        theta = ((x[J] + zero_tol) / d[J]).min()
        J = [j for j in J if x[j] / d[j] <= theta]
        if not J:
            return False, log, it
        tp = bas.index(t)
        lv = tp if tp in J else J[0]
        leaving = bas[lv]; ratio = x[lv] / d[lv]; x = x - ratio * d; x[lv] = ratio; B[:, lv] = Be; bas[lv] = entering
        log.append(leaving)
        if Binv is not None:
            row = Binv[lv, :] / d[lv]
            Binv = Binv - np.outer(d, row); Binv[lv, :] = row
    return False, log, min(1000, 50 * n)


def main():
    import hostsim_api as H
    import oracle_api as O
    from moby_b200 import scenes
    O.build(); H.build()
    out = {}
    for name, make, dt, steps in (("parts-feeder (mu = 0.01, box-box face contact)", lambda: scenes.parts_feeder(256), 1e-3, 100),
                                  ("configs[1] sitting boxes", lambda: scenes.small_lcp_batch(1024, seed=0xB200), 1e-3, 300)):
        sc = make()
        hs = H.HostSim(sc, taps=True)
        hs.step(dt, steps)
        r = dict(problems=0, n_max=0, cond_median=None, oracle_solved=0, lapack_solved=0, tableau_solved=0, oracle_pivots=0, lapack_pivots=0,
                 tableau_pivots=0, same_path_lapack_vs_oracle=0, same_path_tableau_vs_oracle=0, same_path_tableau_vs_lapack=0)
        conds = []
        for e in range(sc.n_envs):
            n, MM, qq, _ = hs.last_lcp(e)
            if n < 16:
                continue
            ok, _, info = O.lcp_lemke(MM, qq)
            lo = list(info["log"])
            okl, ll, pl = lemke(MM, qq, "lapack")
            okt, lt, pt = lemke(MM, qq, "tableau")
            r["problems"] += 1; r["n_max"] = max(r["n_max"], n); conds.append(np.linalg.cond(MM))
            r["oracle_solved"] += bool(ok); r["lapack_solved"] += bool(okl); r["tableau_solved"] += bool(okt)
            r["oracle_pivots"] += int(info["pivots"]); r["lapack_pivots"] += pl; r["tableau_pivots"] += pt
            r["same_path_lapack_vs_oracle"] += ll == lo; r["same_path_tableau_vs_oracle"] += lt == lo; r["same_path_tableau_vs_lapack"] += lt == ll
        r["cond_median"] = float(np.median(conds))
        out[name] = r
        print(name, r)
    with open(os.path.join(ROOT, "profiles", "r02_lemke_path_sensitivity.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
