"""Env-by-env comparison of a stepped batch with the oracle (test infrastructure, shared by the CPU tests that drive the
kernels' device code on the host and the GPU tests that go through the C ABI).

Two statements are checked, the ones BASELINE.json's north star makes:

* per step, from identical inputs (state in, empty warm starts on both sides): the same number of impact problems, of
  LCP::lcp_fast calls, of LCP::lcp_lemke calls and of LCPSolverException equivalents for every env, and post-step states
  within 1e-9 relative.  The one documented exception: on a degenerate LCP the tableau form of Lemke used by the kernels
  and the LU-per-pivot form of the reference (LCP.cpp:834-838) round differently, may take different pivot paths and may
  then be accepted at different rungs of the regularisation ladder (LCP.cpp:419-477).  Both results pass the wrapper's own
  acceptance test, whose tolerance is n * |M|_inf * sqrt(eps) ~ 1e-6 (LCP.cpp:369,381-390) -- so such envs agree only to
  that tolerance.  Their share is bounded and reported.
* over many steps: drift, reported as the share of envs above 1e-9.
"""
import copy

import numpy as np

STAT_KEYS = ("lcp_failures", "lemke_calls", "lcp_fast_calls", "lcp_solves", "pivots")


def rel_state_error(q, v, qo, vo):
    """Per-env max |difference| over all state entries, relative to max(1, largest |entry| of the env)."""
    scale = np.maximum(1.0, np.maximum(np.abs(qo).max(axis=(0, 1)), np.abs(vo).max(axis=(0, 1))))
    return np.maximum(np.abs(q - qo).max(axis=(0, 1)), np.abs(v - vo).max(axis=(0, 1))) / scale


def scene_at_state(scene, q, v, joints=None):
    """A copy of `scene` whose initial state is (q, v) [, (jq, jqd)]."""
    s = copy.copy(scene)
    s.q, s.v = np.ascontiguousarray(q), np.ascontiguousarray(v)
    if joints is not None and getattr(scene, "rc", None) is not None:
        rc = copy.copy(scene.rc)
        rc.scene, rc.jq, rc.jqd = s, np.ascontiguousarray(joints[0]), np.ascontiguousarray(joints[1])
        s.rc = rc
    return s


def oracle_run(oracle, scene, idx, dt, steps, threads):
    """Oracle on envs `idx` of `scene`: returns (stats dict of [len(idx)] arrays, q, v SoA over idx[, jq, jqd])."""
    from moby_b200 import sharding
    sub = sharding.select_envs(scene, idx)
    ob = oracle.OracleBatch(sub)
    ob.run(dt, steps, threads=threads)
    q, v = ob.get_state_soa()
    return ob.env_stats(), q, v


def compare(stat, q, v, ostat, qo, vo, idx, tol=1e-9, ladder_tol=1e-4):
    """stat/q/v: the batch under test (all envs); ostat/qo/vo: the oracle on envs `idx`.  Returns a report dict."""
    idx = np.asarray(idx)
    err = rel_state_error(q[..., idx], v[..., idx], qo, vo)
    same = {k: stat[k][idx] == ostat[k] for k in STAT_KEYS}
    # envs whose Lemke runs took a different pivot path (another number of pivots) or stopped at a different rung of the ladder
    ladder = ~same["lemke_calls"] | ~same["pivots"]
    rep = dict(n=int(idx.size), err_max=float(err.max()) if idx.size else 0.0,
               above_tol=int((err > tol).sum()), above_tol_same_path=int(((err > tol) & ~ladder).sum()),
               ladder_mismatch=int(ladder.sum()), err_max_ladder=float(err[ladder].max()) if ladder.any() else 0.0,
               err_max_same_path=float(err[~ladder].max()) if (~ladder).any() else 0.0)
    for k in STAT_KEYS:
        rep["mismatch_" + k] = int((~same[k]).sum())
        rep["sum_" + k] = (int(stat[k][idx].sum()), int(ostat[k].sum()))
    rep["mismatch_outside_ladder"] = int(((~same["lcp_fast_calls"] | ~same["lcp_solves"] | ~same["lcp_failures"]) & ~ladder).sum())
    rep["ok_ladder_tol"] = bool((err[ladder] <= ladder_tol).all()) if ladder.any() else True
    return rep
