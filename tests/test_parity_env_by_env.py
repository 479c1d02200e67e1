"""VERDICT r1 item 1a at CPU scale: the kernels' device code (host build) against the oracle, env by env -- solver-call
counts, failures and states, per step from identical inputs and over a horizon.  tests/test_gpu_parity_full.py repeats it
on the GPU at BASELINE's full size with the configuration bench.py times."""
import numpy as np
import pytest

import parity_util as PU
from moby_b200 import scenes


class _HostStepper:
    """TimeSteppingSimulator-shaped adapter over tests/hostsim (phased schedule, as the GPU launches it)."""

    def __init__(self, hostsim, scene):
        self.hs = hostsim.HostSim(scene)

    def step(self, dt, n):
        self.hs.step_phased(dt, n)

    def get_state(self):
        return self.hs.q.copy(), self.hs.v.copy()

    def env_stats(self):
        out = {k: a.copy() for k, a in self.hs.env_stats().items()}
        self.hs.stat[:] = 0
        return out


@pytest.fixture(scope="module")
def hostsim():
    import hostsim_api
    hostsim_api.build()
    return hostsim_api


def test_per_step_parity_from_identical_inputs(hostsim, oracle):
    ne, pre = 2048, 200
    sc = scenes.small_lcp_batch(ne, seed=0xB200)          # bench.py's configuration: per-scene min-step-size
    sim = _HostStepper(hostsim, sc)
    sim.step(1e-3, pre)
    q, v = sim.get_state()
    sc2 = PU.scene_at_state(sc, q, v)
    sim2 = _HostStepper(hostsim, sc2)
    sim2.step(1e-3, 1)
    st = sim2.env_stats()
    q2, v2 = sim2.get_state()
    idx = np.arange(ne)
    ost, qo, vo = PU.oracle_run(oracle, sc2, idx, 1e-3, 1, threads=4)
    rep = PU.compare(st, q2, v2, ost, qo, vo, idx)
    assert rep["sum_lcp_solves"][0] > 200 and rep["sum_lemke_calls"][0] > 20, rep     # the step exercises both solvers
    assert rep["mismatch_lcp_failures"] == 0 and rep["mismatch_outside_ladder"] == 0, rep
    assert rep["above_tol_same_path"] == 0, rep                                       # 1e-9 wherever the ladder agrees
    assert rep["ladder_mismatch"] <= max(2, 3 * ne // 1000) and rep["err_max_ladder"] < 0.1, rep   # <= 0.3 % of envs (see tests/test_gpu_parity_full.py)


def test_horizon_statistics_and_failures(hostsim, oracle):
    """300 steps from the scene's initial state: every env that reports an unsolved LCP does so in the oracle too, the
    totals agree, and the share of envs above 1e-9 is small (the drift bench.py reports)."""
    ne, steps = 1024, 300
    sc = scenes.small_lcp_batch(ne, seed=0xB200 + 5)
    sim = _HostStepper(hostsim, sc)
    sim.step(1e-3, steps)
    st = sim.env_stats()
    q, v = sim.get_state()
    idx = np.arange(ne)
    ost, qo, vo = PU.oracle_run(oracle, sc, idx, 1e-3, steps, threads=4)
    rep = PU.compare(st, q, v, ost, qo, vo, idx)
    assert np.array_equal(st["lcp_failures"] > 0, ost["lcp_failures"] > 0), rep
    assert rep["mismatch_lcp_solves"] <= ne // 100, rep
    for k in ("lemke_calls", "lcp_fast_calls"):
        a, b = rep["sum_" + k]
        assert abs(a - b) <= 0.05 * max(a, b), rep
    assert rep["above_tol"] <= ne // 20, rep
