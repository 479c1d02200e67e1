"""CPU checks of the kernels' device code (compiled for the host with a single-thread group, tests/hostsim)
against the oracle: this is how kernel logic is validated where no GPU exists.  The GPU tests repeat the same
comparisons on the device."""
import numpy as np
import pytest

from lcp_problems import random_batch
from moby_b200 import scenes


@pytest.fixture(scope="module")
def hostsim():
    import hostsim_api
    hostsim_api.build()
    return hostsim_api


@pytest.mark.parametrize("n", [3, 8, 24, 40])
def test_lcp_device_code_matches_oracle(hostsim, oracle, n):
    M, q = random_batch(10, n, seed=31 + n)
    for b in range(10):
        ok, zo, info = oracle.lcp_lemke(M[b], q[b])
        st, z, piv, log = hostsim.lcp(0, M[b], q[b])
        assert ok and st in (0, 1) and piv == info["pivots"] and list(log) == list(info["log"])
        assert np.allclose(z, zo, rtol=0, atol=1e-10 * max(1, np.abs(zo).max()))
        ok, zo, info = oracle.lcp_fast(M[b], q[b])            # may cycle to the 2n cap on some problems: then both must
        st, z, piv, log = hostsim.lcp(1, M[b], q[b])
        assert ok == (st in (0, 1)) and st == info["status"] and piv == info["pivots"] and list(log) == list(info["log"])
        assert np.array_equal(z, zo)                      # same arithmetic order => same bits


@pytest.mark.parametrize("name,dt,steps,tol", [("box", 1e-3, 300, 0.0), ("boxmu", 1e-3, 200, 0.0), ("stack", 1e-3, 200, 0.0),
                                              ("ball", 0.025, 80, 1e-14), ("box_ap", 1e-3, 100, 1e-13)])
def test_scene_trajectories(hostsim, oracle, name, dt, steps, tol):
    sc = {"box": lambda: scenes.sitting_box(1, NK=8, y0=0.50001), "boxmu": lambda: scenes.sitting_box(1, NK=4, mu=0.5),
          "stack": lambda: scenes.sphere_stack(1), "ball": lambda: scenes.bouncing_ball(1),
          "box_ap": lambda: scenes.sitting_box(1, NK=8, mu=0.3)}[name]()
    if name == "box_ap":
        sc.impact_model = scenes.MODEL_AP
    hs, osim = hostsim.HostSim(sc), oracle.OracleSim(sc)
    for _ in range(steps):
        hs.step(dt)
        osim.step(dt)
        qo, vo = osim.get_state()
        assert np.abs(hs.q[:, :, 0] - qo).max() <= tol and np.abs(hs.v[:, :, 0] - vo).max() <= tol
    co, ch = osim.counters(), hs.counters_dict()
    for k in ("env_steps", "mini_steps", "lcp_solves", "contacts", "max_lcp_n", "lcp_failures"):
        assert co[k] == ch[k], (k, co[k], ch[k])


def test_random_batch_short_horizon(hostsim, oracle):
    """SURVEY 8(d) case 2 at 48 envs: 60 steps; 1e-9 relative per the north star (the tumbling boxes hit the Lemke
    fallback, where tableau and LU-per-pivot round differently, so this is a tolerance, not bit, comparison)."""
    sc = scenes.small_lcp_batch(48, seed=5)
    hs = hostsim.HostSim(sc)
    hs.step(1e-3, 60)
    worst = 0.0
    for e in range(48):
        osim = oracle.OracleSim(sc, env=e)
        osim.step(1e-3, 60)
        qo, vo = osim.get_state()
        scale = max(1.0, np.abs(qo).max(), np.abs(vo).max())
        worst = max(worst, np.abs(hs.q[:, :, e] - qo).max() / scale, np.abs(hs.v[:, :, e] - vo).max() / scale)
    assert worst < 1e-9, worst
    assert hs.counters_dict()["overflow"] == 0


@pytest.mark.parametrize("rounds,budget,min_step", [(1, 0, "scene"), (2, 0, "scene"), (2, 8, "scene"), (3, 0, "default"), (2, 5, "default")])
def test_phased_schedule_is_bit_identical_to_fused(hostsim, rounds, budget, min_step):
    """advance / impact-class / straggler / finish phases (the GPU launch schedule) against the fused per-env loop:
    same bits in q, v, time, warm start and the same counters, whatever the number of rounds or the pivot budget."""
    sc = scenes.small_lcp_batch(64, seed=11)
    if min_step == "default":
        sc.min_step_size_env = None          # sqrt(eps) everywhere: many conservative-advancement mini-steps per step
    a, b = hostsim.HostSim(sc), hostsim.HostSim(sc)
    a.step(1e-3, 80)
    for _ in range(8):
        b.step_phased(1e-3, 10, rounds=rounds, pivot_budget=budget)
    assert np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v) and np.array_equal(a.time, b.time)
    assert np.array_equal(a.zlast_n, b.zlast_n)
    for e in range(64):
        assert np.array_equal(a.zlast[:a.zlast_n[e], e], b.zlast[:b.zlast_n[e], e])
    ca, cb = a.counters_dict(), b.counters_dict()
    assert ca["lcp_solves"] > 100 and ca["mini_steps"] >= ca["env_steps"] == 64 * 80
    assert ca == cb, (ca, cb)


def test_phased_sphere_stack_multibody(hostsim):
    sc = scenes.sphere_stack(3)
    a, b = hostsim.HostSim(sc), hostsim.HostSim(sc)
    a.step(1e-3, 120)
    b.step_phased(1e-3, 120, rounds=2, pivot_budget=0)
    assert np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v)
    assert a.counters_dict() == b.counters_dict()
