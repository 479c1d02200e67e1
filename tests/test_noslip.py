"""No-slip impact model (all contacts of an island with mu >= 100; ImpactConstraintHandler.cpp:122-135, 236-293,
1009-1417): the kernels' device code on the host against the oracle, plus the model's defining properties."""
import numpy as np
import pytest

from moby_b200 import scenes


@pytest.fixture(scope="module")
def hostsim():
    import hostsim_api
    hostsim_api.build()
    return hostsim_api


def _sliding_box(eps=0.0):
    sc = scenes.sitting_box(1, NK=4, mu=100.0, eps=eps)
    sc.v[0, 0, :] = 0.7          # sliding along x and z, spinning about y
    sc.v[0, 2, :] = -0.3
    sc.v[0, 4, :] = 0.5
    return sc


def _noslip_batch(n, seed):
    sc = scenes.small_lcp_batch(n, seed=seed)
    sc.mu_coulomb[:] = 100.0
    return sc


def test_sliding_box_sticks(oracle):
    """The first impact removes the tangential velocity of every contact point (the sliding cube trips over its leading
    edge instead of sliding on), through an nc x nc LCP rather than the 32-variable QP one."""
    osim = oracle.OracleSim(_sliding_box())
    osim.step(1e-3, 1)
    q, v = osim.get_state()
    x, vl, om = q[0, :3], v[0, :3], v[0, 3:]
    assert np.abs(vl).max() > 0.1                                     # still moving (tipping), not simply stopped
    for sx in (-0.5, 0.5):
        for sz in (-0.5, 0.5):
            r = np.array([sx, -0.5, sz])                              # bottom vertices; rotation after 1 ms is negligible
            pv = vl + np.cross(om, r)
            assert abs(pv[0]) < 1e-3 and abs(pv[2]) < 1e-3 and pv[1] > -1e-2, pv
    c = osim.counters()
    assert c["max_lcp_n"] == 4 and c["lcp_failures"] == 0


@pytest.mark.parametrize("name,dt,steps", [("slide", 1e-3, 200), ("slide_eps", 1e-3, 200), ("ball", 0.01, 200)])
def test_scene_trajectories(hostsim, oracle, name, dt, steps):
    if name == "ball":
        sc = scenes.bouncing_ball(1, eps=0.8)
        sc.mu_coulomb[:] = 100.0
        sc.v[0, 0, :] = 1.0
    else:
        sc = _sliding_box(eps=0.5 if name == "slide_eps" else 0.0)
    hs, osim = hostsim.HostSim(sc), oracle.OracleSim(sc)
    for _ in range(steps):
        hs.step(dt)
        osim.step(dt)
        qo, vo = osim.get_state()
        assert np.array_equal(hs.q[:, :, 0], qo) and np.array_equal(hs.v[:, :, 0], vo)      # lcp_fast only: same bits
    co, ch = osim.counters(), hs.counters_dict()
    assert co["lcp_solves"] > 0
    for k in ("env_steps", "mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "pivots", "contacts", "max_lcp_n", "lcp_failures"):
        assert co[k] == ch[k], (k, co[k], ch[k])


def test_random_batch(hostsim, oracle):
    sc = _noslip_batch(32, seed=17)
    hs = hostsim.HostSim(sc)
    hs.step(1e-3, 80)
    worst = 0.0
    for e in range(32):
        osim = oracle.OracleSim(sc, env=e)
        osim.step(1e-3, 80)
        qo, vo = osim.get_state()
        scale = max(1.0, np.abs(qo).max(), np.abs(vo).max())
        worst = max(worst, np.abs(hs.q[:, :, e] - qo).max() / scale, np.abs(hs.v[:, :, e] - vo).max() / scale)
    assert worst < 1e-9, worst
    c = hs.counters_dict()
    assert c["lcp_solves"] > 50 and c["max_lcp_n"] <= 8 and c["overflow"] == 0


def test_phased_schedule_matches_fused(hostsim):
    sc = _noslip_batch(32, seed=3)
    a, b = hostsim.HostSim(sc), hostsim.HostSim(sc)
    a.step(1e-3, 60)
    for _ in range(6):
        b.step_phased(1e-3, 10, rounds=2, pivot_budget=8)
    assert np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v)
    assert np.array_equal(a.zlast_n, b.zlast_n) and np.array_equal(a.zlast, b.zlast)
    assert a.counters_dict() == b.counters_dict()


@pytest.mark.gpu
def test_gpu_noslip_scenes_match_oracle(oracle):
    """The CUDA path through the C ABI, single scenes step by step: same bits as the oracle (only lcp_fast runs)."""
    import torch
    assert torch.cuda.is_available()
    from moby_b200 import TimeSteppingSimulator
    for eps in (0.0, 0.5):
        sc = _sliding_box(eps=eps)
        sim, osim = TimeSteppingSimulator(sc), oracle.OracleSim(sc)
        for _ in range(150):
            sim.step(1e-3)
            osim.step(1e-3)
            q, v = sim.get_state()
            qo, vo = osim.get_state()
            assert np.array_equal(q[:, :, 0], qo) and np.array_equal(v[:, :, 0], vo)
        cg, co = sim.counters(), osim.counters()
        for k in ("env_steps", "mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "contacts", "max_lcp_n", "lcp_failures"):
            assert cg[k] == co[k], (k, cg[k], co[k])
        assert cg["max_lcp_n"] == 4


@pytest.mark.gpu
def test_gpu_noslip_batch_matches_oracle_and_host_build(oracle, hostsim):
    import torch
    assert torch.cuda.is_available()
    from moby_b200 import TimeSteppingSimulator
    ne = 203
    sc = _noslip_batch(ne, seed=23)
    sim = TimeSteppingSimulator(sc)
    sim.step(1e-3, 40)
    sim.step(1e-3, 40)
    q, v = sim.get_state()
    qo, vo = sc.q.copy(), sc.v.copy()
    co = oracle.batch_step(sc, qo, vo, 1e-3, 80, threads=8)
    scale = np.maximum(1.0, np.maximum(np.abs(qo).max(axis=(0, 1)), np.abs(vo).max(axis=(0, 1))))
    err = np.maximum(np.abs(q - qo).max(axis=(0, 1)), np.abs(v - vo).max(axis=(0, 1))) / scale
    assert err.max() < 1e-9, (err.max(), int(err.argmax()))
    cg = sim.counters()
    for k in ("env_steps", "mini_steps", "lcp_solves", "contacts", "max_lcp_n", "lcp_failures"):
        assert cg[k] == co[k], (k, cg[k], co[k])
    hs = hostsim.HostSim(sc)
    hs.step(1e-3, 80)
    assert np.array_equal(q, hs.q) and np.array_equal(v, hs.v)
