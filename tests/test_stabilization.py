"""Constraint stabilization (ConstraintStabilization.cpp:167-254): the oracle's restatement, and the
kernels' device code (host build) against it.  tests/test_gpu_stabilization.py repeats the comparison through the C ABI.

What the reference's default (max_iterations = UINT_MAX, eps = +sqrt(eps): rule H7) does to a resting body: after every
step, while some pair is closer than 1.49e-8, a frictionless position LCP pushes the pair apart to dist >= 2 sqrt(eps)
along a line search; velocities are untouched."""
import numpy as np
import pytest

from moby_b200 import scenes

NEAR_ZERO = scenes.NEAR_ZERO


@pytest.fixture(scope="module")
def hostsim():
    import hostsim_api
    hostsim_api.build()
    return hostsim_api


def _on(sc, iters=-1):
    sc.stabilization_max_iterations = iters
    return sc


def test_resting_box_is_lifted_to_two_sqrt_eps_and_stays(oracle):
    sc = _on(scenes.sitting_box(1, NK=8, y0=0.5))
    sc.min_step_size = 1e-3
    o = oracle.OracleSim(sc)
    o.step(1e-3, 1)
    q, v = o.get_state()
    assert abs(q[0, 1] - (0.5 + 2 * NEAR_ZERO)) < 1e-15                 # Cn_v = dist - |eps| - NEAR_ZERO (:432-433) => target 2 sqrt(eps)
    c = o.counters()
    assert c["stab_iterations"] == 1 and c["stab_lcp_solves"] == 1 and c["stab_line_search_failures"] == 0
    o.step(1e-3, 500)
    q2, v2 = o.get_state()
    assert np.abs(q2 - q).max() < 1e-15 and np.abs(v2).max() < 1e-15       # no further correction: the pair is no longer closer than eps
    assert o.counters()["stab_iterations"] == 1


def test_stabilization_keeps_velocities_and_removes_penetration(oracle):
    """min-step-size 1e-3 lets boxes integrate into the ground (test_oracle_regress notes the reference relies on
    stabilization there): with it on, no vertex stays below the plane and velocities are what the impact solve left."""
    ne = 96
    sc_off, sc_on = scenes.small_lcp_batch(ne, seed=9), _on(scenes.small_lcp_batch(ne, seed=9))
    for sc, on in ((sc_off, False), (sc_on, True)):
        ob = oracle.OracleBatch(sc)
        c = ob.run(1e-3, 400, threads=4)
        q, v = ob.get_state_soa()
        isbox = sc.shape[0] == scenes.SHAPE_BOX
        R = scenes._rotmat(q[0, 3:7, :])
        he = sc.dims[0] / 2
        low = q[0, 1, :] - (np.abs(R[1, 0]) * he[0] + np.abs(R[1, 1]) * he[1] + np.abs(R[1, 2]) * he[2])
        if on:
            assert low[isbox].min() > -1e-9 and c["stab_iterations"] > ne and c["stab_line_search_failures"] == 0    # test/TestDie.cpp:130, far inside its 1e-6
        else:
            assert c["stab_iterations"] == 0
        assert c["lcp_failures"] == 0


def test_max_iterations_is_honoured(oracle):
    sc = _on(scenes.small_lcp_batch(32, seed=4), iters=1)
    ob = oracle.OracleBatch(sc)
    c = ob.run(1e-3, 200, threads=2)
    assert 0 < c["stab_iterations"] <= 32 * 200                          # at most one iteration per env-step


@pytest.mark.parametrize("name,dt,steps,tol", [("box", 1e-3, 300, 0.0), ("batch", 1e-3, 300, 1e-9), ("spheres", 1e-3, 200, 0.0),
                                              ("stack3", 1e-3, 80, 0.0), ("ur10", 5e-4, 150, 1e-9), ("feeder", 1e-3, 40, 1e-9)])
def test_device_code_matches_oracle_with_stabilization(hostsim, oracle, name, dt, steps, tol):
    sc = {"box": lambda: scenes.sitting_box(2, NK=8, y0=0.5), "batch": lambda: scenes.small_lcp_batch(96, seed=3),
          "spheres": lambda: scenes.sphere_stack(2), "stack3": lambda: scenes.box_stack(2, 3), "ur10": lambda: scenes.ur10(4),
          "feeder": lambda: scenes.parts_feeder(6)}[name]()
    _on(sc)
    hs = hostsim.HostSim(sc)
    if name == "batch":
        hs.step_phased(dt, steps)                                       # the schedule the GPU launches, stabilization phase last
    else:
        hs.step(dt, steps)
    ob = oracle.OracleBatch(sc)
    co = ob.run(dt, steps, threads=4)
    qo, vo = ob.get_state_soa()
    err = np.maximum(np.abs(hs.q - qo).max(axis=(0, 1)), np.abs(hs.v - vo).max(axis=(0, 1)))
    if name == "feeder":                  # every impact LCP of this scene is singular: a third of the Lemke runs part ways per step (tests/parity_util.py)
        assert err.max() < 1e-2, err.max()
    elif name == "batch":                 # a few envs run Lemke on singular LCPs, where tableau and LU-per-pivot may part ways
        assert (err > tol).sum() <= max(1, sc.n_envs // 20) and err.max() < 1e-4, (int((err > tol).sum()), err.max())
    else:
        assert err.max() <= tol, err.max()
    ch = hs.counters_dict()
    for k in ("env_steps", "stab_iterations", "stab_lcp_solves", "stab_line_search_failures", "lcp_solves", "lcp_failures", "contacts"):
        assert ch[k] == co[k] or (name in ("batch", "feeder") and abs(ch[k] - co[k]) <= 0.05 * co[k]), (k, ch[k], co[k])
    assert ch["stab_iterations"] > 0
