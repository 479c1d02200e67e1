"""Constraint stabilization on the GPU (k_stabilize.cu through the C ABI) against the oracle: the reference's default
configuration (ConstraintStabilization.cpp:53-59) instead of constraint-stabilization-max-iterations=0."""
import os

import numpy as np
import pytest

from moby_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _make(name):
    sc = {"box": lambda: scenes.sitting_box(3, NK=8, y0=0.5), "batch": lambda: scenes.small_lcp_batch(4096, seed=3),
          "spheres": lambda: scenes.sphere_stack(3), "stack3": lambda: scenes.box_stack(3, 3), "ur10": lambda: scenes.ur10(8)}[name]()
    sc.stabilization_max_iterations = -1
    return sc


@pytest.mark.parametrize("name,dt,steps,tol", [("box", 1e-3, 300, 0.0), ("spheres", 1e-3, 200, 0.0), ("stack3", 1e-3, 80, 0.0),
                                              ("ur10", 5e-4, 150, 1e-9), ("batch", 1e-3, 300, 1e-9)])
@pytest.mark.parametrize("thread", [1, 0])
def test_stabilized_stepping_matches_oracle(torch_cuda, oracle, name, dt, steps, tol, thread):
    """thread = 1: thread-per-env stabilization kernel; 0: the warp-per-env one (B200MOBY_STAB_THREAD=0)."""
    from moby_b200 import TimeSteppingSimulator
    if thread == 0 and name in ("spheres", "ur10"):
        pytest.skip("covered by the thread variant; the warp variant is exercised on box / stack3 / batch")
    sc = _make(name)
    os.environ["B200MOBY_STAB_THREAD"] = str(thread)
    try:
        sim = TimeSteppingSimulator(sc)
    finally:
        del os.environ["B200MOBY_STAB_THREAD"]
    sim.step(dt, steps)
    q, v = sim.get_state()
    ob = oracle.OracleBatch(sc)
    co = ob.run(dt, steps, threads=min(16, os.cpu_count() or 1))
    qo, vo = ob.get_state_soa()
    err = np.maximum(np.abs(q - qo).max(axis=(0, 1)), np.abs(v - vo).max(axis=(0, 1)))
    cg = sim.counters()
    if name == "batch":       # a few envs run Lemke on singular LCPs (tests/parity_util.py): bounded share, everything else at 1e-9
        assert (err > tol).sum() <= sc.n_envs // 50 and err.max() < 1e-3, (int((err > tol).sum()), err.max())
        for k in ("stab_iterations", "stab_lcp_solves", "lcp_solves"):
            assert abs(cg[k] - co[k]) <= 0.01 * co[k], (k, cg[k], co[k])
        isbox = sc.shape[0] == scenes.SHAPE_BOX
        R = scenes._rotmat(q[0, 3:7, :])
        he = sc.dims[0] / 2
        low = q[0, 1, :] - (np.abs(R[1, 0]) * he[0] + np.abs(R[1, 1]) * he[1] + np.abs(R[1, 2]) * he[2])
        assert low[isbox].min() > -1e-9                                  # test/TestDie.cpp:130 with a wide margin
    else:
        assert err.max() <= tol, err.max()
        for k in ("env_steps", "stab_iterations", "stab_lcp_solves", "stab_line_search_failures", "lcp_solves", "lcp_failures", "contacts"):
            assert cg[k] == co[k], (k, cg[k], co[k])
    assert cg["stab_iterations"] > 0 and cg["stab_line_search_failures"] == 0 and cg["lcp_failures"] == 0
