"""Rimless wheel on the GPU (SHAPE_WHEEL through the C ABI) against the oracle and against the reference's own golden
trajectory regress/rimless-wheel.dat (see tests/test_rimless_wheel.py for how its configuration was reconstructed)."""
import os

import numpy as np
import pytest

from moby_b200 import scenes
from test_rimless_wheel import GOLDEN, golden_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def test_gpu_matches_regress_rimless_wheel(torch_cuda):
    from moby_b200 import TimeSteppingSimulator
    gold = np.loadtxt(os.path.join(GOLDEN, "regress_rimless_wheel.txt"))
    sim = TimeSteppingSimulator(golden_scene(2))
    step, worst = 0, 0.0
    for row in gold:
        target = int(round(row[0] / 1e-3))
        if target > step:
            sim.step(1e-3, target - step)
        step = target
        q, _ = sim.get_state()
        worst = max(worst, np.abs(q[1, :, 0] - row[1:8]).max(), np.abs(q[1, :, 1] - row[1:8]).max())
    assert worst < 2e-5, worst
    c = sim.counters()
    assert c["lcp_failures"] == 0 and c["lcp_solves"] > 10000 and c["stab_line_search_failures"] == 0


@pytest.mark.parametrize("stab", [0, -1])
def test_gpu_wheel_batch_matches_oracle(torch_cuda, oracle, stab):
    """4,096 wheels with random spins, 600 steps (every wheel passes at least one spoke-to-spoke impact): states within
    1e-9 of the oracle, counters equal."""
    from moby_b200 import TimeSteppingSimulator
    ne = 4096
    s = scenes.rimless_wheel(ne, theta_dot=3.0, seed=5, stabilization=stab)
    sim = TimeSteppingSimulator(s)
    sim.step(1e-3, 600)
    q, v = sim.get_state()
    ob = oracle.OracleBatch(s)
    co = ob.run(1e-3, 600, threads=min(16, os.cpu_count() or 1))
    qo, vo = ob.get_state_soa()
    err = max(np.abs(q - qo).max(), np.abs(v - vo).max())
    assert err < 1e-9, err
    cg = sim.counters()
    for k in ("env_steps", "mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "lcp_failures", "contacts", "stab_iterations", "stab_lcp_solves"):
        assert cg[k] == co[k], (k, cg[k], co[k])
    assert cg["lcp_failures"] == 0 and (q[1, 0, :] > 0.5).all()
