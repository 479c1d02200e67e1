"""The contact-constrained pendulum on the GPU (SHAPE_PIN / SHAPE_PINWORLD through the C ABI) against the reference's golden
trajectory and against the oracle (see tests/test_contact_constrained_pendulum.py)."""
import os

import numpy as np
import pytest

from moby_b200 import scenes
from test_contact_constrained_pendulum import _compare_with_golden

pytestmark = pytest.mark.gpu


def test_gpu_matches_regress_contact_constrained_pendulum():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from moby_b200 import TimeSteppingSimulator
    sim = TimeSteppingSimulator(scenes.contact_constrained_pendulum(2))
    first, worst = _compare_with_golden(lambda n: sim.step(1e-3, n), lambda: sim.get_state()[0][0, :, 1])
    assert first < 2e-3 and worst < 3e-2, (first, worst)
    c = sim.counters()
    assert c["lcp_failures"] == 0 and c["max_lcp_n"] == 48


def test_gpu_pendulum_batch_matches_oracle(oracle):
    """256 pendulums released from different heights: 1,000 steps, states within 1e-8 of the oracle (singular LCPs: the pivot
    paths differ, the net impulses do not)."""
    import torch
    assert torch.cuda.is_available()
    from moby_b200 import TimeSteppingSimulator
    ne = 256
    s = scenes.contact_constrained_pendulum(ne)
    ang = np.linspace(0.2, 2.9, ne)                      # body y-axis points from the body to the anchor: rotate about z
    s.q[0, 0, :], s.q[0, 1, :] = np.sin(ang), -np.cos(ang)
    for k, val in zip(range(4), scenes.quat_from_rpy(0.0, 0.0, ang)):
        s.q[0, 3 + k, :] = val
    sim = TimeSteppingSimulator(s)
    sim.step(1e-3, 1000)
    q, v = sim.get_state()
    ob = oracle.OracleBatch(s)
    co = ob.run(1e-3, 1000, threads=min(16, os.cpu_count() or 1))
    qo, vo = ob.get_state_soa()
    assert max(np.abs(q - qo).max(), np.abs(v - vo).max()) < 1e-8
    cg = sim.counters()
    assert cg["lcp_failures"] == 0 and co["lcp_failures"] == 0 and cg["lcp_solves"] == co["lcp_solves"] and cg["contacts"] == co["contacts"]
