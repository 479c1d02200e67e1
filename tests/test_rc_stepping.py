"""Stepped path with a reduced-coordinate articulated body (SURVEY.md 8 rows a6-a12b with RC links): the product's device
code compiled for the host (world-coordinate dynamics, motion-subspace Jacobians) against the oracle (link-coordinate
dynamics, link Jacobians as the reference builds them, ImpactConstraintHandler.cpp:1869-1878).  1e-9 relative is the
north-star tolerance; observed ~1e-12."""
import numpy as np
import pytest

import hostsim_api as H
import oracle_api as O
from moby_b200 import scenes


def pendulum_on_plane(n_envs=2, eps=0.5, mu=0.3):
    s = scenes.SceneBatch(n_envs, 3)
    for b in range(2):
        s.mass[b, :] = 1.0
        s.inertia[b, :, :] = 0.4 * 1.5811 ** 2
    s.set_sphere(1, 0.2, mass=1.0)
    s.inertia[1, :, :] = 0.4 * 1.5811 ** 2
    s.set_plane(2, pos=(0, -0.9, 0))
    s.set_contact(1, 2, mu_coulomb=mu, epsilon=eps, NK=4)
    rc = scenes.ArticulatedBody(s, 0, 2)
    rc.set_joint(1, 0, scenes.JOINT_REVOLUTE, (0, 0, 1), (0, 0, 0), (-1.0, 0, 0))
    rc.jq[0, :] = np.linspace(0.3, -0.2, n_envs)
    rc.jqd[0, :] = np.linspace(0.0, -1.0, n_envs)
    return s


def _compare(sc, dt, steps, tol, phased=False):
    hs = H.HostSim(sc)
    sims = [O.OracleSim(sc, e) for e in range(sc.n_envs)]
    if phased:
        hs.step_phased(dt, steps)
    else:
        hs.step(dt, steps)
    for e, sm in enumerate(sims):
        sm.step(dt, steps)
        jq, jqd = sm.get_joint_state()
        qo, vo = sm.get_state()
        scale = max(1.0, np.abs(jqd).max(), np.abs(vo).max())
        assert np.abs(hs.jq[:, e] - jq).max() < tol and np.abs(hs.jqd[:, e] - jqd).max() < tol * scale
        assert np.abs(hs.q[:, :, e] - qo).max() < tol and np.abs(hs.v[:, :, e] - vo).max() < tol * scale
    ch = hs.counters_dict()
    for k in ("env_steps", "mini_steps", "lcp_solves", "contacts", "lcp_fast_calls", "lemke_calls"):
        assert ch[k] == sum(sm.counters()[k] for sm in sims), k
    assert ch["lcp_failures"] == 0 and ch["overflow"] == 0
    return hs, sims


def test_pendulum_bouncing_on_a_plane():
    hs, sims = _compare(pendulum_on_plane(), 1e-3, 1500, 1e-10)
    assert hs.counters_dict()["lcp_solves"] >= 3          # the bob really hit the plane


def test_pendulum_phased_schedule_is_identical_to_fused():
    sc = pendulum_on_plane(4)
    a, b = H.HostSim(sc), H.HostSim(sc)
    a.step(1e-3, 800)
    b.step_phased(1e-3, 800)
    assert np.array_equal(a.jq, b.jq) and np.array_equal(a.jqd, b.jqd) and np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v)


@pytest.mark.parametrize("fdyn", [scenes.FDYN_CRB, scenes.FDYN_FSAB])
def test_ur10_with_block_and_table(fdyn):
    """example/ur10 benchmark variant: 9-DoF arm + gripper under the controller.cpp PD law, block resting on the table,
    arm proxies touching the table; CRB (what SDFReader wires for the UR10) and ABA."""
    hs, sims = _compare(scenes.ur10(2, fdyn=fdyn), 5e-4, 300, 1e-9)
    assert hs.counters_dict()["lcp_solves"] > 600         # block island every step plus arm impacts


def test_ur10_aba_and_crb_trajectories_agree():
    a = H.HostSim(scenes.ur10(1, fdyn=scenes.FDYN_CRB))
    b = H.HostSim(scenes.ur10(1, fdyn=scenes.FDYN_FSAB))
    a.step(5e-4, 200)
    b.step(5e-4, 200)
    assert np.abs(a.jq - b.jq).max() < 1e-8 and np.abs(a.jqd - b.jqd).max() < 1e-6


def test_parts_feeder_like_scene():
    """SURVEY 8(d) case 5 variant: prismatic shaker tray (articulated link, box geometry) + a free box part, mu = 0.01:
    box-box contacts between an articulated link and a free body, QP model, PD-driven prismatic joint.  The four-contact
    face LCP is degenerate; where lcp_fast fails and Lemke runs, the tableau (kernels) and the LU-per-pivot (oracle) forms
    can break a ratio-test tie differently and return another valid solution (a few 1e-6 of lateral slip): such envs are
    recognised by their solver-call counts and held to 1e-4, all others to 1e-9."""
    sc = scenes.parts_feeder(4)
    exact = 0
    for e in range(sc.n_envs):
        hs, osim = H.HostSim(sc), O.OracleSim(sc, e)
        hs.step(1e-3, 400, e0=e, e1=e + 1)
        osim.step(1e-3, 400)
        jq, jqd = osim.get_joint_state()
        qo, vo = osim.get_state()
        ch, co = hs.counters_dict(), osim.counters()
        same_path = all(ch[k] == co[k] for k in ("lcp_solves", "lcp_fast_calls", "lemke_calls", "pivots"))
        tol = 1e-9 if same_path else 1e-4
        exact += same_path
        err = max(np.abs(hs.jq[:, e] - jq).max(), np.abs(hs.jqd[:, e] - jqd).max(), np.abs(hs.q[:, :, e] - qo).max(), np.abs(hs.v[:, :, e] - vo).max())
        assert err < tol, (e, err, same_path)
        assert ch["env_steps"] == co["env_steps"] == 400 and ch["lcp_failures"] == 0 and ch["max_lcp_n"] == 32 and ch["lcp_solves"] > 300
        assert hs.q[2, 0, e] > sc.q[2, 0, e]                                   # the part slides down the tilted tray
    assert exact >= 1


@pytest.mark.parametrize("make,dt,steps", [(lambda: scenes.ur10(3, mu=100.0), 5e-4, 120), (lambda: scenes.parts_feeder(5), 1e-3, 150),
                                           (lambda: scenes.ur10(2), 5e-4, 102)])
@pytest.mark.parametrize("budget", [0, 6])
def test_phased_schedule_with_queues_is_identical_to_fused(make, dt, steps, budget):
    """The launch schedule of the GPU path (advance / hard queue / impact classes with a per-env solver budget / stragglers /
    finish) on scenes with an articulated body: bit-identical to the fused per-env loop, whatever the budget."""
    sc = make()
    a, b = H.HostSim(sc), H.HostSim(sc)
    a.step(dt, steps)
    for _ in range(3):
        b.step_phased(dt, steps // 3, rounds=2, pivot_budget=budget)
    assert np.array_equal(a.jq, b.jq) and np.array_equal(a.jqd, b.jqd) and np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v)
    ca, cb = a.counters_dict(), b.counters_dict()
    assert ca == cb, (ca, cb)
    assert ca["lcp_solves"] > 50


def test_ur10_under_the_anitescu_potra_model():
    """-DUSE_AP_MODEL build of the reference (ImpactConstraintHandlerLCP.cpp) on the articulated scene: dense problem data +
    A-P LCP + Lemke only."""
    sc = scenes.ur10(2)
    sc.impact_model = scenes.MODEL_AP
    hs, sims = _compare(sc, 5e-4, 200, 1e-9)
    c = hs.counters_dict()
    assert c["lemke_calls"] >= c["lcp_solves"] > 300 and c["lcp_fast_calls"] == 0 and c["max_lcp_n"] == 24
