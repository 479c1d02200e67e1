"""Articulated-body forward dynamics: the oracle's link-coordinate ABA and CRB against
each other, against first principles (pendulum, sum of J^T M J, energy conservation), and the product's world-coordinate
device code (compiled for the host) against the oracle.  GPU parity of the same kernels is in test_gpu_rc.py."""
import numpy as np
import pytest

import hostsim_api as H
import oracle_api as O
from moby_b200 import scenes

G = (0.0, -9.81, 0.0)


def _state(sc, env=0):
    return sc.rc.jq[:, env].copy(), sc.rc.jqd[:, env].copy()


@pytest.mark.parametrize("n_links,branch,seed", [(2, False, 1), (4, False, 2), (7, False, 3), (10, False, 4), (10, True, 5), (16, True, 6)])
def test_oracle_aba_equals_crb(n_links, branch, seed):
    sc = scenes.chain(3, n_links, seed=seed, branch=branch)
    rng = np.random.default_rng(seed)
    for env in range(3):
        q, qd = _state(sc, env)
        tau = rng.normal(size=n_links - 1)
        a = O.rc_fwd_dyn(sc.rc, 0, q, qd, tau, G, env)
        c = O.rc_fwd_dyn(sc.rc, 1, q, qd, tau, G, env)
        assert np.allclose(a, c, rtol=1e-9, atol=1e-9 * np.abs(c).max())


def test_oracle_pendulum_closed_form():
    sc = scenes.pendulum(1)
    m, l, Ic = 1.0, 1.0, 0.4 * 1.5811 ** 2
    for th, w in [(0.3, 0.0), (1.2, -2.0), (-2.5, 5.0)]:
        # link COM hangs at (cos th, sin th) * l from the joint (loc_child = (-1,0,0)), gravity -y
        qdd = O.rc_fwd_dyn(sc.rc, 0, [th], [w], None, G)[0]
        expect = -m * 9.81 * l * np.cos(th) / (Ic + m * l * l)
        assert abs(qdd - expect) < 1e-12 * max(1.0, abs(expect))


@pytest.mark.parametrize("seed", [11, 12])
def test_oracle_inertia_is_sum_of_JtMJ(seed):
    sc = scenes.chain(1, 8, seed=seed, branch=True)
    q, qd = _state(sc)
    Hm = O.rc_inertia(sc.rc, q)
    L = O.rc_links(sc.rc, q, qd)
    mass, J, _ = sc.rc.env_mass_props(0)
    S = np.zeros_like(Hm)
    for i in range(1, sc.rc.n_links):
        M6 = np.zeros((6, 6))
        M6[:3, :3] = mass[i] * np.eye(3)
        M6[3:, 3:] = L["R"][i] @ np.diag(J[i]) @ L["R"][i].T
        S += L["jac"][i].T @ M6 @ L["jac"][i]
    assert np.allclose(Hm, S, rtol=1e-11, atol=1e-12)
    assert np.allclose(Hm, Hm.T) and np.all(np.linalg.eigvalsh(Hm) > 0)
    # link velocities are J qd
    for i in range(1, sc.rc.n_links):
        v = L["jac"][i] @ qd
        assert np.allclose(v[:3], L["vl"][i], atol=1e-12) and np.allclose(v[3:], L["va"][i], atol=1e-12)


def test_oracle_energy_conserved_rk4():
    sc = scenes.chain(1, 5, seed=21)
    sc.rc.joint_type[:] = scenes.JOINT_REVOLUTE
    q, qd = _state(sc)
    e0 = O.rc_energy(sc.rc, q, qd, G)
    f = lambda q_, qd_: (qd_, O.rc_fwd_dyn(sc.rc, 0, q_, qd_, None, G))  # noqa: E731
    h = 1e-3
    for _ in range(300):
        k1 = f(q, qd); k2 = f(q + 0.5 * h * k1[0], qd + 0.5 * h * k1[1]); k3 = f(q + 0.5 * h * k2[0], qd + 0.5 * h * k2[1])
        k4 = f(q + h * k3[0], qd + h * k3[1])
        q = q + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0]); qd = qd + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
    e1 = O.rc_energy(sc.rc, q, qd, G)
    assert abs(e1 - e0) < 1e-7 * max(1.0, abs(e0))


@pytest.mark.parametrize("n_links,branch,seed", [(2, False, 1), (5, False, 2), (10, False, 3), (10, True, 4), (16, True, 5)])
def test_device_code_matches_oracle(n_links, branch, seed):
    """rc_device.cuh (world-coordinate formulation) on the host vs the oracle (link coordinates with transforms)."""
    sc = scenes.chain(2, n_links, seed=seed, branch=branch)
    sc.q[0, :3, :] = np.array([[0.3, -0.2], [0.1, 0.4], [-0.5, 0.2]])          # base away from the origin, rotated
    sc.q[0, 3:, 1] = scenes.quat_from_rpy(np.float64(0.3), np.float64(-0.7), np.float64(1.1))
    rng = np.random.default_rng(seed)
    for env in range(2):
        q, qd = _state(sc, env)
        tau = rng.normal(size=n_links - 1)
        for algo in (0, 1):
            ref = O.rc_fwd_dyn(sc.rc, algo, q, qd, tau, G, env)
            out, links = H.rc_eval(sc.rc, algo, q, qd, tau, G, env)
            assert np.allclose(out, ref, rtol=1e-9, atol=1e-10 * max(1.0, np.abs(ref).max())), (algo, np.abs(out - ref).max())
        Hd, _ = H.rc_eval(sc.rc, 2, q, qd, None, G, env)
        assert np.allclose(Hd, O.rc_inertia(sc.rc, q, env), rtol=1e-10, atol=1e-12)
        L = O.rc_links(sc.rc, q, qd, env)
        assert np.allclose(links["x"], L["x"], atol=1e-12)
        assert np.allclose(links["vl"], L["vl"], atol=1e-11) and np.allclose(links["va"], L["va"], atol=1e-12)
        for i in range(n_links):
            Rq = scenes._rotmat(links["quat"][i])
            assert np.allclose(Rq, L["R"][i], atol=1e-12)


def test_ur10_model_is_consistent():
    sc = scenes.ur10(1, q_jitter=0.0)
    q = np.zeros(9)
    L = O.rc_links(sc.rc, q, q)
    # at q = 0 the link COM frames sit where model.sdf puts them
    assert np.allclose(L["x"][1], [0, 0, 0.1273], atol=1e-9)
    assert np.allclose(L["x"][7], [1.1843, 0.256 + 0.035, 0.0116], atol=1e-5)
    assert np.allclose(L["x"][8], [1.1843 - 0.0205, 0.256 + 0.0798, 0.0116], atol=1e-5)
    Hm = O.rc_inertia(sc.rc, q)
    assert np.all(np.linalg.eigvalsh(Hm) > 0)
    a = O.rc_fwd_dyn(sc.rc, 0, q, q, None, sc.gravity)
    c = O.rc_fwd_dyn(sc.rc, 1, q, q, None, sc.gravity)
    assert np.allclose(a, c, rtol=1e-8, atol=1e-8 * np.abs(c).max())
