"""Box-box narrowphase under rule H5 (SURVEY.md 8a'): geometric properties of the oracle's routine, and the product's
device code (compiled for the host) against the oracle on stacks (contacts, LCP sizes, trajectories)."""
import numpy as np
import pytest

import hostsim_api as H
import oracle_api as O
from moby_b200 import scenes


def _two_boxes(dimsA, dimsB, posB, quatB=(0, 0, 0, 1), posA=(0, 0, 0)):
    s = scenes.SceneBatch(1, 2)
    s.set_box(0, *dimsA, density=1.0)
    s.set_box(1, *dimsB, density=1.0)
    s.set_contact(0, 1, NK=4)
    s.q[0, :3, 0] = posA
    s.q[1, :3, 0] = posB
    s.q[1, 3:, 0] = quatB
    s.gravity = (0.0, 0.0, 0.0)
    return s


def _contacts(s):
    sim = O.OracleSim(s)
    n, MM, qq, nc = sim.assemble()
    sim.step(1e-9)      # populates last_contacts without moving anything measurably
    return sim.last_contacts()


def test_face_face_resting_gives_the_top_face_corners():
    s = _two_boxes((1, 1, 1), (0.9, 1, 0.9), (0.01, 1.0, -0.02))
    c = _contacts(s)
    assert c["count"] == 4
    assert np.allclose(c["normal"], [[0, -1, 0]] * 4)            # from geom2 (upper box) toward geom1 (lower box)
    assert np.allclose(c["dist"], 0.0, atol=1e-15)
    pts = sorted(map(tuple, np.round(c["point"], 12)))
    exp = sorted((0.01 + sx * 0.45, 0.5, -0.02 + sz * 0.45) for sx in (-1, 1) for sz in (-1, 1))
    assert np.allclose(pts, exp)
    assert np.all(c["pair"] == 0 * 2 + 1)


def test_overhanging_face_is_clipped():
    s = _two_boxes((1, 1, 1), (1, 1, 1), (0.6, 1.0, 0.0))
    c = _contacts(s)
    assert c["count"] == 4
    xs = sorted(np.round(c["point"][:, 0], 12))
    assert np.allclose(xs, [0.1, 0.1, 0.5, 0.5])                   # overlap strip x in [0.1, 0.5]


def test_yawed_face_gives_an_octagon():
    q = scenes.quat_from_rpy(np.float64(0.0), np.float64(np.pi / 4), np.float64(0.0))
    s = _two_boxes((1, 1, 1), (1, 1, 1), (0.0, 1.0, 0.0), tuple(q))
    c = _contacts(s)
    assert c["count"] == 8
    assert np.allclose(c["point"][:, 1], 0.5)
    r = np.hypot(c["point"][:, 0], c["point"][:, 2])
    assert np.allclose(r, r[0])                                   # regular octagon


def test_gap_and_penetration_distances():
    for gap in (0.25, 1e-7, -1e-3):
        s = _two_boxes((1, 1, 1), (0.8, 0.6, 0.8), (0.05, 0.5 + 0.3 + gap, 0.0))
        sim = O.OracleSim(s)
        pairs = sim.last_contacts()
        hs = H.HostSim(s)
        # signed distance through the stage path of the host-compiled device code: contact generation threshold
        sim.step(1e-9)
        c = sim.last_contacts()
        if gap > 1e-6:
            assert c["count"] == 0
        else:
            assert c["count"] == 4 and np.allclose(c["dist"], gap, atol=1e-12)


def test_edge_edge_contact():
    # upper box rolled 45 deg about x and yawed 90 deg: its lowest edge (along world x after the yaw... ) crosses an edge of the tilted lower box
    qA = scenes.quat_from_rpy(np.float64(0.0), np.float64(0.0), np.float64(np.pi / 4))   # lower box rolled about z: top edge along z
    qB = scenes.quat_from_rpy(np.float64(np.pi / 4), np.float64(0.0), np.float64(0.0))   # upper box rolled about x: bottom edge along x
    h = np.sqrt(0.5)
    s = _two_boxes((1, 1, 1), (1, 1, 1), (0.0, 2 * h + 5e-7, 0.0), tuple(qB))
    s.q[0, 3:, 0] = qA
    c = _contacts(s)
    assert c["count"] == 1
    assert np.allclose(c["point"][0], [0, h + 2.5e-7, 0], atol=1e-9)
    assert np.allclose(np.abs(c["normal"][0]), [0, 1, 0], atol=1e-9) and c["normal"][0][1] < 0
    assert abs(c["dist"][0] - 5e-7) < 1e-12


@pytest.mark.parametrize("n_boxes,steps,yaw", [(2, 60, 0.0), (3, 60, 0.0), (3, 40, 0.3)])
def test_stack_device_code_matches_oracle(n_boxes, steps, yaw):
    """stack.xml as shipped (3 registered boxes): same contact counts and LCP sizes, trajectories within 1e-9."""
    s = scenes.box_stack(2, n_boxes, yaw_jitter=yaw)
    hs = H.HostSim(s)
    sims = [O.OracleSim(s, e) for e in range(2)]
    for _ in range(steps):
        hs.step(1e-3)
        for sm in sims:
            sm.step(1e-3)
    for e, sm in enumerate(sims):
        qo, vo = sm.get_state()
        assert np.abs(hs.q[:, :, e] - qo).max() < 1e-9 and np.abs(hs.v[:, :, e] - vo).max() < 1e-9
        assert np.abs(qo[:n_boxes, 1] - (0.5 + np.arange(n_boxes))).max() < 1e-5      # the stack stands
    ch = hs.counters_dict()
    co = {k: sum(sm.counters()[k] for sm in sims) for k in ("contacts", "lcp_solves", "mini_steps")}
    for k in co:
        assert ch[k] == co[k], (k, ch[k], co[k])
    assert ch["max_lcp_n"] == max(sm.counters()["max_lcp_n"] for sm in sims)
    if yaw == 0.0:
        assert ch["max_lcp_n"] == 4 * n_boxes * 8
    assert ch["overflow"] == 0 and ch["lcp_failures"] == 0


# ---- test/VClipTest.cpp re-expressed: box-box signed distance against an independent computation ----
def _rand_pose(rng, lo, hi):
    q = rng.uniform(-1, 1, 4)                       # VClipTest.cpp:60-68: quaternion from four uniforms, normalised
    q /= np.linalg.norm(q)
    return rng.uniform(lo, hi, 3), scenes._rotmat(np.array([q[0], q[1], q[2], q[3]]))


def _independent_distance(cB, RB):
    """Distance of two separated 2x2x2 boxes (A at the origin, axis-aligned) by bound-constrained minimisation of
    |a - (cB + RB b)|^2 over a, b in [-1,1]^3 -- a convex problem; no feature enumeration involved."""
    from scipy.optimize import minimize
    def f(u):
        d = u[:3] - (cB + RB @ u[3:])
        J = np.concatenate([2 * d, -2 * (RB.T @ d)])
        return d @ d, J
    best = np.inf
    for start in (np.zeros(6), np.concatenate([np.sign(cB), -np.sign(RB.T @ cB)])):
        r = minimize(f, start, jac=True, bounds=[(-1, 1)] * 6, method="L-BFGS-B", options=dict(maxiter=2000, ftol=1e-18, gtol=1e-14))
        best = min(best, r.fun)
    return np.sqrt(best)


def test_vclip_apart_boxes_distance():
    """test/VClipTest.cpp:24-107 (Apart_BB_VClip): two 2x2x2 boxes, the second translated by U[2.5,10]^3 and rotated at
    random; the signed distance must equal the polytope distance to 1e-6 (the gtest's TOL).  Oracle and device code."""
    rng = np.random.default_rng(24)
    I, ext = np.eye(3), np.full(3, 2.0)
    worst = 0.0
    kinds = set()
    for _ in range(300):
        cB, RB = _rand_pose(rng, 2.5, 10.0)
        ref = _independent_distance(cB, RB)
        d_o, pA, pB = O.boxbox_dist(np.zeros(3), I, ext, cB, RB, ext)
        d_h, pAh, pBh = H.boxbox_dist(np.zeros(3), I, ext, cB, RB, ext)
        assert d_o == d_h and np.array_equal(pA, pAh) and np.array_equal(pB, pBh)              # device code == oracle, bit for bit
        assert abs(d_o - ref) < 1e-6, (d_o, ref)
        assert abs(np.linalg.norm(pA - pB) - d_o) < 1e-9                                         # the closest points realise the distance
        assert np.abs(pA).max() <= 1 + 1e-12 and np.abs(RB.T @ (pB - cB)).max() <= 1 + 1e-12     # ... and lie on the boxes
        worst = max(worst, abs(d_o - ref))
        onA = np.isclose(np.abs(pA), 1.0).sum()
        onB = np.isclose(np.abs(RB.T @ (pB - cB)), 1.0).sum()
        kinds.add((int(onA), int(onB)))
    assert (3, 3) in kinds or (3, 2) in kinds or (2, 3) in kinds         # vertex-vertex / vertex-edge cases were among them


def test_vclip_penetrating_boxes_depth():
    """test/VClipTest.cpp:177-247 (Penetrating_BB_Vclip): translation U[-1,1]^3; the signed distance of overlapping boxes is
    the signed distance of the origin to their Minkowski difference (here: its convex hull through scipy / qhull, facet
    equations n.x + c <= 0 inside)."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(177)
    I, ext = np.eye(3), np.full(3, 2.0)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
    n_pen = 0
    for _ in range(200):
        cB, RB = _rand_pose(rng, -1.0, 1.0)
        VB = cB + corners @ RB.T
        md = (corners[:, None, :] - VB[None, :, :]).reshape(-1, 3)
        hull = ConvexHull(md)
        ref = hull.equations[:, 3].max()                              # signed distance of the origin (negative inside)
        d_o, _, _ = O.boxbox_dist(np.zeros(3), I, ext, cB, RB, ext)
        d_h, _, _ = H.boxbox_dist(np.zeros(3), I, ext, cB, RB, ext)
        assert d_o == d_h
        if ref < 0:
            n_pen += 1
            assert abs(d_o - ref) < 1e-6, (d_o, ref)
    assert n_pen > 150
