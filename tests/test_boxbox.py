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
