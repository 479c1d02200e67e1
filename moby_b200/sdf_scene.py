"""SDF model -> kinematic / inertial tables of a fixed-base RCArticulatedBody (the subset of src/SDFReader.cpp the accelerated
path can use): <link> pose / inertial (mass, principal moments, COM offset) and <joint> parent / child / type / axis
(`use_parent_model_frame` as SDFReader.cpp:556-562 reads it: the parent link's frame) / limits, as example/ur10/model.sdf uses them.

What is NOT read: collision and visual geometry (the UR10's are triangle meshes, outside the path: scenes.ur10 puts sphere /
box proxies on the gripper), off-diagonal inertia terms and rotated inertial frames (refused), joint dynamics.  Joint limits are
returned but not enforced.  The link order is breadth-first from the link whose parent is `world`, so that
parents precede children as b200moby_rc_desc requires; SDF joints sit at the child link's frame origin (SDFReader.cpp:973-978).
"""
import xml.etree.ElementTree as ET

import numpy as np

from . import scenes


def _nums(text, n):
    v = [float(x) for x in (text or "").split()]
    if len(v) != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return v


def load_sdf_model(path):
    """Returns (name, links, joints): links = [(name, pose6, com3, mass, (ixx, iyy, izz))] in tree order, joints[i] for link
    i > 0 = dict(name, parent (index), type (scenes.JOINT_*), axis (unit, child link frame), lower, upper); joints[0] is the
    joint that welds the root to the world (or None)."""
    root = ET.parse(path).getroot()
    model = root.find("model")
    if model is None:
        raise ValueError("no <model>")
    raw = {}
    for l in model.findall("link"):
        pose = _nums(l.findtext("pose"), 6) if l.find("pose") is not None else [0.0] * 6
        ine = l.find("inertial")
        if ine is None:
            raise ValueError(f"link {l.get('name')!r}: no <inertial>")
        ipose = _nums(ine.findtext("pose"), 6) if ine.find("pose") is not None else [0.0] * 6
        if any(ipose[3:]):
            raise ValueError(f"link {l.get('name')!r}: rotated inertial frames are not supported (the link frame must be principal)")
        I = ine.find("inertia")
        if any(float(I.findtext(k, "0")) != 0.0 for k in ("ixy", "ixz", "iyz")):
            raise ValueError(f"link {l.get('name')!r}: off-diagonal inertia terms are not supported")
        raw[l.get("name")] = (l.get("name"), tuple(pose), tuple(ipose[:3]), float(ine.findtext("mass")),
                              tuple(float(I.findtext(k)) for k in ("ixx", "iyy", "izz")))
    jraw = []
    for j in model.findall("joint"):
        t = j.get("type")
        if t not in ("revolute", "prismatic"):
            raise ValueError(f"joint {j.get('name')!r}: type {t!r} is not on the accelerated path (revolute / prismatic)")
        ax = j.find("axis")
        xyz = np.array(_nums(ax.findtext("xyz"), 3))
        lim = ax.find("limit")
        jraw.append(dict(name=j.get("name"), parent=j.findtext("parent"), child=j.findtext("child"),
                         type=scenes.JOINT_REVOLUTE if t == "revolute" else scenes.JOINT_PRISMATIC, xyz=xyz / np.linalg.norm(xyz),
                         model_frame=(ax.findtext("use_parent_model_frame", "0").strip() == "1"),
                         lower=float(lim.findtext("lower", "-inf")) if lim is not None else -np.inf,
                         upper=float(lim.findtext("upper", "inf")) if lim is not None else np.inf))
    roots = [j for j in jraw if j["parent"] == "world"]
    children = {j["child"] for j in jraw}
    base = roots[0]["child"] if roots else next(n for n in raw if n not in children)
    order, joints = [base], [roots[0] if roots else None]
    i = 0
    while i < len(order):
        for j in jraw:
            if j["parent"] == order[i] and j["child"] not in order:
                order.append(j["child"]); joints.append(j)
        i += 1
    if len(order) != len(raw):
        raise ValueError("links not connected to the base (kinematic loop or stray link)")
    index = {n: k for k, n in enumerate(order)}
    links = [raw[n] for n in order]
    out = [joints[0]]
    for k in range(1, len(order)):
        j = dict(joints[k])
        pose = links[k][1]
        R = scenes._rotmat(scenes.quat_from_rpy(np.float64(pose[3]), np.float64(pose[4]), np.float64(pose[5])))
        j["parent"] = index[j["parent"]]
        pp = links[j["parent"]][1]
        Rp = scenes._rotmat(scenes.quat_from_rpy(np.float64(pp[3]), np.float64(pp[4]), np.float64(pp[5])))
        # SDFReader.cpp:556-562: with use_parent_model_frame the axis is given in the PARENT LINK's frame (axis.pose =
        # parent->get_pose()), otherwise in the joint's own frame, which sits on the child link; stored here in the child frame
        j["axis"] = (R.T @ (Rp @ j["xyz"])) if j["model_frame"] else j["xyz"]
        out.append(j)
    return model.get("name"), links, out
