"""Moby XML scene -> batch descriptor: the subset of the reference's scene format that the accelerated
path covers -- free rigid bodies with Sphere / Box / Plane collision geometry (or the rimless wheel's collision-detection
plugin, mapped to the built-in spoke-tip shape), a GravityForce, and a
TimeSteppingSimulator with ContactParameters / DisabledPair children -- so existing scenes such as
example/simple-contact/simplest.xml, example/bouncing-ball/bouncing-ball.xml and example/stacks/*.xml load unmodified.

Attribute semantics follow the reference's loaders:
  Primitive            mass | density, position, rpy | quat (w x y z)          src/Primitive.cpp:240-300, XMLTree.cpp:407-421
  Box / Sphere         xlen ylen zlen / radius                                 src/BoxPrimitive.cpp:649-651, SpherePrimitive.cpp:368
  RigidBody            enabled, mass, position, rpy | quat, linear-velocity, angular-velocity; InertiaFromPrimitive and
                       CollisionGeometry children by primitive-id             src/RigidBody.cpp:165-330
  GravityForce         accel                                                   src/GravityForce.cpp:81
  ContactParameters    object1-id object2-id epsilon mu-coulomb mu-viscous compliance friction-cone-edges
                                                                               src/ContactParameters.cpp:57-136
  TimeSteppingSimulator min-step-size, contact-dist-thresh, constraint-stabilization-max-iterations; DynamicBody,
                       RecurrentForce, DisabledPair children                   src/TimeSteppingSimulator.cpp:470,
                                                                               ConstraintSimulator.cpp:585-611, Simulator.cpp:860-928
Bodies are numbered in the order of the simulator's DynamicBody tags (rule H4: bodies by scene index).  Pairs without a
ContactParameters entry keep the constraint defaults (4 cone edges, mu = epsilon = 0, UnilateralConstraint.cpp:52,
ConstraintSimulator.cpp:402-404).  constraint-stabilization-max-iterations is carried into the descriptor; a file that omits
it gets the reference's default -- stabilize after every step, no iteration limit (ConstraintStabilization.cpp:53-59,
TimeSteppingSimulator.cpp:474) -- which the kernels run (stab_device.cuh).  Anything outside the subset (joints,
articulated bodies, other primitives, geometry offsets on moving bodies) raises ValueError naming the construct --
nothing is silently dropped.
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

from . import scenes

_PRIMS = ("Box", "Sphere", "Plane")
# Collision-detection plugins (XMLReader.cpp:334-388, ConstraintSimulator.cpp:562-572) the accelerated path has a built-in
# equivalent for: plugin file name -> {body id: (shape setter arguments)}.  The rimless wheel's plugin finds its bodies by
# the ids "WHEEL" and "GROUND" (coldet-plugin.cpp:29-36) and takes R, W, N_SPOKES from params.h:4-6.
_COLDET_PLUGINS = {"librimless-wheel-coldet-plugin.so": {"WHEEL": dict(radius=1.0, width=0.0, n_spokes=6), "GROUND": None},
                   # contact-constrained-pendulum-coldet-plugin.cpp:21-37,60-75: bodies "l1" (anchor point (0,1,0)) and "world"
                   "libcontact-constrained-pendulum-coldet-plugin.so": {"l1": dict(pin=(0.0, 1.0, 0.0)), "world": dict(pinworld=True)}}
_UNSUPPORTED_PRIMS = ("Cone", "Cylinder", "Torus", "Heightmap", "TriangleMesh", "Polyhedron", "CSG", "GaussianMixture")


def _vec(s, n):
    v = [float(x) for x in s.replace(",", " ").split()]
    if len(v) != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return np.array(v, np.float64)


def _pose(node):
    """(position, quaternion x y z w) of a node carrying position / rpy / quat attributes."""
    x = _vec(node.get("position"), 3) if node.get("position") is not None else np.zeros(3)
    if node.get("quat") is not None:
        w, qx, qy, qz = _vec(node.get("quat"), 4)                     # XMLTree.cpp:415-419: w x y z
        q = np.array([qx, qy, qz, w])
    elif node.get("rpy") is not None:
        r, p, y = _vec(node.get("rpy"), 3)
        q = np.array(scenes.quat_from_rpy(np.float64(r), np.float64(p), np.float64(y)), np.float64)
    elif node.get("aangle") is not None:
        ax, ay, az, ang = _vec(node.get("aangle"), 4)
        nrm = math.sqrt(ax * ax + ay * ay + az * az)
        s = math.sin(0.5 * ang) / nrm
        q = np.array([ax * s, ay * s, az * s, math.cos(0.5 * ang)])
    else:
        q = np.array([0.0, 0.0, 0.0, 1.0])
    return x, q / np.linalg.norm(q)


def _identity(x, q):
    return not np.any(x) and abs(abs(q[3]) - 1.0) < 1e-15


def load_xml(source, n_envs=1):
    """Parse a Moby XML scene (file path or XML text) into a scenes.SceneBatch replicated over n_envs envs.

    Returns (scene, info); info maps body ids to body indices and carries the DRIVER step size if present."""
    text = source if source.lstrip().startswith("<") else open(source).read()
    root = ET.fromstring(text)
    moby = root if root.tag == "MOBY" else root.find("MOBY")
    if moby is None:
        raise ValueError("no <MOBY> element")
    for tag in _UNSUPPORTED_PRIMS:
        if moby.find(f".//{tag}") is not None and any(cg.get("primitive-id") == n.get("id") for n in moby.iter(tag) for cg in moby.iter("CollisionGeometry")):
            raise ValueError(f"<{tag}> collision geometry is outside the accelerated path (SURVEY.md section 2)")
    for tag in ("RCArticulatedBody", "MCArticulatedBody", "RevoluteJoint", "PrismaticJoint", "FixedJoint", "SphericalJoint", "UniversalJoint"):
        if moby.find(f".//{tag}") is not None:
            raise ValueError(f"<{tag}>: articulated bodies are built with scenes.ArticulatedBody, not loaded from XML yet")
    prims = {n.get("id"): n for tag in _PRIMS for n in moby.iter(tag)}
    bodies = {n.get("id"): n for n in moby.iter("RigidBody")}
    forces = {n.get("id"): n for n in moby.iter("GravityForce")}
    sims = list(moby.iter("TimeSteppingSimulator"))
    if len(sims) != 1:
        raise ValueError("exactly one <TimeSteppingSimulator> expected (EventDrivenSimulator and others are out of scope)")
    sim = sims[0]
    plugin_bodies = {}
    if sim.get("collision-detection-plugin") is not None:
        pl = {n.get("id"): n.get("plugin") for n in moby.iter("CollisionDetectionPlugin")}.get(sim.get("collision-detection-plugin"))
        if pl not in _COLDET_PLUGINS:
            raise ValueError(f"collision-detection-plugin {pl!r}: no built-in equivalent on the accelerated path (known: {sorted(_COLDET_PLUGINS)})")
        plugin_bodies = _COLDET_PLUGINS[pl]
    order = [n.get("dynamic-body-id") for n in sim.findall("DynamicBody")]
    for b in order:
        if b not in bodies:
            raise ValueError(f"DynamicBody {b!r} is not a <RigidBody> of this file")
    index = {b: i for i, b in enumerate(order)}
    s = scenes.SceneBatch(n_envs, len(order))
    s.name = "xml"
    # simulator attributes
    if sim.get("min-step-size") is not None:
        s.min_step_size = float(sim.get("min-step-size"))
    if sim.get("contact-dist-thresh") is not None:
        s.contact_dist_thresh = float(sim.get("contact-dist-thresh"))
    stab = sim.get("constraint-stabilization-max-iterations")
    # absent: the reference's default, UINT_MAX = no limit (-1 in the descriptor); present: that many (0 = off)
    s.stabilization_max_iterations = -1 if stab is None else min(int(stab), 2 ** 31 - 1)
    info = {"bodies": index, "stabilization_max_iterations": None if stab is None else int(stab)}
    drv = root.find("DRIVER")
    if drv is not None and drv.get("step-size") is not None:
        info["step_size"] = float(drv.get("step-size"))
    g = np.zeros(3)
    for rf in sim.findall("RecurrentForce"):
        f = forces.get(rf.get("recurrent-force-id"))
        if f is None:
            raise ValueError(f"RecurrentForce {rf.get('recurrent-force-id')!r}: only GravityForce is on the accelerated path")
        g = g + _vec(f.get("accel"), 3)
    s.gravity = tuple(g)
    # bodies
    for bid, i in index.items():
        node = bodies[bid]
        enabled = node.get("enabled", "true").strip().lower() in ("true", "1")
        bx, bq = _pose(node)
        cgs = node.findall("CollisionGeometry")
        if len(cgs) > 1:
            raise ValueError(f"body {bid!r}: one CollisionGeometry per body on the accelerated path")
        shape_set = False
        if cgs:
            cg = cgs[0]
            if any(cg.get(k) is not None for k in ("relative-origin", "relative-rpy", "relative-quat")):
                raise ValueError(f"body {bid!r}: CollisionGeometry offsets are not supported")
            p = prims.get(cg.get("primitive-id"))
            if p is None and cg.get("primitive-id") is None and plugin_bodies.get(bid) is not None:
                spec = plugin_bodies[bid]                            # geometry without a primitive: the plugin's own shape
                if "pin" in spec:
                    s.shape[i, :] = scenes.SHAPE_PIN
                    for k in range(3):
                        s.dims[i, k, :] = spec["pin"][k]
                elif "pinworld" in spec:
                    s.shape[i, :] = scenes.SHAPE_PINWORLD
                else:
                    s.set_wheel(i, **spec)
                cgs = []
            elif p is None:
                raise ValueError(f"body {bid!r}: primitive {cg.get('primitive-id')!r} is not a Box / Sphere / Plane of this file")
        if cgs:
            px, pq = _pose(p)
            if p.tag == "Plane":
                if enabled:
                    raise ValueError(f"body {bid!r}: a Plane on an enabled body is not supported")
                # static half-space: the primitive's pose composes with the body's (Primitive.cpp:270-300)
                R = scenes._rotmat(bq)
                s.set_plane(i, quat=tuple(scenes.quat_mul(bq, pq)), pos=tuple(bx + R @ px))
                shape_set = True
            else:
                if not _identity(px, pq):
                    raise ValueError(f"body {bid!r}: a posed {p.tag} primitive (geometry offset from the body frame) is not supported")
                if p.tag == "Box":
                    s.set_box(i, float(p.get("xlen")), float(p.get("ylen")), float(p.get("zlen")), mass=1.0)
                else:
                    s.set_sphere(i, float(p.get("radius")), mass=1.0)
        if not shape_set:
            s.enabled[i, :] = 1 if enabled else 0
            for k in range(3):
                s.q[i, k, :] = bx[k]
            for k in range(4):
                s.q[i, 3 + k, :] = bq[k]
        # inertia: InertiaFromPrimitive (mass or density of the primitive), then the body's own mass attribute
        ifp = node.findall("InertiaFromPrimitive")
        if len(ifp) > 1:
            raise ValueError(f"body {bid!r}: one InertiaFromPrimitive per body on the accelerated path")
        if ifp:
            p = prims.get(ifp[0].get("primitive-id"))
            if p is None or p.tag == "Plane":
                raise ValueError(f"body {bid!r}: InertiaFromPrimitive needs a Box or Sphere of this file")
            kw = {"mass": float(p.get("mass"))} if p.get("mass") is not None else {"density": float(p.get("density", "1.0"))}
            shape, dims = s.shape[i].copy(), s.dims[i].copy()
            if p.tag == "Box":
                s.set_box(i, float(p.get("xlen")), float(p.get("ylen")), float(p.get("zlen")), **kw)
            else:
                s.set_sphere(i, float(p.get("radius")), **kw)
            s.shape[i], s.dims[i] = shape, dims                     # the collision shape stays what CollisionGeometry said
        elif enabled and (node.get("inertia") is None or node.get("mass") is None):
            raise ValueError(f"body {bid!r}: an enabled body needs InertiaFromPrimitive or both mass and inertia attributes")
        if node.get("mass") is not None:
            s.mass[i, :] = float(node.get("mass"))                   # RigidBody.cpp:182-188: J.m only
        if node.get("inertia") is not None:                          # RigidBody.cpp:191-197: J.J, rows separated by ';'
            J = np.array([_vec(r, 3) for r in node.get("inertia").split(";") if r.strip()])
            if J.shape != (3, 3) or np.abs(J - np.diag(np.diag(J))).max() != 0.0:
                raise ValueError(f"body {bid!r}: the body frame must be the principal frame (diagonal inertia matrix)")
            for k in range(3):
                s.inertia[i, k, :] = J[k, k]
        s.enabled[i, :] = 1 if enabled else 0
        if node.get("linear-velocity") is not None:
            s.v[i, 0:3, :] = _vec(node.get("linear-velocity"), 3)[:, None]
        if node.get("angular-velocity") is not None:
            s.v[i, 3:6, :] = _vec(node.get("angular-velocity"), 3)[:, None]
    # contact parameters: defaults for every pair, then the file's entries, then disabled pairs
    nb = len(order)
    for i in range(nb):
        for j in range(i + 1, nb):
            s.set_contact(i, j)
    for cp in sim.findall("ContactParameters"):
        a, b = cp.get("object1-id"), cp.get("object2-id")
        if a not in index or b not in index:
            if a in bodies and b in bodies:
                continue                                             # parameters for bodies the simulator does not register (stack.xml:85-97)
            raise ValueError(f"ContactParameters {a!r} / {b!r}: objects must be rigid bodies of this file")
        nk = int(cp.get("friction-cone-edges", "4"))
        if nk < 4:
            nk = 4                                                   # ContactParameters.cpp:132-136
        s.set_contact(index[a], index[b], mu_coulomb=float(cp.get("mu-coulomb", "0")), mu_viscous=float(cp.get("mu-viscous", "0")),
                      epsilon=float(cp.get("epsilon", "0")), NK=nk, compliance=float(cp.get("compliance", "0")))
    for dp in sim.findall("DisabledPair"):
        a, b = dp.get("object1-id"), dp.get("object2-id")
        if a in index and b in index and a != b:
            s.set_contact(index[a], index[b], NK=0)
    return s, info
