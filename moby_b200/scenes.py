"""Batch scene descriptors: the host-side stand-in for the object graph Moby's XMLReader builds
(src/XMLReader.cpp:60-132) for the benchmark scenes of BASELINE.json / SURVEY.md 8(d).

Arrays follow include/b200moby.h: per-env parameters are SoA across envs ([body][env], [body][3][env],
[body_i*nb+body_j][env]); state q is [body][7][env] (x y z qx qy qz qw), v is [body][6][env].
"""
import ctypes as C
import math

import numpy as np

from .capi import RcDesc, SceneDesc

SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_PLANE, SHAPE_WHEEL, SHAPE_PIN, SHAPE_PINWORLD = 0, 1, 2, 3, 4, 5, 6
MODEL_QP, MODEL_AP = 0, 1
JOINT_REVOLUTE, JOINT_PRISMATIC = 1, 2
FDYN_FSAB, FDYN_CRB = 0, 1
NEAR_ZERO = math.sqrt(np.finfo(np.float64).eps)  # Constants.h:21


def quat_from_rpy(roll, pitch, yaw):
    """x y z w quaternion of R = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


class SceneBatch:
    def __init__(self, n_envs, n_bodies):
        ne, nb = n_envs, n_bodies
        self.n_envs, self.n_bodies = ne, nb
        self.shape = np.zeros((nb, ne), np.int32)
        self.enabled = np.ones((nb, ne), np.int32)
        self.mass = np.ones((nb, ne), np.float64)
        self.dims = np.zeros((nb, 3, ne), np.float64)
        self.inertia = np.ones((nb, 3, ne), np.float64)
        self.mu_coulomb = np.zeros((nb * nb, ne), np.float64)
        self.mu_viscous = np.zeros((nb * nb, ne), np.float64)
        self.epsilon = np.zeros((nb * nb, ne), np.float64)
        self.compliance = np.zeros((nb * nb, ne), np.float64)
        self.NK = np.zeros((nb * nb, ne), np.int32)  # 0 = pair disabled
        self.gravity = (0.0, -9.81, 0.0)
        self.contact_dist_thresh = 1e-6      # ConstraintSimulator.cpp:56
        self.min_step_size = NEAR_ZERO       # TimeSteppingSimulator.cpp:48
        self.min_step_size_env = None        # optional per-env override, [env]
        self.impact_model = MODEL_QP
        # constraint-stabilization-max-iterations: 0 = off, < 0 = the reference's default (unlimited), > 0 = that many
        # (ConstraintStabilization.cpp:53-59).  The builders below keep 0 unless told otherwise: BASELINE's numbers of round 1
        # were taken without it; bench.py --stabilization turns it on.
        self.stabilization_max_iterations = 0
        self.q = np.zeros((nb, 7, ne), np.float64)
        self.q[:, 6, :] = 1.0
        self.v = np.zeros((nb, 6, ne), np.float64)
        self.name = "custom"
        self.rc = None                        # optional ArticulatedBody
        self.max_contacts = 0                 # working-set bounds per env (0 = worst case over all pairs)
        self.max_lcp_n = 0

    # ---- primitives (InertiaFromPrimitive: BoxPrimitive / SpherePrimitive::calc_mass_properties) ----
    def set_box(self, b, xlen, ylen, zlen, density=None, mass=None, envs=slice(None)):
        xlen, ylen, zlen = (np.broadcast_to(np.asarray(a, np.float64), self.shape[b, envs].shape) for a in (xlen, ylen, zlen))
        self.shape[b, envs] = SHAPE_BOX
        self.dims[b, 0, envs], self.dims[b, 1, envs], self.dims[b, 2, envs] = xlen, ylen, zlen
        m = np.asarray(mass, np.float64) if mass is not None else np.asarray(density, np.float64) * xlen * ylen * zlen
        self.mass[b, envs] = m
        self.inertia[b, 0, envs] = m * (ylen ** 2 + zlen ** 2) / 12.0
        self.inertia[b, 1, envs] = m * (xlen ** 2 + zlen ** 2) / 12.0
        self.inertia[b, 2, envs] = m * (xlen ** 2 + ylen ** 2) / 12.0

    def set_sphere(self, b, radius, density=None, mass=None, envs=slice(None)):
        radius = np.broadcast_to(np.asarray(radius, np.float64), self.shape[b, envs].shape)
        self.shape[b, envs] = SHAPE_SPHERE
        self.dims[b, 0, envs] = radius
        m = np.asarray(mass, np.float64) if mass is not None else np.asarray(density, np.float64) * (4.0 / 3.0) * math.pi * radius ** 3
        self.mass[b, envs] = m
        for k in range(3):
            self.inertia[b, k, envs] = 0.4 * m * radius ** 2

    def set_wheel(self, b, radius=1.0, width=0.0, n_spokes=6, mass=1.0, inertia=(2.0, 1.0, 2.0), envs=slice(None)):
        """Rimless wheel of example/rimless-wheel (params.h:4-6, wheel.xml:42): spokes in the body's x-z plane, tips at
        (cos(2 pi i / N) R, +-W/2, sin(2 pi i / N) R); it collides with planes only, through the spoke-tip generator of
        coldet-plugin.cpp:86-137,222-310."""
        self.shape[b, envs] = SHAPE_WHEEL
        self.dims[b, 0, envs], self.dims[b, 1, envs], self.dims[b, 2, envs] = radius, width, n_spokes
        self.mass[b, envs] = mass
        for k in range(3):
            self.inertia[b, k, envs] = inertia[k]

    def set_plane(self, b, quat=(0, 0, 0, 1), pos=(0, 0, 0)):
        """Static half-space y<=0 of the body frame (PlanePrimitive)."""
        self.shape[b, :] = SHAPE_PLANE
        self.enabled[b, :] = 0
        for k in range(3):
            self.q[b, k, :] = pos[k]
        for k in range(4):
            self.q[b, 3 + k, :] = quat[k]

    def set_contact(self, i, j, mu_coulomb=0.0, mu_viscous=0.0, epsilon=0.0, NK=4, compliance=0.0, envs=slice(None)):
        i, j = min(i, j), max(i, j)
        p = i * self.n_bodies + j
        self.mu_coulomb[p, envs], self.mu_viscous[p, envs] = mu_coulomb, mu_viscous
        self.epsilon[p, envs], self.compliance[p, envs], self.NK[p, envs] = epsilon, compliance, NK

    def cdesc(self):
        """ctypes descriptor; keeps the numpy arrays alive on the returned object."""
        d = SceneDesc()
        d.n_envs, d.n_bodies = self.n_envs, self.n_bodies
        keep = []
        for name, ct in (("shape", C.c_int), ("enabled", C.c_int), ("mass", C.c_double), ("dims", C.c_double),
                         ("inertia", C.c_double), ("mu_coulomb", C.c_double), ("mu_viscous", C.c_double),
                         ("epsilon", C.c_double), ("compliance", C.c_double), ("NK", C.c_int)):
            a = np.ascontiguousarray(getattr(self, name))
            keep.append(a)
            setattr(d, name, a.ctypes.data_as(C.POINTER(ct)))
        d.gravity = (C.c_double * 3)(*self.gravity)
        d.contact_dist_thresh, d.min_step_size = self.contact_dist_thresh, self.min_step_size
        d.impact_model, d.stabilization_max_iterations = self.impact_model, self.stabilization_max_iterations
        if self.min_step_size_env is not None:
            a = np.ascontiguousarray(self.min_step_size_env, np.float64)
            keep.append(a)
            d.min_step_size_env = a.ctypes.data_as(C.POINTER(C.c_double))
        d.max_contacts, d.max_lcp_n = self.max_contacts, self.max_lcp_n
        if self.rc is not None:
            rd = self.rc.cdesc()
            keep.append(rd)
            d.rc = C.pointer(rd)
        d._keep = keep
        return d


class ArticulatedBody:
    """Fixed-base reduced-coordinate body (Moby RCArticulatedBody) whose links are bodies
    [first_body, first_body + n_links) of `scene`; mirrors b200moby_rc_desc.  Link body frames are COM frames with
    principal axes; joint k = link k+1's inboard joint.  jq / jqd: [dof][env] initial joint state."""

    def __init__(self, scene, first_body, n_links, fdyn=FDYN_FSAB):
        self.scene, self.first_body, self.n_links = scene, first_body, n_links
        self.parent = np.zeros(n_links, np.int32)
        self.joint_type = np.full(n_links, JOINT_REVOLUTE, np.int32)
        self.joint_axis = np.zeros((n_links, 3)); self.joint_axis[:, 2] = 1.0
        self.loc_parent = np.zeros((n_links, 3))
        self.loc_child = np.zeros((n_links, 3))
        self.rel_quat = np.zeros((n_links, 4)); self.rel_quat[:, 3] = 1.0
        self.fdyn = fdyn
        self.ctrl = None                      # (kp, kv, amp, freq), each [dof]
        nd = n_links - 1
        self.jq = np.zeros((nd, scene.n_envs))
        self.jqd = np.zeros((nd, scene.n_envs))
        scene.enabled[first_body, :] = 0      # the base is welded to the world
        scene.rc = self

    @property
    def n_dof(self):
        return self.n_links - 1

    def set_joint(self, link, parent, jtype, axis, loc_parent, loc_child, rel_quat=(0, 0, 0, 1)):
        self.parent[link], self.joint_type[link] = parent, jtype
        self.joint_axis[link], self.loc_parent[link], self.loc_child[link], self.rel_quat[link] = axis, loc_parent, loc_child, rel_quat

    def set_controller(self, kp, kv, amp, freq):
        self.ctrl = tuple(np.ascontiguousarray(a, np.float64) for a in (kp, kv, amp, freq))

    def cdesc(self):
        d = RcDesc()
        d.n_links, d.first_body, d.fdyn_algorithm = self.n_links, self.first_body, self.fdyn
        keep = []
        for name, ct in (("parent", C.c_int), ("joint_type", C.c_int), ("joint_axis", C.c_double), ("loc_parent", C.c_double),
                         ("loc_child", C.c_double), ("rel_quat", C.c_double)):
            a = np.ascontiguousarray(getattr(self, name), np.int32 if ct is C.c_int else np.float64)
            keep.append(a)
            setattr(d, name, a.ctypes.data_as(C.POINTER(ct)))
        if self.ctrl is not None:
            for name, a in zip(("ctrl_kp", "ctrl_kv", "ctrl_amp", "ctrl_freq"), self.ctrl):
                keep.append(a)
                setattr(d, name, a.ctypes.data_as(C.POINTER(C.c_double)))
        d._keep = keep
        return d

    def link_poses(self, env):
        """World pose (x [link][3], R [link][3][3]) of every link COM frame of env `env` at the stored joint state."""
        sc, b0 = self.scene, self.first_body
        x = np.zeros((self.n_links, 3))
        R = np.zeros((self.n_links, 3, 3))
        x[0], R[0] = sc.q[b0, :3, env], _rotmat(sc.q[b0, 3:, env] / np.linalg.norm(sc.q[b0, 3:, env]))
        for i in range(1, self.n_links):
            p = self.parent[i]
            a = self.joint_axis[i] / np.linalg.norm(self.joint_axis[i])
            R0 = _rotmat(self.rel_quat[i] / np.linalg.norm(self.rel_quat[i]))
            qi = self.jq[i - 1, env]
            if self.joint_type[i] == JOINT_REVOLUTE:
                K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
                Rrel = R0 @ (np.eye(3) + np.sin(qi) * K + (1 - np.cos(qi)) * (K @ K))
                r = self.loc_parent[i] - Rrel @ self.loc_child[i]
            else:
                Rrel = R0
                r = self.loc_parent[i] + Rrel @ (a * qi - self.loc_child[i])
            R[i] = R[p] @ Rrel
            x[i] = x[p] + R[p] @ r
        return x, R

    def env_mass_props(self, env):
        """(mass [link], J [link][3], base pose [7]) of env `env` (what the oracle / host-compiled checks take)."""
        b0, nl, sc = self.first_body, self.n_links, self.scene
        mass = np.ascontiguousarray(sc.mass[b0:b0 + nl, env], np.float64)
        J = np.ascontiguousarray(sc.inertia[b0:b0 + nl, :, env], np.float64)
        pose = np.ascontiguousarray(sc.q[b0, :, env], np.float64)
        return mass, J, pose


# ---------------- the scenes of SURVEY.md 8(d) ----------------
def sitting_box(n_envs=1, NK=8, mu=0.0, eps=0.0, y0=0.5):
    """example/simple-contact/simplest.xml: unit box (density 1) on the plane y=0, body order box, ground."""
    s = SceneBatch(n_envs, 2)
    s.name = "sitting-box"
    s.set_box(0, 1.0, 1.0, 1.0, density=1.0)
    s.set_plane(1)
    s.set_contact(0, 1, mu_coulomb=mu, epsilon=eps, NK=NK)
    s.q[0, 1, :] = y0
    return s


def bouncing_ball(n_envs=1, eps=1.0, y0=1.5):
    """example/bouncing-ball/bouncing-ball.xml: r=1 sphere, density 1, omega_y = 10, epsilon = 1, 4 cone edges."""
    s = SceneBatch(n_envs, 2)
    s.name = "bouncing-ball"
    s.set_sphere(0, 1.0, density=1.0)
    s.set_plane(1)
    s.set_contact(0, 1, mu_coulomb=0.0, epsilon=eps, NK=4)
    s.q[0, 1, :] = y0
    s.v[0, 4, :] = 10.0
    return s


def sphere_stack(n_envs=1):
    """example/stacks/sphere-stack.xml: three unit spheres (mass 1) at z=1,3,5 on the plane z=0, g=(0,0,-9.81), 16 edges."""
    s = SceneBatch(n_envs, 4)
    s.name = "sphere-stack"
    for b in range(3):
        s.set_sphere(b, 1.0, mass=1.0)
        s.q[b, 2, :] = 1.0 + 2.0 * b
    s.set_plane(3, quat=tuple(quat_from_rpy(np.float64(1.5707963267949), 0.0, 0.0)))
    s.gravity = (0.0, 0.0, -9.81)
    s.set_contact(0, 3, NK=16)
    s.set_contact(0, 1, NK=16)
    s.set_contact(1, 2, NK=16)
    # pairs without <ContactParameters> never touch in this scene; keep them checked with defaults (NK=4)
    s.set_contact(0, 2, NK=4)
    s.set_contact(1, 3, NK=4)
    s.set_contact(2, 3, NK=4)
    return s


def contact_constrained_pendulum(n_envs=1, stabilization=25):
    """example/contact-constrained-pendulum/contact-constrained-pendulum.xml: body l1 (sphere inertia r = 1.5811, mass 1) at
    (1, 0, 0) turned 90 degrees about z, its point (0, 1, 0) held at the origin of the fixed body `world` by the six frictionless
    contacts of the scene's collision-detection plugin; gravity -y; constraint-stabilization-max-iterations = 25."""
    s = SceneBatch(n_envs, 2)
    s.name = "contact-constrained-pendulum"
    s.gravity = (0.0, -9.81, 0.0)
    s.set_sphere(0, 1.5811, mass=1.0)
    s.shape[0, :] = SHAPE_PIN
    s.dims[0, 0, :], s.dims[0, 1, :], s.dims[0, 2, :] = 0.0, 1.0, 0.0
    s.q[0, 0, :] = 1.0
    qz = quat_from_rpy(0.0, 0.0, np.float64(1.57079632679490))
    for k in range(4):
        s.q[0, 3 + k, :] = qz[k]
    s.shape[1, :] = SHAPE_PINWORLD
    s.enabled[1, :] = 0
    s.set_contact(0, 1, mu_coulomb=0.0, epsilon=0.0, NK=4)
    s.stabilization_max_iterations = stabilization
    return s


def rimless_wheel(n_envs=1, theta_dot=0.3, seed=None, alpha_gravity=(0.099833, 0.0, -0.995), stabilization=-1, inertia=(2.0, 1.0, 2.0)):
    """example/rimless-wheel/wheel.xml + init.cpp:150-175: bodies GROUND (plane z = 0), WHEEL (6 spokes, R = 1, W = 0, mass 1,
    inertia diag(2,1,2)), gravity tilted by alpha = 0.1 (downhill along +x), mu = 100 (no-slip model), epsilon = 0, the
    reference's default constraint stabilization.  The initializer puts the wheel on two spokes (z = sin 60 deg) rolling with
    angular velocity theta_dot about y and linear velocity theta_dot * R along x (RIMLESS_WHEEL_THETAD).
    seed != None: theta_dot is drawn per env from [0.5, 1.5] x theta_dot (randomised-initial-state batch)."""
    s = SceneBatch(n_envs, 2)
    s.name = "rimless-wheel"
    s.set_plane(0, quat=tuple(quat_from_rpy(np.float64(1.570796326949), 0.0, 0.0)))   # wheel.xml:56
    s.set_wheel(1, inertia=inertia)
    s.gravity = tuple(alpha_gravity)
    s.set_contact(0, 1, mu_coulomb=100.0, epsilon=0.0, NK=4)
    s.stabilization_max_iterations = stabilization
    td = np.full(n_envs, float(theta_dot))
    if seed is not None:
        td = td * np.random.default_rng(seed).uniform(0.5, 1.5, n_envs)
    s.q[1, 2, :] = 0.866025403784439            # init.cpp:163
    s.v[1, 0, :] = 2 * math.pi * 1.0 * (td / (math.pi * 2.0))   # DIST_PER_REV * REV_PER_SEC (init.cpp:154-156)
    s.v[1, 4, :] = td
    return s


def small_lcp_batch(n_envs, seed=0xB200, NK_box=8):
    """SURVEY.md 8(d) case 2: 50% sitting boxes with the test/TestDie.cpp:70-91 perturbation, 50% bouncing balls.

    Even envs are boxes, odd envs are balls.  Box: half-extents U[0.25,0.75], mu U[0,1], yaw U[0,2pi), tilt <= 5 deg,
    drop height U[0,0.05], v, omega U[-1,1]^3.  Ball: r=1, eps U[0.5,1], omega_y=10, height U[1.5,3].
    """
    rng = np.random.default_rng(seed)
    s = SceneBatch(n_envs, 2)
    s.name = "sitting-box/bouncing-ball batch"
    ne = n_envs
    isbox = (np.arange(ne) % 2) == 0
    nbx, nbl = int(isbox.sum()), int((~isbox).sum())
    he = rng.uniform(0.25, 0.75, (3, ne))
    s.set_box(0, 2 * he[0], 2 * he[1], 2 * he[2], density=1.0)
    # balls overwrite the odd envs
    shape_box, dims_box, mass_box, in_box = s.shape.copy(), s.dims.copy(), s.mass.copy(), s.inertia.copy()
    s.set_sphere(0, 1.0, density=1.0)
    s.shape[0, isbox], s.dims[0][:, isbox] = shape_box[0, isbox], dims_box[0][:, isbox]
    s.mass[0, isbox], s.inertia[0][:, isbox] = mass_box[0, isbox], in_box[0][:, isbox]
    s.set_plane(1)
    mu = rng.uniform(0.0, 1.0, ne)
    eps = rng.uniform(0.5, 1.0, ne)
    s.set_contact(0, 1, mu_coulomb=np.where(isbox, mu, 0.0), epsilon=np.where(isbox, 0.0, eps),
                  NK=np.where(isbox, NK_box, 4).astype(np.int32))
    yaw = rng.uniform(0, 2 * np.pi, ne)
    tilt = np.deg2rad(5.0) * rng.uniform(-1, 1, (2, ne))
    quat = quat_from_rpy(tilt[0], yaw, tilt[1])  # yaw about the vertical (y) axis
    drop = rng.uniform(0.0, 0.05, ne)
    # lowest box vertex sits `drop` above the plane
    R = _rotmat(quat)
    ext = np.abs(R[1, 0]) * he[0] + np.abs(R[1, 1]) * he[1] + np.abs(R[1, 2]) * he[2]
    s.q[0, 1, :] = np.where(isbox, ext + drop, rng.uniform(1.5, 3.0, ne))
    for k in range(4):
        s.q[0, 3 + k, :] = np.where(isbox, quat[k], 1.0 if k == 3 else 0.0)
    vel = rng.uniform(-1, 1, (6, ne))
    for k in range(6):
        s.v[0, k, :] = np.where(isbox, vel[k], 10.0 if k == 4 else 0.0)
    # simulator attributes of the two source scenes: test/box.xml sets min-step-size="1e-3" (the TestDie scene),
    # bouncing-ball.xml keeps the default sqrt(eps) (TimeSteppingSimulator.cpp:48,470-472)
    s.min_step_size_env = np.where(isbox, 1e-3, NEAR_ZERO)
    return s


def _rotmat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def quat_conj(a):
    return np.array([-a[0], -a[1], -a[2], a[3]])


def pendulum(n_envs=1, fdyn=FDYN_FSAB):
    """example/reduced-coords/pendulum.xml: fixed base + one link (mass 1, sphere inertia r=1.5811) on a revolute joint
    about z at the origin, link COM at distance 1 from the joint, q = pi/2, qd = 100 in the file (tests override)."""
    s = SceneBatch(n_envs, 2)
    s.name = "pendulum"
    for b in range(2):
        s.mass[b, :] = 1.0
        s.inertia[b, :, :] = 0.4 * 1.5811 ** 2
    rc = ArticulatedBody(s, 0, 2, fdyn)
    rc.set_joint(1, 0, JOINT_REVOLUTE, (0, 0, 1), (0, 0, 0), (-1.0, 0, 0))
    return s


def chain(n_envs=1, n_links=5, fdyn=FDYN_FSAB, seed=1, branch=False):
    """A randomised fixed-base chain (or tree when `branch`) of revolute and prismatic joints with skew axes: the
    forward-dynamics parity workload (no geometry).  Gravity (0,-9.81,0)."""
    rng = np.random.default_rng(seed)
    s = SceneBatch(n_envs, n_links)
    s.name = "chain"
    rc = ArticulatedBody(s, 0, n_links, fdyn)
    for i in range(n_links):
        s.mass[i, :] = rng.uniform(0.5, 2.0, n_envs)
        s.inertia[i, :, :] = rng.uniform(0.01, 0.2, (3, n_envs))
    for i in range(1, n_links):
        parent = int(rng.integers(0, i)) if branch else i - 1
        jt = JOINT_PRISMATIC if (i % 4 == 3) else JOINT_REVOLUTE
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        qr = rng.normal(size=4)
        qr /= np.linalg.norm(qr)
        rc.set_joint(i, parent, jt, ax, rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3), qr)
    rc.jq[:] = rng.uniform(-1.0, 1.0, rc.jq.shape)
    rc.jqd[:] = rng.uniform(-2.0, 2.0, rc.jqd.shape)
    return s


# example/ur10/model.sdf:5-524 -- link pose (xyz rpy, model frame), COM offset in the link frame, mass, (ixx, iyy, izz)
_UR10_LINKS = [
    ("base_link",      (0, 0, 0, 0, 0, 0),                                   (0, 0, 0),        4.0,   (0.00610633, 0.00610633, 0.01125)),
    ("shoulder_link",  (0, 0, 0.1273, 0, 0, 0),                              (0, 0, 0),        7.778, (0.0314743, 0.0314743, 0.0218756)),
    ("upper_arm_link", (0, 0.220941, 0.1273, 3.14159, 1.57079, 3.14159),     (0, 0, 0.306),    12.93, (0.421754, 0.421754, 0.0363656)),
    ("forearm_link",   (0.612, 0.049041, 0.1273, 3.14159, 1.57079, 3.14159), (0, 0, 0.28615),  3.87,  (0.11107, 0.11107, 0.0108844)),
    ("wrist_1_link",   (1.1843, 0.049041, 0.1273, 3.14159, 3.58979e-09, 3.14159), (0, 0, 0),   1.96,  (0.00510825, 0.00510825, 0.0055125)),
    ("wrist_2_link",   (1.1843, 0.163941, 0.1273, 3.14159, 3.58979e-09, 3.14159), (0, 0, 0),   1.96,  (0.00510825, 0.00510825, 0.0055125)),
    ("wrist_3_link",   (1.1843, 0.163941, 0.0116, 3.14159, 3.58979e-09, 3.14159), (0, 0, 0),   0.202, (0.000526462, 0.000526462, 0.000568125)),
    ("hand",           (1.1843, 0.256, 0.0116, 0, 0, 0),                     (0, 0.035, 0),    0.96,  (0.00053312, 0.00065312, 0.000904)),
    ("l_finger",       (1.1843, 0.256, 0.0116, 0, 0, 0),                     (-0.0205, 0.0798, 0), 0.12, (0.0000095236, 0.0000072, 0.0000052036)),
    ("r_finger",       (1.1843, 0.256, 0.0116, 0, 0, 0),                     (0.0205, 0.0798, 0),  0.12, (0.0000095236, 0.0000072, 0.0000052036)),
]
# joint of link i: (parent, type, axis in the child link frame)  (model.sdf:78-90,123-135,...,361-371,508-524)
_UR10_JOINTS = [None, (0, JOINT_REVOLUTE, (0, 0, 1)), (1, JOINT_REVOLUTE, (0, 1, 0)), (2, JOINT_REVOLUTE, (0, 1, 0)),
                (3, JOINT_REVOLUTE, (0, 1, 0)), (4, JOINT_REVOLUTE, (0, 0, -1)), (5, JOINT_REVOLUTE, (0, 1, 0)),
                (6, JOINT_REVOLUTE, (0, 0, 1)), (7, JOINT_PRISMATIC, (1, 0, 0)), (7, JOINT_PRISMATIC, (1, 0, 0))]


def ur10(n_envs=1, fdyn=FDYN_CRB, table_z=None, with_block=True, seed=0xB200, q_jitter=0.1, controller=True, mu=0.5, NK=4, table_gap=2e-3):
    """SURVEY.md 8(d) case 4: the UR10 + Schunk gripper of example/ur10/model.sdf as a fixed-base chain.

    Benchmark variant (the shipped ur10.xml has mesh geometry, mu = 100, joint limits and no table): `world_joint` is a
    weld (base_link is the fixed base); `fixed_hand_to_wrist` and the two finger prismatics -- held by +-1e-5 joint
    limits in the file, which are outside this round's scope -- are ordinary joints held at 0 by PD gains; the wrist-3
    link, hand and fingers carry sphere / box proxies; a plane "table" (normal +z) sits at `table_z`; optionally the
    block of ur10.xml:22-25 (box .02825 x .025 x .025, mass 1) rests on the table.  Joint PD targets follow
    example/ur10/controller.cpp:46-96.  dt = 5e-4 (ur10.xml:2), gravity (0,0,-9.81).  Bodies: 0..9 links, 10 table,
    11 block."""
    rng = np.random.default_rng(seed)
    nl = len(_UR10_LINKS)
    nb = nl + 1 + (1 if with_block else 0)
    s = SceneBatch(n_envs, nb)
    s.name = "ur10"
    s.gravity = (0.0, 0.0, -9.81)
    s.min_step_size = 5e-4
    rc = ArticulatedBody(s, 0, nl, fdyn)
    pose_q, pose_x, com_w = [], [], []
    for i, (_, pose, com, mass, (ixx, iyy, izz)) in enumerate(_UR10_LINKS):
        ql = quat_from_rpy(np.float64(pose[3]), np.float64(pose[4]), np.float64(pose[5]))
        R = _rotmat(ql)
        pose_q.append(ql); pose_x.append(np.array(pose[:3], np.float64)); com_w.append(pose_x[i] + R @ np.array(com, np.float64))
        s.mass[i, :] = mass
        s.inertia[i, 0, :], s.inertia[i, 1, :], s.inertia[i, 2, :] = ixx, iyy, izz
    # base pose (COM frame of base_link)
    for k in range(3):
        s.q[0, k, :] = com_w[0][k]
    for i in range(1, nl):
        parent, jt, ax = _UR10_JOINTS[i]
        Rp, Rc = _rotmat(pose_q[parent]), _rotmat(pose_q[i])
        joint_w = pose_x[i]                                  # SDF joints sit at the child link's frame origin
        rc.set_joint(i, parent, jt, ax, Rp.T @ (joint_w - com_w[parent]), Rc.T @ (joint_w - com_w[i]),
                     quat_mul(quat_conj(pose_q[parent]), pose_q[i]))
    if controller:
        PERIOD, AMP = 5.0, 0.5
        SMALL = AMP * 0.1
        amp = np.array([AMP * PERIOD, SMALL * PERIOD * 2.0, AMP * PERIOD * 2.0 / 3.0, AMP * PERIOD / 7.0, AMP * PERIOD * 2.0 / 11.0,
                        AMP * PERIOD * 3.0 / 13.0, 0.0, 0.0, 0.0])
        freq = np.array([1.0, 2.0, 2.0 / 3.0, 1.0 / 7.0, 2.0 / 11.0, 3.0 / 13.0, 0.0, 0.0, 0.0])
        # controller.cpp:75-78 gains for the six arm joints, except wrist 3: with the hand no longer locked to it by joint
        # limits its axis carries ~1e-3 kg m^2 and kv = 6 is unstable under the explicit velocity update at dt = 5e-4
        # (kv dt / I > 2); wrist 3, the hand joint and the fingers get gains inside the stability bound.
        kp = np.array([300.0, 300.0, 60.0, 15.0, 15.0, 15.0, 5.0, 1000.0, 1000.0])
        kv = np.array([120.0, 120.0, 24.0, 6.0, 6.0, 1.0, 0.3, 20.0, 20.0])
        rc.set_controller(kp, kv, amp, freq)
        rc.jqd[:6, :] = amp[:6, None]                        # controller.cpp:124-140: qd(0) = cos(0) * amplitude
    rc.jq[:6, :] = rng.uniform(-q_jitter, q_jitter, (6, n_envs))
    # collision proxies (centred on the link COM frames)
    s.set_sphere(6, 0.045, mass=_UR10_LINKS[6][3]); s.inertia[6, :, :] = np.array(_UR10_LINKS[6][4])[:, None]
    s.set_box(7, 0.10, 0.07, 0.05, mass=_UR10_LINKS[7][3]); s.inertia[7, :, :] = np.array(_UR10_LINKS[7][4])[:, None]
    for f in (8, 9):
        s.set_sphere(f, 0.012, mass=_UR10_LINKS[f][3]); s.inertia[f, :, :] = np.array(_UR10_LINKS[f][4])[:, None]
    table = nl
    s.set_plane(table, quat=tuple(quat_from_rpy(np.float64(1.5707963267948966), 0.0, 0.0)), pos=(0.0, 0.0, 0.0))
    if table_z is None:
        # per env: `table_gap` below the lowest point of the proxies at the initial joint state, so nothing starts in
        # penetration and the first contacts come from gravity sag and the commanded motion
        tz = np.zeros(n_envs)
        for e in range(n_envs):
            x, R = rc.link_poses(e)
            low = min(x[k][2] - s.dims[k, 0, e] for k in (6, 8, 9))
            low = min(low, x[7][2] - 0.5 * (abs(R[7][2, 0]) * s.dims[7, 0, e] + abs(R[7][2, 1]) * s.dims[7, 1, e] + abs(R[7][2, 2]) * s.dims[7, 2, e]))
            tz[e] = low - table_gap
        table_z = tz
    s.q[table, 2, :] = table_z
    for link in (6, 7, 8, 9):
        s.set_contact(link, table, mu_coulomb=mu, NK=NK)
    if with_block:
        blk = nl + 1
        s.set_box(blk, 0.02825, 0.025, 0.025, mass=1.0)
        s.q[blk, 0, :] = 1.185 + rng.uniform(-0.05, 0.05, n_envs)
        s.q[blk, 1, :] = 0.336 + rng.uniform(-0.05, 0.05, n_envs)
        s.q[blk, 2, :] = table_z + 0.0125
        s.set_contact(blk, table, mu_coulomb=mu, NK=NK)
        for f in (8, 9):
            s.set_contact(f, blk, mu_coulomb=mu, NK=NK)
    return s


def parts_feeder(n_envs=1, seed=0xB200, mu=0.01, NK=4, tilt=0.05, shake_hz=30.0, shake_amp=1e-3, fdyn=FDYN_CRB):
    """SURVEY.md 8(d) case 5, parts-feeder-like (example/parts-feeder/feeder.xml): a fixed-base RCArticulatedBody whose
    single prismatic joint `shaker` (axis x, feeder.xml:45) carries a flat tray, tilted by `tilt` rad about y
    (rotate="0 0.05 0", feeder.xml:38), driven by a joint-space PD law towards a sinusoid (feeder.cpp:12-21: Kp = 1e3,
    Kv = 10, 1 mm amplitude), with one free box part (0.05 x 0.025 x 0.01, mass 0.1, feeder.xml:72-77) lying on it,
    mu = 0.01 (feeder.xml:22), gravity (0,0,-9.81), dt = 1e-3.  Benchmark variant: the tray is one box link (the file
    builds it from three planes attached to moving links and a fixed bar), the shake frequency is 30 Hz instead of the
    file's 500 Hz (which aliases at dt = 1 ms), the part's position and yaw on the tray are randomised per env.
    Bodies: 0 base (no geometry), 1 tray, 2 part."""
    rng = np.random.default_rng(seed)
    s = SceneBatch(n_envs, 3)
    s.name = "parts-feeder"
    s.gravity = (0.0, 0.0, -9.81)
    s.min_step_size = 1e-3
    qb = quat_from_rpy(np.float64(0.0), np.float64(tilt), np.float64(0.0))
    for k in range(4):
        s.q[0, 3 + k, :] = qb[k]
    s.mass[0, :] = 1.0
    s.set_box(1, 1.0, 0.11, 0.02, mass=1.0)
    s.set_box(2, 0.05, 0.025, 0.01, mass=0.1)
    rc = ArticulatedBody(s, 0, 2, fdyn)
    rc.set_joint(1, 0, JOINT_PRISMATIC, (1, 0, 0), (0, 0, 0), (0, 0, 0))
    w = 2.0 * math.pi * shake_hz
    rc.set_controller(np.array([1e3]), np.array([1e1]), np.array([shake_amp]), np.array([w]))
    # the part rests on the tray's top face (tray frame), placed and yawed at random
    R = _rotmat(np.array(qb, np.float64))
    px = rng.uniform(-0.3, 0.3, n_envs)
    py = rng.uniform(-0.03, 0.03, n_envs)
    yaw = rng.uniform(-np.pi, np.pi, n_envs)
    for e in range(n_envs):
        local = np.array([px[e], py[e], 0.01 + 0.005])
        s.q[2, :3, e] = R @ local
        qy = quat_from_rpy(np.float64(0.0), np.float64(0.0), np.float64(yaw[e]))
        s.q[2, 3:, e] = quat_mul(np.array(qb, np.float64), np.array(qy, np.float64))
    s.set_contact(1, 2, mu_coulomb=mu, NK=NK)
    s.max_contacts = 8
    s.max_lcp_n = 8 * (6 + NK // 2)
    return s


def box_stack(n_envs=1, n_boxes=3, jitter=1e-3, seed=0xB200, adjacent_only=True, mu=1e-4, NK=4, yaw_jitter=0.0):
    """example/stacks/stack.xml (SURVEY.md 8(d) case 3): boxes of height 1 shrinking by 0.05 per level (x and z), density 10,
    centres at y = 0.5, 1.5, ..., mu = 1e-4 between neighbours and on the ground, default 4 cone edges.  The file
    registers 3 of its 7 boxes (stack.xml:85-88); BASELINE's config extends the pattern to 10.  Bodies 0..n-1 are the
    boxes bottom-up, body n is the ground plane.  `jitter`: lateral offset U[-j, j] per box and env.
    `adjacent_only`: only neighbouring boxes (and box 0 / ground) are collision pairs -- what the reference's broad phase
    (CCD.cpp:702-874, swept bounding spheres) leaves for this scene; False keeps every pair with default parameters."""
    rng = np.random.default_rng(seed)
    nb = n_boxes + 1
    s = SceneBatch(n_envs, nb)
    s.name = f"stack-{n_boxes}"
    for k in range(n_boxes):
        w = 1.0 - 0.05 * k
        s.set_box(k, w, 1.0, w, density=10.0)
        s.q[k, 1, :] = 0.5 + k
        s.q[k, 0, :] = rng.uniform(-jitter, jitter, n_envs)
        s.q[k, 2, :] = rng.uniform(-jitter, jitter, n_envs)
        if yaw_jitter:
            yaw = rng.uniform(-yaw_jitter, yaw_jitter, n_envs)
            quat = quat_from_rpy(np.zeros(n_envs), yaw, np.zeros(n_envs))
            for c in range(4):
                s.q[k, 3 + c, :] = quat[c]
    s.set_plane(n_boxes)
    if not adjacent_only:
        for i in range(nb):
            for j in range(i + 1, nb):
                s.set_contact(i, j)
    s.set_contact(0, n_boxes, mu_coulomb=mu, NK=NK)
    for k in range(n_boxes - 1):
        s.set_contact(k, k + 1, mu_coulomb=mu, NK=NK)
    s.max_contacts = (8 if yaw_jitter else 4) * n_boxes + 8
    s.max_lcp_n = s.max_contacts * (6 + NK // 2)
    return s
