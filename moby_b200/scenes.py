"""Batch scene descriptors: the host-side stand-in for the object graph Moby's XMLReader builds
(src/XMLReader.cpp:60-132) for the benchmark scenes of BASELINE.json / SURVEY.md 8(d).

Arrays follow include/b200moby.h: per-env parameters are SoA across envs ([body][env], [body][3][env],
[body_i*nb+body_j][env]); state q is [body][7][env] (x y z qx qy qz qw), v is [body][6][env].
"""
import ctypes as C
import math

import numpy as np

from .capi import SceneDesc

SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_PLANE = 0, 1, 2, 3
MODEL_QP, MODEL_AP = 0, 1
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
        self.q = np.zeros((nb, 7, ne), np.float64)
        self.q[:, 6, :] = 1.0
        self.v = np.zeros((nb, 6, ne), np.float64)
        self.name = "custom"

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
        d.impact_model, d.stabilization_max_iterations = self.impact_model, 0
        if self.min_step_size_env is not None:
            a = np.ascontiguousarray(self.min_step_size_env, np.float64)
            keep.append(a)
            d.min_step_size_env = a.ctypes.data_as(C.POINTER(C.c_double))
        d._keep = keep
        return d


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
