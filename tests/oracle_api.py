"""ctypes wrapper of oracle/liboracle.so -- the CPU checker.  Test infrastructure only: the product
package never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

from moby_b200.capi import Counters, RcDesc, SceneDesc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
TIE_LOWEST, TIE_RAND = 0, 1

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_sim_create.restype = C.c_void_p
        L.oracle_sim_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.c_int]
        L.oracle_sim_destroy.argtypes = [C.c_void_p]
        for f in (L.oracle_sim_set_state, L.oracle_sim_get_state):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_sim_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.oracle_sim_time.restype = C.c_double
        L.oracle_sim_time.argtypes = [C.c_void_p]
        L.oracle_sim_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.oracle_sim_last_lcp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_sim_last_contacts.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.oracle_sim_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_batch_step.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int,
                                        C.c_int, C.c_int, C.POINTER(Counters)]
        L.oracle_lcp_lemke.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_lcp_fast.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_lcp_fast_regularized.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                  C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_lcp_lemke_regularized.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_batch_create.restype = C.c_void_p
        L.oracle_batch_create.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_batch_destroy.argtypes = [C.c_void_p]
        L.oracle_batch_run.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.POINTER(Counters)]
        L.oracle_batch_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_batch_get_state_soa.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_batch_env_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_batch_set_joint_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_batch_get_joint_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        for f in (L.oracle_sim_set_joint_state, L.oracle_sim_get_joint_state):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_sim_set_joint_forces.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_boxbox_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_rc_fwd_dyn.argtypes = [C.POINTER(RcDesc)] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4
        L.oracle_rc_inertia.argtypes = [C.POINTER(RcDesc)] + [C.c_void_p] * 5
        L.oracle_rc_links.argtypes = [C.POINTER(RcDesc)] + [C.c_void_p] * 10
        L.oracle_rc_energy.restype = C.c_double
        L.oracle_rc_energy.argtypes = [C.POINTER(RcDesc)] + [C.c_void_p] * 6
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _colmajor(M):
    return np.asfortranarray(np.asarray(M, np.float64))


def lcp_lemke(M, q, piv_tol=-1.0, zero_tol=-1.0, tie=TIE_LOWEST, log_cap=4096):
    n = len(q)
    Mf, q = _colmajor(M), np.ascontiguousarray(q, np.float64)
    z = np.zeros(n)
    piv, st, ll = C.c_int(), C.c_int(), C.c_int()
    log = np.zeros(log_cap, np.int32)
    ok = lib().oracle_lcp_lemke(n, _p(Mf), _p(q), _p(z), piv_tol, zero_tol, tie, C.byref(piv), C.byref(st), _p(log), log_cap,
                                C.byref(ll))
    return bool(ok), z, dict(pivots=piv.value, status=st.value, log=log[:min(ll.value, log_cap)].copy())


def lcp_fast(M, q, z0=None, zero_tol=-1.0, tie=TIE_LOWEST, log_cap=4096):
    n = len(q)
    Mf, q = _colmajor(M), np.ascontiguousarray(q, np.float64)
    z = np.zeros(n) if z0 is None else np.array(z0, np.float64)
    piv, st, ll = C.c_int(), C.c_int(), C.c_int()
    log = np.zeros(log_cap, np.int32)
    ok = lib().oracle_lcp_fast(n, _p(Mf), _p(q), _p(z), 0 if z0 is None else 1, zero_tol, tie, C.byref(piv), C.byref(st),
                               _p(log), log_cap, C.byref(ll))
    return bool(ok), z, dict(pivots=piv.value, status=st.value, log=log[:min(ll.value, log_cap)].copy())


def lcp_fast_regularized(M, q, z0=None, min_exp=-20, step_exp=4, max_exp=20, zero_tol=-1.0, tie=TIE_LOWEST):
    n = len(q)
    Mf, q = _colmajor(M), np.ascontiguousarray(q, np.float64)
    z = np.zeros(n) if z0 is None else np.array(z0, np.float64)
    piv, st = C.c_int(), C.c_int()
    ok = lib().oracle_lcp_fast_regularized(n, _p(Mf), _p(q), _p(z), 0 if z0 is None else 1, min_exp, step_exp, max_exp,
                                           zero_tol, tie, C.byref(piv), C.byref(st))
    return bool(ok), z, dict(pivots=piv.value, status=st.value)


def lcp_lemke_regularized(M, q, min_exp=-20, step_exp=1, max_exp=1, piv_tol=-1.0, zero_tol=-1.0, tie=TIE_LOWEST):
    n = len(q)
    Mf, q = _colmajor(M), np.ascontiguousarray(q, np.float64)
    z = np.zeros(n)
    piv, st = C.c_int(), C.c_int()
    ok = lib().oracle_lcp_lemke_regularized(n, _p(Mf), _p(q), _p(z), min_exp, step_exp, max_exp, piv_tol, zero_tol, tie,
                                            C.byref(piv), C.byref(st))
    return bool(ok), z, dict(pivots=piv.value, status=st.value)


class OracleSim:
    """One Moby TimeSteppingSimulator instance (env `env` of a SceneBatch) on the CPU oracle."""

    def __init__(self, scene, env=0, tie=TIE_LOWEST):
        self.scene, self.env = scene, env
        self._d = scene.cdesc()
        self.h = lib().oracle_sim_create(C.byref(self._d), env, tie)
        self.nb = scene.n_bodies
        self.set_state(scene.q[:, :, env], scene.v[:, :, env])
        self.rc = getattr(scene, "rc", None)
        if self.rc is not None:
            self.set_joint_state(self.rc.jq[:, env], self.rc.jqd[:, env])

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_sim_destroy(self.h)
            self.h = None

    def set_state(self, q, v):
        q, v = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(v, np.float64)
        lib().oracle_sim_set_state(self.h, _p(q), _p(v))

    def get_state(self):
        q, v = np.zeros((self.nb, 7)), np.zeros((self.nb, 6))
        lib().oracle_sim_get_state(self.h, _p(q), _p(v))
        return q, v

    def set_joint_state(self, jq, jqd):
        jq, jqd = np.ascontiguousarray(jq, np.float64), np.ascontiguousarray(jqd, np.float64)
        lib().oracle_sim_set_joint_state(self.h, _p(jq), _p(jqd))

    def get_joint_state(self):
        nd = self.rc.n_dof
        jq, jqd = np.zeros(nd), np.zeros(nd)
        lib().oracle_sim_get_joint_state(self.h, _p(jq), _p(jqd))
        return jq, jqd

    def set_joint_forces(self, tau):
        tau = None if tau is None else np.ascontiguousarray(tau, np.float64)
        lib().oracle_sim_set_joint_forces(self.h, None if tau is None else _p(tau))

    def step(self, dt, n=1):
        lib().oracle_sim_step(self.h, dt, n)

    @property
    def time(self):
        return lib().oracle_sim_time(self.h)

    def counters(self):
        c = Counters()
        lib().oracle_sim_counters(self.h, C.byref(c))
        return c.as_dict()

    def last_lcp(self, ncap=512):
        MM, qq, z = np.zeros(ncap * ncap), np.zeros(ncap), np.zeros(ncap)
        n = lib().oracle_sim_last_lcp(self.h, _p(MM), _p(qq), _p(z), ncap)
        return n, MM[:n * n].reshape(n, n).T.copy(), qq[:n].copy(), z[:n].copy()

    def last_contacts(self, cap=64):
        pt, nr, t1, t2 = (np.zeros((cap, 3)) for _ in range(4))
        pair, dist = np.zeros(cap, np.int32), np.zeros(cap)
        m = lib().oracle_sim_last_contacts(self.h, cap, _p(pt), _p(nr), _p(t1), _p(t2), _p(pair), _p(dist))
        return dict(count=m, point=pt[:m], normal=nr[:m], tan1=t1[:m], tan2=t2[:m], pair=pair[:m], dist=dist[:m])

    def assemble(self, ncap=512):
        MM, qq = np.zeros(ncap * ncap), np.zeros(ncap)
        nc = C.c_int()
        n = lib().oracle_sim_assemble(self.h, _p(MM), _p(qq), ncap, C.byref(nc))
        return n, MM[:n * n].reshape(n, n).T.copy(), qq[:n].copy(), nc.value


def batch_step(scene, q, v, dt, n_steps, e0=0, e1=None, tie=TIE_LOWEST, threads=1):
    """Steps envs [e0,e1) in place on SoA arrays q [nb][7][ne], v [nb][6][ne]; returns summed counters."""
    d = scene.cdesc()
    e1 = scene.n_envs if e1 is None else e1
    assert q.flags.c_contiguous and v.flags.c_contiguous
    c = Counters()
    lib().oracle_batch_step(C.byref(d), _p(q), _p(v), e0, e1, dt, n_steps, tie, threads, C.byref(c))
    return c.as_dict()


class OracleBatch:
    """Persistent CPU batch (envs [e0,e1) of a scene), optionally multi-threaded: bench.py's CPU baseline."""

    def __init__(self, scene, e0=0, e1=None, tie=TIE_LOWEST):
        self._d = scene.cdesc()
        self.e0, self.e1 = e0, scene.n_envs if e1 is None else e1
        self.nb = scene.n_bodies
        q, v = np.ascontiguousarray(scene.q), np.ascontiguousarray(scene.v)
        self.h = lib().oracle_batch_create(C.byref(self._d), _p(q), _p(v), self.e0, self.e1, tie)
        self.rc = getattr(scene, "rc", None)
        if self.rc is not None:
            self.set_joint_state(self.rc.jq, self.rc.jqd)

    def set_joint_state(self, jq, jqd):
        jq, jqd = np.ascontiguousarray(jq, np.float64), np.ascontiguousarray(jqd, np.float64)
        lib().oracle_batch_set_joint_state(self.h, _p(jq), _p(jqd), jq.shape[1])

    def get_joint_state(self, i):
        nd = self.rc.n_dof
        jq, jqd = np.zeros(nd), np.zeros(nd)
        lib().oracle_batch_get_joint_state(self.h, i, _p(jq), _p(jqd))
        return jq, jqd

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_batch_destroy(self.h)
            self.h = None

    def run(self, dt, n_steps, threads=1):
        c = Counters()
        lib().oracle_batch_run(self.h, dt, n_steps, threads, C.byref(c))
        return c.as_dict()

    def get_state(self, i):
        q, v = np.zeros((self.nb, 7)), np.zeros((self.nb, 6))
        lib().oracle_batch_get_state(self.h, i, _p(q), _p(v))
        return q, v

    def get_state_soa(self):
        """State of every env of the batch in the product's layout: q [body][7][n], v [body][6][n]."""
        n = self.e1 - self.e0
        q, v = np.zeros((self.nb, 7, n)), np.zeros((self.nb, 6, n))
        lib().oracle_batch_get_state_soa(self.h, _p(q), _p(v))
        return q, v

    def env_stats(self):
        """Per-env (lcp_failures, lemke_calls, lcp_fast_calls, lcp_solves) since creation, as a dict of [n] arrays."""
        n = self.e1 - self.e0
        st = np.zeros((5, n), np.int32)
        lib().oracle_batch_env_stats(self.h, _p(st))
        return dict(lcp_failures=st[0], lemke_calls=st[1], lcp_fast_calls=st[2], lcp_solves=st[3], pivots=st[4])


def boxbox_dist(cA, RA, extA, cB, RB, extB):
    """Signed distance and closest points of two posed boxes (oracle/oracle_boxbox.h): returns (dist, pA, pB)."""
    A = np.concatenate([np.asarray(cA, np.float64), np.asarray(RA, np.float64).ravel(), np.asarray(extA, np.float64)])
    B = np.concatenate([np.asarray(cB, np.float64), np.asarray(RB, np.float64).ravel(), np.asarray(extB, np.float64)])
    out = np.zeros(7)
    lib().oracle_boxbox_dist(_p(A), _p(B), _p(out))
    return out[0], out[1:4].copy(), out[4:7].copy()


# ---- reduced-coordinate articulated body (oracle/oracle_rc.h) ----
def rc_fwd_dyn(body, algo, q, qd, tau=None, gravity=(0.0, -9.81, 0.0), env=0):
    """algo 0: Featherstone ABA, 1: CRB + Cholesky.  `body` is a scenes.ArticulatedBody."""
    d = body.cdesc()
    mass, J, pose = body.env_mass_props(env)
    nd = body.n_links - 1
    q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
    tau = np.zeros(nd) if tau is None else np.ascontiguousarray(tau, np.float64)
    g = np.array(gravity, np.float64)
    qdd = np.zeros(nd)
    ok = lib().oracle_rc_fwd_dyn(C.byref(d), _p(mass), _p(J), _p(pose), _p(g), algo, _p(q), _p(qd), _p(tau), _p(qdd))
    assert ok == 1
    return qdd


def rc_inertia(body, q, env=0):
    d = body.cdesc()
    mass, J, pose = body.env_mass_props(env)
    nd = body.n_links - 1
    q = np.ascontiguousarray(q, np.float64)
    H = np.zeros(nd * nd)
    lib().oracle_rc_inertia(C.byref(d), _p(mass), _p(J), _p(pose), _p(q), _p(H))
    return H.reshape(nd, nd).T.copy()


def rc_links(body, q, qd, env=0):
    d = body.cdesc()
    mass, J, pose = body.env_mass_props(env)
    nl, nd = body.n_links, body.n_links - 1
    q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
    x, R, vl, va, jac = np.zeros((nl, 3)), np.zeros((nl, 3, 3)), np.zeros((nl, 3)), np.zeros((nl, 3)), np.zeros((nl, 6, nd))
    lib().oracle_rc_links(C.byref(d), _p(mass), _p(J), _p(pose), _p(q), _p(qd), _p(x), _p(R), _p(vl), _p(va), _p(jac))
    return dict(x=x, R=R, vl=vl, va=va, jac=jac)


def rc_energy(body, q, qd, gravity=(0.0, -9.81, 0.0), env=0):
    d = body.cdesc()
    mass, J, pose = body.env_mass_props(env)
    q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
    g = np.array(gravity, np.float64)
    return lib().oracle_rc_energy(C.byref(d), _p(mass), _p(J), _p(pose), _p(g), _p(q), _p(qd))
