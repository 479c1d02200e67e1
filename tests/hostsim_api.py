"""ctypes wrapper of tests/hostsim/libhostsim.so: the kernels' device code compiled for the host with a
single-thread group.  Test infrastructure only (lets the CPU suite check kernel logic against the oracle)."""
import ctypes as C
import os
import subprocess

import numpy as np

from moby_b200.capi import RcDesc, SceneDesc

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
CNT = ("env_steps", "mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "pivots", "lcp_failures",
       "impact_tol_events", "contacts", "max_lcp_n", "overflow", "pivot_flops", "assembly_flops", "ca_iterations",
       "stab_iterations", "stab_lcp_solves", "stab_line_search_failures")
_lib = None


def build():
    src = os.path.join(HERE, "hostsim.cpp")
    out = os.path.join(HERE, "libhostsim.so")
    csrc = os.path.join(os.path.dirname(HERE), "..", "moby_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src])
    return out


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hostsim_run.argtypes = [C.POINTER(SceneDesc)] + [C.c_void_p] * 6 + [C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.hostsim_run_phased.argtypes = [C.POINTER(SceneDesc)] + [C.c_void_p] * 6 + [C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 2
        L.hostsim_lcp.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hostsim_rc.argtypes = [C.POINTER(RcDesc)] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 8
        L.hostsim_set_env_stat.argtypes = [C.c_void_p]
        L.hostsim_boxbox_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class HostSim:
    """Steps a SceneBatch with the device code on the host; state arrays use the product's SoA layout."""

    def __init__(self, scene, taps=False):
        self.scene = scene
        self._L = lib()
        self._d = scene.cdesc()
        self.nmax = self._L.hostsim_run(C.byref(self._d), None, None, None, None, None, None, 0.0, 0, 0, 0, None, None, None, None, None, None)
        ne = scene.n_envs
        self.q, self.v = scene.q.copy(), scene.v.copy()
        self.q[:, 3:7, :] /= np.sqrt((self.q[:, 3:7, :] ** 2).sum(axis=1, keepdims=True))   # as b200moby_set_state does
        self.time = np.zeros(ne)
        self.zlast, self.zlast_n = np.zeros((2 * self.nmax, ne)), np.zeros(2 * ne, np.int32)
        self.counters = np.zeros(24, np.uint64)
        self.rc = getattr(scene, "rc", None)
        self.jq = self.rc.jq.copy() if self.rc is not None else None
        self.jqd = self.rc.jqd.copy() if self.rc is not None else None
        self.stat = np.zeros((5, ne), np.int32)     # per-env lcp_failures, lemke_calls, lcp_fast_calls, lcp_solves, pivots (SimParams::env_stat)
        self.taps = taps
        if taps:
            self.tapMM, self.tapqq = np.zeros((ne, self.nmax * self.nmax)), np.zeros((ne, self.nmax))
            self.tapz, self.tapn = np.zeros((ne, self.nmax)), np.zeros(ne, np.int32)

    def step(self, dt, n=1, e0=0, e1=None):
        e1 = self.scene.n_envs if e1 is None else e1
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        t = [p(self.tapMM), p(self.tapqq), p(self.tapz), p(self.tapn)] if self.taps else [None] * 4
        self._L.hostsim_set_env_stat(p(self.stat))
        self._L.hostsim_run(C.byref(self._d), p(self.q), p(self.v), p(self.time), p(self.zlast), p(self.zlast_n), p(self.counters),
                          dt, n, e0, e1, *t, self._jp(self.jq), self._jp(self.jqd))

    def step_phased(self, dt, n=1, rounds=2, pivot_budget=0):
        """The same steps through the phased schedule (advance / impact classes / stragglers / finish)."""
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        self._L.hostsim_set_env_stat(p(self.stat))
        self._L.hostsim_run_phased(C.byref(self._d), p(self.q), p(self.v), p(self.time), p(self.zlast), p(self.zlast_n), p(self.counters),
                                 dt, n, rounds, pivot_budget, self._jp(self.jq), self._jp(self.jqd))

    @staticmethod
    def _jp(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def env_stats(self):
        return dict(lcp_failures=self.stat[0], lemke_calls=self.stat[1], lcp_fast_calls=self.stat[2], lcp_solves=self.stat[3], pivots=self.stat[4])

    def counters_dict(self):
        return {k: int(self.counters[i]) for i, k in enumerate(CNT)}

    def last_lcp(self, e):
        n = int(self.tapn[e])
        return n, self.tapMM[e, :n * n].reshape(n, n).T.copy(), self.tapqq[e, :n].copy(), self.tapz[e, :n].copy()


def lcp(mode, M, q, z0=None, piv_tol=-1.0, zero_tol=-1.0, min_exp=-20, step_exp=1, max_exp=1, log_cap=4096, want_log=True):
    """mode 0 lemke, 1 fast, 2 lemke_regularized, 3 fast_regularized.  Returns (status, z, pivots, log)."""
    n = len(q)
    Mf = np.asfortranarray(np.asarray(M, np.float64))
    q = np.ascontiguousarray(q, np.float64)
    z = np.zeros(n) if z0 is None else np.array(z0, np.float64)
    piv, ll = C.c_int(), C.c_int()
    log = np.zeros(log_cap, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    st = lib().hostsim_lcp(mode, n, p(Mf), p(q), p(z), 0 if z0 is None else 1, piv_tol, zero_tol, min_exp, step_exp, max_exp,
                                  C.byref(piv), p(log) if want_log else None, log_cap, C.byref(ll))
    return st, z, piv.value, log[:min(ll.value, log_cap)].copy()


def boxbox_dist(cA, RA, extA, cB, RB, extB):
    """Signed distance and closest points of two posed boxes through the kernels' device code: (dist, pA, pB)."""
    A = np.concatenate([np.asarray(cA, np.float64), np.asarray(RA, np.float64).ravel(), np.asarray(extA, np.float64)])
    B = np.concatenate([np.asarray(cB, np.float64), np.asarray(RB, np.float64).ravel(), np.asarray(extB, np.float64)])
    out = np.zeros(7)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hostsim_boxbox_dist(p(A), p(B), p(out))
    return out[0], out[1:4].copy(), out[4:7].copy()


def rc_eval(body, what, q, qd, tau=None, gravity=(0.0, -9.81, 0.0), env=0):
    """Device articulated-body code on the host for one env of an ArticulatedBody description.
    what: 0 ABA qdd, 1 CRB qdd, 2 joint-space inertia.  Returns (out, links dict)."""
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    d = body.cdesc()
    mass, J, pose = body.env_mass_props(env)
    nl, nd = body.n_links, body.n_links - 1
    q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
    tau = np.zeros(nd) if tau is None else np.ascontiguousarray(tau, np.float64)
    g = np.array(gravity, np.float64)
    out = np.zeros(nd * nd if what == 2 else nd)
    lx, lq, lvl, lva = np.zeros((nl, 3)), np.zeros((nl, 4)), np.zeros((nl, 3)), np.zeros((nl, 3))
    rc = lib().hostsim_rc(C.byref(d), p(mass), p(J), p(pose), p(g), what, p(q), p(qd), p(tau), p(out), p(lx), p(lq), p(lvl), p(lva))
    assert rc == 0
    if what == 2:
        out = out.reshape(nd, nd).T.copy()
    return out, dict(x=lx, quat=lq, vl=lvl, va=lva)
