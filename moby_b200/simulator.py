"""Host-side mirror of Moby's TimeSteppingSimulator for a batch of independent instances.

Same member names and meaning as the reference (include/Moby/TimeSteppingSimulator.h:36, ConstraintSimulator.h:
51-68, Simulator.h:50,80): `step(dt)` returns dt, `current_time`, `contact_dist_thresh`, `min_step_size`,
`post_step_callback_fn`.  All compute goes through the C ABI (include/b200moby.h); there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import capi


class TimeSteppingSimulator:
    def __init__(self, scene, device=0):
        self.scene = scene
        self.device = device
        self.contact_dist_thresh = scene.contact_dist_thresh
        self.min_step_size = scene.min_step_size
        self.post_step_callback_fn = None          # called as fn(self) after every step (Simulator.h:80)
        self._desc = scene.cdesc()
        self._h = C.c_void_p()
        capi.check(capi.lib().b200moby_create(C.byref(self._desc), device, C.byref(self._h)))
        self.n_envs, self.n_bodies = scene.n_envs, scene.n_bodies
        self.set_state(scene.q, scene.v)
        self.rc = getattr(scene, "rc", None)
        if self.rc is not None:
            self.set_joint_state(self.rc.jq, self.rc.jqd)
        self.steps_taken = 0

    def close(self):
        """Destroy the device-side simulator now (b200moby_destroy); the object is unusable afterwards."""
        self.__del__()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                capi.lib().b200moby_destroy(h)
            except (TypeError, AttributeError):   # interpreter shutdown: module globals are already gone, the process frees the device
                pass
            self._h = None

    # ---- state (host buffers, SoA [body][7|6][env]) ----
    def set_state(self, q, v):
        q, v = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(v, np.float64)
        assert q.shape == (self.n_bodies, 7, self.n_envs) and v.shape == (self.n_bodies, 6, self.n_envs)
        capi.check(capi.lib().b200moby_set_state(self._h, q.ctypes.data, v.ctypes.data))

    def get_state(self):
        q = np.empty((self.n_bodies, 7, self.n_envs))
        v = np.empty((self.n_bodies, 6, self.n_envs))
        capi.check(capi.lib().b200moby_get_state(self._h, q.ctypes.data, v.ctypes.data))
        return q, v

    # ---- articulated body: joint state [dof][env] (RCArticulatedBodyd generalized coordinates / velocities) ----
    def set_joint_state(self, jq, jqd):
        jq, jqd = np.ascontiguousarray(jq, np.float64), np.ascontiguousarray(jqd, np.float64)
        assert jq.shape == (self.rc.n_dof, self.n_envs) and jqd.shape == jq.shape
        capi.check(capi.lib().b200moby_set_joint_state(self._h, jq.ctypes.data, jqd.ctypes.data))

    def get_joint_state(self):
        jq, jqd = np.empty((self.rc.n_dof, self.n_envs)), np.empty((self.rc.n_dof, self.n_envs))
        capi.check(capi.lib().b200moby_get_joint_state(self._h, jq.ctypes.data, jqd.ctypes.data))
        return jq, jqd

    def set_joint_state_dev(self, jq, jqd, stream=None):
        capi.check(capi.lib().b200moby_set_joint_state_dev(self._h, jq.data_ptr(), jqd.data_ptr(), _stream(stream)))

    def get_joint_state_dev(self, jq, jqd, stream=None):
        capi.check(capi.lib().b200moby_get_joint_state_dev(self._h, jq.data_ptr(), jqd.data_ptr(), _stream(stream)))

    def set_joint_forces(self, tau):
        """Generalized joint forces applied every mini-step until changed ([dof][env]); None clears them."""
        if tau is None:
            capi.check(capi.lib().b200moby_set_joint_forces(self._h, None))
        else:
            tau = np.ascontiguousarray(tau, np.float64)
            assert tau.shape == (self.rc.n_dof, self.n_envs)
            capi.check(capi.lib().b200moby_set_joint_forces(self._h, tau.ctypes.data))

    def rc_fwd_dyn(self, algorithm, jq, jqd, tau=None, stream=None):
        """Batched articulated-body forward dynamics on torch CUDA tensors [dof][env]; returns qdd."""
        import torch
        qdd = torch.empty_like(jq)
        capi.check(capi.lib().b200moby_rc_fwd_dyn_batched(self._h, int(algorithm), jq.data_ptr(), jqd.data_ptr(),
                                                          None if tau is None else tau.data_ptr(), qdd.data_ptr(), _stream(stream)))
        return qdd

    def rc_inertia(self, jq, stream=None):
        import torch
        nd = self.rc.n_dof
        Hm = torch.empty((nd * nd, self.n_envs), dtype=torch.float64, device=jq.device)
        capi.check(capi.lib().b200moby_rc_inertia_batched(self._h, jq.data_ptr(), Hm.data_ptr(), _stream(stream)))
        return Hm

    # ---- state (torch CUDA tensors, no host round trip) ----
    def set_state_dev(self, q, v, stream=None):
        capi.check(capi.lib().b200moby_set_state_dev(self._h, q.data_ptr(), v.data_ptr(), _stream(stream)))

    def get_state_dev(self, q, v, stream=None):
        capi.check(capi.lib().b200moby_get_state_dev(self._h, q.data_ptr(), v.data_ptr(), _stream(stream)))

    def step(self, dt, n_steps=1, stream=None):
        """TimeSteppingSimulator::step for every env (asynchronous on `stream`); returns dt like the reference."""
        capi.check(capi.lib().b200moby_step(self._h, float(dt), int(n_steps), _stream(stream)))
        self.steps_taken += n_steps
        if self.post_step_callback_fn is not None:
            self.post_step_callback_fn(self)
        return dt

    @property
    def current_time(self):
        t = np.empty(self.n_envs)
        capi.check(capi.lib().b200moby_get_time(self._h, t.ctypes.data))
        return t

    def counters(self):
        c = capi.Counters()
        capi.check(capi.lib().b200moby_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def set_pivot_budget(self, budget):
        """Scheduling knob only (results are identical): see b200moby_set_pivot_budget."""
        capi.check(capi.lib().b200moby_set_pivot_budget(self._h, int(budget)))

    def launch_count(self):
        n = C.c_longlong()
        capi.check(capi.lib().b200moby_get_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def reset_counters(self):
        capi.check(capi.lib().b200moby_reset_counters(self._h))

    def kernel_profile(self, enable=True, reset=True):
        """Per-kernel durations / envs / algorithmic flops of the steps since the last reset (synchronises)."""
        kp = capi.KernelProfile()
        capi.check(capi.lib().b200moby_get_kernel_profile(self._h, int(enable), int(reset), C.byref(kp)))
        return [dict(name=kp.k[i].name.decode(), ms=kp.k[i].ms, launches=kp.k[i].launches, envs=kp.k[i].envs, flops=kp.k[i].flops,
                     lcp_solves=kp.k[i].lcp_solves, lcp_nmax=kp.k[i].lcp_nmax, threads_per_env=kp.k[i].threads_per_env)
                for i in range(kp.n_kernels)]

    def impact_profile(self):
        """Debug tap: (cycles, pivots, executed iterations, n) per env of the last impact phase; first call arms it."""
        prof = np.zeros((13, self.n_envs), np.int64)
        capi.check(capi.lib().b200moby_get_impact_profile(self._h, prof.ctypes.data))
        return prof

    def env_stats(self):
        """Debug tap: per-env (lcp_failures, lemke_calls, lcp_fast_calls, lcp_solves, pivots) since the previous call, as a dict of
        [env] arrays; the first call arms the tap and returns zeros."""
        st = np.zeros((5, self.n_envs), np.int32)
        capi.check(capi.lib().b200moby_get_env_stats(self._h, st.ctypes.data))
        return dict(lcp_failures=st[0], lemke_calls=st[1], lcp_fast_calls=st[2], lcp_solves=st[3], pivots=st[4])

    def last_lcp_z(self, zcap):
        n = np.zeros(self.n_envs, np.int32)
        z = np.zeros((self.n_envs, zcap))
        capi.check(capi.lib().b200moby_get_last_lcp(self._h, n.ctypes.data, z.ctypes.data, zcap))
        return n, z

    # ---- stage kernels (torch CUDA tensors) ----
    def fwd_dyn(self, q, v, dt, stream=None):
        capi.check(capi.lib().b200moby_fwd_dyn_batched(self._h, q.data_ptr(), v.data_ptr(), float(dt), _stream(stream)))

    def find_contacts(self, q, v, cap, stream=None):
        import torch
        ne, dev = self.n_envs, q.device
        count = torch.zeros(ne, dtype=torch.int32, device=dev)
        pt, nr, t1, t2 = (torch.zeros((cap, 3, ne), dtype=torch.float64, device=dev) for _ in range(4))
        pair = torch.zeros((cap, ne), dtype=torch.int32, device=dev)
        dist = torch.zeros((cap, ne), dtype=torch.float64, device=dev)
        capi.check(capi.lib().b200moby_find_contacts_batched(self._h, q.data_ptr(), v.data_ptr(), cap, count.data_ptr(), pt.data_ptr(),
                                                             nr.data_ptr(), t1.data_ptr(), t2.data_ptr(), pair.data_ptr(),
                                                             dist.data_ptr(), _stream(stream)))
        return dict(count=count, point=pt, normal=nr, tan1=t1, tan2=t2, pair=pair, dist=dist)

    def delassus(self, q, v, nmax, stream=None):
        import torch
        ne, dev = self.n_envs, q.device
        MM = torch.zeros((ne, nmax * nmax), dtype=torch.float64, device=dev)
        qq = torch.zeros((ne, nmax), dtype=torch.float64, device=dev)
        n = torch.zeros(ne, dtype=torch.int32, device=dev)
        capi.check(capi.lib().b200moby_delassus_batched(self._h, q.data_ptr(), v.data_ptr(), nmax, MM.data_ptr(), qq.data_ptr(),
                                                        n.data_ptr(), _stream(stream)))
        return MM, qq, n


def _stream(stream):
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        except ImportError:
            pass
        return None
    return C.c_void_p(stream)
