"""ctypes binding of include/b200moby.h.  Fails loudly when libb200moby.so is absent (no CPU fallback)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MOBY_LIB", os.path.join(_HERE, "libb200moby.so"))   # override: build experiments only


class B200MobyError(RuntimeError):
    pass


class RcDesc(C.Structure):
    """b200moby_rc_desc: the fixed-base articulated body of each env."""
    _fields_ = [
        ("n_links", C.c_int), ("first_body", C.c_int),
        ("parent", C.POINTER(C.c_int)), ("joint_type", C.POINTER(C.c_int)),
        ("joint_axis", C.POINTER(C.c_double)), ("loc_parent", C.POINTER(C.c_double)), ("loc_child", C.POINTER(C.c_double)),
        ("rel_quat", C.POINTER(C.c_double)),
        ("fdyn_algorithm", C.c_int),
        ("ctrl_kp", C.POINTER(C.c_double)), ("ctrl_kv", C.POINTER(C.c_double)), ("ctrl_amp", C.POINTER(C.c_double)),
        ("ctrl_freq", C.POINTER(C.c_double)),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int), ("n_bodies", C.c_int),
        ("shape", C.POINTER(C.c_int)), ("enabled", C.POINTER(C.c_int)), ("mass", C.POINTER(C.c_double)),
        ("dims", C.POINTER(C.c_double)), ("inertia", C.POINTER(C.c_double)),
        ("mu_coulomb", C.POINTER(C.c_double)), ("mu_viscous", C.POINTER(C.c_double)),
        ("epsilon", C.POINTER(C.c_double)), ("compliance", C.POINTER(C.c_double)), ("NK", C.POINTER(C.c_int)),
        ("gravity", C.c_double * 3), ("contact_dist_thresh", C.c_double), ("min_step_size", C.c_double),
        ("min_step_size_env", C.POINTER(C.c_double)),
        ("impact_model", C.c_int), ("stabilization_max_iterations", C.c_int),
        ("rc", C.POINTER(RcDesc)),
        ("max_contacts", C.c_int), ("max_lcp_n", C.c_int),
    ]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_double), ("launches", C.c_longlong), ("envs", C.c_longlong),
                ("flops", C.c_longlong), ("lcp_solves", C.c_longlong), ("lcp_nmax", C.c_int), ("threads_per_env", C.c_int)]


class KernelProfile(C.Structure):
    _fields_ = [("n_kernels", C.c_int), ("k", KernelStat * 20)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_longlong) for k in (
        "env_steps", "mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "pivots", "lcp_failures",
        "impact_tol_events", "contacts", "max_lcp_n", "pivot_flops", "ca_iterations", "assembly_flops", "stab_iterations",
        "stab_lcp_solves", "stab_line_search_failures")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# every symbol include/b200moby.h declares (tests check the library exports all of them)
SYMBOLS = [
    "b200moby_last_error", "b200moby_abi_version", "b200moby_device_count",
    "b200moby_create", "b200moby_destroy", "b200moby_set_state", "b200moby_get_state",
    "b200moby_set_state_dev", "b200moby_get_state_dev", "b200moby_step", "b200moby_set_pivot_budget", "b200moby_get_counters",
    "b200moby_reset_counters", "b200moby_get_launch_count", "b200moby_get_time", "b200moby_get_last_lcp",
    "b200moby_get_impact_profile", "b200moby_get_kernel_profile", "b200moby_get_env_stats",
    "b200moby_lcp_lemke_batched", "b200moby_lcp_fast_batched", "b200moby_lcp_lemke_regularized_batched",
    "b200moby_lcp_fast_regularized_batched", "b200moby_lcp_lemke_host", "b200moby_lcp_fast_host", "b200moby_lcp_solve_host",
    "b200moby_selftest_div",
    "b200moby_fwd_dyn_batched", "b200moby_find_contacts_batched", "b200moby_find_contacts_host", "b200moby_delassus_batched",
    "b200moby_set_joint_state", "b200moby_get_joint_state", "b200moby_set_joint_state_dev", "b200moby_get_joint_state_dev",
    "b200moby_set_joint_forces", "b200moby_rc_fwd_dyn_batched", "b200moby_rc_inertia_batched",
]

_lib = None


def lib():
    """Load libb200moby.so (raises B200MobyError if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200MobyError(f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback for the hot path)")
    L = C.CDLL(LIB_PATH)
    L.b200moby_last_error.restype = C.c_char_p
    vp, ip, dp = C.c_void_p, C.c_void_p, C.c_void_p  # device/host pointers passed as integers
    L.b200moby_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(C.c_void_p)]
    L.b200moby_destroy.argtypes = [C.c_void_p]
    L.b200moby_set_state.argtypes = [C.c_void_p, dp, dp]
    L.b200moby_get_state.argtypes = [C.c_void_p, dp, dp]
    L.b200moby_set_state_dev.argtypes = [C.c_void_p, dp, dp, vp]
    L.b200moby_get_state_dev.argtypes = [C.c_void_p, dp, dp, vp]
    L.b200moby_step.argtypes = [C.c_void_p, C.c_double, C.c_int, vp]
    L.b200moby_set_pivot_budget.argtypes = [C.c_void_p, C.c_int]
    L.b200moby_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    L.b200moby_reset_counters.argtypes = [C.c_void_p]
    L.b200moby_get_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    L.b200moby_get_time.argtypes = [C.c_void_p, dp]
    L.b200moby_get_last_lcp.argtypes = [C.c_void_p, ip, dp, C.c_int]
    L.b200moby_get_impact_profile.argtypes = [C.c_void_p, ip]
    L.b200moby_get_env_stats.argtypes = [C.c_void_p, ip]
    L.b200moby_get_kernel_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(KernelProfile)]
    L.b200moby_lcp_lemke_batched.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, ip, ip, ip, C.c_int, vp]
    L.b200moby_lcp_fast_batched.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_int, C.c_double, ip, ip, ip, C.c_int, vp]
    L.b200moby_lcp_lemke_regularized_batched.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_int, C.c_int, C.c_int,
                                                         C.c_double, C.c_double, ip, ip, vp]
    L.b200moby_lcp_fast_regularized_batched.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                        C.c_double, ip, ip, vp]
    L.b200moby_lcp_lemke_host.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, ip, ip, C.c_int]
    L.b200moby_lcp_fast_host.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_int, C.c_double, ip, ip, C.c_int]
    L.b200moby_lcp_solve_host.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                          C.c_int, ip, ip, C.c_int]
    L.b200moby_selftest_div.argtypes = [C.c_int, dp, dp, dp, dp, vp]
    L.b200moby_fwd_dyn_batched.argtypes = [C.c_void_p, dp, dp, C.c_double, vp]
    L.b200moby_find_contacts_batched.argtypes = [C.c_void_p, dp, dp, C.c_int, ip, dp, dp, dp, dp, ip, dp, vp]
    L.b200moby_find_contacts_host.argtypes = [C.c_void_p, C.c_int, ip, dp, dp, dp, dp, ip, dp]
    L.b200moby_delassus_batched.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, dp, ip, vp]
    L.b200moby_set_joint_state.argtypes = [C.c_void_p, dp, dp]
    L.b200moby_get_joint_state.argtypes = [C.c_void_p, dp, dp]
    L.b200moby_set_joint_state_dev.argtypes = [C.c_void_p, dp, dp, vp]
    L.b200moby_get_joint_state_dev.argtypes = [C.c_void_p, dp, dp, vp]
    L.b200moby_set_joint_forces.argtypes = [C.c_void_p, dp]
    L.b200moby_rc_fwd_dyn_batched.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, vp]
    L.b200moby_rc_inertia_batched.argtypes = [C.c_void_p, dp, dp, vp]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise B200MobyError(f"b200moby status {status}: {lib().b200moby_last_error().decode()}")
