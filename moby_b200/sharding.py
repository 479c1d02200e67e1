"""Env sharding across the GPUs of one box: envs are independent and no island spans two envs, so rank
g owns the contiguous env range [g*E/G, (g+1)*E/G) and runs the identical kernel sequence.  There is NO collective on
the step path; a run ends with one all-gather of the final state and one all-reduce of the counters (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  Host-side plumbing only."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous env range of `rank`; sizes differ by at most one env."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world)
    e0 = rank * base + min(rank, rem)
    return e0, e0 + base + (1 if rank < rem else 0)


def shard_scene(scene, rank, world):
    """The sub-batch of `scene` that `rank` owns (a copy; per-env arrays sliced on their last axis)."""
    import copy
    e0, e1 = shard_range(scene.n_envs, rank, world)
    s = copy.copy(scene)
    s.n_envs = e1 - e0
    for name in ("shape", "enabled", "mass", "dims", "inertia", "mu_coulomb", "mu_viscous", "epsilon", "compliance", "NK", "q", "v"):
        setattr(s, name, np.ascontiguousarray(getattr(scene, name)[..., e0:e1]))
    if scene.min_step_size_env is not None:
        s.min_step_size_env = np.ascontiguousarray(scene.min_step_size_env[e0:e1])
    for name in getattr(scene, "per_env_extra", ()):          # per-env arrays a scene builder added (joint limits, ...)
        setattr(s, name, np.ascontiguousarray(getattr(scene, name)[..., e0:e1]))
    if getattr(scene, "rc", None) is not None:                # the articulated body carries per-env joint state and points at its scene
        rc = copy.copy(scene.rc)
        rc.scene = s
        rc.jq = np.ascontiguousarray(scene.rc.jq[:, e0:e1])
        rc.jqd = np.ascontiguousarray(scene.rc.jqd[:, e0:e1])
        s.rc = rc
    return s


def select_envs(scene, idx):
    """The sub-batch made of envs `idx` (any order, any subset) of `scene`: what shard_scene does for a contiguous range."""
    import copy
    idx = np.asarray(idx, np.int64)
    s = copy.copy(scene)
    s.n_envs = int(idx.size)
    for name in ("shape", "enabled", "mass", "dims", "inertia", "mu_coulomb", "mu_viscous", "epsilon", "compliance", "NK", "q", "v"):
        setattr(s, name, np.ascontiguousarray(getattr(scene, name)[..., idx]))
    if scene.min_step_size_env is not None:
        s.min_step_size_env = np.ascontiguousarray(scene.min_step_size_env[idx])
    for name in getattr(scene, "per_env_extra", ()):
        setattr(s, name, np.ascontiguousarray(getattr(scene, name)[..., idx]))
    if getattr(scene, "rc", None) is not None:
        rc = copy.copy(scene.rc)
        rc.scene = s
        rc.jq = np.ascontiguousarray(scene.rc.jq[:, idx])
        rc.jqd = np.ascontiguousarray(scene.rc.jqd[:, idx])
        s.rc = rc
    return s


def gather_state(q, v, n_total, group=None):
    """All-gather of the final state: q [body][7][env_local], v [body][6][env_local] (torch tensors, CPU for gloo or CUDA
    for nccl) -> full [body][7][n_total], [body][6][n_total] on every rank.  Shards may differ by one env, so the
    gather is padded to the largest shard."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    widest = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    out = []
    for t in (q, v):
        pad = torch.zeros(t.shape[:-1] + (widest,), dtype=t.dtype, device=t.device)
        pad[..., :t.shape[-1]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.cat([p[..., :shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]] for r, p in enumerate(parts)], dim=-1)
        out.append(full)
    return out[0], out[1]


def reduce_counters(counters, device=None, group=None):
    """Sum (max for max_lcp_n) of the per-rank counter dicts on every rank."""
    import torch
    import torch.distributed as dist
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    mx = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return {k: int(mx[i].item() if k == "max_lcp_n" else t[i].item()) for i, k in enumerate(keys)}
