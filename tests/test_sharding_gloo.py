"""N > 1 host logic on CPU: world-size-2 gloo.  Each rank steps its env shard (with the kernels' device code compiled
for the host, tests/hostsim -- test infrastructure standing in for the GPU), then the final state is all-gathered and
the counters all-reduced exactly as a multi-GPU run does; the result must equal the single-process full batch."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_envs, steps, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch
    import torch.distributed as dist
    import hostsim_api
    from moby_b200 import scenes, sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = scenes.small_lcp_batch(n_envs, seed=21)
    mine = sharding.shard_scene(full, rank, world)
    hs = hostsim_api.HostSim(mine)
    hs.step_phased(1e-3, steps)
    q, v = sharding.gather_state(torch.from_numpy(hs.q), torch.from_numpy(hs.v), n_envs)
    cnt = sharding.reduce_counters(hs.counters_dict())
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), q=q.numpy(), v=v.numpy(), **{"c_" + k: v_ for k, v_ in cnt.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    from moby_b200 import sharding
    for n, w in ((65536, 8), (10, 3), (7, 8), (1, 1)):
        r = [sharding.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_envs", [24, 25])
def test_two_rank_gloo_matches_single_process(tmp_path, n_envs):
    import torch.multiprocessing as mp
    import hostsim_api
    from moby_b200 import scenes
    hostsim_api.build()
    steps, port = 60, 29500 + (os.getpid() % 2000) + n_envs
    mp.spawn(_worker, args=(2, port, n_envs, steps, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    ref = hostsim_api.HostSim(scenes.small_lcp_batch(n_envs, seed=21))
    ref.step_phased(1e-3, steps)
    assert np.array_equal(got["q"], ref.q) and np.array_equal(got["v"], ref.v)
    rc = ref.counters_dict()
    for k in ("env_steps", "mini_steps", "lcp_solves", "pivots", "contacts", "lcp_fast_calls", "lemke_calls"):
        assert int(got["c_" + k]) == rc[k], k
    assert int(got["c_env_steps"]) == n_envs * steps


def test_shard_scene_slices_the_articulated_body():
    """A shard owns its own ArticulatedBody: joint state sliced, rc.scene pointing at the shard, and
    stepping the two shards equals stepping the full batch."""
    import hostsim_api
    from moby_b200 import scenes, sharding
    hostsim_api.build()
    full = scenes.ur10(8, seed=5)
    parts = [sharding.shard_scene(full, r, 2) for r in range(2)]
    for r, s in enumerate(parts):
        assert s.rc is not full.rc and s.rc.scene is s
        assert s.rc.jq.shape == (9, 4) and s.rc.jqd.shape == (9, 4)
        assert np.array_equal(s.rc.jq, full.rc.jq[:, 4 * r:4 * r + 4])
        for e in range(4):                                   # link_poses / env_mass_props index the shard's own envs
            x_s, R_s = s.rc.link_poses(e)
            x_f, R_f = full.rc.link_poses(4 * r + e)
            assert np.array_equal(x_s, x_f) and np.array_equal(R_s, R_f)
            assert all(np.array_equal(a, b) for a, b in zip(s.rc.env_mass_props(e), full.rc.env_mass_props(4 * r + e)))
    ref = hostsim_api.HostSim(full)
    ref.step(5e-4, 20)
    for r, s in enumerate(parts):
        hs = hostsim_api.HostSim(s)
        hs.step(5e-4, 20)
        assert np.array_equal(hs.q, ref.q[..., 4 * r:4 * r + 4]) and np.array_equal(hs.jq, ref.jq[:, 4 * r:4 * r + 4])
