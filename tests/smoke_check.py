"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0 checked against the oracle."""
import numpy as np


def run():
    import torch
    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    import oracle_api as O
    from lcp_problems import random_batch
    from moby_b200 import TimeSteppingSimulator, scenes
    from moby_b200.lcp import LCP
    O.build()
    # stepped path: 64 randomized envs, 40 steps
    sc = scenes.small_lcp_batch(64, seed=1)
    sim = TimeSteppingSimulator(sc)
    sim.step(1e-3, 40)
    q, v = sim.get_state()
    qo, vo = sc.q.copy(), sc.v.copy()
    O.batch_step(sc, qo, vo, 1e-3, 40, threads=4)
    err = max(np.abs(q - qo).max(), np.abs(v - vo).max())
    assert err < 1e-9, err
    # batched solvers
    M, q_ = random_batch(128, 32, seed=2)
    Md, qd = torch.from_numpy(M).cuda(), torch.from_numpy(q_).cuda()
    zl, sl, _, _ = LCP().lcp_lemke(Md, qd)
    zf, sf, _, _ = LCP().lcp_fast(Md, qd)
    zl, zf = zl.cpu().numpy(), zf.cpu().numpy()
    for b in range(0, 128, 16):
        ok, zo, _ = O.lcp_lemke(M[b], q_[b])
        assert ok and np.allclose(zl[b], zo, atol=1e-9)
        ok, zo, _ = O.lcp_fast(M[b], q_[b])
        assert ok and np.array_equal(zf[b], zo)
    print(f"smoke ok: stepped-path max |gpu - oracle| = {err:.3e}; counters {sim.counters()}")
