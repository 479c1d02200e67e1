"""Freezes known-answer (M, q, z, pivot log) vectors from the oracle: random PSD problems plus the LCPs the
oracle's sitting-box / bouncing-ball / sphere-stack runs hand to the solver.  Run from the repo root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_api as O  # noqa: E402
from lcp_problems import random_psd_lcp  # noqa: E402
from moby_b200 import scenes  # noqa: E402

out, k = {}, 0
rng = np.random.default_rng(20261017)
probs = [random_psd_lcp(n, rng) for n in (4, 8, 8, 16, 24, 32, 40, 64, 96)]
for sc, dt, steps in ((scenes.sitting_box(1, NK=8), 1e-3, 3), (scenes.sitting_box(1, NK=4, mu=0.5), 1e-3, 3),
                      (scenes.sphere_stack(1), 1e-3, 3)):
    sim = O.OracleSim(sc)
    sim.step(dt, steps)
    n, MM, qq, z = sim.last_lcp()
    probs.append((MM, qq))
for M, q in probs:
    ok, z, li = O.lcp_lemke(M, q)
    ok2, z2, fi = O.lcp_fast(M, q)
    assert ok and ok2
    out[f"M{k}"], out[f"q{k}"], out[f"z{k}"], out[f"zfast{k}"] = M, q, z, z2
    out[f"lemke_log{k}"], out[f"fast_log{k}"] = li["log"], fi["log"]
    k += 1
out["count"] = k
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lcp_kats.npz"), **out)
print("wrote", k, "KATs")
