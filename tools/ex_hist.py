"""GPU experiment: executed-iteration histogram of the thread-per-env impact classes and how well the previous step's
count predicts it (is sorting a class queue by cost history worth it?)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moby_b200 import TimeSteppingSimulator, scenes

ne = 65536
sim = TimeSteppingSimulator(scenes.small_lcp_batch(ne))
sim.step(1e-3, 300)
sim.impact_profile()
prev = None
out = []
for s in range(4):
    sim.step(1e-3, 1)
    torch.cuda.synchronize()
    p = sim.impact_profile()
    cyc, piv, ex, n = p[:4]
    kslot, n = n // 1000, n % 1000
    r = {}
    for ks in (3, 4, 5):
        k = (cyc > 0) & (kslot == ks)
        h = np.bincount(np.minimum(ex[k], 13).astype(int), minlength=14).tolist()
        r[ks] = dict(envs=int(k.sum()), ex_hist=h, cyc_pct=[float(np.percentile(cyc[k], x)) for x in (10, 50, 90, 100)])
        if prev is not None:
            both = k & (prev[0] > 0)
            a, b = prev[1][both], ex[both]
            r[ks]["pred"] = dict(n=int(both.sum()), hi_now=int((b >= 3).sum()), hi_prev=int((a >= 3).sum()), hi_both=int(((a >= 3) & (b >= 3)).sum()))
    out.append(r)
    prev = (cyc.copy(), ex.copy())
print(json.dumps(out))
