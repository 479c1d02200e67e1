"""GPU experiment: timeline of the hard-queue launch (B200MOBY_TAP_TIMES=1): when each env starts and ends on the global
timer, how many envs are in flight over time, where the long envs sit."""
import json, sys, os
os.environ["B200MOBY_TAP_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moby_b200 import TimeSteppingSimulator, scenes

sim = TimeSteppingSimulator(scenes.small_lcp_batch(65536))
sim.step(1e-3, 300)
sim.impact_profile()
out = []
for s in range(3):
    sim.step(1e-3, 1)
    torch.cuda.synchronize()
    p = sim.impact_profile()
    cyc, piv, ex, n = p[:4]
    kslot, n = n // 1000, n % 1000
    t0, t1 = p[4 + 2], p[4 + 8]
    r = {}
    for ks in np.unique(kslot[cyc > 0]):
        k = (cyc > 0) & (kslot == ks)
        base = t0[k].min()
        a, b = (t0[k] - base) / 1e6, (t1[k] - base) / 1e6        # ms
        grid = np.linspace(0, b.max(), 14)
        inflight = [int(((a <= x) & (b > x)).sum()) for x in grid]
        started = [int((a <= x).sum()) for x in grid]
        lng = np.where(ex[k] >= 500)[0]
        r[int(ks)] = dict(envs=int(k.sum()), span_ms=float(b.max()), grid_ms=[round(float(x), 2) for x in grid], inflight=inflight, started=started,
                          dur_ms_pct=[round(float(np.percentile(b - a, q)), 3) for q in (50, 90, 99, 100)],
                          long=[dict(start=round(float(a[i]), 2), end=round(float(b[i]), 2), ex=int(ex[k][i]), n=int(n[k][i]),
                                     fast0=round(float((p[4 + 1][k][i] - base) / 1e6), 2), fast1=round(float((p[4 + 3][k][i] - base) / 1e6), 2),
                                     lemke1=round(float((p[4 + 4][k][i] - base) / 1e6), 2), max_pick_ms=float(p[4 + 7][k][i]) / 1e3, max_run_ms=float(p[4 + 0][k][i]) / 1e3,
                                     rungs=int(p[4 + 6][k][i] % 1000), rungs_1000=int(p[4 + 6][k][i] // 1000)) for i in lng])
    out.append(r)
print(json.dumps(out))
