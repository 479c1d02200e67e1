"""GPU experiment: where a step of the 10-box stack (BASELINE configs[2], QP-LCP n = 320) spends its time.
Per env of the hard queue: SM cycles per phase of the impact (tap_prof rows), pivots, executed iterations."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moby_b200 import TimeSteppingSimulator, scenes

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
PH = ["load", "contacts", "islands", "problem", "build", "fast", "lemke", "apply", "store"]
sim = TimeSteppingSimulator(scenes.box_stack(ne, 10, seed=0xB200))
sim.step(1e-3, 2)
sim.impact_profile()
out = []
for s in range(steps):
    torch.cuda.synchronize(); t0 = time.time()
    sim.step(1e-3, 1)
    torch.cuda.synchronize(); wall = time.time() - t0
    p = sim.impact_profile()
    cyc, piv, ex, n = p[:4]
    n = n % 1000
    ph = p[4:]
    m = cyc > 0
    order = np.argsort(cyc)[::-1]
    r = dict(step=s, wall_ms=1e3 * wall, envs=int(m.sum()), n=sorted(set(int(x) for x in n[m])),
             mean_phase_Mcyc={PH[j]: round(float(ph[j][m].mean()) / 1e6, 3) for j in range(9)},
             cyc_pct_M=[round(float(np.percentile(cyc[m], x)) / 1e6, 2) for x in (50, 90, 99, 100)],
             piv_pct=[float(np.percentile(piv[m], x)) for x in (50, 90, 99, 100)],
             top=[dict(e=int(e), Mcyc=round(int(cyc[e]) / 1e6, 2), piv=int(piv[e]), ex=int(ex[e]), n=int(n[e]),
                       phases_M={PH[j]: round(int(ph[j][e]) / 1e6, 2) for j in range(9)}) for e in order[:4]])
    out.append(r)
print(json.dumps(out, indent=1))
print(json.dumps(sim.counters() if hasattr(sim, "counters") else {}, default=str)[:600])
