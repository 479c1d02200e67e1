"""GPU experiment: per-env cost of the impact phase on the bench workload (cycles / pivots histogram)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moby_b200 import TimeSteppingSimulator, scenes

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sc = scenes.small_lcp_batch(ne)
sim = TimeSteppingSimulator(sc)
sim.step(1e-3, 300)
sim.impact_profile()
res = []
for s in range(5):
    sim.step(1e-3, 1)
    torch.cuda.synchronize()
    p = sim.impact_profile()
    cyc, piv, ex, n = p[:4]
    kslot, n = n // 1000, n % 1000
    ph = p[4:]
    m = cyc > 0
    r = dict(step=s, envs=int(m.sum()), cyc_sum=int(cyc.sum()), cyc_pct=[float(np.percentile(cyc[m], x)) for x in (50, 90, 99, 99.9, 100)],
             piv_pct=[float(np.percentile(piv[m], x)) for x in (50, 90, 99, 99.9, 100)], ex_pct=[float(np.percentile(ex[m], x)) for x in (50, 90, 99, 99.9, 100)])
    by_n = {}
    for nn in np.unique(n[m]):
        k = m & (n == nn)
        by_n[int(nn)] = dict(phases=[round(float(ph[j][k].mean())) for j in range(9)], envs=int(k.sum()), cyc_mean=float(cyc[k].mean()), cyc_max=int(cyc[k].max()), cyc_sum=int(cyc[k].sum()), piv_mean=float(piv[k].mean()), ex_mean=float(ex[k].mean()))
    r["by_n"] = by_n
    r["by_kslot_n"] = {f"{int(ks)}:{int(nn)}": dict(phases=[round(float(ph[j][k2].mean())) for j in range(9)], envs=int(k2.sum()), cyc_mean=float(cyc[k2].mean()), cyc_max=int(cyc[k2].max()), ex_mean=float(ex[k2].mean()), ex_max=int(ex[k2].max()))
                       for ks in np.unique(kslot[m]) for nn in np.unique(n[m & (kslot == ks)]) for k2 in [m & (kslot == ks) & (n == nn)]}
    top = np.argsort(cyc)[-8:]
    r["top"] = [dict(e=int(e), cyc=int(cyc[e]), piv=int(piv[e]), ex=int(ex[e]), n=int(n[e]), kslot=int(kslot[e]), fast_cyc=int(ph[5][e]), lemke_cyc=int(ph[6][e])) for e in top]
    r["by_kslot"] = {int(k): dict(envs=int((m & (kslot == k)).sum()), cyc_max=int(cyc[m & (kslot == k)].max()), cyc_sum=int(cyc[m & (kslot == k)].sum()), ex_max=int(ex[m & (kslot == k)].max())) for k in np.unique(kslot[m])}
    big = np.where(m & (ex >= 500))[0]
    r["long"] = [dict(e=int(e), ex=int(ex[e]), cyc=int(cyc[e]), kslot=int(kslot[e]), n=int(n[e])) for e in big]
    res.append(r)
print(json.dumps(res, indent=1))
