"""GPU experiment: drift of the CUDA path against the CPU oracle over N steps (north star: 1e-9 relative per step, drift
reported).  Prints one JSON object; profiles/ keeps the output."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import oracle_api as O
from moby_b200 import TimeSteppingSimulator, scenes

O.build()
marks = [1, 10, 100, 300, 1000]
out = {}
for name, make, dt in (("small (boxes + balls, QP model)", lambda: scenes.small_lcp_batch(256, seed=77), 1e-3),
                       ("small, mu = 100 (no-slip model)", lambda: (lambda s: (s.mu_coulomb.__setitem__(slice(None), 100.0), s)[1])(scenes.small_lcp_batch(256, seed=78)), 1e-3),
                       ("ur10 + block + table, mu = 100", lambda: scenes.ur10(32, mu=100.0), 5e-4)):
    sc = make()
    sim = TimeSteppingSimulator(sc)
    ob = O.OracleBatch(sc, 0, sc.n_envs)
    rows, done = [], 0
    for mk in marks:
        sim.step(dt, mk - done)
        ob.run(dt, mk - done, threads=os.cpu_count() or 1)
        done = mk
        q, v = sim.get_state()
        qs = [ob.get_state(i) for i in range(sc.n_envs)]
        qo = np.stack([a for a, _ in qs], axis=2); vo = np.stack([b for _, b in qs], axis=2)
        scale = np.maximum(1.0, np.maximum(np.abs(qo).max(axis=(0, 1)), np.abs(vo).max(axis=(0, 1))))
        err = np.maximum(np.abs(q - qo).max(axis=(0, 1)), np.abs(v - vo).max(axis=(0, 1))) / scale
        rows.append({"steps": mk, "median_rel_err": float(np.median(err)), "p99_rel_err": float(np.percentile(err, 99)), "max_rel_err": float(err.max()),
                     "envs_above_1e-9": int((err > 1e-9).sum()), "envs_bit_identical": int((err == 0).sum()), "envs": int(sc.n_envs)})
    cg = sim.counters()
    out[name] = {"dt": dt, "drift": rows, "gpu_counters": {k: cg[k] for k in ("env_steps", "mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "lcp_failures")}}
print(json.dumps(out, indent=1))
