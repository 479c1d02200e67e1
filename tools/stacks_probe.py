"""GPU experiment: per-step cost of the 10-box stack (BASELINE configs[2]) at a small batch, with the per-kernel profile."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
from moby_b200 import TimeSteppingSimulator, scenes

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sc = scenes.box_stack(ne, 10, seed=0xB200)
sim = TimeSteppingSimulator(sc)
prev = sim.counters()
for s in range(steps):
    sim.kernel_profile(enable=True, reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step(1e-3, 1)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    kp = sim.kernel_profile(enable=False, reset=True)
    c = sim.counters()
    d = {k: c[k] - prev[k] for k in ("mini_steps", "lcp_solves", "lcp_fast_calls", "lemke_calls", "pivots", "lcp_failures", "ca_iterations")}
    prev = c
    print(f"step {s}: {el:.3f} s", d, [(k["name"], round(k["ms"], 1), k["envs"]) for k in kp if k["launches"] and k["ms"] > 0.5], flush=True)
