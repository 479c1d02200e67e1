"""Steps a bench workload for ncu captures: python tools/profile_steps.py <workload> <envs> <preroll> <steps>
(the same scenes and seeds as bench.py; no timing here -- numbers taken under a profiler are never bench values)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from moby_b200 import TimeSteppingSimulator, scenes  # noqa: E402

wl, ne, pre, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
W = bench.WORKLOADS[wl]
sc = W["make"](scenes, ne, 0xB200)
if "--stabilization" in sys.argv:
    sc.stabilization_max_iterations = -1
sim = TimeSteppingSimulator(sc)
sim.step(W["dt"], pre)
torch.cuda.synchronize()
for _ in range(steps):
    sim.step(W["dt"], 1)
torch.cuda.synchronize()
print(sim.counters())
