"""GPU experiment: UR10 step time over consecutive windows of 20 steps, graph and plain launches (is the workload drifting?)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moby_b200 import TimeSteppingSimulator, scenes
W = bench.WORKLOADS["ur10"]
out = {}
for mode in ("1", "0"):
    os.environ["B200MOBY_GRAPH"] = mode
    sc = W["make"](scenes, W["envs"], 0xB200)
    sc.stabilization_max_iterations = 0
    sim = TimeSteppingSimulator(sc)
    sim.step(W["dt"], W["preroll"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    res = []
    for k in range(5):
        t = bench._device_timed(sim, W["dt"], 20, 3 if k == 0 else 0, flush, stream)
        c = sim.counters()
        res.append((round(1e3 * t / 20, 3), c["contacts"], c["lcp_solves"]))
    out["graph" if mode == "1" else "plain"] = res
    sim.close()
print(json.dumps(out))
