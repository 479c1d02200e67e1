"""GPU experiment: one step as a CUDA graph (default) against plain launches (B200MOBY_GRAPH=0), kernel profiling off."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpy as np
import bench
from moby_b200 import TimeSteppingSimulator, scenes

wl = sys.argv[1] if len(sys.argv) > 1 else "small"
W = bench.WORKLOADS[wl]
out = {}
for mode in ("1", "0"):
    os.environ["B200MOBY_GRAPH"] = mode
    sc = W["make"](scenes, W["envs"], 0xB200)
    sc.stabilization_max_iterations = 0 if wl == "ur10" else -1
    sim = TimeSteppingSimulator(sc)
    sim.step(W["dt"], W["preroll"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    t = bench._device_timed(sim, W["dt"], 30, 5, flush, stream)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter(); sim.step(W["dt"], 30); t_enq = time.perf_counter() - t0; torch.cuda.synchronize()
    q, v = sim.get_state()
    out["graph" if mode == "1" else "plain"] = dict(ms_per_step=1e3 * t / 30, env_steps_per_s=W["envs"] * 30 / t, host_enqueue_ms_per_step=1e3 * t_enq / 30,
                                                    checksum=float(np.abs(q).sum() + np.abs(v).sum()), counters={k: sim.counters()[k] for k in ("env_steps", "lcp_solves", "pivots", "lcp_failures")})
print(json.dumps(out))
