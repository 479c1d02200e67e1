"""A few steps of a small batch of 10-box stacks (for profiling the block-per-env launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moby_b200 import TimeSteppingSimulator, scenes
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sim = TimeSteppingSimulator(scenes.box_stack(ne, 10, seed=0xB200))
sim.step(1e-3, steps)
torch.cuda.synchronize()
print(sim.counters())
