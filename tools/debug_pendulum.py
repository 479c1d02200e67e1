"""GPU experiment: where does the pendulum-on-plane batch leave the host build of the same code?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import hostsim_api as H
from test_rc_stepping import pendulum_on_plane
from moby_b200 import TimeSteppingSimulator

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 5
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sc = pendulum_on_plane(ne)
sim, hs = TimeSteppingSimulator(sc), H.HostSim(sc)
done = set()
for blk in range(1200 // chunk):
    sim.step(1e-3, chunk); hs.step(1e-3, chunk)
    jq, jqd = sim.get_joint_state()
    q, v = sim.get_state()
    for e in range(ne):
        d = max(abs(jq[0, e] - hs.jq[0, e]), abs(jqd[0, e] - hs.jqd[0, e]), np.abs(q[:, :, e] - hs.q[:, :, e]).max(), np.abs(v[:, :, e] - hs.v[:, :, e]).max())
        if d > 0 and e not in done:
            done.add(e)
            print(f"env {e}: first difference after step {(blk + 1) * chunk}: {d:.3e}  jq {jq[0, e]!r} vs {hs.jq[0, e]!r}  jqd {jqd[0, e]!r} vs {hs.jqd[0, e]!r}")
            print("   q gpu", q[1, :, e], "host", hs.q[1, :, e])
            print("   v gpu", v[1, :, e], "host", hs.v[1, :, e])
print("gpu counters", sim.counters())
print("host counters", hs.counters_dict())
jq, jqd = sim.get_joint_state()
print("final max diff", np.abs(jq - hs.jq).max(), np.abs(jqd - hs.jqd).max())
