"""The C++ host facade (include/b200moby.hpp, Moby's class names over the C ABI): builds with plain g++ against
libb200moby.so; without a GPU the program fails loudly (no CPU fallback); on the GPU it reproduces the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def sitting_box():
    import __graft_entry__ as g
    from moby_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        g.build()
    subprocess.check_call(["make", "-s", "-C", CPP])
    return os.path.join(CPP, "sitting_box")


def test_facade_builds_and_refuses_without_gpu(sitting_box):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sitting_box, "10", "5"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_facade_sitting_box_matches_oracle(sitting_box, oracle):
    from moby_b200 import scenes
    r = subprocess.run([sitting_box, "400", "100", "3"], capture_output=True, text=True, check=True)
    rows = [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")]
    osim = oracle.OracleSim(scenes.sitting_box(1, NK=8, y0=0.50001))
    k = 0
    for step in range(1, 401):
        osim.step(1e-3)
        if step % 100 == 0:
            qo, _ = osim.get_state()
            for _ in range(2):                       # env 0 and env n-1 are printed; both are the same scene
                row = rows[k]; k += 1
                assert abs(float(row[0]) - step * 1e-3) < 1e-12
                assert np.array_equal(np.array([float(x) for x in row[2:9]]), qo[0])
    meta = {l.split()[1]: l.split()[2:] for l in r.stdout.splitlines() if l.startswith("#")}
    assert meta["env_steps"][0] == "1200" and meta["env_steps"][-1] == "400"          # 3 envs x 400 steps; 400 callbacks
    zl = [float(x) for x in meta["lcp_lemke"][1:]]
    zf = [float(x) for x in meta["lcp_fast"][1:]]
    assert meta["lcp_lemke"][0] == "1" and np.allclose(zl, [4 / 3, 7 / 3], atol=1e-14)
    assert meta["lcp_fast"][0] == "1" and np.allclose(zf, [4 / 3, 7 / 3], atol=1e-14)
