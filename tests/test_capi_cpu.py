"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200moby.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    from moby_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        g.build()
    return capi.lib()


def test_exports_every_declared_symbol(L):
    from moby_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "b200moby.h")).read()
    declared = set(re.findall(r"\b(b200moby_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.b200moby_abi_version() == 3


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.b200moby_device_count() == 0
    M, q, z = np.eye(2), -np.ones(2), np.zeros(2)
    st = np.zeros(1, np.int32)
    rc = L.b200moby_lcp_lemke_host(1, 2, M.ctypes.data, q.ctypes.data, z.ctypes.data, -1.0, -1.0, st.ctypes.data, None, 0)
    assert rc == 4 and b"no CPU fallback" in L.b200moby_last_error()      # B200MOBY_ERR_NO_DEVICE
    assert not z.any()


def test_oracle_is_not_linked_into_the_product():
    """The oracle is test infrastructure: the product library and package must not link, load or import it."""
    from moby_b200 import capi
    out = os.popen(f"nm -D {capi.LIB_PATH}").read()
    assert "oracle_" not in out
    pat = re.compile(r"liboracle|oracle_api|import\s+oracle|from\s+oracle|oracle/|oracle_[a-z]+\.h")
    for root, _, files in os.walk(os.path.join(ROOT, "moby_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(root, f)).read()), f
