"""GPU parity tests of the batched LCP kernels against the CPU oracle (through the C ABI)."""
import os

import numpy as np
import pytest

from lcp_problems import lcp_residuals, random_batch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _log_list(row):
    row = list(row)
    return row[:row.index(-1)] if -1 in row else row


@pytest.mark.parametrize("n,batch", [(1, 8), (3, 64), (8, 256), (24, 128), (40, 128), (96, 48), (200, 6), (320, 3)])
def test_lemke_matches_oracle(torch_cuda, oracle, n, batch):
    """Tableau Lemke vs the oracle's LU-per-pivot Lemke: same leaving-variable sequence on tie-free well-conditioned
    problems, z within 1e-9 relative."""
    torch = torch_cuda
    from moby_b200.lcp import LCP
    M, q = random_batch(batch, n, seed=1000 + n)
    cap = 50 * n + 8
    z, st, piv, log = LCP(log_cap=cap).lcp_lemke(_dev(torch, M), _dev(torch, q))
    z, st, piv, log = z.cpu().numpy(), st.cpu().numpy(), piv.cpu().numpy(), log.cpu().numpy()
    same_seq = 0
    for b in range(batch):
        ok, zo, info = oracle.lcp_lemke(M[b], q[b], log_cap=cap)
        assert ok and st[b] in (0, 1), (b, st[b], info)
        scale = max(1.0, np.abs(zo).max())
        assert np.allclose(z[b], zo, rtol=0, atol=1e-9 * scale), (b, np.abs(z[b] - zo).max())
        if _log_list(log[b]) == list(info["log"]) and piv[b] == info["pivots"]:
            same_seq += 1
    assert same_seq == batch, f"pivot sequences identical on {same_seq}/{batch}"


@pytest.mark.parametrize("n,batch", [(1, 8), (3, 64), (8, 256), (24, 128), (40, 128), (96, 48), (200, 6)])
def test_fast_matches_oracle_bitwise(torch_cuda, oracle, n, batch):
    """lcp_fast follows the oracle's arithmetic order: identical index sequences and bit-identical z."""
    torch = torch_cuda
    from moby_b200.lcp import LCP
    M, q = random_batch(batch, n, seed=2000 + n)
    cap = 4 * n + 8
    z, st, piv, log = LCP(log_cap=cap).lcp_fast(_dev(torch, M), _dev(torch, q))
    z, st, piv, log = z.cpu().numpy(), st.cpu().numpy(), piv.cpu().numpy(), log.cpu().numpy()
    for b in range(batch):
        ok, zo, info = oracle.lcp_fast(M[b], q[b], log_cap=cap)
        assert ok == (st[b] in (0, 1)), (b, st[b], info)
        assert piv[b] == info["pivots"] and _log_list(log[b]) == list(info["log"]), b
        if ok:
            assert np.array_equal(z[b], zo), (b, np.abs(z[b] - zo).max())


def test_fast_warm_start(torch_cuda, oracle):
    torch = torch_cuda
    from moby_b200.lcp import LCP
    M, q = random_batch(64, 32, seed=77)
    rng = np.random.default_rng(3)
    z0 = np.where(rng.random((64, 32)) < 0.3, rng.random((64, 32)), 0.0)      # arbitrary (stale) warm start, rule H1
    z, st, piv, _ = LCP().lcp_fast(_dev(torch, M), _dev(torch, q), z0=_dev(torch, z0))
    z, st, piv = z.cpu().numpy(), st.cpu().numpy(), piv.cpu().numpy()
    for b in range(64):
        ok, zo, info = oracle.lcp_fast(M[b], q[b], z0=z0[b])
        assert ok == (st[b] in (0, 1)) and piv[b] == info["pivots"]
        assert np.array_equal(z[b], zo)


def test_regularized_wrappers(torch_cuda, oracle):
    """Includes rank-deficient problems with a zero diagonal block (the [[H,-A'],[A,0]] shape of the QP-as-LCP)."""
    torch = torch_cuda
    from moby_b200.lcp import LCP
    rng = np.random.default_rng(9)
    Ms, qs = [], []
    for _ in range(64):
        m, k = 6, 4
        J = rng.standard_normal((m, 3))
        H = J @ J.T                                   # rank 3 < 6
        A = np.abs(rng.standard_normal((k, m)))
        MM = np.block([[H, -A.T], [A, np.zeros((k, k))]])
        Ms.append(MM)
        qs.append(np.concatenate([rng.standard_normal(m), np.abs(rng.standard_normal(k))]))
    M, q = np.stack(Ms), np.stack(qs)
    lcp = LCP()
    zf, sf, pf = lcp.lcp_fast_regularized(_dev(torch, M), _dev(torch, q), min_exp=-20, step_exp=4, max_exp=-8)
    zl, sl, pl = lcp.lcp_lemke_regularized(_dev(torch, M), _dev(torch, q))
    zf, sf, pf, zl, sl, pl = (t.cpu().numpy() for t in (zf, sf, pf, zl, sl, pl))
    n = M.shape[1]
    for b in range(64):
        ok, zo, info = oracle.lcp_fast_regularized(M[b], q[b], min_exp=-20, step_exp=4, max_exp=-8)
        assert info["status"] == sf[b] and info["pivots"] == pf[b], (b, info, sf[b], pf[b])
        if ok:
            assert np.array_equal(zf[b], zo)
        ok, zo, info = oracle.lcp_lemke_regularized(M[b], q[b])
        # Rank-deficient problems: the tableau and the oracle's LU-per-pivot Lemke round differently, and the wrapper's
        # acceptance test (z.w < T with |z| ~ 1e3) sits at the rounding-noise level, so the accepted lambda may differ
        # by one notch.  What must hold: both accept, and the accepted z solves the matrix that was accepted.
        assert ok == (sl[b] != 6) and abs(int(info["status"]) - int(sl[b])) <= 1, (b, info, sl[b])
        if ok:
            lam = 0.0 if sl[b] < 16 else 10.0 ** (-20 + (sl[b] - 16))          # attempt k = status - 16 (LCP.cpp:419-444)
            Mr = M[b] + lam * np.eye(n)
            T = n * np.abs(M[b]).max() * np.sqrt(np.finfo(float).eps)
            r = lcp_residuals(Mr, q[b], zl[b])
            assert r["min_z"] >= -10 * T and r["min_w"] >= -10 * T and r["max_zw"] < 10 * T, (b, r, T)


def test_frozen_kats(torch_cuda):
    torch = torch_cuda
    from moby_b200.lcp import LCP
    d = np.load(os.path.join(GOLDEN, "lcp_kats.npz"))
    for k in range(int(d["count"])):
        M, q = d[f"M{k}"][None], d[f"q{k}"][None]
        cap = 50 * M.shape[1] + 8
        z, st, piv, log = LCP(log_cap=cap).lcp_fast(_dev(torch, M), _dev(torch, q))
        assert st.item() in (0, 1)
        assert np.array_equal(z.cpu().numpy()[0], d[f"zfast{k}"])
        assert _log_list(log.cpu().numpy()[0]) == list(d[f"fast_log{k}"])
        z, st, piv, log = LCP(log_cap=cap).lcp_lemke(_dev(torch, M), _dev(torch, q))
        assert st.item() in (0, 1)
        r = lcp_residuals(M[0], q[0], z.cpu().numpy()[0])
        T = M.shape[1] * np.abs(M).max() * np.sqrt(np.finfo(float).eps)
        assert r["min_z"] >= -T and r["min_w"] >= -T and r["max_zw"] < T


def test_host_forms_and_block_path(torch_cuda, oracle):
    """Host-buffer entry points (copies inside) and the block-per-LCP path (n too large for one warp's shared memory)."""
    from moby_b200.lcp import lcp_fast_host, lcp_lemke_host
    M, q = random_batch(4, 180, seed=5)
    z, st, piv = lcp_lemke_host(M, q)
    for b in range(4):
        ok, zo, info = oracle.lcp_lemke(M[b], q[b])
        assert ok and st[b] in (0, 1) and np.allclose(z[b], zo, rtol=0, atol=1e-8)
    z, st, piv = lcp_fast_host(M, q)
    for b in range(4):
        ok, zo, info = oracle.lcp_fast(M[b], q[b])       # lcp_fast may hit its 2n cap on a cold start: then both must
        assert ok == (st[b] in (0, 1)) and st[b] == info["status"] and piv[b] == info["pivots"]
        if ok:
            assert np.array_equal(z[b], zo)


@pytest.mark.parametrize("n,batch", [(170, 5), (200, 40), (320, 50), (401, 3)])
def test_cluster_tableau_matches_block_path(torch_cuda, n, batch):
    """n in the hundreds: the Lemke tableau in the distributed shared memory of an 8-CTA cluster (lcp_cluster_kernel) against
    the block-per-LCP kernel that keeps it in global scratch (B200MOBY_LCP_CLUSTER=0): z, status, pivot count and the whole
    pivot log bit for bit; more problems than resident clusters, an n that is not a multiple of the cluster size, and
    problems that end on a ray (q < 0 against a matrix with a zero row)."""
    import os
    torch = torch_cuda
    from moby_b200.lcp import LCP
    M, q = random_batch(batch, n, seed=77 + n)
    M[1, 3, :] = 0.0; M[1, :, 3] = 0.0; q[1, 3] = -1.0            # no z can lift w_3 = q_3 < 0: the solve cannot succeed
    cap = 50 * n + 8
    res = []
    for val in ("1", "0"):
        os.environ["B200MOBY_LCP_CLUSTER"] = val
        try:
            z, st, piv, log = LCP(log_cap=cap).lcp_lemke(_dev(torch, M), _dev(torch, q))
            res.append((z.cpu().numpy(), st.cpu().numpy(), piv.cpu().numpy(), log.cpu().numpy()))
        finally:
            del os.environ["B200MOBY_LCP_CLUSTER"]
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)
    assert res[0][1][1] not in (0, 1) and (res[0][1][[0, 2]] == 0).all()


def test_empty_and_ragged(torch_cuda):
    torch = torch_cuda
    from moby_b200.lcp import LCP
    z, st, piv, _ = LCP().lcp_lemke(torch.zeros((0, 4, 4), dtype=torch.float64, device="cuda"),
                                    torch.zeros((0, 4), dtype=torch.float64, device="cuda"))
    assert z.shape == (0, 4)
    # trivial problems: q >= 0 -> z = 0, status TRIVIAL
    M, q = random_batch(5, 6, seed=1)
    z, st, piv, _ = LCP().lcp_lemke(_dev(torch, M), _dev(torch, np.abs(q)))
    assert (st.cpu().numpy() == 1).all() and not z.cpu().numpy().any()


def test_lockstep_division_is_ieee_division(torch_cuda):
    """b2m_divn (the lock-step division of Lemke's ratio test) against the compiler's `/`, bit for bit: 2^27 random operand
    pairs over 600 binades and over one binade (where misrounding of a Newton-Raphson quotient would show: a seed with a
    zero low word instead of the compiler's 1 failed the stepped-path parity tests at a rate this test now catches),
    small-integer ratios (exact ties), powers of two, zeros, denormals, infinities and NaNs (the last take the fallback)."""
    import torch
    from moby_b200 import capi
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    n = 1 << 25
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 3.0, 1e-310, -4e-320, 2.0 ** -500, 2.0 ** 500, 1e300, 1e-300, float("inf"), float("-inf"),
                            float("nan"), 2.0 ** -1022, 1.7976931348623157e308], dtype=torch.float64, device="cuda")
    m = special.numel()
    for kind in ("wide", "binade", "binade", "integers"):
        if kind == "integers":
            x = torch.randint(1, 4096, (n,), device="cuda", generator=g).double() * 2.0 ** -12
            y = torch.randint(1, 4096, (n,), device="cuda", generator=g).double() * 2.0 ** -9
        else:
            mant = torch.rand(2, n, dtype=torch.float64, device="cuda", generator=g) + 1.0
            span = 300 if kind == "wide" else 1
            ex = torch.randint(-span, span, (2, n), device="cuda", generator=g).double()
            sign = torch.randint(0, 2, (2, n), device="cuda", generator=g).double() * 2 - 1
            x, y = (sign * mant * torch.exp2(ex)).unbind(0)
            x, y = x.clone(), y.clone()
            del mant, ex, sign
        x[:m * m] = special.repeat_interleave(m)
        y[:m * m] = special.repeat(m)
        x[m * m:m * m + 4096] = y[m * m:m * m + 4096] * 3.0            # exact quotients
        q, qref = torch.empty_like(x), torch.empty_like(x)
        capi.check(capi.lib().b200moby_selftest_div(n, x.data_ptr(), y.data_ptr(), q.data_ptr(), qref.data_ptr(), None))
        torch.cuda.synchronize()
        same = (q.view(torch.int64) == qref.view(torch.int64)) | (q.isnan() & qref.isnan())
        assert bool(same.all()), (kind, x[~same][:4], y[~same][:4], q[~same][:4], qref[~same][:4])
        del x, y, q, qref, same


@pytest.mark.parametrize("n,batch", [(2, 1000), (5, 777), (8, 4096)])
def test_small_lcps_several_per_warp(torch_cuda, oracle, n, batch):
    """n <= 8 without a pivot log: eight lanes per problem, four problems per warp (lcp_subwarp_kernel).  Same arithmetic as
    the warp-per-problem kernels: z, status and pivot counts bit-identical to them, for
    all four solver families, ragged batch sizes included; Lemke also against the oracle."""
    torch = torch_cuda
    from moby_b200.lcp import LCP
    M, q = random_batch(batch, n, seed=4000 + n)
    Md, qd = _dev(torch, M), _dev(torch, q)
    solver = LCP(log_cap=0)
    for name in ("lcp_lemke", "lcp_fast", "lcp_lemke_regularized", "lcp_fast_regularized"):
        a = getattr(solver, name)(Md, qd)
        os.environ["B200MOBY_LCP_SUBWARP_NMAX"] = "0"            # the warp-per-problem kernel (read at every launch)
        try:
            b = getattr(solver, name)(Md, qd)
        finally:
            del os.environ["B200MOBY_LCP_SUBWARP_NMAX"]
        za, sa, pa = (x.cpu().numpy() for x in a[:3])
        zb, sb, pb = (x.cpu().numpy() for x in b[:3])
        assert np.array_equal(sa, sb) and np.array_equal(pa, pb), name
        okm = (sa == 0) | (sa == 1) | (sa >= 16)
        assert np.array_equal(za[okm], zb[okm]), name
    z = solver.lcp_lemke(Md, qd)[0].cpu().numpy()
    for b in range(0, batch, max(1, batch // 64)):
        ok, zo, _ = oracle.lcp_lemke(M[b], q[b])
        assert ok and np.allclose(z[b], zo, rtol=0, atol=1e-9 * max(1.0, np.abs(zo).max()))
