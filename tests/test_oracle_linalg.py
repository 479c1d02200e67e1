"""Pins of the oracle's dense helpers against LAPACK through scipy / numpy.  SURVEY.md 8(c) records these as assumptions
about the absent Ravelin dependency (LinAlgd::solve_fast = dgesv: partial pivoting on the first maximum; factor_chol =
dpotrf, false when not positive definite; inverse_SPD through the Cholesky factor); the pivot order matters because
lcp_fast and the reference's Lemke re-solve with it at every pivot."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla


@pytest.fixture(scope="module")
def L(oracle):
    lib = oracle.lib()
    lib.oracle_solve_fast.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_factor_chol.argtypes = [C.c_int, C.c_void_p]
    lib.oracle_inverse_spd.argtypes = [C.c_int, C.c_void_p]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [1, 2, 5, 17, 40, 96])
def test_solve_fast_pivots_like_lapack(L, n):
    rng = np.random.default_rng(100 + n)
    for trial in range(20):
        A = rng.normal(size=(n, n))
        if trial % 4 == 0 and n > 2:                      # rows of equal magnitude in a column: the first maximum must win
            A[1, 0] = -A[0, 0]
            A[2, 0] = A[0, 0]
        b = rng.normal(size=n)
        Af, x, piv = np.asfortranarray(A.copy()), b.copy(), np.zeros(n, np.int32)
        assert L.oracle_solve_fast(n, _p(Af), _p(x), _p(piv)) == 1
        lu, ipiv = sla.lu_factor(A)
        assert np.array_equal(piv, ipiv), (n, trial)
        assert np.allclose(x, sla.lu_solve((lu, ipiv), b), rtol=1e-9, atol=1e-11 * np.abs(x).max())
        assert np.allclose(np.tril(Af, -1), np.tril(lu, -1), rtol=1e-10, atol=1e-13) and np.allclose(np.triu(Af), np.triu(lu), rtol=1e-10, atol=1e-12)


def test_solve_fast_reports_an_exactly_singular_matrix(L):
    A = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 4.0]]))
    b = np.array([1.0, 1.0])
    assert L.oracle_solve_fast(2, _p(A), _p(b), None) == 0


@pytest.mark.parametrize("n", [1, 3, 6, 9, 24])
def test_cholesky_and_spd_inverse(L, n):
    rng = np.random.default_rng(7 + n)
    G = rng.normal(size=(n, n + 2))
    A = G @ G.T + 1e-3 * np.eye(n)
    F = np.asfortranarray(A.copy())
    assert L.oracle_factor_chol(n, _p(F)) == 1
    assert np.allclose(np.tril(F), np.linalg.cholesky(A), rtol=1e-11, atol=1e-13)
    Ai = np.asfortranarray(A.copy())
    assert L.oracle_inverse_spd(n, _p(Ai)) == 1
    assert np.allclose(Ai, np.linalg.inv(A), rtol=1e-8, atol=1e-10 * np.abs(Ai).max())
    B = np.asfortranarray(A - (np.linalg.eigvalsh(A).min() + 1e-6) * np.eye(n))       # not positive definite any more
    assert L.oracle_factor_chol(n, _p(B)) == 0
