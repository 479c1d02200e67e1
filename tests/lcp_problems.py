"""Seeded LCP generators shared by the oracle tests, the GPU parity tests and bench.py (SURVEY.md 8d case 6)."""
import itertools

import numpy as np


def random_psd_lcp(n, rng, reg=1e-3):
    """M = A A^T / n + reg I, q ~ N(0,1)."""
    A = rng.standard_normal((n, n))
    return A @ A.T / n + reg * np.eye(n), rng.standard_normal(n)


def random_batch(batch, n, seed, reg=1e-3):
    rng = np.random.default_rng(seed)
    Ms, qs = zip(*(random_psd_lcp(n, rng, reg) for _ in range(batch)))
    return np.stack(Ms), np.stack(qs)


def brute_force_lcp(M, q, tol=1e-9):
    """All solutions found by enumerating complementary bases (n <= 12)."""
    n = len(q)
    sols = []
    for k in range(n + 1):
        for S in itertools.combinations(range(n), k):
            S = list(S)
            z = np.zeros(n)
            if S:
                try:
                    z[S] = np.linalg.solve(M[np.ix_(S, S)], -q[S])
                except np.linalg.LinAlgError:
                    continue
            w = M @ z + q
            if z.min(initial=0) >= -tol and w.min() >= -tol:
                sols.append(z)
    return sols


def lcp_residuals(M, q, z):
    w = M @ z + q
    return dict(min_z=z.min(), min_w=w.min(), max_zw=np.abs(z * w).max())
