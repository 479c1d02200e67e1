"""Host-side mirror of Moby's LCP object (include/Moby/LCP.h:21-27) for batches: same names, argument meaning and
failure reporting (a per-problem status instead of a bool), over the C ABI.  Inputs are torch CUDA tensors
(device pointers go straight to the kernels) or numpy arrays (host form: copies inside the call)."""
import ctypes as C

import numpy as np

from . import capi


def _stream_ptr(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(stream)


def _colmajor_batch(M):
    """[batch, n, n] row-indexed (M[b][r][c]) -> contiguous column-major blocks."""
    return M.transpose(-1, -2).contiguous()


class LCP:
    """Batched LCP solver object.  M: [batch, n, n] (M[b, r, c]), q: [batch, n]; returns (z, status, pivots[, log])."""

    def __init__(self, log_cap=0):
        self.log_cap = log_cap

    def _prep(self, M, q, z0):
        import torch
        assert M.is_cuda and q.is_cuda and M.dtype == torch.float64
        batch, n = q.shape
        Mc = _colmajor_batch(M)
        qc = q.contiguous()
        z = torch.zeros_like(qc) if z0 is None else z0.clone().contiguous()
        status = torch.zeros(batch, dtype=torch.int32, device=q.device)
        pivots = torch.zeros(batch, dtype=torch.int32, device=q.device)
        log = torch.full((batch, max(self.log_cap, 1)), -1, dtype=torch.int32, device=q.device)
        return batch, n, Mc, qc, z, status, pivots, log

    def lcp_lemke(self, M, q, piv_tol=-1.0, zero_tol=-1.0, stream=None):
        batch, n, Mc, qc, z, status, pivots, log = self._prep(M, q, None)
        capi.check(capi.lib().b200moby_lcp_lemke_batched(batch, n, Mc.data_ptr(), qc.data_ptr(), z.data_ptr(), piv_tol, zero_tol,
                                                         status.data_ptr(), pivots.data_ptr(),
                                                         log.data_ptr() if self.log_cap else None, self.log_cap, _stream_ptr(stream)))
        return z, status, pivots, log

    def lcp_fast(self, M, q, z0=None, zero_tol=-1.0, stream=None):
        batch, n, Mc, qc, z, status, pivots, log = self._prep(M, q, z0)
        capi.check(capi.lib().b200moby_lcp_fast_batched(batch, n, Mc.data_ptr(), qc.data_ptr(), z.data_ptr(), 0 if z0 is None else 1,
                                                        zero_tol, status.data_ptr(), pivots.data_ptr(),
                                                        log.data_ptr() if self.log_cap else None, self.log_cap, _stream_ptr(stream)))
        return z, status, pivots, log

    def lcp_lemke_regularized(self, M, q, min_exp=-20, step_exp=1, max_exp=1, piv_tol=-1.0, zero_tol=-1.0, stream=None):
        batch, n, Mc, qc, z, status, pivots, _ = self._prep(M, q, None)
        capi.check(capi.lib().b200moby_lcp_lemke_regularized_batched(batch, n, Mc.data_ptr(), qc.data_ptr(), z.data_ptr(), min_exp,
                                                                     step_exp, max_exp, piv_tol, zero_tol, status.data_ptr(),
                                                                     pivots.data_ptr(), _stream_ptr(stream)))
        return z, status, pivots

    def lcp_fast_regularized(self, M, q, z0=None, min_exp=-20, step_exp=4, max_exp=20, zero_tol=-1.0, stream=None):
        batch, n, Mc, qc, z, status, pivots, _ = self._prep(M, q, z0)
        capi.check(capi.lib().b200moby_lcp_fast_regularized_batched(batch, n, Mc.data_ptr(), qc.data_ptr(), z.data_ptr(),
                                                                    0 if z0 is None else 1, min_exp, step_exp, max_exp, zero_tol,
                                                                    status.data_ptr(), pivots.data_ptr(), _stream_ptr(stream)))
        return z, status, pivots


def lcp_lemke_host(M, q, piv_tol=-1.0, zero_tol=-1.0, device=0):
    """Host buffers in, host buffers out (H2D + solve + D2H inside): numpy [batch,n,n], [batch,n]."""
    batch, n = q.shape
    Mc = np.ascontiguousarray(np.swapaxes(M, 1, 2), np.float64)
    qc = np.ascontiguousarray(q, np.float64)
    z = np.zeros_like(qc)
    status, pivots = np.zeros(batch, np.int32), np.zeros(batch, np.int32)
    capi.check(capi.lib().b200moby_lcp_lemke_host(batch, n, Mc.ctypes.data, qc.ctypes.data, z.ctypes.data, piv_tol, zero_tol,
                                                  status.ctypes.data, pivots.ctypes.data, device))
    return z, status, pivots


def lcp_fast_host(M, q, z0=None, zero_tol=-1.0, device=0):
    batch, n = q.shape
    Mc = np.ascontiguousarray(np.swapaxes(M, 1, 2), np.float64)
    qc = np.ascontiguousarray(q, np.float64)
    z = np.zeros_like(qc) if z0 is None else np.array(z0, np.float64)
    status, pivots = np.zeros(batch, np.int32), np.zeros(batch, np.int32)
    capi.check(capi.lib().b200moby_lcp_fast_host(batch, n, Mc.ctypes.data, qc.ctypes.data, z.ctypes.data, 0 if z0 is None else 1,
                                                 zero_tol, status.ctypes.data, pivots.ctypes.data, device))
    return z, status, pivots
