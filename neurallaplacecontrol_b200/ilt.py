"""Generic Fourier-series inverse Laplace transform on the GPU (BASELINE config 2).

Replaces the ``fourier`` ILT of ``torchlaplace`` that ``w_nl.py:137-144`` reaches through
``laplace_reconstruct`` (PARITY UNPINNED: that package is not in the reference tree; the CPU statement
is ``oracle/ilt.py:fourier_line_integrate``)."""
from __future__ import annotations

import torch

from . import _lib


def fourier_ilt(F: torch.Tensor, t: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``F``: complex64 ``(N, n_t, S)`` Laplace-domain samples at the Fourier s-points of each time;
    ``t``: float32 ``(n_t,)`` shared grid or ``(N, n_t)`` per-trajectory times.  Returns float32 ``(N, n_t)``."""
    if not F.is_cuda:
        raise RuntimeError("fourier_ilt runs on the GPU only (no CPU fallback)")
    if F.dtype != torch.complex64 or F.dim() != 3:
        raise TypeError("F must be complex64 of shape (N, n_t, S)")
    N, n_t, S = F.shape
    t = t.to(device=F.device, dtype=torch.float32).contiguous()
    if t.shape == (n_t,):
        per_row = 0
    elif t.shape == (N, n_t):
        per_row = 1
    else:
        raise ValueError("t must have shape (n_t,) or (N, n_t)")
    F = F.contiguous()
    if out is None:
        out = torch.empty((N, n_t), dtype=torch.float32, device=F.device)
    lib = _lib.load()
    with torch.cuda.device(F.device):
        _lib.check(lib.nlc_ilt_fourier(torch.view_as_real(F).data_ptr(), t.data_ptr(), per_row, N, n_t, S,
                                       out.data_ptr(), _lib.current_stream_ptr()), "nlc_ilt_fourier")
    return out
