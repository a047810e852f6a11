"""Philox4x32-10 counter-based generator + Box-Muller normals (CPU oracle of the on-device sampler).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference samples its action noise with
``torch.distributions.MultivariateNormal`` (``planners/mppi_delay.py:158,319``); a K-sharded planner needs samples
that do not depend on how K is split, so the CUDA path draws them from Philox4x32-10 (Salmon et al., SC'11) keyed on
``(seed, call index, global sample k, step t)``.  This module restates that generator in numpy integer arithmetic
(bit-exact against the Random123 known-answer vectors in ``tests/test_oracle_philox.py``) and the fp32 Box-Muller
transform of ``neurallaplacecontrol_b200/csrc/stage1_perturb.cu:philox_normals``.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """``counter``: uint32 (..., 4); ``key``: uint32 (2,) or (..., 2).  Returns uint32 (..., 4)."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.broadcast_to(np.asarray(key, dtype=np.uint32), c.shape[:-1] + (2,)).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[..., 0].astype(np.uint64)
            p1 = M1 * c[..., 2].astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[..., 1] ^ k[..., 0]
            n1 = (p1 & MASK).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[..., 3] ^ k[..., 1]
            n3 = (p0 & MASK).astype(np.uint32)
            c = np.stack([n0, n1, n2, n3], axis=-1)
            k = np.stack([k[..., 0] + W0, k[..., 1] + W1], axis=-1)
    return c


def standard_normals(K, T, k_offset, seed, call_index):
    """z (K, T, 4) float32: the four standard normals the kernel derives for each (global sample, t)."""
    gk = np.arange(k_offset, k_offset + K, dtype=np.uint64)[:, None]
    idx = gk * np.uint64(T) + np.arange(T, dtype=np.uint64)[None, :]
    ctr = np.zeros((K, T, 4), dtype=np.uint32)
    ctr[..., 0] = (idx & MASK).astype(np.uint32)
    ctr[..., 1] = (idx >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = np.uint32(call_index & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((call_index >> 32) & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32_10(ctr, key)
    z = np.zeros((K, T, 4), dtype=np.float32)
    inv = np.float32(2.3283064365386963e-10)
    for h in range(2):
        u1 = (x[..., 2 * h].astype(np.float32) + np.float32(0.5)) * inv
        u2 = x[..., 2 * h + 1].astype(np.float32) * inv
        r = np.sqrt(np.float32(-2.0) * np.log(u1.astype(np.float64))).astype(np.float32)
        ang = 2.0 * np.pi * u2.astype(np.float64)
        z[..., 2 * h] = (r * np.cos(ang)).astype(np.float32)
        z[..., 2 * h + 1] = (r * np.sin(ang)).astype(np.float32)
    return z


def sampled_noise(K, T, nu, k_offset, seed, call_index, chol, mu):
    """noise (K, T, nu) = mu + L z  with L the lower Cholesky factor of the covariance."""
    z = standard_normals(K, T, k_offset, seed, call_index)[..., :nu].astype(np.float64)
    return np.asarray(mu, dtype=np.float64) + z @ np.asarray(chol, dtype=np.float64).T
