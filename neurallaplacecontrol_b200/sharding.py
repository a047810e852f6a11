"""Host-side logic of K-sharding over the GPUs of one node.

Rank g owns the global samples ``[g*K/G, (g+1)*K/G)`` - rollouts and costs are independent per sample
(``mppi_delay.py:271-296``) so stages 1-3 need no communication.  Stage 4 exchanges ONE small message per control
step: every rank contributes its ``(beta_g, eta_g, W_g[T][nu])`` triple (``2 + T*nu`` floats) to an all-gather and
every rank then runs the same deterministic log-sum-exp combine (``nlc_softmax_combine``), so ``U`` stays replicated
bit-identically without a broadcast.
"""
from __future__ import annotations

import torch


def shard_range(K: int, G: int, rank: int):
    """(offset, count) of rank's contiguous slice of the K samples; K must divide evenly so that every rank
    launches identical grids (``global k = offset + local k``)."""
    if G < 1 or not 0 <= rank < G:
        raise ValueError("bad shard spec")
    if K % G != 0:
        raise ValueError(f"num_samples={K} must divide over {G} shards")
    n = K // G
    return rank * n, n


def gather_triples(triple: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """All-gather the per-shard triple (``[2+T*nu]``) into ``[G][2+T*nu]`` on every rank (NCCL for CUDA tensors,
    gloo for CPU tensors in the tests)."""
    import torch.distributed as dist

    G = dist.get_world_size(group)
    if out is None:
        out = torch.empty((G, triple.numel()), dtype=triple.dtype, device=triple.device)
    dist.all_gather_into_tensor(out.view(-1), triple.contiguous().view(-1), group=group)
    return out
