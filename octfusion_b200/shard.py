"""Batch sharding across GPUs (one process per GPU, torch.distributed).

Every operator of the U-Net is per-shape (edges never cross shapes, norm statistics are per sample,
attention and dense convolutions are per batch element -- SURVEY.md 8e), so the denoising loop needs NO
collective: shapes are dealt to ranks exactly like the reference's `generate`
(train.py:168: result_index = iter_i * world_size + rank) and each rank samples its own.  The only exchange is
one ragged all-gather of the final latents before VAE decode / metrics.
"""
from __future__ import annotations
import torch
import torch.distributed as dist


def shard_range(num_shapes: int, rank: int, world: int):
    """contiguous block of shapes for `rank` (balanced to within one shape)."""
    base, rem = divmod(num_shapes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def strided_indices(num_shapes: int, rank: int, world: int):
    """the reference's dealing rule, train.py:168."""
    return list(range(rank, num_shapes, world))


def all_gather_latents(x: torch.Tensor, group=None):
    """Ragged all-gather of per-rank latents [n_r, C] -> list of world tensors (on every rank).
    Row counts are exchanged first, payloads are padded to the max and gathered with one collective."""
    if not (dist.is_available() and dist.is_initialized()):
        return [x]
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx, x.shape[1]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:c] for b, c in zip(bufs, counts)]
