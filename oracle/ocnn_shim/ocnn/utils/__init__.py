"""Restated `ocnn.utils` subset (test infrastructure)."""
import torch


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    """Same contract as the reference's local scatter.py:24-39 (modules.py:17 imports this one)."""
    if dim < 0:
        dim = src.dim() + dim
    idx = index
    if idx.dim() == 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        idx = idx.view(shape)
    idx = idx.expand_as(src)
    if out is None:
        size = list(src.shape)
        size[dim] = dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, idx, src)


def cumsum(data, dim, exclusive=False):
    """exclusive => leading zero, length n+1 (dual_octree.py:30,190 index ncum[d] for d<=depth)."""
    out = torch.cumsum(data, dim)
    if exclusive:
        size = list(data.shape)
        size[dim] = 1
        out = torch.cat([torch.zeros(size, dtype=out.dtype, device=out.device), out], dim)
    return out
