"""Restated `ocnn.nn` subset (test infrastructure)."""
import torch


def octree2voxel(data, octree, depth, nempty=False):
    """zeros [B,2^d,2^d,2^d,C]; out[b,x,y,z,:] = data  (reference graph_unet_lr.py:176-181)."""
    x, y, z, b = octree.xyzb(depth, nempty)
    s = 2 ** depth
    out = data.new_zeros(octree.batch_size, s, s, s, data.shape[1])
    out[b, x, y, z] = data
    return out


def octree_pad(data, octree, depth, val=0.0):
    """out = full([nnum[d], C], val); out[nempty_mask] = data (util_dualoctree.py:204,218)."""
    mask = octree.nempty_mask(depth)
    out = torch.full((mask.shape[0], data.shape[1]), val, dtype=data.dtype, device=data.device)
    out[mask] = data
    return out
