"""TEST INFRASTRUCTURE.  Build the shim Octree the way stage 1 hands it to stage 2
(reference utils/util_dualoctree.py:225-250 `split2octree_small`, ldm_diffusion_util.py:318-325
`create_full_octree`): full layers 0..full_depth, split at full_depth and full_depth+1."""
import torch
from .ref_import import ensure_shim


def octree_from_splits(label_fd, label_fd1, batch_size, full_depth=4, device='cpu'):
    ensure_shim()
    from ocnn.octree import Octree
    depth = full_depth + 2
    oct_ = Octree(depth, full_depth, batch_size, device)
    for d in range(full_depth + 1):
        oct_.octree_grow_full(d)
    oct_.depth = full_depth
    oct_.octree_split(label_fd, full_depth)
    oct_.octree_grow(full_depth + 1)
    oct_.depth += 1
    oct_.octree_split(label_fd1, full_depth + 1)
    oct_.octree_grow(full_depth + 2)
    oct_.depth += 1
    return oct_
