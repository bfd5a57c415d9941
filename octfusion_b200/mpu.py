"""NeuralMPU: evaluates the implicit function defined by the GraphVAE decoder's per-node regression values at
arbitrary query points.  Drop-in for reference models/networks/dualoctree_networks/mpu.py
`NeuralMPU`: same constructor and call signature, returns {depth: (fval [P], flgs [P] bool)}.  One CUDA kernel per
target depth (csrc/mpu.cu) instead of the reference's key search + two sparse-matrix products per depth."""
from __future__ import annotations
import ctypes as C
import torch

from ._lib import lib, ptr, stream, check, OctreeLevels, require_cuda


def _levels(octree, depth):
    lv = OctreeLevels()
    lv.full_depth, lv.depth, lv.batch = octree.full_depth, depth, octree.batch_size
    keep = []
    for d in range(octree.full_depth, depth + 1):
        ch = octree.children[d].contiguous()
        assert ch.dtype == torch.int32 and ch.is_cuda
        lv.children[d] = ch.data_ptr()
        lv.nnum[d] = int(octree.nnum[d])
        keep.append(ch)
    return lv, keep


class NeuralMPU:
    def __init__(self, full_depth, depth_stop, depth):
        self.full_depth, self.depth_stop, self.depth = full_depth, depth_stop, depth

    @torch.no_grad()
    def __call__(self, pos, reg_voxs, octree_out):
        require_cuda(pos)
        pos = pos.float().contiguous()
        assert pos.dim() == 2 and pos.shape[1] == 4
        n = pos.shape[0]
        mpus = {}
        for d in range(self.depth_stop, self.depth + 1):
            reg = reg_voxs[d].float().contiguous()
            want = int(octree_out.nnum[self.full_depth:d + 1].sum())
            assert reg.shape == (want, 4), 'reg_voxs[%d] must hold one row per octree node of depths %d..%d' % (
                d, self.full_depth, d)
            lv, keep = _levels(octree_out, d)
            fval = torch.empty(n, dtype=torch.float32, device=pos.device)
            hit = torch.empty(n, dtype=torch.uint8, device=pos.device)
            check(lib.of_mpu_eval(C.byref(lv), d, ptr(pos), n, ptr(reg), ptr(fval), ptr(hit), stream()), 'of_mpu_eval')
            del keep
            mpus[d] = (fval, hit.bool())
        return mpus
