"""-m gpu parity of the NeuralMPU kernel against the oracle, on the grown depth-8 octree of the
VAE fixture with random per-node regression values."""
import os
import numpy as np
import pytest
import torch

from oracle import restate as R
from tests import util as U
from tests.util import relerr, GOLDEN

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_neural_mpu_matches_oracle():
    from octfusion_b200.mpu import NeuralMPU
    from octfusion_b200 import graph_vae
    g = np.load(os.path.join(GOLDEN, 'vae_decode.npz'))
    labels = {d: torch.from_numpy(np.unpackbits(g['label%d' % d])[: int(g['nnum'][d])].astype(np.int64)) for d in (6, 7, 8)}
    octree = U.oracle_grown_octree(labels)
    # the same octree on the device
    net = graph_vae.GraphVAE(**U.VAE)
    doc6 = U.product_doctree(1, 0)
    po = net.create_child_octree(doc6.octree)
    for d in (6, 7, 8):
        po.octree_split(labels[d].to(DEV), d)
        if d < 8:
            po.octree_grow(d + 1)
            po.depth += 1
    gen = torch.Generator().manual_seed(21)
    reg = {d: torch.randn(int(octree.nnum[4:d + 1].sum()), 4, generator=gen) for d in (6, 7, 8)}
    x, y, z, b = octree.xyzb(8)
    pick = torch.randperm(x.numel(), generator=gen)[:20000]
    near = (torch.stack([x, y, z], 1)[pick].float() + torch.rand(20000, 3, generator=gen)) / 128.0 - 1.0
    uni = torch.rand(20000, 3, generator=gen) * 2 - 1
    edge = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [-1.0, 1.0, 0.999]])
    pts = torch.cat([near, uni, edge])
    pos = torch.cat([pts, torch.zeros(pts.shape[0], 1)], 1)
    want = R.mpu_eval(pos, reg, octree, 4, 6, 8)
    got = NeuralMPU(4, 6, 8)(pos.to(DEV), {d: v.to(DEV) for d, v in reg.items()}, po)
    for d in (6, 7, 8):
        assert torch.equal(got[d][1].cpu(), want[d][1]), d
        assert relerr(got[d][0].cpu(), want[d][0]) < 1e-4, d
