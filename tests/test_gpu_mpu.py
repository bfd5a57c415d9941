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


def test_calc_sdf_grid_matches_oracle_and_generic_path():
    """calc_sdf (reference utils/util_dualoctree.py:99-118) on a 64^3 grid of the golden depth-8 octree: the in-kernel
    grid generator (of_mpu_eval_grid) equals (i) the explicit-point path bit for bit and (ii) the oracle NeuralMPU on
    the reference's own grid expression."""
    from octfusion_b200.mpu import NeuralMPU, calc_sdf, get_mgrid
    from octfusion_b200 import graph_vae
    g = np.load(os.path.join(GOLDEN, 'vae_decode.npz'))
    labels = {d: torch.from_numpy(np.unpackbits(g['label%d' % d])[: int(g['nnum'][d])].astype(np.int64)) for d in (6, 7, 8)}
    octree = U.oracle_grown_octree(labels)
    net = graph_vae.GraphVAE(**U.VAE)
    po = net.create_child_octree(U.product_doctree(1, 0).octree)
    for d in (6, 7, 8):
        po.octree_split(labels[d].to(DEV), d)
        if d < 8:
            po.octree_grow(d + 1)
            po.depth += 1
    gen = torch.Generator().manual_seed(33)
    reg = {d: torch.randn(int(octree.nnum[4:d + 1].sum()), 4, generator=gen) for d in (6, 7, 8)}
    regd = {d: v.to(DEV) for d, v in reg.items()}
    mpu = NeuralMPU(4, 6, 8)

    def model(pos):
        return mpu(pos, regd, po)[8][0]
    size, bbmin, bbmax = 64, -0.9, 0.9
    generic = calc_sdf(model, 1, size, 50000, bbmin, bbmax)
    model.mpu_args = (mpu, regd, po)
    fast = calc_sdf(model, 1, size, 50000, bbmin, bbmax)
    assert fast.shape == (1, size, size, size) and torch.equal(fast, generic)
    # the reference's grid expression (numpy float32 arithmetic) on the CPU, through the oracle
    samples = np.stack(np.meshgrid(*([np.arange(0, size, dtype=np.float32)] * 3), indexing='ij'), -1).reshape(-1, 3)
    samples = torch.from_numpy(samples * ((bbmax - bbmin) / size) + bbmin)
    assert torch.equal((get_mgrid(size, 3, DEV) * ((bbmax - bbmin) / size) + bbmin).cpu(), samples)
    pos = torch.cat([samples, torch.zeros(samples.shape[0], 1)], 1)
    want = R.mpu_eval(pos, reg, octree, 4, 6, 8)[8][0]
    assert relerr(fast.reshape(-1).cpu(), want) < 1e-4
