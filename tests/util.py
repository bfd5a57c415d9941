"""Shared helpers of the parity tests: seeded inputs built once for the oracle (CPU) and the product
(CUDA) from the same synthetic split labels."""
from __future__ import annotations
import functools
import os
import torch

from octfusion_b200.synth import synth_splits
from oracle import restate as R
from oracle.octree_util import octree_from_splits as oracle_octree

UNCOND = dict(
    image_size=[16, 64], input_depth=[4, 6], unet_type=['lr', 'hr'], df_type=['x0', 'eps'], full_depth=4,
    input_channels=[8, 3], out_channels=[8, 3], model_channels=[64, 128], num_res_blocks=[[1, 1, 1], [1, 1, 0]],
    attention_resolutions=[2, 4], channel_mult=[[1, 2, 4], [1, 2, 4]], num_heads=4, use_checkpoint=False, dims=3)
COND = dict(UNCOND, num_res_blocks=[[1, 1, 1], [2, 2, 0]], attention_resolutions=[2, 4, 8],
            channel_mult=[[1, 2, 4, 8], [1, 2, 4]], num_classes=5)
# a narrow net with the same topology: fast enough for CPU-side tests and golden fixtures
SMALL = dict(UNCOND, model_channels=[64, 64], channel_mult=[[1, 2], [1, 1, 2]], attention_resolutions=[2])


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@functools.lru_cache(maxsize=8)
def oracle_doctree(batch, seed=0):
    l4, l5 = synth_splits(batch, seed)
    oct_ = oracle_octree(l4, l5, batch)
    return R.DualGraph(oct_), (l4, l5)


def product_doctree(batch, seed=0, device='cuda'):
    from octfusion_b200 import octree_from_splits, DualOctree
    l4, l5 = synth_splits(batch, seed)
    return DualOctree(octree_from_splits(l4, l5, batch, device=device))


def model_shapes(cfg, stage='hr'):
    """state_dict shapes of the product model (== the reference's; checked by test_state_dict_parity)."""
    from octfusion_b200 import graph_unet_union
    with torch.device('meta'):
        net = graph_unet_union.UNet3DModel(stage, **cfg)
    return {k: tuple(v.shape) for k, v in net.state_dict().items()}


def build_product(cfg, sd, device='cuda', stage='hr'):
    from octfusion_b200 import graph_unet_union
    net = graph_unet_union.UNet3DModel(stage, **cfg)
    net.load_state_dict(sd)
    return net.to(device).eval()


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
