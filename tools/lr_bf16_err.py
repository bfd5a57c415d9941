"""bf16-vs-fp32 error of the dense LR U-Net used as the middle block (product only, no oracle): diagnostic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import UNCOND, product_doctree, relerr
from octfusion_b200 import graph_unet_union
import bench
net = bench.randomise_(graph_unet_union.UNet3DModel('hr', **UNCOND), 0).cuda().eval()
doc = product_doctree(2, 0)
for seed in range(3):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn((2 * 4096, 64), generator=g).cuda()
    ts = torch.tensor([1.5, -0.5]).cuda()
    ref = net.unet_lr.forward_as_middle(h, doc, ts, None, None).float()
    errs = []
    for rep in range(3):
        y = net.unet_lr.forward_as_middle(h.bfloat16(), doc, ts, None, None).float()
        errs.append(relerr(y, ref))
    print('seed %d  bf16-vs-fp32 relerr %s' % (seed, ' '.join('%.4f' % e for e in errs)))
