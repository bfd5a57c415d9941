"""Element-wise run-to-run divergence per op (B=4), debug aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from octfusion_b200 import octree_from_splits, DualOctree, ops, graph_unet_union
from octfusion_b200.synth import synth_splits
from tests.util import UNCOND
import bench
B = 4
l4, l5 = synth_splits(B, 1000)
doc = DualOctree(octree_from_splits(l4, l5, B, device='cuda'))
net = bench.randomise_(graph_unet_union.UNet3DModel('hr', **UNCOND), 0).cuda().eval()
x = torch.randn((doc.total_num, 3), device='cuda').bfloat16()
ts = torch.full((B,), 1.5, device='cuda')
ops._TRACE_KEEP = True
tr = []
for i in range(2):
    sink = []
    ops.set_trace(sink)
    net(unet_type='hr', x=x, doctree=doc, timesteps=ts, unet_lr=net.unet_lr, label=None)
    ops.set_trace(None)
    tr.append(sink)
for k, (a, b) in enumerate(zip(tr[0], tr[1])):
    d = float((a[3].float() - b[3].float()).norm() / a[3].float().norm().clamp_min(1e-30))
    mx = float((a[3].float() - b[3].float()).abs().max())
    nz = int((a[3] != b[3]).sum())
    print(k, a[0], a[1], 'rel-l2 %.2e  maxabs %.3e  differing elems %d of %d  absmax %.2f' % (d, mx, nz, a[3].numel(), float(a[3].float().abs().max())))
