"""Run-to-run reproducibility probe of the individual kernels at full size (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from octfusion_b200 import octree_from_splits, DualOctree, ops
from octfusion_b200.synth import synth_splits
from octfusion_b200.modules import GraphConv, DualOctreeGroupNorm, GraphDownsample, GraphUpsample

B = 32
l4, l5 = synth_splits(B, 1000)
doc = DualOctree(octree_from_splits(l4, l5, B, device='cuda'))


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


for d, cin, cout in ((6, 128, 128), (5, 256, 256), (4, 512, 512), (6, 128, 8)):
    n = doc.plan[d].rows
    x = torch.randn((n, cin), device='cuda').bfloat16()
    conv = GraphConv(cin, cout, 7, 7, d - 1).cuda()
    ys = [conv(x, doc, d).clone() for _ in range(4)]
    print('conv d%d %d->%d bit-identical:' % (d, cin, cout), all(torch.equal(ys[0], y) for y in ys[1:]),
          max(rel(y, ys[0]) for y in ys[1:]))
for d, c in ((6, 128), (5, 768)):
    n = doc.plan[d].rows
    x = (torch.randn((n, c), device='cuda') * 2 + 1).bfloat16()
    gn = DualOctreeGroupNorm(c).cuda()
    ys = [gn.run(x, doc.plan[d], B, act=True).clone() for _ in range(4)]
    print('gn d%d C=%d bit-identical:' % (d, c), all(torch.equal(ys[0], y) for y in ys[1:]), max(rel(y, ys[0]) for y in ys[1:]))
x = torch.randn((doc.plan[6].rows, 128), device='cuda').bfloat16()
m = GraphDownsample(128, 128, 7, 7, 4).cuda()
ys = [m(x, doc, 6).clone() for _ in range(3)]
print('down bit-identical:', all(torch.equal(ys[0], y) for y in ys[1:]), max(rel(y, ys[0]) for y in ys[1:]))
x = torch.randn((doc.plan[5].rows, 256), device='cuda').bfloat16()
m = GraphUpsample(256, 256, 7, 7, 5).cuda()
ys = [m(x, doc, 5).clone() for _ in range(3)]
print('up bit-identical:', all(torch.equal(ys[0], y) for y in ys[1:]), max(rel(y, ys[0]) for y in ys[1:]))
qkv = torch.randn((B * 512, 384), device='cuda').bfloat16()
ys = [ops.attention(qkv, B, 512, 4).clone() for _ in range(3)]
print('attention bit-identical:', all(torch.equal(ys[0], y) for y in ys[1:]))
# full forward twice
import bench
from octfusion_b200 import graph_unet_union
from tests.util import UNCOND
net = bench.randomise_(graph_unet_union.UNet3DModel('hr', **UNCOND), 0).cuda().eval()
x = torch.randn((doc.total_num, 3), device='cuda').bfloat16()
ts = torch.full((B,), 1.5, device='cuda')
ys = [net(unet_type='hr', x=x, doctree=doc, timesteps=ts, unet_lr=net.unet_lr, label=None).clone() for _ in range(3)]
print('forward rel diffs:', [rel(y, ys[0]) for y in ys[1:]], 'max abs', [float((y - ys[0]).abs().max()) for y in ys[1:]], 'out absmax', float(ys[0].abs().max()))

tr = []
for i in range(2):
    sink = []
    ops.set_trace(sink)
    net(unet_type='hr', x=x, doctree=doc, timesteps=ts, unet_lr=net.unet_lr, label=None)
    ops.set_trace(None)
    tr.append(sink)
print('ops traced', len(tr[0]))
for k, (a, b) in enumerate(zip(tr[0], tr[1])):
    d = abs(a[2] - b[2]) / max(abs(a[2]), 1e-30)
    if d > 1e-7 or k < 3:
        print(k, a[0], a[1], 'checksum rel diff %.3e' % d, 'value %.4e' % a[2])
