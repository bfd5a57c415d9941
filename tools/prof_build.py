"""Stage-1 -> stage-2 handoff timing: split signal [B, 8, 16, 16, 16] on the device -> split2octree_small ->
DualOctree (tap tables, node types, multi-neighbour index, statistics plans) at B=32.  usage: python tools/prof_build.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from octfusion_b200 import octree_from_splits, DualOctree, split2octree_small, octree2split_small
from octfusion_b200.synth import synth_splits

B = int(os.environ.get('BATCH', 32))
l4, l5 = synth_splits(B, 0)
oc = octree_from_splits(l4, l5, B, device='cuda')
split = octree2split_small(oc, 4)                    # the tensor stage 1 hands over (reference util_dualoctree.py:198-211)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o2 = split2octree_small(split, 6, 4)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    doc = DualOctree(o2)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if it:
        print('B=%d: split2octree_small %.2f ms, DualOctree build %.2f ms (rows d4/d5/d6 = %d / %d / %d); amortised over %d denoising steps'
              % (B, (t1 - t0) * 1e3, (t2 - t1) * 1e3, doc.plan[4].rows, doc.plan[5].rows, doc.plan[6].rows, 200))
