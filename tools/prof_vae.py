"""GraphVAE.decode_code(update_octree=True) timing on synthetic depth-6 octrees (SURVEY.md 8d config 4 analogue,
decoder half): latents [N6, 3] -> split logits + MPU values at depths 6..8, octree grown on the device.
usage: python tools/prof_vae.py ; env BATCH (8), REPS (3)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import util as U
from octfusion_b200 import octree_from_splits, DualOctree
from octfusion_b200.synth import synth_splits

B = int(os.environ.get('BATCH', 8))
reps = int(os.environ.get('REPS', 3))
net = U.build_vae(U.vae_state_dict())
l4, l5 = synth_splits(B, 0)
for dtype in (torch.bfloat16, torch.float32):
    ts = []
    for r in range(reps + 1):
        doc = DualOctree(octree_from_splits(l4, l5, B, device='cuda'))
        code = U.vae_code(doc.plan[6].rows).cuda().to(dtype)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = net.decode_code(code, doc, update_octree=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    o = out['octree_out']
    print('%s B=%d: nodes d6/d7/d8 = %d / %d / %d, decode_code %.1f ms (best of %d; includes octree growth + 3 graph builds)'
          % (str(dtype).split('.')[-1], B, int(o.nnum[6]), int(o.nnum[7]), int(o.nnum[8]), min(ts[1:]) * 1e3, reps))
