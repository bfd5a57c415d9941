"""calc_sdf on the 256^3 grid (reference utils/util_dualoctree.py:99-118) through of_mpu_eval_grid: points/s and the
bytes the kernel touches.  Octree: synthetic ellipsoid-shell shapes refined to depth 8 (tools/prof_vae.deep_octree), random
per-node regression values (the NeuralMPU arithmetic does not depend on their values).
usage: python tools/prof_sdf.py ; env BATCH (2), SIZE (256)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.prof_vae import deep_octree
from octfusion_b200.mpu import NeuralMPU, calc_sdf

B = int(os.environ.get('BATCH', 2))
size = int(os.environ.get('SIZE', 256))
octree = deep_octree(B, 0, 8, 'cuda')
g = torch.Generator(device='cuda').manual_seed(1)
reg = {d: torch.randn((int(octree.nnum[4:d + 1].sum()), 4), generator=g, device='cuda') for d in (6, 7, 8)}
mpu = NeuralMPU(4, 6, 8)


def model(pos):
    return mpu(pos, reg, octree)[8][0]


for fast in (True, False):
    if fast:
        model.mpu_args = (mpu, reg, octree)
    elif hasattr(model, 'mpu_args'):
        del model.mpu_args
    calc_sdf(model, 1, 64, 64 ** 3, -0.9, 0.9)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sdf = calc_sdf(model, B, size, 64 ** 3, -0.9, 0.9)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pts = B * size ** 3
    print('%s: %d shapes x %d^3 = %.1f M points in %.1f ms -> %.2f G points/s (finest depth 8: %d nodes; inside-shell fraction %.3f)'
          % ('grid kernel (coordinates generated in the kernel)' if fast else 'explicit point tensors (reference call pattern)',
             B, size, pts / 1e6, dt * 1e3, pts / dt / 1e9, int(octree.nnum[8]), float((sdf.abs() > 0).float().mean())))
