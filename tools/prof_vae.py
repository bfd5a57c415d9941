"""GraphVAE at depth 8 (BASELINE.json configs[3]): encoder network (`extract_code` without the
point-cloud feature extraction: input features are given) + decoder `decode_code(update_octree=False)` on a synthetic
depth-8 octree -- the sparse-conv-only, HBM-bound regime (24..32 channels at 16x the depth-6 node count).
`build_case` is the workload of `bench.py --workload vae`; run as a script it prints the timing.
usage: python tools/prof_vae.py ; env BATCH (8), REPS (3)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

# reference configs/vae_snet_train.yaml:6-22
VAE = dict(depth=8, channel_in=4, nout=4, full_depth=4, depth_stop=6, depth_out=8, resblk_type='basic', bottleneck=4,
           resblk_num=2, code_channel=16, embed_dim=3)
HALFWIDTH = {4: 1.55, 5: 1.1, 6: 0.9, 7: 0.8}      # shell half-width in cells per level (4, 5: octfusion_b200/synth.py)


def deep_octree(batch, seed, depth, device):
    """the synthetic ellipsoid-shell shapes of octfusion_b200/synth.py refined to `depth` (a closed surface: the node
    count grows ~4x per level, like a ShapeNet mesh), built level by level on the device."""
    from octfusion_b200.octree import create_full_octree
    g = torch.Generator().manual_seed(seed)
    u = torch.stack([torch.rand(8, generator=g, dtype=torch.float64) for _ in range(batch)]).to(device)
    centre = (u[:, 0:3] * 2 - 1) * 0.15
    axes = (0.6 + 0.8 * u[:, 4:7]) * (0.45 + 0.30 * u[:, 3:4])
    octree = create_full_octree(depth, 4, batch, device)
    for d in range(4, depth):
        x, y, z, b = octree.xyzb(d)
        h = 2.0 / (1 << d)
        p = (torch.stack([x, y, z], 1).double() + 0.5) * h - 1.0
        f = torch.sqrt((((p - centre[b]) / axes[b]) ** 2).sum(1))
        lab = ((f - 1.0).abs() * axes.mean(1)[b] < HALFWIDTH[d] * h)
        octree.octree_split(lab.long(), d)
        octree.octree_grow(d + 1)
        octree.depth += 1
    return octree


def build_case(batch, device, dtype):
    import bench
    from octfusion_b200 import DualOctree, ops, graph_vae
    net = bench.randomise_(graph_vae.GraphVAE(**VAE), 3).to(device).eval()
    doc = DualOctree(deep_octree(batch, 0, 8, device))
    rows = {d: doc.plan[d].rows for d in range(4, 9)}
    g = torch.Generator(device=device).manual_seed(1)
    data = torch.randn((rows[8], 4), generator=g, device=device).to(dtype)

    def step():
        feat = net.octree_encoder_step(data, doc)              # [rows6, 32] -> posterior moments -> latent code
        code = net.KL_conv(feat)[:, :3].contiguous()
        return net.decode_code(code, doc, update_octree=False)

    step()
    sink = []
    ops.set_profile(sink)
    step()
    ops.set_profile(None)
    torch.cuda.synchronize()
    nbytes = sum(r['bytes'] for r in sink)
    return {'step': step, 'algorithmic_bytes': float(nbytes), 'nodes': rows, 'batch': batch,
            'note': 'synthetic ellipsoid-shell octrees refined to depth 8; encoder on random [N8, 4] input features, mean of '
                    'the posterior as the code, decode_code(update_octree=False); algorithmic bytes = sum over the %d GEMM / '
                    'norm calls of operands read once + results written once' % len(sink)}


if __name__ == '__main__':
    B = int(os.environ.get('BATCH', 8))
    reps = int(os.environ.get('REPS', 3))
    for dtype in (torch.bfloat16, torch.float32):
        case = build_case(B, torch.device('cuda'), dtype)
        ts = []
        for r in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            case['step']()
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print('%s B=%d nodes %s: encode+decode %.2f ms (best of %d), %.1f GB algorithmic -> %.0f GB/s'
              % (str(dtype).split('.')[-1], B, case['nodes'], min(ts) * 1e3, reps, case['algorithmic_bytes'] / 1e9,
                 case['algorithmic_bytes'] / min(ts) / 1e9))
