#!/usr/bin/env python
"""Benchmark of the OctFusion denoising U-Net hot path (BASELINE.json: "U-Net denoise steps/sec
(depth-6, B=32)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: the stage-2 ("hr") U-Net forward on B=32 synthetic
depth-6 ShapeNet-shaped octrees (uncond yaml: model_channels [64,128], channel_mult [[1,2,4],[1,2,4]],
num_res_blocks [[1,1,1],[1,1,0]], 4 heads, LR middle block included) + the eps-DDIM update, exactly what
reference sample_loop does per time step (models/octfusion_model_union.py:313-350).  bf16 activations,
fp32 accumulation / statistics / latent.  Random-init weights (no checkpoint is reachable), synthetic
octrees (no dataset is reachable).

One JSON line on stdout (rank 0): value = whole-job steps/s with inputs resident in HBM (CUDA-graph
replay, CUDA-event timed, max over ranks); e2e = the same metric through the public module API with the
latent coming from pinned host memory and the result going back every step; roofline = the tcgen05 tap-gather
GEMM aggregated over all its launches of one step; cpu_baseline = the oracle port of the reference on the
host cores (bounded sample).  Multi-GPU: each rank owns its own 32 shapes (weak scaling, no data-path
collective -- SURVEY.md 8e); one ragged all-gather of the final latents after the timed region.
"""
from __future__ import annotations
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNCOND = dict(
    image_size=[16, 64], input_depth=[4, 6], unet_type=['lr', 'hr'], df_type=['x0', 'eps'], full_depth=4,
    input_channels=[8, 3], out_channels=[8, 3], model_channels=[64, 128], num_res_blocks=[[1, 1, 1], [1, 1, 0]],
    attention_resolutions=[2, 4], channel_mult=[[1, 2, 4], [1, 2, 4]], num_heads=4, use_checkpoint=False, dims=3)
METRIC = 'unet_denoise_steps_per_sec_depth6_B32'
UNIT = 'steps/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='shapes per GPU per step')
    ap.add_argument('--code-channels', type=int, default=8,
                    help='feature channels of the latent (BASELINE.json: 8; the shipped snet yaml uses 3)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    return ap.parse_args()


def config_for(args):
    cfg = dict(UNCOND)
    cfg['input_channels'] = [8, args.code_channels]
    cfg['out_channels'] = [8, args.code_channels]
    return cfg


def randomise_(net, seed):
    """default init + N(0, 1/fan_in) for the reference's zero-initialised tensors (modules.py:719,525,499;
    graph_unet_hr.py:209), which would otherwise make the whole net output 0 and let kernels skip nothing
    but also prove nothing."""
    g = torch.Generator().manual_seed(seed)
    for _, p in net.named_parameters():
        if float(p.detach().abs().max()) == 0.0 and p.dim() > 1:
            fan = p[..., 0].numel() if p.dim() == 2 and p.shape[0] > p.shape[1] else p[0].numel()
            p.data.copy_(torch.randn(p.shape, generator=g) / max(fan, 1) ** 0.5)
    return net


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for t, line in self.samples:
            f = [s.strip() for s in line.split(',')]
            if len(f) < 6 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_reference_steps_per_sec(cfg, batch_full, sample_batch=1, timed=1):
    from oracle import restate as R
    from oracle.octree_util import octree_from_splits
    from octfusion_b200.synth import synth_splits
    from octfusion_b200 import graph_unet_union
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.device('meta'):
        shapes = {k: tuple(v.shape) for k, v in graph_unet_union.UNet3DModel('hr', **cfg).state_dict().items()}
    sd = R.seeded_state_dict(shapes, 0)
    l4, l5 = synth_splits(sample_batch, 0)
    dg = R.DualGraph(octree_from_splits(l4, l5, sample_batch))
    lr_cfg, hr_cfg = R.split_cfg(cfg)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(dg.total_num, hr_cfg['in_channels'], generator=g)
    ts = torch.full((sample_batch,), 1.5)
    ls, lsn = torch.tensor(1.5), torch.tensor(1.9)

    def step():
        eps = R.hr_forward(x, dg, ts, sd, hr_cfg, lr_cfg)
        return R.ddim_eps_update(x, eps, ls, lsn)

    with torch.no_grad():
        step()
        t0 = time.perf_counter()
        for _ in range(timed):
            step()
        dt = (time.perf_counter() - t0) / timed
    value = 1.0 / (dt * batch_full / sample_batch)
    return {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle port (torch fp32, %d threads) of the reference U-Net step on B=%d of the %d shapes, '
                      '1 warm-up + %d timed, %.2f s/step, scaled linearly to B=%d' %
                      (cores, sample_batch, batch_full, timed, dt, batch_full)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = config_for(args)
    k = max(1, min(args.steps, 3))
    base = cpu_reference_steps_per_sec(cfg, args.batch, 1, k)
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 / base['value'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, None), 'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(args, nodes):
    c = {'workload': 'OctFusion stage-2 (hr) denoising U-Net step: uncond snet yaml, depth-6 synthetic ShapeNet-shaped '
                     'octrees, full_depth 4, LR middle block + attention included, eps-DDIM update',
         'batch_per_gpu': args.batch, 'code_channels': args.code_channels, 'parallelism': 'dp%d (batch shard, no '
         'collective in the step loop)' % args.gpus,
         'l2': 'no flush: one step streams several GB of activations (>> 126 MB L2)'}
    if nodes:
        c['nodes_per_gpu'] = nodes
    return c


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    # exactly ONE line may reach stdout (the JSON): NCCL / torch banners are sent to stderr for the whole run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device: octfusion_b200 has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from octfusion_b200 import graph_unet_union, octree_from_splits, DualOctree, ops, _lib, shard
    from octfusion_b200.synth import synth_splits
    from octfusion_b200.sampler import HRStepper, sampling_log_snr

    cfg = config_for(args)
    act = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    net = randomise_(graph_unet_union.UNet3DModel('hr', **cfg), 0).to(dev).eval()
    l4, l5 = synth_splits(args.batch, seed=0)         # same 32 shapes on every rank: identical per-GPU work (weak scaling)
    doc = DualOctree(octree_from_splits(l4, l5, args.batch, device=dev))
    nodes = {d: doc.plan[d].rows for d in range(4, 7)}
    n6, cc = doc.total_num, args.code_channels
    total_steps = args.warmup + args.steps
    ls = sampling_log_snr(max(total_steps, 50))
    g = torch.Generator(device=dev).manual_seed(rank)
    noise = torch.randn((n6, cc), generator=g, device=dev)

    st = HRStepper(net.unet_hr, net.unet_lr, doc, act, None, use_cuda_graph=True)
    st.set_latent(noise)
    for i in range(args.warmup):                                   # includes the eager pass + graph capture
        st.step(ls[i], ls[i + 1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for i in range(args.warmup, total_steps):
        st.step(ls[i], ls[i + 1])
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clk = clocks.stop(t0, t1) if rank == 0 else None
    value = world * args.steps / (ms_total / 1000.0)
    kernels_per_step = int(st.kernels_per_step)

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region -------------------------
    x_host = torch.randn((n6, cc)).pin_memory()
    y_host = torch.empty((n6, cc)).pin_memory()
    ts_host = torch.empty(args.batch).pin_memory()
    ls_dev = torch.zeros(1, device=dev); lsn_dev = torch.zeros(1, device=dev)

    def e2e_step(i):
        ts_host.fill_(ls[i])
        x = x_host.to(dev, non_blocking=True)
        ts = ts_host.to(dev, non_blocking=True)
        xin = x if act == torch.float32 else x.to(act)
        eps = net(unet_type='hr', x=xin, doctree=doc, timesteps=ts, unet_lr=net.unet_lr, label=None)
        ls_dev.fill_(ls[i]); lsn_dev.fill_(ls[i + 1])
        ops.ddim_eps_update(x, eps, ls_dev, lsn_dev)
        y_host.copy_(x, non_blocking=True)
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k2 = max(3, min(args.steps, 10))
    e0.record()
    for i in range(k2):
        e2e_step(i)
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e = {'value': world * k2 / (float(ms2.item()) / 1000.0), 'unit': UNIT,
           'h2d_bytes_per_step': n6 * cc * 4 + args.batch * 4, 'd2h_bytes_per_step': n6 * cc * 4,
           'steps': k2, 'path': 'graph_unet_union.UNet3DModel.forward + ops.ddim_eps_update, eager launches, '
                                'pinned host latent in, result out'}

    # ---- roofline of the dominant kernel (per-launch CUDA events, eager pass) -----------------------
    roofline, per_layer = None, None
    if rank == 0 and not args.no_roofline and act == torch.bfloat16:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:  # noqa: BLE001
            pass
        peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else '1400 TF/s (of fallback)'
        agg = {}
        reps = 3
        for rep in range(reps + 1):
            sink = []
            ops.set_profile(sink)
            st.use_cuda_graph = False
            st.step(ls[0], ls[1])
            ops.set_profile(None)
            torch.cuda.synchronize()
            if rep == 0:
                continue
            for r in sink:
                if r['kind'] != 'tc':
                    continue
                key = (r['M'], r['K'], r['N'], r['taps'])
                a = agg.setdefault(key, {'ms': 0.0, 'flops': 0.0, 'bytes': 0.0, 'launches': 0})
                a['ms'] += r['start'].elapsed_time(r['end']); a['flops'] += r['flops']; a['bytes'] += r['bytes']
                a['launches'] += 1
        st.use_cuda_graph = True
        tot_ms = sum(a['ms'] for a in agg.values()); tot_fl = sum(a['flops'] for a in agg.values())
        tot_by = sum(a['bytes'] for a in agg.values())
        if tot_ms > 0:
            ach = tot_fl / (tot_ms * 1e-3) / 1e12
            traffic = None
            try:        # DRAM bytes of the same launches from the committed ncu capture (profiles/tc_traffic_r01.json)
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'tc_traffic_r01.json')))
                traffic = tj['dram_bytes_per_step'] / tj['launches_per_step']
            except Exception:  # noqa: BLE001
                pass
            nl = sum(a['launches'] for a in agg.values()) / reps
            roofline = {'kernel': 'gather_gemm_tc_kernel (all launches of one step)', 'bound': 'tensor',
                        'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': traffic,
                        'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read+write, mean over the step)',
                        'launches_per_step': nl, 'algorithmic_bytes_per_launch': tot_by / reps / max(nl, 1),
                        'peak_source': peak_src, 'ms_per_step_in_kernel': tot_ms / reps,
                        'algorithmic_gflop_per_step': tot_fl / reps / 1e9, 'algorithmic_gb_per_step': tot_by / reps / 1e9}
            per_layer = []
            for (m, k, n, taps), a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
                per_layer.append({'M': m, 'K': k, 'N': n, 'taps': taps, 'launches_per_step': a['launches'] // reps,
                                  'us': 1000.0 * a['ms'] / a['launches'],
                                  'tflops': a['flops'] / (a['ms'] * 1e-3) / 1e12,
                                  'gbps_algorithmic': a['bytes'] / (a['ms'] * 1e-3) / 1e9})

    # ---- the one collective of the path: ragged all-gather of the final latents ---------------------
    gathered = shard.all_gather_latents(st.x)
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu = cpu_reference_steps_per_sec(cfg, args.batch, 1, 1)
        except Exception as e:  # noqa: BLE001
            cpu = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'failed: %r' % (e,)}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic', 'config': workload_config(args, nodes), 'clocks': clk,
            'e2e': e2e, 'gpu_launches': kernels_per_step * args.steps, 'kernels_per_step': kernels_per_step,
            'roofline': roofline, 'cpu_baseline': cpu, 'gathered_latent_rows': [int(t.shape[0]) for t in gathered]}
    if per_layer:
        line['roofline_per_layer'] = per_layer[:12]
    os.write(real_stdout, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
