"""Pipeline timeline of ONE CTA of the tcgen05 tap-gather GEMM (of_tc_trace_set): where each warp role waits.
usage: python tools/trace_tc.py "6,128,128;6,256,256" [block]      (depth, cin, cout per layer; env BATCH)
Regions (clock64 stamps, include/octfusion_b200.h): 0 MMA warp (per stage: loop top, ready, MMAs issued, committed), 1 weight loader (slot free, issued),
2..5 producer groups (loop top, slot free, issued), 6 epilogue warp 0 (accumulator full, drained), 7 scout (stage seen
full).  Each stamp costs the stamping warp ~60-130 cycles (clock read + global store) and the SM clock under tensor load is
~1.5 GHz, not the 1.965 GHz nvidia-smi shows: compare shares and orderings, not absolute cycles; use prof_conv.py for times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from octfusion_b200 import octree_from_splits, DualOctree
from octfusion_b200._lib import lib, ptr, check
from octfusion_b200.synth import synth_splits
from octfusion_b200.modules import GraphConv

B = int(os.environ.get('BATCH', 32))
CAP = 8192
shapes = [tuple(int(v) for v in sh.split(',')) for sh in (sys.argv[1] if len(sys.argv) > 1 else '6,128,128').split(';')]
block = int(sys.argv[2]) if len(sys.argv) > 2 else 5
l4, l5 = synth_splits(B, 0)
doc = DualOctree(octree_from_splits(l4, l5, B, device='cuda'))


def stat(name, v):
    v = np.asarray(v, dtype=np.float64)
    if v.size == 0:
        return
    print('    %-34s n=%5d  mean %7.0f  p50 %7.0f  p90 %7.0f  max %7.0f  sum %9.0f' %
          (name, v.size, v.mean(), np.percentile(v, 50), np.percentile(v, 90), v.max(), v.sum()))


for d, cin, cout in shapes:
    n = doc.plan[d].rows
    x = torch.randn((n, cin), device='cuda').bfloat16()
    conv = GraphConv(cin, cout, 7, 7, d - 1).cuda()
    if os.environ.get('EPI'):
        epi = {}
        pl = doc.plan[d]
        if 'stats' in os.environ['EPI']: epi['stats'] = pl.stat
        if 'emb' in os.environ['EPI']: epi.update(row_add=torch.randn((B, cout), device='cuda'), row_add_idx=pl.batch_id)
        if 'resid' in os.environ['EPI']: epi['resid'] = torch.randn((n, cout), device='cuda').bfloat16()
        conv_ = conv
        conv = lambda x, doc, d: conv_.run(x, doc.plan[d], **epi)
    for _ in range(2):
        conv(x, doc, d)
    torch.cuda.synchronize()
    buf = torch.zeros((8, CAP), dtype=torch.int64, device='cuda')
    check(lib.of_tc_trace_set(ptr(buf), CAP, block), 'trace')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); conv(x, doc, d); e1.record()
    torch.cuda.synchronize()
    check(lib.of_tc_trace_set(None, 0, 0), 'trace')
    t = buf.cpu().numpy()
    k = 7 * (cin + d - 1)
    print('== depth %d rows %d K %d N %d: %.1f us (traced launch), block %d' % (d, n, k, cout, e0.elapsed_time(e1) * 1e3, block))
    m4 = t[0][t[0] > 0]; m4 = m4[: len(m4) // 4 * 4].reshape(-1, 4); m = m4[:, 1:]
    if len(m) > 2:
        print('  MMA warp: %d stages, span %d clk' % (len(m), m[-1, 2] - m[0, 0]))
        stat('stage ready (from prev commit)', m[1:, 0] - m[:-1, 2])
        stat('  commit -> loop top', m4[1:, 0] - m4[:-1, 3])
        stat('  loop top -> flag seen + elected', m4[:, 1] - m4[:, 0])
        stat('issue MMAs', m[:, 1] - m[:, 0])
        stat('commits', m[:, 2] - m[:, 1])
        stat('stage period', m[1:, 2] - m[:-1, 2])
    w = t[1][t[1] > 0]; w = w[: len(w) // 2 * 2].reshape(-1, 2)
    if len(w) > 2:
        print('  weight loader:')
        stat('wait slot free (from prev issue)', w[1:, 0] - w[:-1, 1])
        stat('issue', w[:, 1] - w[:, 0])
    for gq in range(4):
        p = t[2 + gq][t[2 + gq] > 0]; p = p[: len(p) // 3 * 3].reshape(-1, 3)
        if len(p) > 2 and gq in (0, 3):
            print('  producer group %d: %d slots' % (gq, len(p)))
            stat('table fetch + bookkeeping', p[:, 1] * 0 + (p[:, 1] - p[:, 0]))
            stat('issue 16 cp.async + arrive', p[:, 2] - p[:, 1])
            stat('slot period', p[1:, 0] - p[:-1, 0])
    # cross-role latencies of the unified ring (4 stages; 256-wide: 1 sub-tile per stage, 128-wide: 2): how long after the
    # MMA warp's commit of a stage its producers see the slot free, and how long after the producers' issue the MMA warp
    # sees the refilled stage ready
    if len(m) > 8 and cout % 128 == 0 and os.environ.get('OCTFUSION_TC_UNI', '1') == '1':
        subs, nst = (1 if cout % 256 == 0 else 2), 4
        free_lat, fill_lat = [], {}
        for gq in range(4):
            p = t[2 + gq][t[2 + gq] > 0]; p = p[: len(p) // 3 * 3].reshape(-1, 3)
            for j in range(len(p)):
                k = (gq + 4 * j) // subs
                if nst <= k < len(m):
                    free_lat.append(p[j, 1] - m[k - nst, 2])
                    fill_lat[k] = max(fill_lat.get(k, -10 ** 9), m[k, 0] - p[j, 2])
        sc = t[7]
        det = [sc[k] - (m[k, 0] - fill_lat[k]) for k in fill_lat if k < CAP and sc[k] > 0]
        stat('producer issued -> scout sees stage full', det)
        stat('scout sees stage full -> MMA warp ready', [m[k, 0] - sc[k] for k in fill_lat if k < CAP and sc[k] > 0])
        stat('commit(stage k) -> producer sees slot free', free_lat)
        stat('producer issued -> MMA warp sees stage ready', list(fill_lat.values()))
        w2 = t[1][t[1] > 0]; w2 = w2[: len(w2) // 2 * 2].reshape(-1, 2)
        n2 = min(len(w2), len(m))
        stat('commit(stage k) -> loader sees slot free', [w2[k, 0] - m[k - nst, 2] for k in range(nst, n2)])
        stat('loader issued -> MMA warp sees stage ready', [m[k, 0] - w2[k, 1] for k in range(nst, n2)])
    e = t[6][t[6] > 0]; e = e[: len(e) // 2 * 2].reshape(-1, 2)
    if len(e) > 2:
        print('  epilogue warp 0: %d tiles' % len(e))
        stat('drain one CTA tile', e[:, 1] - e[:, 0])
        stat('wait accumulator (from prev drain)', e[1:, 0] - e[:-1, 1])
        stat('tile period', e[1:, 0] - e[:-1, 0])
