"""Summarise an ncu launch list of one denoising step (CSV with gpu__time_duration.sum and, optionally,
dram__bytes_read.sum / dram__bytes_write.sum per launch): per-kernel time share and DRAM traffic; optionally write the
tcgen05 GEMM traffic as JSON for bench.py's roofline.traffic.
usage: python tools/launch_summary.py launches.csv [tc_traffic.json]"""
import csv, collections, json, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iid, iu, im = (hdr.index(c) for c in ('Kernel Name', 'Metric Value', 'ID', 'Metric Unit', 'Metric Name'))
launch = collections.OrderedDict()          # id -> [name, us, dram bytes]
for x in r:
    rec = launch.setdefault(x[iid], [x[ik], 0.0, 0.0])
    v, u, m = float(x[iv].replace(',', '')), x[iu], x[im]
    if m.startswith('gpu__time_duration'):
        rec[1] = v / 1000.0 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000.0)
    elif m.startswith('dram__bytes'):
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
        rec[2] += v * scale
data = list(launch.values())
idx = [i for i, d in enumerate(data) if 'ddim_eps' in d[0]]
seg = data[idx[0] + 1: idx[1] + 1] if len(idx) >= 2 else data
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, v, b in seg:
    name = k.split('(')[0].replace('void ', '')[:60]
    a = agg[name]; a[0] += 1; a[1] += v; a[2] += b
tot = sum(a[1] for a in agg.values())
print('# one denoising step: %d kernel launches, %.2f ms summed device time under ncu (serialised, cold caches: compare SHARES)'
      % (len(seg), tot / 1000))
print('#       time    share  launches  DRAM GB   kernel')
for k, (n, v, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%9.1f us %5.1f%%  x%3d  %7.3f   %s' % (v, 100 * v / tot, n, b / 1e9, k))
if len(sys.argv) > 2:
    tc = [d for d in seg if 'gather_gemm_tc_kernel' in d[0]]
    json.dump({'dram_bytes_per_step': sum(d[2] for d in tc), 'launches_per_step': len(tc),
               'source': 'ncu dram__bytes_read.sum+dram__bytes_write.sum over the gather_gemm_tc_kernel launches of one '
                         'step (bench.py --steps 1 --warmup 1), round 1'}, open(sys.argv[2], 'w'), indent=1)
