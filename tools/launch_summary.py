"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one denoising step, per kernel."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
iu = hdr.index('Metric Unit')
data = []
for x in r:
    v = float(x[iv].replace(',', ''))
    u = x[iu]
    v_us = v / 1000.0 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000.0)
    data.append((x[ik], v_us))
idx = [i for i, (k, v) in enumerate(data) if 'ddim_eps' in k]
seg = data[idx[0] + 1: idx[1] + 1] if len(idx) >= 2 else data
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in seg:
    name = k.split('(')[0].replace('void ', '')[:60]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v for _, v in agg.values())
print('# one denoising step: %d kernel launches, %.2f ms summed device time (serialised under ncu)' % (len(seg), tot / 1000))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%9.1f us %5.1f%%  x%3d  %s' % (v, 100 * v / tot, n, k))
