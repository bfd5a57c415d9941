#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== errors"
timeout 1500 python -m pytest tests/test_gpu_model.py -q -s --timeout 600 2>&1 | grep -E "ERR |passed|failed"
echo "=== bench (all layers)"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_14.json 2> gpurun_out/bench_14.err; tail -3 gpurun_out/bench_14.err
python tools/show_bench.py gpurun_out/bench_14.json 2>&1 | head -50
echo "=== launch list + DRAM bytes (one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2500 -c 700 --csv --log-file gpurun_out/launches_14.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline > /dev/null 2> gpurun_out/ncu_14.err; tail -2 gpurun_out/ncu_14.err
python tools/launch_summary.py gpurun_out/launches_14.csv gpurun_out/tc_traffic_14.json 2>&1 | head -30
echo "=== ncu gn_apply"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 1 -o gpurun_out/prof_gn_apply python tools/prof_gn.py 2>&1 | tail -2
echo "=== vae launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_vae.csv python bench.py --workload vae --batch 8 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/ncu_vae.err; tail -2 gpurun_out/ncu_vae.err
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_vae.csv') if l.startswith('"')]
r=csv.reader(lines); h=next(r)
ik,iv,iu=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit')
rows=[(x[ik], (float(x[iv].replace(',',''))/1000 if x[iu] in ('ns','nsecond') else float(x[iv].replace(',','')))) for x in r]
print('total launches', len(rows))
# last pass = kernels after the last KL_conv-like marker: take the last 215 launches
seg=rows[-215:]
agg=collections.defaultdict(lambda:[0,0.0])
for k,v in seg:
    kk=k.split('(')[0][:70]; agg[kk][0]+=1; agg[kk][1]+=v
tot=sum(a[1] for a in agg.values())
print('last 215 launches: %.2f ms' % (tot/1000))
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print('%9.1f us %5.1f%% x%3d %s'%(v,100*v/tot,n,k))
PY
