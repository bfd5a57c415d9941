#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
for U in 1 0; do
for e in "" "stats,emb" "stats,resid"; do
  echo "=== UNI=$U EPI=$e"
  OCTFUSION_TC_UNI=$U EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
done
echo "=== bench UNI=1"
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_13a.json 2> gpurun_out/bench_13a.err; tail -3 gpurun_out/bench_13a.err
python tools/show_bench.py gpurun_out/bench_13a.json 2>&1 | head -18
echo "=== bench UNI=0"
OCTFUSION_TC_UNI=0 timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_13b.json 2> gpurun_out/bench_13b.err; tail -3 gpurun_out/bench_13b.err
python tools/show_bench.py gpurun_out/bench_13b.json 2>&1 | head -18
echo "=== errors"
timeout 1500 python -m pytest tests/test_gpu_model.py -q -s --timeout 600 2>&1 | grep -E "^ERR|passed|failed"
echo "=== vae launch summary"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 230 --csv --log-file gpurun_out/launches_vae.csv python bench.py --workload vae --batch 8 --steps 2 --warmup 1 > /dev/null 2> gpurun_out/ncu_vae.err; tail -2 gpurun_out/ncu_vae.err
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_vae.csv') if l.startswith('"')]
r=csv.reader(lines); h=next(r)
ik,iv,iu=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for x in r:
    v=float(x[iv].replace(',','')); v = v/1000 if x[iu] in ('ns','nsecond') else v
    k=x[ik].split('(')[0][:70]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(a[1] for a in agg.values())
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print('%9.1f us %5.1f%% x%3d %s'%(v,100*v/tot,n,k))
PY
