#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== gpu tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 2>&1 | tail -6
for L in 0 1; do
for e in "" "stats,emb"; do
  echo "=== LAYOUT=$L EPI=$e"
  OCTFUSION_TC_LAYOUT=$L EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
done
echo "=== EPI= CG=2"
OCTFUSION_TC_CG=2 SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
echo "=== timeline CG=2"
OCTFUSION_TC_CG=2 timeout 600 python tools/trace_tc.py "4,512,512" 0 2>&1 | tail -30
echo "=== bench LAYOUT=0"
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_9a.json 2> gpurun_out/bench_9a.err; tail -3 gpurun_out/bench_9a.err
python tools/show_bench.py gpurun_out/bench_9a.json 2>&1 | head -20
echo "=== bench LAYOUT=1"
OCTFUSION_TC_LAYOUT=1 timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_9b.json 2> gpurun_out/bench_9b.err; tail -3 gpurun_out/bench_9b.err
python tools/show_bench.py gpurun_out/bench_9b.json 2>&1 | head -20
