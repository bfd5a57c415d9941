#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -x 2>&1 | tail -3
for e in "" "stats,emb" "stats,resid"; do
  echo "=== EPI=$e"
  EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
echo "=== bench"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_17.json 2> gpurun_out/bench_17.err; tail -3 gpurun_out/bench_17.err
python tools/show_bench.py gpurun_out/bench_17.json 2>&1 | head -12
