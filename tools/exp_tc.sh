#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== all gpu tests"; tools/run_gpu_tests.sh 2>&1 | grep -E "^==|passed|failed|error" | head -30
for e in "" "stats,emb" "stats,resid"; do
  echo "=== EPI=$e"
  EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
echo "=== bench (full line)"
timeout 2400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_12.json 2> gpurun_out/bench_12.err; tail -3 gpurun_out/bench_12.err
python tools/show_bench.py gpurun_out/bench_12.json 2>&1 | tail -24
echo "=== bench cond"
timeout 1500 python bench.py --workload cond --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_12_cond.json 2> gpurun_out/bench_12_cond.err; tail -3 gpurun_out/bench_12_cond.err
python tools/show_bench.py gpurun_out/bench_12_cond.json 2>&1 | head -4
echo "=== bench vae"
timeout 1500 python bench.py --workload vae --batch 8 --steps 5 --warmup 2 > gpurun_out/bench_12_vae.json 2> gpurun_out/bench_12_vae.err; tail -5 gpurun_out/bench_12_vae.err; cat gpurun_out/bench_12_vae.json | cut -c1-1500
echo "=== smoke x2"
python __graft_entry__.py smoke 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -4
