#!/bin/bash
# TMA tile::gather4 producers (OCTFUSION_TC_TMAG=1) against the cp.async producers: parity tests, per-layer timing, the step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== kernel tests, TMA gather"; OCTFUSION_TC_TMAG=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -x 2>&1 | tail -15
echo "=== model tests, TMA gather"; OCTFUSION_TC_TMAG=1 timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -5
SH="6,128,128;6,256,128;6,128,256;5,256,256;5,768,256;4,512,512;6,128,8"
for t in 0 1; do for e in "" "stats,resid"; do
  echo "=== TMAG=$t EPI=$e"
  OCTFUSION_TC_TMAG=$t EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -7
done; done
echo "=== TMAG=1 no gather / no weights (d4)"
for d in 1 2; do OCTFUSION_TC_TMAG=1 OCTFUSION_TC_DEBUG=$d SHAPES="4,512,512;6,128,128" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -2; done
echo "=== bench TMAG=1"
OCTFUSION_TC_TMAG=1 timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_tmag.json 2> gpurun_out/bench_tmag.err; tail -3 gpurun_out/bench_tmag.err
python tools/show_bench.py gpurun_out/bench_tmag.json 2>&1 | head -14
echo "=== bench TMAG=0"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_tmag0.json 2> gpurun_out/bench_tmag0.err; tail -3 gpurun_out/bench_tmag0.err
python tools/show_bench.py gpurun_out/bench_tmag0.json 2>&1 | head -4
