#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== gpu tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q --timeout 600 2>&1 | tail -12
for e in "" "stats,emb" "stats,resid"; do
  echo "=== EPI=$e"
  EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
echo "=== EPI= CG=2"
OCTFUSION_TC_CG=2 SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
echo "=== EPI=stats,emb CG=2"
OCTFUSION_TC_CG=2 EPI="stats,emb" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
echo "=== timeline EPI=stats,emb"
EPI="stats,emb" timeout 600 python tools/trace_tc.py "6,128,128;5,128,256" 2>&1 | grep -E "==|epilogue|drain|wait acc|tile period|stage period|MMA warp|stage ready|issue MMAs|commits"
echo "=== timeline CG=2"
OCTFUSION_TC_CG=2 timeout 600 python tools/trace_tc.py "4,512,512" 0 2>&1 | tail -30
echo "=== bench CG=1"
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_8a.json 2> gpurun_out/bench_8a.err; tail -3 gpurun_out/bench_8a.err
python tools/show_bench.py gpurun_out/bench_8a.json 2>&1 | head -20
echo "=== bench CG=2"
OCTFUSION_TC_CG=2 timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_8b.json 2> gpurun_out/bench_8b.err; tail -3 gpurun_out/bench_8b.err
python tools/show_bench.py gpurun_out/bench_8b.json 2>&1 | head -20
echo "=== prof_gn"
timeout 600 python tools/prof_gn.py 2>&1 | tail -8
