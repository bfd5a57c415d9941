#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== gpu tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 2>&1 | tail -6
for U in 0 1 2; do
  echo "=== UNI=$U EPI="
  OCTFUSION_TC_UNI=$U SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
for e in "stats,emb" "stats,resid"; do
  echo "=== UNI=0 EPI=$e"
  EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
echo "=== EPI= CG=2"
OCTFUSION_TC_CG=2 SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
echo "=== timeline EPI=stats,emb"
EPI="stats,emb" timeout 600 python tools/trace_tc.py "6,128,128" 2>&1 | grep -E "==|epilogue|drain|wait acc|tile period|stage period|MMA warp|stage ready|issue MMAs|commits"
echo "=== timeline CG=2"
OCTFUSION_TC_CG=2 timeout 600 python tools/trace_tc.py "4,512,512" 0 2>&1 | grep -E "==|epilogue|drain|wait acc|tile period|stage period|MMA warp|stage ready|issue MMAs|commits|loader|slot free|issue  "
echo "=== bench UNI=0"
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_10a.json 2> gpurun_out/bench_10a.err; tail -3 gpurun_out/bench_10a.err
python tools/show_bench.py gpurun_out/bench_10a.json 2>&1 | head -20
echo "=== bench UNI=2"
OCTFUSION_TC_UNI=2 timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_10b.json 2> gpurun_out/bench_10b.err; tail -3 gpurun_out/bench_10b.err
python tools/show_bench.py gpurun_out/bench_10b.json 2>&1 | head -3
echo "=== bench CG=2"
OCTFUSION_TC_CG=2 timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_10c.json 2> gpurun_out/bench_10c.err; tail -3 gpurun_out/bench_10c.err
python tools/show_bench.py gpurun_out/bench_10c.json 2>&1 | head -3
