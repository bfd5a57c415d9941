#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,256,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== gpu tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_octree_conv.py tests/test_gpu_mpu.py -q --timeout 600 2>&1 | tail -8
for e in "" "stats,emb" "stats,resid"; do
  echo "=== EPI=$e"
  EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -6
done
echo "=== timeline EPI=stats,emb"
EPI="stats,emb" timeout 600 python tools/trace_tc.py "6,128,128" 2>&1 | grep -E "==|epilogue|drain|wait acc|tile period|stage period|MMA warp"
echo "=== bench fused"
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_6a.json 2> gpurun_out/bench_6a.err; tail -3 gpurun_out/bench_6a.err
python tools/show_bench.py gpurun_out/bench_6a.json 2>&1 | head -20
echo "=== launch list (one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 700 --csv --log-file gpurun_out/launches_6.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline > /dev/null 2> gpurun_out/ncu_6.err; tail -2 gpurun_out/ncu_6.err
python tools/launch_summary.py gpurun_out/launches_6.csv 2>&1 | head -30
