#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,128,256;5,128,256;5,768,256;4,512,512"
echo "=== gpu tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q --timeout 600 2>&1 | tail -8
for e in "" "stats,emb" "stats,resid"; do
  echo "=== EPI=$e"
  EPI="$e" SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -5
done
echo "=== timeline EPI=stats,emb"
EPI="stats,emb" timeout 600 python tools/trace_tc.py "6,128,128" 2>&1 | grep -E "==|epilogue|drain|wait acc|tile period|stage period|MMA warp"
echo "=== timeline EPI=stats,resid"
EPI="stats,resid" timeout 600 python tools/trace_tc.py "6,128,128" 2>&1 | grep -E "==|epilogue|drain|wait acc|tile period|stage period|MMA warp"
echo "=== bench fused"
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_7a.json 2> gpurun_out/bench_7a.err; tail -3 gpurun_out/bench_7a.err
python tools/show_bench.py gpurun_out/bench_7a.json 2>&1 | head -20
echo "=== prof_gn"
timeout 600 python tools/prof_gn.py 2>&1 | tail -8
echo "=== prof_gn chunk 32KB"
OCTFUSION_GN_CHUNK_KB=32 timeout 600 python tools/prof_gn.py 2>&1 | tail -8
echo "=== prof_gn chunk 128KB"
OCTFUSION_GN_CHUNK_KB=128 timeout 600 python tools/prof_gn.py 2>&1 | tail -8
