#!/bin/bash
# round-2 experiment batch for the tcgen05 tap-gather GEMM: correctness first, then variants, then the timeline.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="6,128,128;6,256,128;6,384,128;6,128,256;5,256,256;5,128,256;5,768,256;5,512,512;4,512,512;4,256,256;6,64,128"
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_octree_conv.py tests/test_gpu_mpu.py -q --timeout 300 2>&1 | tail -15
echo "=== model tests"; timeout 1200 python -m pytest tests/test_gpu_model.py -q --timeout 600 2>&1 | tail -8
for v in "UNI=0 MT=2" "UNI=1 MT=2"; do
  set -- $v
  echo "=== variant $v"
  env OCTFUSION_TC_${1%%=*}=${1##*=} OCTFUSION_TC_${2%%=*}=${2##*=} SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -12
done
echo "=== timeline UNI=0"
timeout 600 python tools/trace_tc.py "6,128,128;4,512,512" 2>&1 | tail -60
echo "=== bench"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_4.json 2> gpurun_out/bench_4.err; tail -5 gpurun_out/bench_4.err
python tools/show_bench.py gpurun_out/bench_4.json 2>&1 | tail -60
