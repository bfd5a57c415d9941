#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 300 2>&1 | tail -5
echo "=== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -x --timeout 600 2>&1 | tail -5
echo "=== bench"; timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_26.json 2> gpurun_out/bench_26.err; tail -3 gpurun_out/bench_26.err
python tools/show_bench.py gpurun_out/bench_26.json 2>&1 | head -60
} > gpurun_out/small_26.log 2>&1
cat gpurun_out/small_26.log
