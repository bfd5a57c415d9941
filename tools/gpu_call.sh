#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "=== kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 200 2>&1 | tail -3
echo "=== model tests (golden, full U-Net, reproducibility)"; timeout 300 python -m pytest tests/test_gpu_model.py -q -x --timeout 200 2>&1 | tail -3
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_35.json 2> gpurun_out/bench_35.err; tail -2 gpurun_out/bench_35.err
python tools/show_bench.py gpurun_out/bench_35.json 2>&1 | head -14
} > gpurun_out/last_35.log 2>&1
cat gpurun_out/last_35.log
