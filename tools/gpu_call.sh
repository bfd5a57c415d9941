#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 300 2>&1 | tail -8
echo "=== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_vae.py -q -x --timeout 600 2>&1 | tail -5
echo "=== bench"; timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_29.json 2> gpurun_out/bench_29.err; tail -3 gpurun_out/bench_29.err
python tools/show_bench.py gpurun_out/bench_29.json 2>&1 | grep -E "steps/s|TC kernel|M=   2048|M=  16384|clocks"
echo "=== bench, split-K off"; OCTFUSION_TC_SPLITK=0 timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_29b.json 2> gpurun_out/bench_29b.err; tail -3 gpurun_out/bench_29b.err
python tools/show_bench.py gpurun_out/bench_29b.json 2>&1 | grep -E "steps/s|TC kernel|M=   2048|M=  16384|clocks"
} > gpurun_out/splitk_29.log 2>&1
cat gpurun_out/splitk_29.log
