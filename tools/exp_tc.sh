#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== prof_gn OCC=5"
timeout 600 python tools/prof_gn.py 2>&1 | tail -6
echo "=== prof_gn OCC=4"
OCTFUSION_GN_OCC=4 timeout 600 python tools/prof_gn.py 2>&1 | tail -6
echo "=== bench OCC=5"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline > gpurun_out/bench_15a.json 2> gpurun_out/bench_15a.err; tail -3 gpurun_out/bench_15a.err
python tools/show_bench.py gpurun_out/bench_15a.json 2>&1 | head -2
echo "=== bench OCC=4"
OCTFUSION_GN_OCC=4 timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline > gpurun_out/bench_15b.json 2> gpurun_out/bench_15b.err; tail -3 gpurun_out/bench_15b.err
python tools/show_bench.py gpurun_out/bench_15b.json 2>&1 | head -2
echo "=== model tests (tolerances)"
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --timeout 600 2>&1 | tail -5
