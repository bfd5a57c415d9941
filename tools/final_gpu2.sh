#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vae.py -x -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -f -o gpurun_out/tc_d5_k5404 python tools/prof_conv.py 5 768 256 2>&1 | tail -2
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -f -o gpurun_out/tc_d6_k931 python tools/prof_conv.py 6 128 128 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench13.err > gpurun_out/bench13.json; python tools/show_bench.py gpurun_out/bench13.json 2>&1 | head -3
