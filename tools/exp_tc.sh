#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 2>&1 | tail -4
echo "=== ncu source-level: d6 128->128 stats,resid"
EPI="stats,resid" REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -o gpurun_out/prof_epi python tools/prof_conv.py 6 128 128 2>&1 | tail -3
echo "=== ncu source-level: d4 512->512 plain"
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -o gpurun_out/prof_d4 python tools/prof_conv.py 4 512 512 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
