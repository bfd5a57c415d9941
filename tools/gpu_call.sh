#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 300 2>&1 | tail -6
echo "=== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_vae.py tests/test_gpu_mpu.py tests/test_octree_conv.py -q -x --timeout 600 2>&1 | tail -5
echo "=== model tests with split-K on (informational)"; OCTFUSION_TC_SPLITK=1 timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 2>&1 | tail -6
} > gpurun_out/tests_31.log 2>&1
cat gpurun_out/tests_31.log
