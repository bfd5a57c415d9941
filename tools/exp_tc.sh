#!/bin/bash
timeout 250 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
REPS=10 SHAPES="6,128,128;6,256,128;6,256,256;5,256,256;5,768,256;4,512,512" timeout 120 python tools/prof_conv.py 2>&1 | tail -6
