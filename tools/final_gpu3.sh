#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_tests.txt
timeout 400 python bench.py --steps 20 --warmup 3 2>gpurun_out/final_bench.err >gpurun_out/final_bench.json
python tools/show_bench.py gpurun_out/final_bench.json 2>&1 | head -16
