#!/bin/bash
# round-end consolidation: full GPU suite, bench line, ncu launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/final_bench.err >gpurun_out/final_bench.json
python tools/show_bench.py gpurun_out/final_bench.json 2>&1 | head -20
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline >gpurun_out/final_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/final_launches.csv gpurun_out/tc_traffic.json | head -30
timeout 300 python tools/prof_vae.py 2>&1 | tail -2 | tee gpurun_out/final_vae.txt
