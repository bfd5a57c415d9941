#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 2300 tools/exp_tmag.sh > gpurun_out/exp_tmag_18.log 2>&1
tail -n 150 gpurun_out/exp_tmag_18.log
