#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
tools/exp_tc.sh > gpurun_out/exp_tc_17.log 2>&1
tail -n 300 gpurun_out/exp_tc_17.log
