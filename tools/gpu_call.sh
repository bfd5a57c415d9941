#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
tools/exp_tc.sh > gpurun_out/exp_tc_17.log 2>&1
tools/exp_l2.sh > gpurun_out/exp_l2_17.log 2>&1
tail -n 80 gpurun_out/exp_tc_17.log
tail -n 120 gpurun_out/exp_l2_17.log
