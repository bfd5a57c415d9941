#!/bin/bash
# wrapper handed to gpurun: edit per experiment.  The round-end evidence run is tools/final_gpu_r02.sh.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 3300 tools/final_gpu_r02.sh > gpurun_out/final_r02.log 2>&1
tail -n 200 gpurun_out/final_r02.log
