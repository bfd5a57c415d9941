#!/bin/bash
# 4 GPUs: BASELINE configs[4] (class-conditional, B=32 sharded 8 per GPU) and the default workload, strong scaling
N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== cond --gpus $N"
timeout 600 $TR --master-port 29502 bench.py --gpus $N --workload cond --steps 20 --warmup 3 --no-roofline > gpurun_out/bench_cond_n${N}.json 2> gpurun_out/bench_cond_n${N}.err; tail -2 gpurun_out/bench_cond_n${N}.err
python tools/show_bench.py gpurun_out/bench_cond_n${N}.json 2>&1 | head -2
echo "=== unet --gpus $N (strong)"
timeout 600 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 --no-roofline > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; tail -2 gpurun_out/bench_n${N}.err
python tools/show_bench.py gpurun_out/bench_n${N}.json 2>&1 | head -2
python -c "import json;d=json.load(open('gpurun_out/bench_n${N}.json'));print(d['scaling'], d['config']['batch_per_gpu'], d['gathered_latent_rows'], d['clocks'], d['e2e'])"
