#!/bin/bash
# multi-GPU check: N ranks, B=32 sharded (strong scaling), uncond + cond; the CPU arm under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
echo "=== model parity (1 GPU) on this build"; timeout 900 python -m pytest tests/test_gpu_model.py -q -x --timeout 600 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== unet --gpus $N (strong)"
timeout 900 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; tail -2 gpurun_out/bench_n${N}.err
python tools/show_bench.py gpurun_out/bench_n${N}.json 2>&1 | head -3
python -c "import json;d=json.load(open('gpurun_out/bench_n${N}.json'));print(d['scaling'], d['config']['batch_per_gpu'], d['config']['nodes_per_gpu'], d['gathered_latent_rows'], d['clocks'])"
echo "=== cond --gpus $N"
timeout 900 $TR --master-port 29502 bench.py --gpus $N --workload cond --steps 20 --warmup 3 > gpurun_out/bench_cond_n${N}.json 2> gpurun_out/bench_cond_n${N}.err; tail -2 gpurun_out/bench_cond_n${N}.err
python tools/show_bench.py gpurun_out/bench_cond_n${N}.json 2>&1 | head -3
echo "=== unet --gpus $N --weak"
timeout 900 $TR --master-port 29503 bench.py --gpus $N --weak --steps 10 --warmup 3 --no-roofline > gpurun_out/bench_weak_n${N}.json 2> gpurun_out/bench_weak_n${N}.err; tail -2 gpurun_out/bench_weak_n${N}.err
python tools/show_bench.py gpurun_out/bench_weak_n${N}.json 2>&1 | head -2
