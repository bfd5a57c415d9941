#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "=== stepper tests"; timeout 900 python -m pytest tests/test_gpu_model.py -q -x --timeout 600 -k "stepper or sampler" 2>&1 | tail -3
echo "=== bench"; timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_25.json 2> gpurun_out/bench_25.err; tail -3 gpurun_out/bench_25.err
python tools/show_bench.py gpurun_out/bench_25.json 2>&1 | head -3
python -c "import json;d=json.load(open('gpurun_out/bench_25.json'));print(d['e2e'])"
} > gpurun_out/e2e_25.log 2>&1
cat gpurun_out/e2e_25.log
