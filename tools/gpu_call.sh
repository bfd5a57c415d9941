#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_36.json 2> gpurun_out/bench_36.err; tail -3 gpurun_out/bench_36.err
python tools/show_bench.py gpurun_out/bench_36.json 2>&1 | head -3
python -c "import json;d=json.load(open('gpurun_out/bench_36.json'));print(d['graph_build_ms']); print(d['e2e']['value'], d['e2e']['serial_value'])"
} > gpurun_out/last_36.log 2>&1
cat gpurun_out/last_36.log
