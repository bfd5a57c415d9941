#!/bin/bash
# round-2 end-of-round evidence run (1 GPU): full -m gpu suite, smoke x2, the bench line with both baselines, the ncu
# launch list of one step with DRAM bytes (-> profiles/launches_step_r02.txt, profiles/tc_traffic_r02.json), ncu --set full
# of the widest and the narrowest GraphConv, bench lines of the cond and vae workloads, the calc_sdf grid timing.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== all gpu tests"; tools/run_gpu_tests.sh 2>&1 | grep -E "^==|passed|failed|error" | head -30
echo "=== smoke x2"
python __graft_entry__.py smoke 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench (full line)"
timeout 2400 python bench.py > gpurun_out/bench_n1_r02.json 2> gpurun_out/bench_n1_r02.err; tail -3 gpurun_out/bench_n1_r02.err
python tools/show_bench.py gpurun_out/bench_n1_r02.json 2>&1 | tail -50
echo "=== bench --impl reference"
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; tail -2 gpurun_out/bench_ref_r02.err; cut -c1-600 gpurun_out/bench_ref_r02.json
echo "=== launch list + DRAM bytes (one step)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2500 -c 700 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline > /dev/null 2> gpurun_out/ncu_r02.err; tail -2 gpurun_out/ncu_r02.err
python tools/launch_summary.py gpurun_out/launches_r02.csv gpurun_out/tc_traffic_r02.json > gpurun_out/launches_step_r02.txt 2>&1; head -30 gpurun_out/launches_step_r02.txt
echo "=== ncu full: d4 512->512, d6 128->128 (plain), d6 128->128 conv2 epilogue"
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -o gpurun_out/prof_tc_d4_r02 python tools/prof_conv.py 4 512 512 2>&1 | tail -1
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -o gpurun_out/prof_tc_d6_r02 python tools/prof_conv.py 6 128 128 2>&1 | tail -1
EPI="stats,resid" REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_tc -s 2 -c 1 -o gpurun_out/prof_tc_d6_epi_r02 python tools/prof_conv.py 6 128 128 2>&1 | tail -1
REPS=1 timeout 900 ncu --set full --clock-control none -k regex:attention_tc -c 1 -o gpurun_out/prof_attention_r02 python -m pytest tests/test_gpu_kernels.py -q -k "test_attention and 512-4-32 and dtype1" 2>&1 | tail -1
echo "=== cond / vae bench lines"
timeout 1500 python bench.py --workload cond --no-cpu-baseline --no-library-baseline > gpurun_out/bench_cond_n1_r02.json 2> gpurun_out/bench_cond_r02.err; python tools/show_bench.py gpurun_out/bench_cond_n1_r02.json 2>&1 | head -3
timeout 1500 python bench.py --workload vae --batch 8 --steps 5 --warmup 2 > gpurun_out/bench_vae_n1_r02.json 2> gpurun_out/bench_vae_r02.err; cut -c1-400 gpurun_out/bench_vae_n1_r02.json
echo "=== calc_sdf 256^3"
timeout 900 python tools/prof_sdf.py 2>&1 | tail -6
echo "=== graph build"
timeout 600 python tools/prof_build.py 2>&1 | tail -6
