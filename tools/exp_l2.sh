#!/bin/bash
# Is the 256-wide tile bound by L2 -> SM traffic?  One GraphConv with the gather / the weight stream switched off
# (OCTFUSION_TC_DEBUG 1 / 2 / 3: results are garbage, timing only), the gather routed through L1 (64), CTA pairs with 6 / 8
# weight slots.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SH="5,768,256;4,512,512;5,512,512"
run() { echo "--- $1"; env $1 SHAPES="$SH" REPS=10 timeout 600 python tools/prof_conv.py 2>&1 | tail -3; }
run "X=0"
run "OCTFUSION_TC_DEBUG=1"
run "OCTFUSION_TC_DEBUG=2"
run "OCTFUSION_TC_DEBUG=3"
run "OCTFUSION_TC_DEBUG=64"
run "OCTFUSION_TC_CG=2 OCTFUSION_TC_UNI=0"
run "OCTFUSION_TC_CG=2 OCTFUSION_TC_UNI=2"
run "OCTFUSION_TC_CG=2 OCTFUSION_TC_UNI=2 OCTFUSION_TC_DEBUG=1"
run "OCTFUSION_TC_CG=2 OCTFUSION_TC_UNI=2 OCTFUSION_TC_DEBUG=2"
run "OCTFUSION_TC_CG=2 OCTFUSION_TC_UNI=2 OCTFUSION_TC_DEBUG=64"
echo "=== N=128 layers, gather through L1"
SH="6,128,128;6,256,128"
run "X=0"
run "OCTFUSION_TC_DEBUG=64"
run "OCTFUSION_TC_DEBUG=1"
run "OCTFUSION_TC_DEBUG=2"
echo "=== L2 metrics (ncu) of d4 512->512: single CTA vs pair"
for v in "OCTFUSION_TC_CG=1" "OCTFUSION_TC_CG=2 OCTFUSION_TC_UNI=2" "OCTFUSION_TC_DEBUG=64"; do
  echo "--- $v"
  env $v REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors.sum.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,lts__cycles_elapsed.avg.per_second --clock-control none -k regex:gather_gemm_tc -s 2 -c 1 python tools/prof_conv.py 4 512 512 2>&1 | grep -E "lts__|l1tex__|gpu__time|sm__" 
done
