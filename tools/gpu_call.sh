#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "##### DEBUG=7"; OCTFUSION_TC_DEBUG=7 python tools/trace_tc.py "6,128,128;4,512,512" 2>&1 | grep -E "==|stage period|stage ready|commit ->|loop top|issue MMAs|commits"
echo "##### DEBUG=15 (no MMA either)"; OCTFUSION_TC_DEBUG=15 python tools/trace_tc.py "6,128,128;4,512,512" 2>&1 | grep -E "==|stage period|stage ready|commit ->|loop top|issue MMAs|commits"
echo "##### plain"; python tools/trace_tc.py "6,128,128;4,512,512" 2>&1 | grep -E "==|stage period|stage ready|commit ->|loop top|issue MMAs|commits"
} > gpurun_out/trace_23.log 2>&1
cat gpurun_out/trace_23.log
