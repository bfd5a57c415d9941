#!/bin/bash
for g in 148 111 74; do
  echo "== OCTFUSION_TC_GRID=$g"; OCTFUSION_TC_GRID=$g REPS=10 SHAPES="6,128,128;6,256,256;5,256,256" timeout 120 python tools/prof_conv.py 2>&1 | tail -3
done
echo "== GRID=74 no gather"; OCTFUSION_TC_GRID=74 OCTFUSION_TC_DEBUG=1 REPS=10 SHAPES="6,128,128;6,256,256" timeout 120 python tools/prof_conv.py 2>&1 | tail -2
echo "== GRID=148 no B"; OCTFUSION_TC_DEBUG=2 REPS=10 SHAPES="6,128,128;6,256,256" timeout 120 python tools/prof_conv.py 2>&1 | tail -2
