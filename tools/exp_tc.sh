#!/bin/bash
for dbg in 15 47 79 111 1 33 65; do
  echo "== OCTFUSION_TC_DEBUG=$dbg"; OCTFUSION_TC_DEBUG=$dbg REPS=10 SHAPES="6,128,128;6,256,256" timeout 120 python tools/prof_conv.py 2>&1 | tail -2
done
