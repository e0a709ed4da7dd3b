#!/bin/bash
# development helper: sweep kd search variants with the quick bench pass
for cell in 0.16 0.32 0.64; do
  for warp in "" 1; do
    if [ -n "$warp" ]; then export PLS_KD_WARP=1; else unset PLS_KD_WARP; fi
    PLS_KD_CELL=$cell python bench.py --quick --steps 30 --warmup 24 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cell $cell warp [$warp] ms/frame', round(d['ms_per_step'],4))"
  done
done
