#!/bin/bash
# development helper: sweep kd search variants with the quick bench pass (device-resident frames only)
for g in 0 2 4 8; do
  for cell in 0.16 0.32; do
    PLS_KD_GROUP=$g PLS_KD_CELL=$cell python bench.py --quick --steps 40 --warmup 24 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('group $g cell $cell ms/frame', round(d['ms_per_step'],4), 'launches', d['gpu_launches'])"
  done
done
