#!/bin/bash
# development aid (needs a TUNING=1 build of the library on LD path, see csrc/Makefile): end-to-end pairs/s of the blocking
# vb_pairs_run for several sub-batch schedules (VB_PAIRS_SCHEDULE) and with/without the twin context (option pairs_twin)
for tw in 1 0; do
for s in "32,64,128,256,352,192" "64,128,256,576" "128,128,128,128,128,128,128,128" "64,64,128,128,128,128,128,128,128" "96,160,256,256,256" "64,192,256,256,256" "256,256,256,256"; do
  VB_OPTIONS="pairs_twin=$tw" VB_PAIRS_SCHEDULE=$s python bench.py --steps 6 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('twin=$tw', '$s', round(d['e2e_blocking_call']['value']), round(d['value']))"
done
done
