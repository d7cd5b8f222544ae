#!/bin/bash
# Tuning of the bounded counting on the bench workload: checkpoint schedule (options prune_first_chunks, prune_first16,
# prune_rounds, prune_growth16) and work-item size (prune_item_chunks), passed through VB_OPTIONS (vslam_b200/lib.py).
for fc in ${FIRSTC:-2}; do for f in ${FIRST16:-20}; do for ic in ${ITEM:-1}; do for r in ${ROUNDS:-8}; do for g in ${GROWTH:-6}; do
  VB_OPTIONS="prune_first_chunks=$fc,prune_first16=$f,prune_item_chunks=$ic,prune_rounds=$r,prune_growth16=$g" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('first_chunks $fc first16 $f item $ic rounds $r growth16 $g', round(d['ms_per_step'],3), 'count', round(d['kernel_ms']['score'],3), 'frac', round(d['bounded_counting']['fraction'],4), 'e2e', round(d['e2e']['value']))"
done; done; done; done; done
