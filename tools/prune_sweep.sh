#!/bin/bash
# Tuning of the bounded counting on the bench workload: checkpoint schedule (VB_PRUNE_ROUNDS x VB_PRUNE_GROWTH16), item size
# (VB_PRUNE_ITEM_CHUNKS), round-per-launch version (VB_PRUNE_QUEUE=0).
for q in ${QUEUE:-1}; do for ic in ${ITEM:-2}; do for r in ${ROUNDS:-6 8}; do for g in ${GROWTH:-8}; do
  VB_PRUNE_QUEUE=$q VB_PRUNE_ITEM_CHUNKS=$ic VB_PRUNE_ROUNDS=$r VB_PRUNE_GROWTH16=$g timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('queue $q item $ic rounds $r growth16 $g', round(d['ms_per_step'],3), 'score', round(d['kernel_ms']['score'],3), 'frac', round(d['bounded_counting']['fraction'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])"
done; done; done; done
