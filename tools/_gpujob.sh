python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --quick --no-cpu-baseline > gpurun_out/r2ao.json 2> gpurun_out/r2ao.err; tail -1 gpurun_out/r2ao.err
python -c "
import json; d=json.loads(open('gpurun_out/r2ao.json').read().strip().splitlines()[-1])
print('value', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['frac_of_device_resident'],4), 'blocking', round(d['e2e_blocking_call']['value']), 'pageable', round(d['e2e_pageable']['value']))"; done
