(time python bench.py) > gpurun_out/r2ay_bench.json 2> gpurun_out/r2ay_bench.err; tail -4 gpurun_out/r2ay_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2ay_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'one_stream', d['value_one_stream']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_device_resident'], 'roof', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['peak'])
print('config3', d['config3']['device_resident_ms'], 'config4', d['config4']['device_resident'], d['config4']['e2e'], 'launches', d['gpu_launches'])
print(d['kernel_ms']); print(d['clocks']); print(d['cpu_baseline']['value'], d['e2e_blocking_call']['value'], d['e2e_pageable']['value'], d['e2e']['copy_ceiling_pairs_per_s'])"
