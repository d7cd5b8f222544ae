python -m pytest tests/test_gpu_bounded_count.py tests/test_gpu_parity.py -x -q -k "not tensor_path_edge" 2>&1 | tail -3
for i in 1 2; do python bench.py --quick --no-cpu-baseline > gpurun_out/r2aq.json 2> gpurun_out/r2aq.err; tail -1 gpurun_out/r2aq.err
python -c "
import json; d=json.loads(open('gpurun_out/r2aq.json').read().strip().splitlines()[-1])
print('value', round(d['value']), round(d['ms_per_step'],3), 'one', round(d['value_one_stream']['ms_per_step'],3), 'score', round(d['kernel_ms']['score'],4), 'e2e', round(d['e2e']['value']))"; done
