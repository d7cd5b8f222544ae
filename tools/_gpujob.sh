python -m pytest tests/test_gpu_bounded_count.py -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --quick --no-cpu-baseline > gpurun_out/r2aj.json 2> gpurun_out/r2aj.err; tail -1 gpurun_out/r2aj.err
python -c "
import json; d=json.loads(open('gpurun_out/r2aj.json').read().strip().splitlines()[-1])
print('value', round(d['value']), round(d['ms_per_step'],3), 'one_stream', round(d['value_one_stream']['ms_per_step'],3), 'score', round(d['kernel_ms']['score'],4))"; done
