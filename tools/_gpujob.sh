for v in lib lib_lbc8 lib_lbc9; do
VB_LIB_PATH=$PWD/vslam_b200/$v/libvslam_b200.so python bench.py --quick --no-cpu-baseline > gpurun_out/r2ah_$v.json 2> gpurun_out/r2ah_$v.err; tail -1 gpurun_out/r2ah_$v.err
python -c "
import json; d=json.loads(open('gpurun_out/r2ah_$v.json').read().strip().splitlines()[-1])
print('$v value', round(d['value']), round(d['ms_per_step'],3), 'one_stream', round(d['value_one_stream']['ms_per_step'],3), 'score', round(d['kernel_ms']['score'],4))"
done
