export VB_LIB_PATH=$PWD/vslam_b200/lib_tuning/libvslam_b200.so
TC_DRAINS=1,6 timeout 300 python tools/tc_variants.py 2>&1 | tail -2 | tee gpurun_out/r2ba_tc_variants.log
TC_DRAINS=1 TC_EXTRA=tc_dbg=2 timeout 300 python tools/tc_variants.py 2>&1 | tail -1 | tee -a gpurun_out/r2ba_tc_variants.log
unset VB_LIB_PATH
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_path or alternative_kernels or packed_drain" 2>&1 | tail -2
python bench.py --quick --no-cpu-baseline > gpurun_out/r2ba.json 2> gpurun_out/r2ba.err; tail -1 gpurun_out/r2ba.err
python -c "
import json; d=json.loads(open('gpurun_out/r2ba.json').read().strip().splitlines()[-1])
print('value', round(d['value']), round(d['ms_per_step'],3), 'one', round(d['value_one_stream']['ms_per_step'],3), 'hamming', round(d['kernel_ms']['hamming'],4), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4))"
