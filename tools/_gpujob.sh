for o in "prune_item_chunks=1" "prune_item_chunks=2" "prune_item_chunks=3" "prune_item_chunks=4" "prune_item_chunks=2,prune_rounds=10" "prune_item_chunks=2,prune_growth16=4" "prune_first16=18" "prune_first_chunks=1"; do
VB_OPTIONS="$o" python bench.py --quick --no-cpu-baseline > gpurun_out/r2af.json 2> gpurun_out/r2af.err; tail -1 gpurun_out/r2af.err
python -c "
import json; d=json.loads(open('gpurun_out/r2af.json').read().strip().splitlines()[-1])
print('$o value', round(d['value']), round(d['ms_per_step'],3), 'one_stream', round(d['value_one_stream']['ms_per_step'],3), 'score', round(d['kernel_ms']['score'],4), 'solve', round(d['kernel_ms']['solve'],3), 'frac', round(d['bounded_counting']['fraction'],4))" | tee -a gpurun_out/r2af_sweep.txt
done
