N=$1; TAG=${2:-r2cm}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -3 gpurun_out/${TAG}_bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
print('n', d['n_gpus'], 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('frac_of_copy_ceiling'), d['e2e']['copy_ceiling_pairs_per_s'], d['e2e']['copy_ceiling_GBps_per_gpu'], 'multi', d.get('e2e_multi',{}).get('value'))
print('config4', d['config4']['device_resident']['value'], d['config4']['e2e']['value'])"
