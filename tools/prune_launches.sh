python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_count|k_select|k_corr" -s 60 -c 40 --csv --log-file gpurun_out/launches_prune.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_prune.csv')) if len(r)>5]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
for r in rows[1:]:
    print(r[ix['Kernel Name']][:40], r[ix['Grid Size']] if 'Grid Size' in ix else '', r[ix['Metric Value']], r[ix['Metric Unit']])
PY
