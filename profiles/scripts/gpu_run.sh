set -x
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?; tail -1 gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().split('\n')[-1])
print('sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['roofline']['frac'], d['clocks'], d['gpu_launches'])
print('full', d['full_sweep']['value']/1e9, d['full_sweep']['wall_s'], d['full_sweep']['parity']['ok'])
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline'], b['parity']['ok'])
print(d.get('extras') or d.get('extra'))
PY
python profiles/scripts/ba_batch_prof.py > gpurun_out/plain_bab.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ba_build_dense_kernel|ba_backsub_kernel' -c 5 -f -o gpurun_out/r02_ba_batch_full python profiles/scripts/ba_batch_prof.py > gpurun_out/ncu_bab_full.log 2>&1
tail -2 gpurun_out/ncu_bab_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ba_batched_launches.csv python bench.py --workload ba_batched --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bab.log 2>&1; echo rc=$?
