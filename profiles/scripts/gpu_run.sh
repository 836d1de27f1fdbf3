set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python profiles/scripts/sweep_probe.py 128 5 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo rc=$?; tail -3 gpurun_out/bench_n1c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1c.json').read().strip().split('\n')[-1])
print('sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['roofline']['frac'], d['roofline']['frac_executed'], d['roofline']['peak'], d['clocks'])
print('full', d['full_sweep']['value']/1e9, d['full_sweep']['wall_s'], d['cpu_baseline']['value']/1e9)
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['kernel'], b['roofline']['frac'], b['roofline']['ms_per_attempt_by_part'], b['parity']['ok'], b['cpu_baseline']['value']/1e6)
print(d['extra'])
PY
