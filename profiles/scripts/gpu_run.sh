set -x
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?; tail -1 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().split('\n')[-1])
print('sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['roofline']['frac'], d['clocks'], d['gpu_launches'])
print('full', d['full_sweep']['value']/1e9, d['full_sweep']['wall_s'], d['full_sweep']['parity']['ok'])
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['parity']['ok'], b.get('cpu_baseline',{}).get('value'))
print(sorted(d.keys()))
r=json.loads(open('gpurun_out/bench_final_ref.json').read().strip().split('\n')[-1])
print('ref', r['value']/1e9, r.get('impl'), sorted(r.keys()))
PY
