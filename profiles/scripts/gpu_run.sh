set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().split('\n')[-1])
print('sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['roofline']['frac'], d['roofline'].get('mma_stream_microbench'), d['clocks'])
for k in ('ba_batched','ba_large'):
    b=d[k]; print(k, b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['kernel'], b['roofline']['frac'], b['parity'], b.get('cpu_baseline',{}).get('value'))
PY
