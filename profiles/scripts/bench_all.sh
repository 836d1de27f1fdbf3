# every bench line of the round on one GPU (default sweep with extras, reference arm, both BA workloads)
python bench.py > gpurun_out/bench_n1_sweep.json 2> gpurun_out/bench_n1_sweep.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_n1_reference.json 2>/dev/null
python bench.py --workload ba_batched --windows 64 --steps 5 --warmup 3 > gpurun_out/bench_n1_ba_batched.json 2> gpurun_out/bench_n1_ba_batched.err
python bench.py --workload ba_large --steps 3 --warmup 3 > gpurun_out/bench_n1_ba_large.json 2> gpurun_out/bench_n1_ba_large.err
python - <<'PY'
import json
for f in ('sweep', 'reference', 'ba_batched', 'ba_large'):
    try:
        d = json.load(open('gpurun_out/bench_n1_%s.json' % f))
    except Exception as e:
        print(f, 'FAILED', e); continue
    print(f, '%.4g' % d['value'], d['unit'], 'e2e %.4g' % d['e2e']['value'], 'roofline', {k: d.get('roofline', {}).get(k) for k in ('kernel', 'achieved', 'peak', 'frac', 'avg_launch_ms')})
    if 'extra' in d: print(json.dumps(d['extra'], indent=1))
PY
