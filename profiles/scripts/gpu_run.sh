set -x
timeout 600 python -m pytest tests/test_match_bf_gpu.py tests/test_edge_cases_gpu.py tests/test_ref_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --workload sweep --no-cpu-baseline > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo rc=$?; tail -2 gpurun_out/bench_sweep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sweep.json').read().strip().split('\n')[-1])
print('sweep', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['roofline']['frac'], d['roofline']['peak'], d['roofline'].get('frac_executed'), d['clocks'])
print('full', d.get('full_sweep',{}).get('value',0)/1e9, d.get('full_sweep',{}).get('wall_s'), d.get('full_sweep',{}).get('parity',{}).get('ok'))
PY
LORB_SOAK_SEED=1301 timeout 300 python profiles/scripts/match_soak.py 150 sweep > gpurun_out/soak_sweep.log 2>&1; echo rc=$?; tail -1 gpurun_out/soak_sweep.log
