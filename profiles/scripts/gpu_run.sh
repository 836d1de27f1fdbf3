set -x
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_host_dropin_gpu.py tests/test_edge_cases_gpu.py tests/test_ref_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
for W in ba_batched; do
timeout 600 python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo rc=$?; tail -2 gpurun_out/bench_$W.err
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_$W.json').read().strip().split('\n')[-1])
print('$W', b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['ms_per_attempt_by_part'], b['parity'])
PY
done
LORB_SOAK_SEED=404 timeout 300 python profiles/scripts/ba_soak.py 80 > gpurun_out/soak_ba.log 2>&1; echo rc=$?; tail -3 gpurun_out/soak_ba.log | cut -c1-300
