set -x
timeout 600 python bench.py --workload ba_large --no-cpu-baseline > gpurun_out/bench_ba_large.json 2> gpurun_out/bench_ba_large.err; echo rc=$?
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_ba_large.json').read().strip().split('\n')[-1])
print('ba_large', b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['ms_per_attempt_by_part'], b['parity']['ok'])
PY
timeout 300 python -m pytest tests/test_ba_gpu.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -2
