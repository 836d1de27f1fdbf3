set -x
for NW in 8 64 128 256 512; do
timeout 600 python bench.py --workload ba_batched --windows-total $NW --no-cpu-baseline > gpurun_out/bench_bb_$NW.json 2> gpurun_out/bench_bb_$NW.err; echo rc=$?
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_bb_$NW.json').read().strip().split('\n')[-1])
print('$NW windows', b['value']/1e9, 'e2e', b['e2e']['value']/1e9, b['roofline']['ms_per_attempt_by_part'], b['parity']['ok'])
PY
done
timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -2
