set -x
LORB_CHOL_CHAIN=1 timeout 300 python -m pytest tests/test_ba_gpu.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
for CH in 1 0 1; do
LORB_CHOL_CHAIN=$CH timeout 300 python bench.py --workload ba_large --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('chain=$CH large', d['value']/1e9, d['roofline']['ms_per_attempt_by_part'], d['parity']['ok'], d['parity']['max_err_over_tolerance'])"
done
