set -x
timeout 600 python -m pytest tests/test_match_bf_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python profiles/scripts/sweep_probe.py 128 5 2>&1 | tail -3
timeout 300 python bench.py --workload sweep --steps 10 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('sweep', d['value']/1e9, d['e2e']['value']/1e9, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
