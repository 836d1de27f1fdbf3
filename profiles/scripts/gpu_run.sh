set -x
LORB_SWEEP_CLUSTER=2 timeout 300 python -m pytest tests/test_match_bf_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu -k sweep 2>&1 | tail -4
LORB_SWEEP_CLUSTER=2 timeout 120 python profiles/scripts/sweep_probe.py 128 5 2>&1 | tail -4
timeout 120 python profiles/scripts/sweep_probe.py 128 5 tensor 2>&1 | tail -2
