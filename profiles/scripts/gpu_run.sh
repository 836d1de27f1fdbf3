set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_match_bf_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python profiles/scripts/sweep_probe.py 32 3 2>&1 | tail -8
timeout 300 python profiles/scripts/sweep_probe.py 128 5 2>&1 | tail -8
