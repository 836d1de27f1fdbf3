set -x
timeout 600 python -m pytest tests/test_fallback_paths_gpu.py -x -q -m gpu 2>&1 | tail -15
