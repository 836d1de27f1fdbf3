set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 120 python profiles/scripts/sweep_probe.py 128 5 tensor 2>&1 | tail -2
