set -x
timeout 300 python profiles/scripts/concurrency_stress.py 30 2>&1 | tail -12
