set -x
timeout 300 python -m pytest tests/test_match_proj_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python profiles/scripts/_soak_old.py 45 > gpurun_out/soak_old.log 2>&1; echo rc=$?; tail -3 gpurun_out/soak_old.log
