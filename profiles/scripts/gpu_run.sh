set -x
LORB_SOAK_SEED=909 timeout 600 python profiles/scripts/ba_soak.py 120 > gpurun_out/soak_ba.log 2>&1; echo rc=$?; tail -8 gpurun_out/soak_ba.log | cut -c1-250
timeout 300 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -2
