set -x
LORB_SOAK_SEED=303 timeout 600 python profiles/scripts/ba_soak.py 80 > gpurun_out/soak_ba.log 2>&1; echo rc=$?; tail -6 gpurun_out/soak_ba.log | cut -c1-300
