set -x
LORB_SOAK_SEED=707 timeout 600 python profiles/scripts/match_soak.py 100 entry > gpurun_out/soak_match.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_match.log
LORB_SOAK_SEED=808 timeout 600 python profiles/scripts/ba_soak.py 100 > gpurun_out/soak_ba.log 2>&1; echo rc=$?; tail -8 gpurun_out/soak_ba.log | cut -c1-250
