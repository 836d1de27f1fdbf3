set -x
LORB_SOAK_SEED=2031 timeout 900 python profiles/scripts/orb_soak.py 200 > gpurun_out/soak_orb.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_orb.log
LORB_SOAK_SEED=9 timeout 600 python profiles/scripts/stereo_soak.py 30 > gpurun_out/soak_stereo.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_stereo.log
LORB_SOAK_SEED=505 timeout 300 python profiles/scripts/match_soak.py 200 > gpurun_out/soak_match.log 2>&1; echo rc=$?; tail -1 gpurun_out/soak_match.log
LORB_SOAK_SEED=606 timeout 300 python profiles/scripts/ba_soak.py 160 > gpurun_out/soak_ba.log 2>&1; echo rc=$?; tail -3 gpurun_out/soak_ba.log | cut -c1-250
