set -x
for S in 101 102 103; do
LORB_SOAK_SEED=$S timeout 300 python profiles/scripts/match_soak.py 150 > gpurun_out/soak_match_$S.log 2>&1; echo rc=$?; tail -1 gpurun_out/soak_match_$S.log
done
LORB_SOAK_SEED=201 timeout 600 python profiles/scripts/ba_soak.py 120 > gpurun_out/soak_ba.log 2>&1; echo rc=$?; tail -3 gpurun_out/soak_ba.log
