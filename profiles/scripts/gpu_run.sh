set -x
for S in 1201 1202; do
LORB_SOAK_SEED=$S timeout 500 python profiles/scripts/match_soak.py 400 > gpurun_out/soak_match_$S.log 2>&1; echo rc=$?; tail -1 gpurun_out/soak_match_$S.log
LORB_SOAK_SEED=$S timeout 500 python profiles/scripts/ba_soak.py 300 > gpurun_out/soak_ba_$S.log 2>&1; echo rc=$?; tail -4 gpurun_out/soak_ba_$S.log | cut -c1-220
done
LORB_SOAK_SEED=2050 timeout 500 python profiles/scripts/orb_soak.py 300 > gpurun_out/soak_orb.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_orb.log
LORB_SOAK_EXTREME=1 LORB_SOAK_SEED=2051 timeout 500 python profiles/scripts/orb_soak.py 200 > gpurun_out/soak_orb_x.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_orb_x.log
LORB_SOAK_SEED=21 timeout 500 python profiles/scripts/stereo_soak.py 60 > gpurun_out/soak_stereo.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_stereo.log
LORB_SOAK_SEED=31 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29711 profiles/scripts/ba_shard_soak.py 150 > gpurun_out/soak_shard.log 2>&1; echo rc=$?; tail -2 gpurun_out/soak_shard.log
timeout 300 python profiles/scripts/concurrency_stress.py 150 2>&1 | tail -4
