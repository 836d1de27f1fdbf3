set -x
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29711 profiles/scripts/ba_shard_soak.py 60 > gpurun_out/soak_shard.log 2>&1; echo rc=$?; tail -8 gpurun_out/soak_shard.log | cut -c1-300
