python -m pytest tests/test_multi_gpu.py tests/test_ba_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras 2>&1 | tail -3
python bench.py --workload ba_large --steps 2 --warmup 1 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --workload ba_large --steps 2 --warmup 1 2>&1 | tail -2
python bench.py --workload ba_batched --windows 64 --steps 2 --warmup 1 2>&1 | tail -2
python ba_prof.py > gpurun_out/plain_ba3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_ba2.csv python ba_prof.py > gpurun_out/ncu_ba2.log 2>&1
echo "ba launch list exit $?"
