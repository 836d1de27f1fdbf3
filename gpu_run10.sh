python -m pytest tests/test_ba_gpu.py tests/test_multi_gpu.py tests/test_host_dropin_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -8
python ba_large_prof.py > gpurun_out/plain_bal4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bal4.csv python ba_large_prof.py > gpurun_out/ncu_bal4.log 2>&1
echo "ba large launch list exit $?"
python bench.py --workload ba_large --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_large_n1.json; cut -c1-300 gpurun_out/bench_ba_large_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --workload ba_large --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_large_n2.json; cut -c1-300 gpurun_out/bench_ba_large_n2.json
