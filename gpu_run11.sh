python -m pytest tests/test_ba_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -8
python ba_large_prof.py > gpurun_out/plain_bal5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bal5.csv python ba_large_prof.py > gpurun_out/ncu_bal5.log 2>&1
echo "ba large launch list exit $?"
python bench.py --workload ba_large --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_large_n1b.json; cut -c1-300 gpurun_out/bench_ba_large_n1b.json
