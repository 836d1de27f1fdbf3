python -m pytest tests/test_ba_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -8
python ba_large_prof.py > gpurun_out/plain_bal3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bal3.csv python ba_large_prof.py > gpurun_out/ncu_bal3.log 2>&1
echo "ba large launch list exit $?"
python bench.py --workload ba_large --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
