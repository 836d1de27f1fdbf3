python -m pytest tests -x -q -m gpu --tb=short 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d['cpu_baseline']); print(json.dumps(d['extra'], indent=1))"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
python bench.py --workload ba_batched --windows 64 --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_batched_n1.json; cut -c1-250 gpurun_out/bench_ba_batched_n1.json
