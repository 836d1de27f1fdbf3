python -m pytest tests/test_ba_gpu.py tests/test_edge_cases_gpu.py tests/test_host_dropin_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -4
python bench.py --workload ba_batched --windows 64 --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_batched_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_ba_batched_n1.json')); print('ba_batched', d['value']/1e6, 'M resident;', d['e2e']['value']/1e6, 'M e2e; ms/window', d['ms_per_local_ba'], 'launches', d['gpu_launches'])"
