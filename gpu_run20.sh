python -m pytest tests/test_ba_gpu.py tests/test_edge_cases_gpu.py tests/test_host_dropin_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -4
python bench.py --workload ba_batched --windows 64 --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ba_batched_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_ba_batched_n1.json')); print('ba_batched', d['value']/1e6, 'M resident;', d['e2e']['value']/1e6, 'M e2e; ms/window', d['ms_per_local_ba'], 'launches', d['gpu_launches'])"
python ba_batch_prof.py > gpurun_out/plain_bab.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_bab2.csv python ba_batch_prof.py > gpurun_out/ncu_bab.log 2>&1
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
pb = synth.make_ba_problem(0, C=10, P=5000)
opt = capi.ba_options(max_num_iterations=10, function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0, max_consecutive_invalid_steps=1<<30)
prob = c.ba_problem(pb); prob.solve(opt)
for _ in range(3):
    prob.reset(); c.sync(); t0 = time.time(); s = prob.solve(opt); dt = time.time() - t0
    print("cfg3 resident solve: %.3f ms, iters %d" % (dt * 1e3, s["iterations"]))
PY
