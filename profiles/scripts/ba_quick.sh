# quick BA regression + timing pass used while optimising the BA kernels
python -m pytest tests/test_ba_gpu.py tests/test_edge_cases_gpu.py tests/test_ref_golden_gpu.py tests/test_host_dropin_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload ba_batched --windows 64 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_ba_batched_q.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_ba_batched_q.json'))
print('ba_batched64: %.1f M obs*iter/s resident; %.1f M e2e; ms/window %.4f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_local_ba']))
import time, sys
sys.path.insert(0, '.')
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
pb = synth.make_ba_problem(0, C=10, P=5000)
opt = capi.ba_options(max_num_iterations=10, function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0, max_consecutive_invalid_steps=1 << 30)
prob = c.ba_problem(pb); prob.solve(opt)
best = 1e9
for _ in range(5):
    prob.reset(); c.sync(); t0 = time.perf_counter(); s = prob.solve(opt); best = min(best, time.perf_counter() - t0)
print('cfg3 resident solve: %.3f ms, iters %d' % (best * 1e3, s['iterations']))
PY
python profiles/scripts/ba_batch_prof.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bab_q.csv python profiles/scripts/ba_batch_prof.py > /dev/null 2>&1
python profiles/scripts/launch_summary.py gpurun_out/launches_bab_q.csv
