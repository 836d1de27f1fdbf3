python -m pytest tests/test_ba_gpu.py tests/test_host_dropin_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -15
python ba_prof.py > gpurun_out/plain_ba4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_ba3.csv python ba_prof.py > gpurun_out/ncu_ba3.log 2>&1
echo "ba launch list exit $?"
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
pb = synth.make_ba_problem(0, C=10, P=5000)
opt = capi.ba_options(max_num_iterations=10, function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0, max_consecutive_invalid_steps=1<<30)
prob = c.ba_problem(pb)
prob.solve(opt)
for _ in range(3):
    prob.reset(); c.sync(); t0 = time.time(); s = prob.solve(opt); dt = time.time() - t0
    print("cfg3 resident solve: %.3f ms, iters %d" % (dt * 1e3, s["iterations"]), s["final_cost"])
c.ba_local(pb, opt)
t0 = time.time(); c.ba_local(pb, opt); print("cfg3 e2e: %.3f ms" % ((time.time() - t0) * 1e3))
PY
cat > /tmp/large_prof.py <<'PY2'
PY2
python ba_large_prof.py > gpurun_out/plain_bal.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bal.csv python ba_large_prof.py > gpurun_out/ncu_bal.log 2>&1
echo "ba large launch list exit $?"
