import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
pbs = [synth.make_ba_problem(i, C=10, P=5000) for i in range(64)]
bt = synth.batch_windows(pbs)
opt = capi.ba_options(max_num_iterations=2, function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0, max_consecutive_invalid_steps=1 << 30)
prob = c.ba_problem_batched(bt)
s = prob.solve(opt)
print(s[0])
