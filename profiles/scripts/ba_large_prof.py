import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
pb = synth.make_ba_problem(0, C=200, P=200000, obs_per_point=(7, 8), traj_len=100.0)
opt = capi.ba_options(max_num_iterations=2, function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0, max_consecutive_invalid_steps=1 << 30)
prob = c.ba_problem(pb)
print(prob.solve(opt))
