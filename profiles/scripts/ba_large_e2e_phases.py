"""Large BA through host buffers: how the end-to-end time splits into problem creation (host work
lists + uploads), the LM solve and the download."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth  # noqa: E402

pb = synth.make_ba_problem(0, C=200, P=200000, obs_per_point=(7, 8), traj_len=100.0)
opt = capi.ba_options(max_num_iterations=10, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
with capi.Context(0) as ctx:
    import os
    for rep in range(4):
        if rep == 3:
            os.environ["LORB_BA_TRACE"] = "1"
        t0 = time.perf_counter()
        prob = ctx.ba_problem(pb)
        ctx.sync()
        t1 = time.perf_counter()
        s = prob.solve(opt)
        ctx.sync()
        t2 = time.perf_counter()
        prob.download()
        t3 = time.perf_counter()
        prob.close()
        t4 = time.perf_counter()
        ctx.ba_local(pb, opt)
        t5 = time.perf_counter()
        print("rep %d: create %.2f ms, solve %.2f ms (%d it), download %.2f ms, close %.2f ms; lorb_ba_local e2e %.2f ms"
              % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, s["iterations"], (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3))
