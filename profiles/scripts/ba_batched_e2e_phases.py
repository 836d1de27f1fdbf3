"""64 batched windows through host buffers (lorb_ba_local_batched): host timeline of the problem
creation (LORB_BA_TRACE) next to the whole call."""
import os
import sys
import time

sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth  # noqa: E402

pbs = [synth.make_ba_problem(i, C=10, P=5000) for i in range(64)]
bt = synth.batch_windows(pbs)
opt = capi.ba_options(max_num_iterations=10, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
with capi.Context(0) as ctx:
    for rep in range(5):
        if rep == 4:
            os.environ["LORB_BA_TRACE"] = "1"
        t0 = time.perf_counter()
        ctx.ba_local_batched(bt, opt)
        print("rep %d: lorb_ba_local_batched e2e %.2f ms" % (rep, (time.perf_counter() - t0) * 1e3))
