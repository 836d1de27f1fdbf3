"""Where the host-buffer batched BA call (lorb_ba_local_batched, BASELINE config 4) spends its time:
create (staging + upload), solve, download, timed separately through the resident-problem API, then the
whole call.
    [LORB_BA_TRACE=1] python profiles/scripts/ba_batched_e2e_phases.py [windows] [host threads]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lorb_slam_b200 import capi, synth  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pbs = [synth.make_ba_problem(i, C=10, P=5000) for i in range(nw)]
bt = synth.batch_windows(pbs)
opt = capi.ba_options(max_num_iterations=10, function_tolerance=-1.0, parameter_tolerance=-1.0,
                      gradient_tolerance=-1.0, max_consecutive_invalid_steps=1 << 30)
with capi.Context(0) as ctx:
    capi.set_host_threads(int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1))
    for rep in range(3):
        t0 = time.perf_counter()
        prob = ctx.ba_problem_batched(bt)
        ctx.sync()
        t1 = time.perf_counter()
        prob.solve(opt)
        t2 = time.perf_counter()
        prob.download()
        t3 = time.perf_counter()
        prob.close()
        t4 = time.perf_counter()
        print("rep %d: create %.1f ms, solve %.1f ms, download %.1f ms, destroy %.1f ms" %
              (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
    for rep in range(4):
        t0 = time.perf_counter()
        ctx.ba_local_batched(bt, opt, inplace=True)
        print("lorb_ba_local_batched, in place: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
