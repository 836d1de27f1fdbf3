set -x
python -m pytest tests/test_match_bf_gpu.py tests/test_match_proj_gpu.py -x -q -m gpu 2>&1 | tail -25
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
for kind in (0, 1):
    for it in (2048, 8192):
        w = c.microbench_popc(kind, it)
        print("microbench kind", kind, "iters", it, "Gwords/s", w / 1e9, "Gpairs/s", w / 8e9)
bank = synth.kf_bank(256, 2000, seed=0, shared_frac=0)
pa, pb = synth.all_pairs(256)
pa, pb = pa[:4096], pb[:4096]
c.bank_upload(bank)
c.sweep_plan_upload(pa, pb)
c.sweep_plan_run(); c.sync()
t0 = time.time()
for _ in range(3):
    c.sweep_plan_run()
c.sync()
dt = (time.time() - t0) / 3
print("CSA sweep 4096 pairs x 2000^2: %.3f ms -> %.3f Gpairs/s" % (dt * 1e3, 4096 * 4e6 / dt / 1e9))
q = synth.descriptors_uniform(1000, np.random.default_rng(0)); t = synth.descriptors_uniform(1000, np.random.default_rng(1))
c.match_bf_crosscheck(q, t)
t0 = time.time()
for _ in range(200): c.match_bf_crosscheck(q, t)
print("bf 1000x1000 e2e: %.1f us" % ((time.time() - t0) / 200 * 1e6))
fr = synth.make_frame(2000, 0); pts = synth.make_proj_points(fr, 5000, 0)
for th in (1.0, 15.0):
    c.search_proj_points(fr, pts, th)
    t0 = time.time()
    for _ in range(50): r = c.search_proj_points(fr, pts, th)
    print("proj points th", th, "e2e: %.1f us" % ((time.time() - t0) / 50 * 1e6), r["n_candidates"])
PY
LORB_HAMMING_CSA=0 python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
bank = synth.kf_bank(256, 2000, seed=0, shared_frac=0)
pa, pb = synth.all_pairs(256)
pa, pb = pa[:4096], pb[:4096]
c.bank_upload(bank)
c.sweep_plan_upload(pa, pb)
c.sweep_plan_run(); c.sync()
t0 = time.time()
for _ in range(3):
    c.sweep_plan_run()
c.sync()
dt = (time.time() - t0) / 3
print("PLAIN popc8 sweep 4096 pairs x 2000^2: %.3f ms -> %.3f Gpairs/s" % (dt * 1e3, 4096 * 4e6 / dt / 1e9))
PY
