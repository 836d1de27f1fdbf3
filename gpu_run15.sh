python -m pytest tests/test_match_proj_gpu.py tests/test_host_dropin_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -5
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
fr = synth.make_frame(2000, 0); pts = synth.make_proj_points(fr, 5000, 0)
for th in (1.0, 15.0):
    c.search_proj_points(fr, pts, th)
    t0 = time.time()
    for _ in range(50): r = c.search_proj_points(fr, pts, th)
    print("proj points th", th, "e2e: %.1f us" % ((time.time() - t0) / 50 * 1e6), r["n_candidates"])
cur, last = synth.make_frame_pair(2000, 0)
for th in (15.0, 30.0):
    c.search_proj_frame(cur, last, th)
    t0 = time.time()
    for _ in range(50): r = c.search_proj_frame(cur, last, th)
    print("proj frame th", th, "e2e: %.1f us" % ((time.time() - t0) / 50 * 1e6), r["n_candidates"])
PY
