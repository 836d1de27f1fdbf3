python -m pytest tests/test_match_bf_gpu.py::test_compute_descriptors tests/test_match_proj_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -5
LORB_DEBUG=1 python - <<'PY' 2>&1 | sort | uniq -c | sort -rn | head -20
import sys
sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
fr = synth.make_frame(2000, 0)
for nobs in (1, 0, (0,1,2)):
    pts = synth.make_proj_points(fr, 5000, 0, nobs=nobs)
    for th in (1.0, 15.0):
        print("case", nobs, th, file=sys.stderr)
        c.search_proj_points(fr, pts, th)
for motion in ("forward", "still"):
    cur, last = synth.make_frame_pair(2000, 0, motion=motion)
    for th in (15.0, 30.0):
        print("frame", motion, th, file=sys.stderr)
        c.search_proj_frame(cur, last, th)
PY
