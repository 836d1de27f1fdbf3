import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from lorb_slam_b200 import capi, synth
c = capi.Context(0)
fr = synth.make_frame(2000, 0); pts = synth.make_proj_points(fr, 5000, 0)
for _ in range(3): r = c.search_proj_points(fr, pts, 15.0)
print(r["n_matches"], r["n_candidates"])
