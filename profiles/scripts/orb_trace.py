import sys, os
sys.path.insert(0, ".")
import numpy as np
from lorb_slam_b200 import capi, synth
g = np.load("tests/golden/orb_golden.npz"); pattern = g["orb/pattern"].astype(np.int32)
img = synth.make_orb_image(0)
with capi.Context(0) as ctx:
    for _ in range(20): ctx.orb_extract(img, pattern)
    os.environ["LORB_ORB_TRACE"] = "1"
    for _ in range(3):
        sys.stderr.write("---- call\n"); ctx.orb_extract(img, pattern)
