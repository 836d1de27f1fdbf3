"""Generate tests/golden/orb_golden.npz: the ORBextractor stages.

Run in the BUILD container (needs /root/reference to compile oracle/_ref, and cv2):
    python tests/golden/make_orb_golden.py

orb/<case>/angle, desc   IC_Angle + computeOrbDescriptor of the REFERENCE'S OWN src/ORBextractor.cpp
                         (compiled unmodified into oracle/_ref/libref.so through
                         oracle/ref_orb_harness.cpp) on the seeded inputs of tests/ref_cases.py
orb/pattern, orb/umax    the tables its ORBextractor constructor builds (bit_pattern_31_, umax)
atan2/*                  cv2.fastAtan2 (cv2 4.13) on integer moments: pins the restated cv::fastAtan2
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_cases as RC  # noqa: E402
from oracle import reflib as R  # noqa: E402


def main():
    import cv2
    out = {}
    for c in RC.ORB_DESCRIBE:
        oi = RC.orb_describe_case(c)
        ang, desc, pattern, umax = R.orb_describe(oi)
        k = "orb/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(oi))
        out[k + "/angle"] = ang
        out[k + "/desc"] = desc
        out["orb/pattern"] = pattern.astype(np.int8)
        out["orb/umax"] = umax.astype(np.int32)
    rng = np.random.default_rng(77)
    n = 20000
    y = rng.integers(-2700000, 2700000, n).astype(np.float32)
    x = rng.integers(-2700000, 2700000, n).astype(np.float32)
    y[:200] = rng.integers(-3, 4, 200)
    x[:200] = rng.integers(-3, 4, 200)
    y[200:300] = x[200:300]  # the |x| == |y| branch boundary
    out["atan2/y"], out["atan2/x"] = y, x
    out["atan2/deg"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "orb_golden.npz"), **out)
    print("orb_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
