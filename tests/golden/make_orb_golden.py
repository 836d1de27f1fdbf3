"""Generate tests/golden/orb_golden.npz: the ORBextractor stages.

Run in the BUILD container (needs /root/reference to compile oracle/_ref, and cv2):
    python tests/golden/make_orb_golden.py

orb/<case>/angle, desc   IC_Angle + computeOrbDescriptor of the REFERENCE'S OWN src/ORBextractor.cpp
                         (compiled unmodified into oracle/_ref/libref.so through
                         oracle/ref_orb_harness.cpp) on the seeded inputs of tests/ref_cases.py
orb/pattern, orb/umax    the tables its ORBextractor constructor builds (bit_pattern_31_, umax)
atan2/*                  cv2.fastAtan2 (cv2 4.13) on integer moments: pins the restated cv::fastAtan2
ext/<case>/*             the REFERENCE'S OWN ORBextractor::operator() (same library; its cv::resize /
                         GaussianBlur / FAST calls go to the cv2-pinned stand-ins of oracle/orb_ref.cpp,
                         its allocations to a bump arena so that the quadtree's address-ordered
                         tie-break is creation order, see oracle/ref_orb_harness.cpp)
cv/*                     cv2 4.13 itself on the image operations: cv2.resize chain, cv2.GaussianBlur
                         (sha256 per level) and the cell loop over cv2.FastFeatureDetector (arrays)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_cases as RC  # noqa: E402
from oracle import reflib as R  # noqa: E402


def cv2_candidates(cv2, img, ini=20, mn=7):
    """The detection loop of ComputeKeyPointsOctTree (:808-862) with cv2's own FAST."""
    rows, cols = img.shape
    min_b, max_bx, max_by = 16, cols - 16, rows - 16
    width, height = np.float32(max_bx - min_b), np.float32(max_by - min_b)
    n_c, n_r = int(width / np.float32(30)), int(height / np.float32(30))
    w_c, h_c = int(np.ceil(width / n_c)), int(np.ceil(height / n_r))
    fi = cv2.FastFeatureDetector_create(threshold=ini, nonmaxSuppression=True)
    fm = cv2.FastFeatureDetector_create(threshold=mn, nonmaxSuppression=True)
    out = []
    for i in range(n_r):
        ini_y = min_b + i * h_c
        if ini_y >= max_by - 3:
            continue
        max_y = min(ini_y + h_c + 6, max_by)
        for j in range(n_c):
            ini_x = min_b + j * w_c
            if ini_x >= max_bx - 6:
                continue
            max_x = min(ini_x + w_c + 6, max_bx)
            sub = np.ascontiguousarray(img[ini_y:max_y, ini_x:max_x])
            k = fi.detect(sub) or fm.detect(sub)
            out += [(p.pt[0] + j * w_c, p.pt[1] + i * h_c, p.response) for p in k]
    return np.array(out, np.float32).reshape(-1, 3).astype(np.uint16)


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
    for c in RC.ORB_EXTRACT:
        img = RC.orb_extract_case(c)
        r = R.orb_extract(img, nfeatures=c[4])
        k = "ext/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(img))
        for f in ("x", "y", "octave", "angle", "response", "size", "desc"):
            out[k + "/" + f] = r[f]
    # cv2 itself on the image operations (first two cases)
    for c in RC.ORB_EXTRACT[:3]:
        img = RC.orb_extract_case(c)
        k = "cv/" + c[0]
        lvl, sums_raw, sums_blur = img, [], []
        for l in range(8):
            if l:
                sf = np.float32(1.0)  # mvScaleFactor[l] as the constructor builds it (:428-436)
                for _ in range(l):
                    sf = np.float32(sf * np.float32(1.2))
                isf = np.float32(1.0) / sf
                sz = (int(np.rint(np.float32(img.shape[1]) * isf)), int(np.rint(np.float32(img.shape[0]) * isf)))
                lvl = cv2.resize(lvl, sz, interpolation=cv2.INTER_LINEAR)
            sums_raw.append(hashlib.sha256(np.ascontiguousarray(lvl).tobytes()).hexdigest())
            sums_blur.append(hashlib.sha256(np.ascontiguousarray(
                cv2.GaussianBlur(lvl, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)).tobytes()).hexdigest())
            if l in (0, 3, 7):
                cand = cv2_candidates(cv2, lvl)
                out[k + "/cand%d" % l] = cand
        out[k + "/sha_raw"] = np.array(sums_raw)
        out[k + "/sha_blur"] = np.array(sums_blur)
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
