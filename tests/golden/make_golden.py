"""Generate tests/golden/*.npz from the executable pins available in this image.

Run in the BUILD container (needs cv2; nothing here reads /root/reference):
    python tests/golden/make_golden.py

Pins produced
  bf_golden.npz    cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match and
                   knnMatch(k=2) results (cv2 4.13 — the OpenCV routine the
                   reference calls at src/matcher.cpp:36-39; the reference pins
                   OpenCV 3.1, CMakeLists.txt:10) on ORB-like, uniform and
                   tie-stress descriptor sets.
  gemm_golden.npz  cv2.gemm float32 results for the two small products the
                   projection search relies on (src/matcher.cpp:74-83, :100-101).
The committed .npz files are what the tests read; the GPU box has no need for cv2.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from lorb_slam_b200 import synth  # noqa: E402


def orb_like_pair(seed=0):
    """SURVEY §8(d) cfg 1: ORB(1000, 1.2, 8) on a synthetic 640x480 image and its warp."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    for _ in range(400):
        x, y = int(rng.integers(0, 600)), int(rng.integers(0, 440))
        w, h = int(rng.integers(8, 60)), int(rng.integers(8, 60))
        cv2.rectangle(img, (x, y), (x + w, y + h), int(rng.integers(0, 256)), -1)
    M = cv2.getRotationMatrix2D((320, 240), 3.0, 1.02)
    M[:, 2] += (4, -3)
    img2 = cv2.warpAffine(img, M, (640, 480))
    orb = cv2.ORB_create(nfeatures=1000, scaleFactor=1.2, nlevels=8, fastThreshold=20)
    _, d1 = orb.detectAndCompute(img, None)
    _, d2 = orb.detectAndCompute(img2, None)
    return d1, d2


def bf_case(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t) if len(q) and len(t) else []
    mq = np.array([x.queryIdx for x in m], np.int32)
    mt = np.array([x.trainIdx for x in m], np.int32)
    md = np.array([int(x.distance) for x in m], np.int32)
    k = min(2, len(t))
    kn_i = np.full((len(q), 2), -1, np.int32)
    kn_d = np.full((len(q), 2), 256, np.int32)
    if k > 0 and len(q):
        kn = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, k=k)
        for i, row in enumerate(kn):
            for j, x in enumerate(row):
                kn_i[i, j] = x.trainIdx
                kn_d[i, j] = int(x.distance)
    return mq, mt, md, kn_i, kn_d


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    cases = []
    d1, d2 = orb_like_pair(0)
    cases.append(("orb", d1, d2))
    cases.append(("uniform", synth.descriptors_uniform(1000, rng), synth.descriptors_uniform(1000, rng)))
    cases.append(("ragged", synth.descriptors_uniform(37, rng), synth.descriptors_uniform(501, rng)))
    cases.append(("one_train", synth.descriptors_uniform(5, rng), synth.descriptors_uniform(1, rng)))
    cases.append(("identical", d1[:300], d1[:300].copy()))
    for i in range(12):
        nq, nt = (int(v) for v in rng.integers(1, 300, 2))
        cases.append((f"tie{i}", synth.descriptors_tie_stress(nq, rng, int(rng.integers(1, 4))),
                      synth.descriptors_tie_stress(nt, rng, int(rng.integers(1, 4)))))
    names = []
    for name, q, t in cases:
        mq, mt, md, ki, kd = bf_case(q, t)
        names.append(name)
        out[f"{name}_q"], out[f"{name}_t"] = q, t
        out[f"{name}_mq"], out[f"{name}_mt"], out[f"{name}_md"] = mq, mt, md
        out[f"{name}_ki"], out[f"{name}_kd"] = ki, kd
        print(name, q.shape, t.shape, "matches", len(mq), "min", md.min() if len(md) else None)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "bf_golden.npz"), **out)

    # --- cv2.gemm pins for the projection arithmetic
    n = 4000
    R = synth.rodrigues(rng.normal(0, 0.4, (n, 3))).astype(np.float32)
    t = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
    x = (rng.normal(0, 6.0, (n, 3)) + [0, 0, 8]).astype(np.float32)
    xc = np.zeros((n, 3), np.float32)
    twc = np.zeros((n, 3), np.float32)
    for i in range(n):
        # x3Dc = Rcw*x3Dw + tcw  (src/matcher.cpp:101) -> gemm(R, x, 1, t, 1)
        xc[i] = cv2.gemm(R[i], x[i].reshape(3, 1), 1.0, t[i].reshape(3, 1), 1.0).ravel()
        # twc = -Rcw.t()*tcw (src/matcher.cpp:77): MatExpr materialises the transpose,
        # then gemm(Rt, tcw, alpha=-1)
        twc[i] = cv2.gemm(np.ascontiguousarray(R[i].T), t[i].reshape(3, 1), -1.0, None, 0.0).ravel()
    np.savez_compressed(os.path.join(HERE, "gemm_golden.npz"), R=R, t=t, x=x, xc=xc, twc=twc)
    print("gemm golden", n)


if __name__ == "__main__":
    main()
