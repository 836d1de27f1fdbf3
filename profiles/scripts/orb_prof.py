"""ORB extractor: e2e time of lorb_orb_extract (host image in, keypoints + descriptors out) next to
the compiled reference on the host; also the driver for ncu launch lists of the extractor kernels."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth  # noqa: E402

g = np.load("tests/golden/orb_golden.npz")
pattern = g["orb/pattern"].astype(np.int32)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
with capi.Context(0) as ctx:
    for (w, h, nf) in ((640, 480, 1000), (752, 480, 2000), (1241, 376, 2000)):
        img = synth.make_orb_image(0, w, h)
        for _ in range(5):
            r = ctx.orb_extract(img, pattern, nfeatures=nf)
        t = time.perf_counter()
        for _ in range(reps):
            r = ctx.orb_extract(img, pattern, nfeatures=nf)
        dt = (time.perf_counter() - t) / reps
        line = "%dx%d nfeatures %d: %d keypoints, %.3f ms per frame (%.0f frames/s)" % (w, h, nf, r["n"], dt * 1e3, 1 / dt)
        try:
            from oracle import reflib
            if reflib.available() and reps > 1:
                t = time.perf_counter()
                for _ in range(3):
                    reflib.orb_extract(img, nfeatures=nf)
                line += "; compiled reference (1 core, stand-in OpenCV) %.1f ms" % ((time.perf_counter() - t) / 3 * 1e3)
        except Exception as e:  # noqa: BLE001
            line += "; reference unavailable (%s)" % e
        print(line)

    st0 = synth.make_stereo_pair(8, 0)
    left, right = st0["pyr_left"][0], st0["pyr_right"][0]
    for _ in range(5):
        r = ctx.stereo_frame(left, right, pattern, float(st0["mbf"]), float(st0["mb"]))
    t = time.perf_counter()
    for _ in range(reps):
        r = ctx.stereo_frame(left, right, pattern, float(st0["mbf"]), float(st0["mb"]))
    dt = (time.perf_counter() - t) / reps
    print("stereo frame 640x480 x2, 1000 features each: %d + %d keypoints, %d stereo matches, %.3f ms per frame pair"
          % (r[0]["n"], r[1]["n"], r[4], dt * 1e3))

# aggregate throughput with several host threads, each with its own context (ctypes drops the GIL)
import threading


def _worker(n_iter, out, k):
    with capi.Context(0) as c2:
        im = synth.make_orb_image(k)
        for _ in range(10):
            c2.orb_extract(im, pattern)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(n_iter):
            c2.orb_extract(im, pattern)
        out[k] = time.perf_counter() - t0


if reps > 1:
    for nt in (1, 2, 4, 8):
        barrier = threading.Barrier(nt)
        out = {}
        th = [threading.Thread(target=_worker, args=(reps, out, k)) for k in range(nt)]
        [t.start() for t in th]
        [t.join() for t in th]
        print("%d host threads x own context: %.0f frames/s aggregate (640x480, 1000 features)"
              % (nt, nt * reps / max(out.values())))
