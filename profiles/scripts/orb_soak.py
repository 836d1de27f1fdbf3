"""Soak: lorb_orb_extract against the compiled reference (oracle/_ref) on many random frames of
varied size, texture, feature budget and thresholds.  Prints the number of frames that differ."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth  # noqa: E402
from oracle import reflib  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
pattern = np.load("tests/golden/orb_golden.npz")["orb/pattern"].astype(np.int32)
rng = np.random.default_rng(int(os.environ.get("LORB_SOAK_SEED", 2026)))
bad = 0
t0 = time.time()
with capi.Context(0) as ctx:
    for f in range(n_frames):
        if os.environ.get("LORB_SOAK_EXTREME"):  # odd, small and large frames, big budgets
            w = int(rng.choice([161, 333, 641, 1601, 1920, 2047]))
            h = int(rng.choice([131, 241, 479, 1080, 1201]))
        else:
            w = int(rng.choice([320, 376, 512, 640, 752, 848, 1024, 1241, 1280]))
            h = int(rng.choice([240, 376, 480, 512, 720]))
        if round(np.float32(w - 32) / np.float32(h - 32)) < 1 or w < h:
            continue  # portrait frames: the reference's own extractor is out of its domain (it crashed on 641 x 1201 noise)
        img = synth.make_orb_image(1000 + f + 7919 * (int(os.environ.get("LORB_SOAK_SEED", 2026)) - 2026), w, h)
        kind = f % 4
        if kind == 1:  # low contrast everywhere: minThFAST cells
            img = (128 + (img.astype(np.float32) - 128) * 0.15).astype(np.uint8)
        elif kind == 2:  # pure noise: very many candidates
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 3:  # half flat
            img = img.copy()
            img[:, : w // 2] = 100
        nf = int(rng.choice([100, 500, 1000, 2000, 4000] + ([17, 8000] if os.environ.get("LORB_SOAK_EXTREME") else [])))
        nl = int(rng.choice([4, 8]))
        ini, mn = (20, 7) if f % 3 else (int(rng.integers(8, 40)), int(rng.integers(2, 8)))
        b = reflib.orb_extract(img, nfeatures=nf, nlevels=nl, ini_th=ini, min_th=mn)
        try:
            a = ctx.orb_extract(img, pattern, nfeatures=nf, nlevels=nl, ini_th=ini, min_th=mn)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("ERROR frame", f, w, h, kind, nf, nl, ini, mn, "reference n =", b["n"], b["n_per_level"], str(e)[-80:])
            continue
        ok = a["n"] == b["n"] and all(np.array_equal(a[k], b[k]) for k in ("x", "y", "octave", "angle", "response", "size", "desc"))
        if not ok:
            bad += 1
            print("MISMATCH frame", f, w, h, kind, nf, nl, ini, mn, a["n"], b["n"])
print("%d frames, %d differ, %.1f s" % (n_frames, bad, time.time() - t0))
