import sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import ref_cases as RC
from lorb_slam_b200 import capi, synth
from oracle import ref
bad = 0
with capi.Context(0) as ctx:
    for c in RC.QUADTREE:
        args = RC.quadtree_case(c)
        a = ref.orb_distribute(*args)
        b = ctx.orb_distribute(*args)
        ok = np.array_equal(a, b)
        print(c[0], len(a), len(b), ok)
        bad += not ok
    rng = np.random.default_rng(7)
    n_bad = 0
    for t in range(300):
        n, w, h = int(rng.integers(1, 3000)), int(rng.integers(100, 1300)), int(rng.integers(60, 700))
        if round(np.float32(w) / np.float32(h)) < 1:
            continue
        span = w - 6 if t % 5 else 30
        x = rng.integers(0, span, n).astype(np.float32); y = rng.integers(0, min(span, h - 6), n).astype(np.float32)
        r = rng.integers(7, 60, n).astype(np.float32); nf = int(rng.integers(1, 1500))
        a = ref.orb_distribute(x, y, r, 16, 16 + w, 16, 16 + h, nf)
        try:
            b = ctx.orb_distribute(x, y, r, 16, 16 + w, 16, 16 + h, nf)
        except Exception as e:
            print("ERROR trial", t, n, w, h, nf, span, str(e)[-60:])
            n_bad += 1
            continue
        if not np.array_equal(a, b):
            n_bad += 1
            if n_bad < 4:
                m = min(len(a), len(b)); d = np.flatnonzero(a[:m] != b[:m])
                print("MISMATCH trial", t, n, w, h, nf, len(a), len(b), d[:5])
    print("random mismatches:", n_bad)
    img = synth.make_orb_image(0)
    x, y, r = ref.orb_level_candidates(img)
    for _ in range(5): ctx.orb_distribute(x, y, r, 16, 624, 16, 464, 217)
    t0 = time.perf_counter()
    for _ in range(200): ctx.orb_distribute(x, y, r, 16, 624, 16, 464, 217)
    print("gpu quadtree e2e (incl. pack/H2D/D2H) %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
