"""Soak: lorb_stereo_frame against the compiled reference's extractor + Frame::ComputeStereoMatches."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth  # noqa: E402
from oracle import reflib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
pattern = np.load("tests/golden/orb_golden.npz")["orb/pattern"].astype(np.int32)
rng = np.random.default_rng(int(os.environ.get("LORB_SOAK_SEED", 5)))
bad = 0
with capi.Context(0) as ctx:
    for s in range(n):
        w, h = int(rng.choice([512, 640, 752, 1024])), int(rng.choice([376, 480]))
        nf = int(rng.choice([500, 1000, 2000]))
        st0 = synth.make_stereo_pair(8, 300 + s + 1009 * (int(os.environ.get("LORB_SOAK_SEED", 5)) - 5), w, h)
        left, right = st0["pyr_left"][0], st0["pyr_right"][0]
        mbf, mb = float(st0["mbf"]), float(st0["mb"])
        L, R, ur, dp, nm = ctx.stereo_frame(left, right, pattern, mbf, mb, nfeatures=nf)
        ra, rb = reflib.orb_extract(left, nfeatures=nf), reflib.orb_extract(right, nfeatures=nf)
        _, _, _, sf = ctx.orb_level_sizes(w, h, nfeatures=nf)
        stg_l, stg_r = ctx.orb_stages(left, nfeatures=nf), ctx.orb_stages(right, nfeatures=nf)
        st = dict(n_levels=8, pyr_left=stg_l["raw"], pyr_right=stg_r["raw"], scale_factors=sf,
                  inv_scale_factors=(np.float32(1.0) / sf).astype(np.float32), mbf=mbf, mb=mb, fx=float(st0["fx"]),
                  n_left=ra["n"], lx=ra["x"], ly=ra["y"], loct=ra["octave"], ldesc=ra["desc"],
                  n_right=rb["n"], rx=rb["x"], ry=rb["y"], roct=rb["octave"], rdesc=rb["desc"])
        rs = reflib.stereo_matches(st)
        ok = (L["n"] == ra["n"] and R["n"] == rb["n"] and np.array_equal(L["desc"], ra["desc"])
              and np.array_equal(R["desc"], rb["desc"]) and np.array_equal(L["x"], ra["x"]) and nm == rs["n_matched"]
              and np.array_equal(ur, rs["uright"]) and np.array_equal(dp, rs["depth"]))
        if not ok:
            bad += 1
            print("MISMATCH", s, w, h, nf, L["n"], ra["n"], nm, rs["n_matched"])
print("%d stereo pairs, %d differ" % (n, bad))
