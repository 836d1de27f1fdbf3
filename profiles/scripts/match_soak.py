"""Soak: the matching entry points against the compiled reference (oracle/_ref) on random inputs."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from lorb_slam_b200 import capi, synth  # noqa: E402
from oracle import ref, reflib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
legs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["entry", "sweep"]  # entry points / keyframe-pair sweep
verbose = len(sys.argv) > 3
rng = np.random.default_rng(int(os.environ.get("LORB_SOAK_SEED", 31)))
bad = {"proj_points": 0, "proj_frame": 0, "bf": 0, "knn2": 0, "mp_desc": 0, "frustum": 0, "stereo": 0, "sweep_tensor": 0, "sweep_popc": 0}
with capi.Context(0) as ctx:
    for s in range(n if "entry" in legs else 0):
        nk, npt = int(rng.integers(50, 3000)), int(rng.integers(50, 6000))
        fr = synth.make_frame(nk, 700 + s, stereo=bool(s % 2), claimed_frac=float(rng.choice([0, 0.2, 0.5])))
        pts = synth.make_proj_points(fr, npt, 700 + s, nobs=(0, 1, 2), inactive_frac=0.1)
        th = float(rng.choice([1.0, 2.0, 7.0, 15.0, 30.0, 60.0]))
        a, b = ctx.search_proj_points(fr, pts, th), reflib.search_proj_points(fr, pts, th)
        bad["proj_points"] += not (a["n_matches"] == b["n_matches"] and np.array_equal(a["point_for_kp"], b["point_for_kp"]))
        cur, last = synth.make_frame_pair(nk, 800 + s, motion=str(rng.choice(["forward", "backward", "still"])))
        cur["kp_claim_obs"] = np.where(rng.random(nk) < 0.2, rng.integers(0, 3, nk), -1).astype(np.int32)
        a, b = ctx.search_proj_frame(cur, last, th), reflib.search_proj_frame(cur, last, th)
        bad["proj_frame"] += not (a["n_matches"] == b["n_matches"] and np.array_equal(
            reflib.final_state_from_oracle(a["state_for_kp"], cur["kp_claim_obs"]), b["state_for_kp"]))
        big = 6000 if s % 5 == 0 else 1500  # every fifth round beyond one tile of the brute-force kernels
        q = synth.descriptors_uniform(int(rng.integers(1, big)), rng)
        t = synth.descriptors_noisy_copy(q[rng.integers(0, len(q), int(rng.integers(1, big)))], rng, 0.05)
        if s % 7 == 3:  # ties everywhere
            q, t = synth.descriptors_tie_stress(len(q), rng, 2, 3), synth.descriptors_tie_stress(len(t), rng, 2, 3)
        a, o = ctx.match_bf_crosscheck(q, t), ref.bf_crosscheck(q, t)
        bad["bf"] += not (np.array_equal(a["q"], o["q"]) and np.array_equal(a["t"], o["t"]) and np.array_equal(a["dist"], o["dist"]))
        ki, kd, _ = ctx.match_knn2(q, t)
        oi, od = ref.knn2(q, t)
        bad["knn2"] += not (np.array_equal(ki, oi) and np.array_equal(kd, od))
        sizes = rng.integers(0, 60, int(rng.integers(1, 400))).astype(np.int32)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        dsc = synth.descriptors_tie_stress(int(offs[-1]), rng, 3, 4) if s % 2 else synth.descriptors_uniform(max(1, int(offs[-1])), rng)[:int(offs[-1])]
        gb, gm = ctx.compute_descriptors(offs, dsc)
        ob, om = ref.compute_descriptors(offs, dsc)
        bad["mp_desc"] += not (np.array_equal(gb, ob) and np.array_equal(gm, om))
        fp = synth.make_frustum_points(int(rng.integers(10, 8000)), 900 + s)
        b, ow, lsf = reflib.frustum_project(fp)
        fp2 = dict(fp, ow=ow, log_sf=float(lsf))  # the reference's own camera centre (mTcw.inv())
        a = ctx.frustum_project(fp2)
        iv = b["in_view"].astype(bool)
        bad["frustum"] += not (np.array_equal(a["in_view"].astype(bool), iv) and all(
            np.array_equal(a[k][iv], b[k][iv]) for k in ("proj_x", "proj_y", "proj_xr", "level", "view_cos")))
        st = synth.make_stereo_pair(int(rng.integers(100, 2500)), 1000 + s)
        a, b = ctx.stereo_matches(st), reflib.stereo_matches(st)
        bad["stereo"] += not (a["n_matched"] == b["n_matched"] and np.array_equal(a["uright"], b["uright"]) and np.array_equal(a["depth"], b["depth"]))
    for s in range(n if "sweep" in legs else 0):
        # keyframe-pair sweep, both kernels: random bank shape, random (repeated, unordered, self) pairs;
        # every third bank is low-entropy (ties everywhere), every third has near-duplicate keyframes
        n_kf, n_desc = int(rng.integers(2, 12)), int(rng.choice([1, 2, 31, 128, 255, 256, 257, 2047, 2048, int(rng.integers(1, 2049))]))
        if s % 3 == 0:
            bank = np.stack([synth.descriptors_tie_stress(n_desc, rng, int(rng.integers(1, 4)), int(rng.integers(2, 5))) for _ in range(n_kf)])
        elif s % 3 == 1:
            base = synth.descriptors_uniform(n_desc, rng)
            bank = np.stack([synth.descriptors_noisy_copy(base[rng.permutation(n_desc)], rng, 0.03) for _ in range(n_kf)])
        else:
            bank = synth.kf_bank(n_kf, n_desc, seed=1100 + s)
        npair = int(rng.integers(1, 40))
        if verbose:
            print("sweep round", s, "n_kf", n_kf, "n_desc", n_desc, "pairs", npair, flush=True)
        pa, pb = rng.integers(0, n_kf, npair).astype(np.int32), rng.integers(0, n_kf, npair).astype(np.int32)
        want = ref.sweep(bank, pa, pb)
        for impl in ("tensor", "popc"):
            ctx.sweep_set_impl(impl)
            got = ctx.match_sweep(bank, pa, pb)
            bad["sweep_" + impl] += not all(np.array_equal(g, w) for g, w in zip(got, want))
        ctx.sweep_set_impl("tensor")
print(n, "rounds; mismatches:", bad)
