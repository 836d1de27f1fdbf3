"""Run by tests/test_fallback_paths_gpu.py in a subprocess whose environment selects the library's
fallback kernels (the ones a device without cooperative launch would get, and the previous Cholesky):
projection search with the single-CTA claim resolution, a 40-camera BA through the round-robin dataflow
Cholesky or the multi-kernel Cholesky.  Prints FALLBACK_PATHS_OK."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorb_slam_b200 import capi, synth  # noqa: E402
from oracle import ref  # noqa: E402

with capi.Context(0) as ctx:
    # projection search, both forms, a small case (claims in shared memory), a large one and one that
    # overflows the first-attempt candidate arena
    for nk, npt, th in ((2000, 5000, 3.0), (3000, 2500, 30.0), (15000, 3000, 1.0)):
        fr = synth.make_frame(nk, seed=nk, stereo=True, claimed_frac=0.2)
        pts = synth.make_proj_points(fr, npt, seed=nk, nobs=(0, 1, 2), inactive_frac=0.1)
        r, o = ctx.search_proj_points(fr, pts, th), ref.search_proj_points(fr, pts, th)
        assert r["n_candidates"] == o["n_candidates"] and r["n_matches"] == o["n_matches"], (nk, npt, th)
        assert np.array_equal(r["kp_for_point"], o["kp_for_point"]) and np.array_equal(r["point_for_kp"], o["point_for_kp"])
    for nk, th in ((2000, 15.0), (3000, 120.0)):
        cur, last = synth.make_frame_pair(nk, seed=5, motion="forward", stereo=True)
        r, o = ctx.search_proj_frame(cur, last, th), ref.search_proj_frame(cur, last, th)
        assert r["n_matches"] == o["n_matches"] and np.array_equal(r["kp_for_item"], o["kp_for_item"])
        assert np.array_equal(r["state_for_kp"], o["state_for_kp"])
    # reduced solve of a 40-camera window (n = 240: the tiled Cholesky)
    pb = synth.make_ba_problem(3, C=40, P=3000, obs_per_point=(5, 6, 7), traj_len=12.0)
    cams, pts, s = ctx.ba_local(pb, capi.ba_options(max_num_iterations=6))
    oc, op, o = ref.ba_local(pb, ref.ba_options(max_num_iterations=6))
    np.testing.assert_allclose(cams, oc, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(pts, op, rtol=1e-6, atol=1e-8)
    assert s["iterations"] == o["iterations"] and s["termination"] == o["termination"]
    # brute-force matching and the popc sweep kernel (LORB_HAMMING_CSA=0 selects the plain 8-popc body)
    rng = np.random.default_rng(2)
    q, t = synth.descriptors_uniform(1300, rng), synth.descriptors_tie_stress(2100, rng, 2, 3)
    a, o = ctx.match_bf_crosscheck(q, t), ref.bf_crosscheck(q, t)
    assert np.array_equal(a["q"], o["q"]) and np.array_equal(a["t"], o["t"]) and np.array_equal(a["dist"], o["dist"])
    ki, kd, _ = ctx.match_knn2(q, t)
    oi, od = ref.knn2(q, t)
    assert np.array_equal(ki, oi) and np.array_equal(kd, od)
    bank = synth.kf_bank(7, 777, seed=4)
    pa, pb2 = synth.all_pairs(7)
    ctx.sweep_set_impl("popc")
    got, want = ctx.match_sweep(bank, pa, pb2), ref.sweep(bank, pa, pb2)
    assert all(np.array_equal(g, w) for g, w in zip(got, want))
print("FALLBACK_PATHS_OK")
