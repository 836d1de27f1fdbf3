"""CPU: oracle/_ref -- the reference's own sources compiled against the OpenCV / Ceres stand-ins.

 * the stand-in arithmetic is pinned bit-exactly against cv2 golden vectors
   (tests/golden/cvshim_golden.npz) and the cv2-pinned BF golden (bf_golden.npz);
 * the committed ref_golden.npz is what this build produces today (guards stale vectors);
 * beyond the committed cases, the restated oracle agrees with the compiled reference on fresh
   random inputs (bit-exact for matching, 1e-9 relative for BA).
Skipped where neither the prebuilt library nor /root/reference is present."""
import os

import numpy as np
import pytest

import ref_cases as RC
import ref_checks as CK
from lorb_slam_b200 import synth
from oracle import ref
from oracle import reflib as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref absent and /root/reference not here")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_standin_arithmetic_pinned_to_cv2():
    g = np.load(os.path.join(GOLD, "cvshim_golden.npz"))
    for i in range(len(g["rvec"])):
        assert np.array_equal(R.cv_rodrigues(g["rvec"][i]), g["rmat"][i]), i
        assert np.array_equal(R.cv_inv4(g["T"][i]), g["Tinv"][i]), i
        assert np.array_equal(R.cv_rt(g["A"][i], g["x"][i], g["t"][i]), g["y"][i]), i


def test_standin_bfmatcher_pinned_to_cv2():
    g = np.load(os.path.join(GOLD, "bf_golden.npz"))
    names = sorted({k[:-2] for k in g.files if k.endswith("_q")})
    assert len(names) >= 5
    for nm in names:
        oq, ot, od = R.cv_bfmatch(g[nm + "_q"], g[nm + "_t"])
        assert np.array_equal(oq, g[nm + "_mq"]) and np.array_equal(ot, g[nm + "_mt"]), nm
        assert np.array_equal(od, g[nm + "_md"].astype(np.int32)), nm


def test_reference_scalars():
    assert R.constants() == dict(TH_LOW=50, TH_HIGH=100, HISTO_LENGTH=30)
    # float(0.998) lies above the double literal 0.998 of src/matcher.cpp:432 -> already the small radius
    assert R.radius_by_viewing_cos(0.9985) == 2.5 and R.radius_by_viewing_cos(0.998) == 2.5
    assert R.radius_by_viewing_cos(np.nextafter(np.float32(0.998), np.float32(0))) == 4.0
    rng = np.random.default_rng(0)
    for _ in range(500):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert R.descriptor_distance(a, b) == ref.hamming256(a, b) == int(np.unpackbits(a ^ b).sum())
        h = rng.integers(0, rng.choice([3, 10, 100]), 30)
        assert R.three_maxima(h) == ref.three_maxima(h)


def test_golden_is_what_the_build_produces():
    g = CK.golden()
    c = RC.PROJ_POINTS[2]
    fr, pts, th = RC.proj_points_case(c)
    r = R.search_proj_points(fr, pts, th)
    assert np.array_equal(r["point_for_kp"], g["pp/" + c[0] + "/point_for_kp"])
    c = RC.PROJ_FRAME[1]
    cur, last, th = RC.proj_frame_case(c)
    assert np.array_equal(R.search_proj_frame(cur, last, th)["state_for_kp"], g["pf/" + c[0] + "/state_for_kp"])
    c = RC.BA_LOCAL[1]
    r = R.ba_local(RC.ba_local_case(c), RC.ba_case_options(c, R.ba_options))
    assert np.array_equal(r["cams_f64"], g["bl/" + c[0] + "/cams_f64"])


@pytest.mark.parametrize("seed", range(3))
def test_oracle_vs_reference_fresh_searches(seed):
    rng = np.random.default_rng(100 + seed)
    fr = synth.make_frame(1500, 50 + seed, stereo=bool(seed % 2), claimed_frac=0.25)
    # border keypoints round to cell 64 / 48 and vanish from the grid (src/frame.cpp:107-112)
    fr["kp_x"][:40] = np.linspace(633.0, 639.99, 40).astype(np.float32)
    fr["kp_y"][40:80] = np.linspace(473.0, 479.99, 40).astype(np.float32)
    pts = synth.make_proj_points(fr, 3000, 50 + seed, nobs=(0, 1, 2), inactive_frac=0.1)
    pts["proj_x"][:30] = rng.uniform(-60, 0, 30).astype(np.float32)
    pts["proj_x"][30:60] = rng.uniform(640, 700, 30).astype(np.float32)
    for th in (1.0, 2.0, 15.0):
        a, b = R.search_proj_points(fr, pts, th), ref.search_proj_points(fr, pts, th)
        assert a["n_matches"] == b["n_matches"]
        assert np.array_equal(a["point_for_kp"], b["point_for_kp"])
    for motion in ("forward", "backward", "still"):
        cur, last = synth.make_frame_pair(1200, 70 + seed, motion=motion)
        cur["kp_claim_obs"] = np.where(rng.random(1200) < 0.2, rng.integers(0, 3, 1200), -1).astype(np.int32)
        for th in (15.0, 30.0):
            a, b = R.search_proj_frame(cur, last, th), ref.search_proj_frame(cur, last, th)
            assert a["n_matches"] == b["n_matches"]
            assert np.array_equal(a["state_for_kp"], R.final_state_from_oracle(b["state_for_kp"], cur["kp_claim_obs"]))
    for _ in range(300):
        x, y = rng.uniform(-30, 670), rng.uniform(-30, 510)
        r, mn, mx = rng.choice([2.5, 4, 10, 30, 80]), int(rng.integers(-1, 8)), int(rng.integers(-1, 8))
        assert np.array_equal(R.features_in_area(fr, x, y, r, mn, mx), ref.features_in_area(fr, x, y, r, mn, mx))


def test_oracle_vs_reference_fresh_ba():
    rel = lambda x, y: np.abs(x - y).max() / np.abs(y).max()  # noqa: E731
    pb = synth.make_ba_problem(11, C=5, P=160, fixed_frac=0.15, obs_per_point=(4,))
    for opt_kw in ({}, dict(max_num_iterations=3), dict(initial_trust_region_radius=1.0)):
        a = R.ba_local(pb, R.ba_options(**opt_kw))
        c, p, s = ref.ba_local(pb, ref.ba_options(**opt_kw))
        assert a["rc"] == 0
        assert rel(a["cams_f64"], c) < 1e-9 and rel(a["pts_f64"], p) < 1e-9
        for f in ("iterations", "num_successful_steps", "num_unsuccessful_steps", "termination"):
            assert a["summary"][f] == s[f]
    po = synth.make_pose_only(5, n=200)
    rt0 = po["rt"].astype(np.float32)
    a = R.ba_pose_only(po["xw"], po["uv"], po["K"], rt0)
    rt, s = ref.ba_pose_only(po["xw"], po["uv"], po["K"], rt0.astype(np.float64))
    assert rel(a["rt_f64"], rt) < 1e-9 and a["summary"]["iterations"] == s["iterations"]
    assert np.array_equal(a["rt_f32"], rt.astype(np.float32))
