"""GPU: edge cases of the hot path (empty / single-element / maximum-size inputs,
degenerate BA structure) — each against the oracle."""
import numpy as np
import pytest

from lorb_slam_b200 import capi, synth
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["popc", "tensor"])
def sweep_impl(ctx, request):
    ctx.sweep_set_impl(request.param)
    yield request.param
    ctx.sweep_set_impl("default")


def test_sweep_maximum_descriptor_count_and_self_pairs(ctx, sweep_impl):
    bank = synth.kf_bank(3, 2048, seed=5)          # the shared-memory resident limit
    pa = np.array([0, 1, 2, 0], np.int32)
    pb = np.array([0, 2, 1, 2], np.int32)          # a self pair, both orders of a pair
    kept, mt, md = ctx.match_sweep(bank, pa, pb)
    ok, om, od = ref.sweep(bank, pa, pb)
    assert np.array_equal(kept, ok) and np.array_equal(mt, om) and np.array_equal(md, od)
    assert mt[0] == 2048 and md[0] == 0            # self pair: everything matches at distance 0
    with pytest.raises(capi.LorbError):
        ctx.match_sweep(np.zeros((2, 2049, 32), np.uint8), [0], [1])
    with pytest.raises(capi.LorbError):
        ctx.match_sweep(bank, [0], [7])            # pair index out of range


def test_sweep_no_pairs(ctx, sweep_impl):
    bank = synth.kf_bank(2, 64, seed=1)
    kept, mt, md = ctx.match_sweep(bank, np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert len(kept) == 0


def test_crosscheck_all_equal_descriptors(ctx):
    """Every distance ties at 0: lowest-index tie-breaks decide everything."""
    q = np.zeros((70, 32), np.uint8)
    t = np.zeros((90, 32), np.uint8)
    r, o = ctx.match_bf_crosscheck(q, t), ref.bf_crosscheck(q, t)
    assert np.array_equal(r["q"], o["q"]) and np.array_equal(r["t"], o["t"])
    assert list(r["q"]) == [0] and list(r["t"]) == [0]
    idx, dist, _ = ctx.match_knn2(q, t)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == 1).all() and (dist == 0).all()


def test_projection_single_keypoint_single_point(ctx):
    fr = synth.make_frame(1, seed=1)
    fr["kp_x"][:] = 100.0
    fr["kp_y"][:] = 100.0
    fr["kp_octave"][:] = 2
    pts = synth.make_proj_points(fr, 1, seed=1, true_frac=1.0, sigma_px=0.1, p_flip=0.0)
    pts["level"][:] = 2
    r, o = ctx.search_proj_points(fr, pts, 1.0), ref.search_proj_points(fr, pts, 1.0)
    assert r["n_matches"] == o["n_matches"] == 1
    assert np.array_equal(r["kp_for_point"], o["kp_for_point"])


def test_projection_everything_already_protected(ctx):
    fr = synth.make_frame(500, seed=2)
    fr["kp_claim_obs"][:] = 5
    pts = synth.make_proj_points(fr, 800, seed=2)
    r = ctx.search_proj_points(fr, pts, 4.0)
    assert r["n_matches"] == 0 and (r["point_for_kp"] == -1).all() and r["n_candidates"] > 0
    assert r["n_candidates"] == ref.search_proj_points(fr, pts, 4.0)["n_candidates"]


def test_ba_points_with_single_and_many_observations(ctx):
    """1 observation per point (rank-deficient 3x3 blocks rescued by the LM diagonal)
    up to 40 observations per point (five rounds of the 8-lane groups)."""
    pb = synth.make_ba_problem(12, C=48, P=300, obs_per_point=(1, 2, 3, 40), traj_len=4.0)
    kw = dict(max_num_iterations=6)
    cams, pts, s = ctx.ba_local(pb, capi.ba_options(**kw))
    oc, op, o = ref.ba_local(pb, ref.ba_options(**kw))
    np.testing.assert_allclose(cams, oc, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(pts, op, rtol=1e-6, atol=1e-8)
    assert s["iterations"] == o["iterations"] and s["termination"] == o["termination"]


def test_ba_zero_iterations_and_no_points(ctx):
    pb = synth.make_ba_problem(13, C=3, P=40, obs_per_point=(3,))
    cams, pts, s = ctx.ba_local(pb, capi.ba_options(max_num_iterations=0))
    assert s["iterations"] == 0 and s["final_cost"] == s["initial_cost"]
    assert np.array_equal(cams, pb["cams"]) and np.array_equal(pts, pb["pts"])
    empty = dict(pb, pts=np.zeros((0, 3)), obs_cam=np.zeros(0, np.int32), obs_pt=np.zeros(0, np.int32),
                 obs_uv=np.zeros((0, 2), np.float32))
    cams, pts, s = ctx.ba_local(empty, capi.ba_options(max_num_iterations=3))
    assert s["initial_cost"] == 0.0 and np.array_equal(cams, pb["cams"])


def test_pose_only_too_few_observations(ctx):
    """Under-determined (2 observations): the damped 6x6 still solves; parity with the oracle."""
    po = synth.make_pose_only(2, 2)
    rt, s = ctx.ba_pose_only(po["xw"], po["uv"], po["K"], po["rt"], capi.ba_options(max_num_iterations=5))
    ort, o = ref.ba_pose_only(po["xw"], po["uv"], po["K"], po["rt"], ref.ba_options(max_num_iterations=5))
    np.testing.assert_allclose(rt, ort, rtol=1e-6, atol=1e-8)
    rt0, s0 = ctx.ba_pose_only(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), po["K"], po["rt"])
    assert np.array_equal(rt0, po["rt"]) and s0["initial_cost"] == 0.0


def test_ba_point_observed_twice_by_one_camera(ctx):
    """Two residual blocks on the same (point, pose) pair — legal for Ceres; the
    atomic-free dense path must hand such windows to the accumulating variant."""
    pb = synth.make_ba_problem(14, C=5, P=120, obs_per_point=(3, 4))
    dup = np.flatnonzero(pb["obs_pt"] < 10)
    pb2 = dict(pb)
    pb2["obs_cam"] = np.concatenate([pb["obs_cam"], pb["obs_cam"][dup]]).astype(np.int32)
    pb2["obs_pt"] = np.concatenate([pb["obs_pt"], pb["obs_pt"][dup]]).astype(np.int32)
    pb2["obs_uv"] = np.concatenate([pb["obs_uv"], pb["obs_uv"][dup] + np.float32(0.5)]).astype(np.float32)
    kw = dict(max_num_iterations=6)
    cams, pts, s = ctx.ba_local(pb2, capi.ba_options(**kw))
    oc, op, o = ref.ba_local(pb2, ref.ba_options(**kw))
    np.testing.assert_allclose(cams, oc, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(pts, op, rtol=1e-6, atol=1e-8)
    assert s["iterations"] == o["iterations"]


def test_profile_brackets_and_fp64_microbench(ctx):
    """lorb_ctx_profile / lorb_microbench_fp64 (bench.py's roofline inputs)."""
    from lorb_slam_b200 import capi
    pb = synth.make_ba_problem(5, C=5, P=200, obs_per_point=(4,))
    ctx.profile(True)
    _, _, s = ctx.ba_local(pb, capi.ba_options(max_num_iterations=3, function_tolerance=-1.0,
                                               parameter_tolerance=-1.0, gradient_tolerance=-1.0))
    for slot in range(3):
        ms, n = ctx.profile_read(slot)
        assert n == s["iterations"] and ms > 0
    ctx.profile(False)
    ctx.ba_local(pb, capi.ba_options(max_num_iterations=2))
    assert ctx.profile_read(0)[1] == s["iterations"]  # nothing recorded while disabled
    assert ctx.microbench_fp64(0, 64) > 1e12 and ctx.microbench_fp64(1, 64) > 1e12


def test_large_path_device_work_lists_edge_cases(ctx):
    """Windows on the work-list path (6C > 96), whose lists are built on the device: duplicate
    (point, camera) observations, fixed out-of-window observers, unsorted observation order,
    cameras without observations, and a window with no points at all."""
    kw = dict(max_num_iterations=5)
    pb = synth.make_ba_problem(21, C=20, P=260, obs_per_point=(2, 3, 9), fixed_frac=0.15, traj_len=6.0)
    dup = np.flatnonzero(pb["obs_pt"] % 17 == 0)
    pb["obs_cam"] = np.concatenate([pb["obs_cam"], pb["obs_cam"][dup]]).astype(np.int32)
    pb["obs_pt"] = np.concatenate([pb["obs_pt"], pb["obs_pt"][dup]]).astype(np.int32)
    pb["obs_uv"] = np.concatenate([pb["obs_uv"], pb["obs_uv"][dup] + np.float32(0.25)]).astype(np.float32)
    perm = np.random.default_rng(5).permutation(len(pb["obs_pt"]))  # not grouped by point any more
    for k in ("obs_cam", "obs_pt", "obs_uv"):
        pb[k] = np.ascontiguousarray(pb[k][perm])
    pb["O"] = len(pb["obs_pt"])
    # cameras 18 and 19 lose all their observations (they stay in the system through the LM diagonal)
    keep = pb["obs_cam"] < 18
    for k in ("obs_cam", "obs_pt", "obs_uv"):
        pb[k] = np.ascontiguousarray(pb[k][keep])
    pb["O"] = len(pb["obs_pt"])
    cams, pts, s = ctx.ba_local(pb, capi.ba_options(**kw))
    oc, op, o = ref.ba_local(pb, ref.ba_options(**kw))
    np.testing.assert_allclose(cams, oc, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(pts, op, rtol=1e-6, atol=1e-8)
    assert s["iterations"] == o["iterations"] and s["termination"] == o["termination"]
    # run-to-run: the lists come out of a stable sort, so the solve is reproducible to the bit
    cams2, pts2, _ = ctx.ba_local(pb, capi.ba_options(**kw))
    assert np.array_equal(cams, cams2) and np.array_equal(pts, pts2)
    empty = dict(pb, pts=np.zeros((0, 3)), obs_cam=np.zeros(0, np.int32), obs_pt=np.zeros(0, np.int32),
                 obs_uv=np.zeros((0, 2), np.float32), fix_pt=np.zeros(0, np.int32),
                 fix_uv=np.zeros((0, 2), np.float32), fix_rt=np.zeros((0, 6), np.float32), O=0)
    cams3, _, s3 = ctx.ba_local(empty, capi.ba_options(max_num_iterations=3))
    assert s3["initial_cost"] == 0.0 and np.array_equal(cams3, pb["cams"])
