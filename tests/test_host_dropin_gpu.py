"""GPU: the C++ drop-in classes (Matcher / BA with the reference's interfaces,
lorb_slam_b200/host) driven through real Frame / MapPoint objects, compared
with the oracle.  This is the path a LORB-SLAM caller takes:
VisualOdometry -> Matcher::* / BA::* -> C ABI -> sm_100a kernels."""
import ctypes as C

import numpy as np
import pytest

from lorb_slam_b200 import synth
from lorb_slam_b200.host import build_host
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    return C.CDLL(build_host.build())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _frame_args(fr):
    return [fr["n_kp"], _p(fr["kp_x"]), _p(fr["kp_y"]), _p(fr["kp_octave"]), _p(fr["kp_angle"]),
            _p(fr["kp_uright"]), _p(fr["desc"]), _p(fr["kp_claim_obs"]), C.c_float(fr["min_x"]),
            C.c_float(fr["max_x"]), C.c_float(fr["min_y"]), C.c_float(fr["max_y"]), fr["n_levels"],
            _p(fr["scale_factors"])]


@pytest.mark.parametrize("which", [0, 1])
def test_bruteforce_entry_points(H, which):
    rng = np.random.default_rng(which)
    q = synth.descriptors_uniform(700, rng)
    t = synth.descriptors_uniform(900, rng)
    t[:400] = synth.descriptors_noisy_copy(q[rng.permutation(700)[:400]], rng, 0.05)
    has = (rng.random(900) < 0.8).astype(np.uint8)
    out = np.full(700, -1, np.int32)
    n = H.harness_bf(which, 700, _p(q), 900, _p(t), _p(has), _p(out))
    idx = np.flatnonzero(has)
    o = ref.bf_crosscheck(q, t[idx])
    exp = np.full(700, -1, np.int32)
    k = o["keep"].astype(bool)
    exp[o["q"][k]] = idx[o["t"][k]]
    assert n == o["n_kept"] > 100
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("th,nobs", [(1.0, 1), (1.0, 0), (15.0, (0, 1, 2))])
def test_search_by_projection_points(H, th, nobs):
    fr = synth.make_frame(2000, seed=3, stereo=True, claimed_frac=0.1)
    pts = synth.make_proj_points(fr, 5000, seed=3, nobs=nobs, inactive_frac=0.05)
    out = np.full(2000, -1, np.int32)
    n = H.harness_proj_points(*_frame_args(fr), 5000, _p(pts["proj_x"]), _p(pts["proj_y"]),
                              _p(pts["proj_xr"]), _p(pts["level"]), _p(pts["view_cos"]),
                              _p(pts["active"]), _p(pts["mp_desc"]), _p(pts["mp_nobs"]),
                              C.c_float(th), _p(out))
    o = ref.search_proj_points(fr, pts, th)
    assert n == o["n_matches"] > 100
    assert np.array_equal(out, o["point_for_kp"])


@pytest.mark.parametrize("motion", ["forward", "still"])
def test_search_by_projection_frame(H, motion):
    cur, last = synth.make_frame_pair(2000, seed=2, motion=motion)
    K = last["K"]
    K6 = np.array([K["fx"], K["fy"], K["cx"], K["cy"], K["mbf"], K["mb"]], np.float32)
    out = np.full(2000, -1, np.int32)
    n = H.harness_proj_frame(*_frame_args(cur), _p(last["tcw_cur"]), _p(last["tcw_last"]), _p(K6),
                             2000, _p(last["valid"]), _p(last["xw"]), _p(last["octave"]),
                             _p(last["angle"]), _p(last["mp_desc"]), _p(last["mp_nobs"]),
                             C.c_float(15.0), _p(out))
    o = ref.search_proj_frame(cur, last, 15.0)
    assert n == o["n_matches"] > 200
    exp = o["state_for_kp"].copy()
    exp[exp == -1] = -2  # the frame had no earlier claims: untouched keypoints hold NULL
    assert np.array_equal(out, exp)


def test_project_pose_optimization(H):
    po = synth.make_pose_only(1, 600)
    rt = po["rt"].astype(np.float32)
    H.harness_pose_opt(600, _p(po["xw"]), _p(po["uv"]), _p(po["K"]), _p(rt))
    ort, _ = ref.ba_pose_only(po["xw"], po["uv"], po["K"], po["rt"])
    np.testing.assert_allclose(rt, ort.astype(np.float32), rtol=2e-6, atol=1e-7)


def test_local_pose_optimization(H):
    pb = synth.make_ba_problem(5, C=6, P=400, obs_per_point=(4, 5), fixed_frac=0.1)
    cams = pb["cams"].astype(np.float32)
    pts = pb["pts"].astype(np.float32)
    H.harness_local_ba(6, _p(cams), 400, _p(pts), pb["O"], _p(pb["obs_cam"]), _p(pb["obs_pt"]),
                       _p(pb["obs_uv"]), pb["F"], _p(pb["fix_pt"]), _p(pb["fix_uv"]),
                       _p(pb["fix_rt"]), _p(pb["K"]))
    oc, op, s = ref.ba_local(pb)
    assert s["final_cost"] < s["initial_cost"]
    np.testing.assert_allclose(cams, oc.astype(np.float32), rtol=3e-6, atol=1e-6)
    np.testing.assert_allclose(pts, op.astype(np.float32), rtol=3e-6, atol=1e-6)


def test_local_mapping_index_sequence(H):
    """keyframes inserted one by one, the local BA of every new keyframe taken from
    the index the mapper maintains per insert (lorb_host::LocalMapIndex) -- against the same sequence
    through BA::LocalPoseOptimization's own std::map walks (bit for bit: the same flat problem must
    reach lorb_ba_local) and against an oracle replay of the whole sequence (reference
    src/local_mapping.cpp:19-78 driving src/bundle_adjust.cpp:207-330)."""
    n_kf, P, first_ba, n_cov = 7, 500, 2, 3
    pb = synth.make_ba_problem(9, C=n_kf, P=P, obs_per_point=(3, 4, 5))
    rng = np.random.default_rng(9)
    kf_rt = pb["cams"].astype(np.float32)
    pts = pb["pts"].astype(np.float32)
    slots = [[] for _ in range(n_kf)]
    for c, p, uv in zip(pb["obs_cam"], pb["obs_pt"], pb["obs_uv"]):
        slots[c].append((int(p), uv))
    for k in range(n_kf):  # a few NULL slots (keypoints without a map point), shuffled slot order
        slots[k] += [(-1, np.zeros(2, np.float32))] * 5
        rng.shuffle(slots[k])
    kf_off = np.cumsum([0] + [len(x) for x in slots]).astype(np.int32)
    slot_pt = np.array([p for x in slots for p, _ in x], np.int32)
    slot_uv = np.array([uv for x in slots for _, uv in x], np.float32)
    cov = [[j for j in range(k - 1, max(-1, k - 1 - n_cov), -1)] for k in range(n_kf)]
    cov_off = np.cumsum([0] + [len(x) for x in cov]).astype(np.int32)
    cov_idx = np.array([j for x in cov for j in x] + [0], np.int32)
    ca, cb = np.zeros((n_kf, 6), np.float32), np.zeros((n_kf, 6), np.float32)
    pa_, pb_ = np.zeros((P, 3), np.float32), np.zeros((P, 3), np.float32)
    us = np.zeros(2)
    n_ba = H.harness_local_mapping(n_kf, _p(kf_rt), _p(kf_off), _p(slot_pt), _p(slot_uv), _p(cov_off),
                                   _p(cov_idx), P, _p(pts), _p(pb["K"]), first_ba, _p(ca), _p(pa_),
                                   _p(cb), _p(pb_), _p(us))
    assert n_ba == n_kf - first_ba
    assert np.array_equal(ca, cb) and np.array_equal(pa_, pb_)
    print("assemble us per BA: index %.0f, std::map walks %.0f" % (us[0] / n_ba, us[1] / n_ba))
    # oracle replay: same inserts, same windows, float write-back after every BA
    cams_o, pts_o = kf_rt.copy(), pts.copy()
    for k in range(first_ba, n_kf):
        window = [k] + cov[k]
        widx = {f: i for i, f in enumerate(window)}
        order, seen = [], set()
        for f in window:
            for p, _ in slots[f]:
                if p >= 0 and p not in seen:
                    seen.add(p)
                    order.append(p)
        pidx = {p: i for i, p in enumerate(order)}
        oc, op, ouv, fp, fuv, frt = [], [], [], [], [], []
        for p in order:
            for f in range(k + 1):  # observers inserted so far, in frame (= address) order
                for q, uv in slots[f]:
                    if q != p:
                        continue
                    if f in widx:
                        oc.append(widx[f]); op.append(pidx[p]); ouv.append(uv)
                    else:
                        fp.append(pidx[p]); fuv.append(uv); frt.append(cams_o[f])
        prob = dict(cams=cams_o[window].astype(np.float64), pts=pts_o[order].astype(np.float64),
                    obs_cam=np.array(oc, np.int32), obs_pt=np.array(op, np.int32),
                    obs_uv=np.array(ouv, np.float32).reshape(-1, 2), fix_pt=np.array(fp, np.int32),
                    fix_uv=np.array(fuv, np.float32).reshape(-1, 2),
                    fix_rt=np.array(frt, np.float32).reshape(-1, 6), K=pb["K"])
        c2, p2, _ = ref.ba_local(prob)
        cams_o[window] = c2.astype(np.float32)
        pts_o[order] = p2.astype(np.float32)
    np.testing.assert_allclose(ca, cams_o, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(pa_, pts_o, rtol=2e-5, atol=2e-6)


def test_orbextractor_class_matches_reference():
    """The drop-in ORBextractor class (host/include/ORBextractor.h) called as Frame's constructor
    calls it, against the compiled reference's extractor output (tests/golden/orb_golden.npz)."""
    import orb_checks as OC
    import ref_cases as RC
    Hl = C.CDLL(build_host.build())
    for c in (RC.ORB_EXTRACT[0], RC.ORB_EXTRACT[3]):
        img = RC.orb_extract_case(c)
        cap = c[4] + 64
        kx, ky, ka, kr, ks = (np.zeros(cap, np.float32) for _ in range(5))
        ko, desc, pyr = np.zeros(cap, np.int32), np.zeros((cap, 32), np.uint8), np.zeros(8, np.int64)
        n = Hl.harness_orb_extract(_p(img), img.shape[1], img.shape[0], c[4], cap, _p(kx), _p(ky), _p(ko), _p(ka),
                                   _p(kr), _p(ks), _p(desc), _p(pyr))

        def as_result(_img, _nf):
            return dict(n=n, x=kx[:n], y=ky[:n], octave=ko[:n], angle=ka[:n], response=kr[:n], size=ks[:n],
                        desc=desc[:n])
        OC.check_extract(as_result, c)
        raw = ref.orb_pyramid(img, [(a.shape[1], a.shape[0]) for a in _ref_level_shapes(img)])
        assert [int(a.astype(np.int64).sum()) for a in raw] == list(pyr)


def _ref_level_shapes(img):
    """Level shapes by the reference's rule (src/ORBextractor.cpp:1161-1163) in float32."""
    out, sf = [], np.float32(1.0)
    for l in range(8):
        if l:
            sf = np.float32(sf * np.float32(1.2))
        isf = np.float32(1.0) / sf
        out.append(np.zeros((int(np.rint(np.float32(img.shape[0]) * isf)), int(np.rint(np.float32(img.shape[1]) * isf))),
                            np.uint8))
    return out


def test_orbextractor_on_fresh_threads_reuses_contexts():
    """The reference extracts left / right on two NEW threads every frame (src/frame.cpp:42-45);
    the drop-in layer's context pool must hand the same two GPU contexts back instead of creating
    (and tearing down) one per thread per frame."""
    import orb_checks as OC
    Hl = C.CDLL(build_host.build())
    st = synth.make_stereo_pair(8, 31)
    left, right = st["pyr_left"][0], st["pyr_right"][0]
    cap = 1064
    nl, nr = C.c_int(0), C.c_int(0)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    created = Hl.harness_stereo_extract_threads(_p(left), _p(right), 640, 480, 1000, 6, cap, C.byref(nl), C.byref(nr),
                                                _p(dl), _p(dr))
    assert created <= 2
    from lorb_slam_b200 import capi
    with capi.Context(0) as ctx:
        a, b = ctx.orb_extract(left, OC.pattern()), ctx.orb_extract(right, OC.pattern())
    assert nl.value == a["n"] and nr.value == b["n"]
    assert np.array_equal(dl[:a["n"]], a["desc"]) and np.array_equal(dr[:b["n"]], b["desc"])
