"""ctypes front-end of the CPU oracles (oracle/match_ref.c, oracle/ba_ref.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(lorb_slam_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "lib")


_LIBS = ("liborc_match.so", "liborc_ba.so", "liborc_orb.so")


def build(force=False):
    """Compile the oracle shared objects with the committed Makefile."""
    need = force or not all(
        os.path.exists(os.path.join(_LIBDIR, n)) for n in _LIBS)
    if not need:
        srcs = [os.path.join(_HERE, "match_ref.c"), os.path.join(_HERE, "ba_ref.cpp"),
                os.path.join(_HERE, "orb_ref.cpp"), os.path.join(_HERE, "orb_quadtree_ref.h"),
                os.path.join(_HERE, "Makefile"),
                os.path.join(_HERE, "..", "lorb_slam_b200", "csrc", "libm_sincosf.cuh"),
                os.path.join(_HERE, "..", "include", "lorb_cuda.h")]
        newest = max(os.path.getmtime(s) for s in srcs)
        oldest = min(os.path.getmtime(os.path.join(_LIBDIR, n))
                     for n in _LIBS)
        need = newest > oldest
    if need:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)


_m = None
_b = None
_o = None


def _match():
    global _m
    if _m is None:
        build()
        _m = C.CDLL(os.path.join(_LIBDIR, "liborc_match.so"))
    return _m


def _orb():
    global _o
    if _o is None:
        build()
        _o = C.CDLL(os.path.join(_LIBDIR, "liborc_orb.so"))
        _o.orc_sincosf_mismatches.restype = C.c_longlong
        _o.orc_sincosf_mismatches.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    return _o


def _ba():
    global _b
    if _b is None:
        build()
        _b = C.CDLL(os.path.join(_LIBDIR, "liborc_ba.so"))
        _b.orc_ba_local_cost.restype = C.c_double
    return _b


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ------------------------------------------------------------------ matching

def hamming256(a, b):
    a, b = _u8(a), _u8(b)
    return int(_match().orc_hamming256(_p(a, C.c_uint8), _p(b, C.c_uint8)))


def bf_crosscheck(q, t, mode=0):
    """-> dict(q, t, dist, keep, n_kept, min_dist) like lorb_match_bf_crosscheck."""
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    cap = max(1, min(len(q), len(t)))
    oq = np.zeros(cap, np.int32)
    ot = np.zeros(cap, np.int32)
    od = np.zeros(cap, np.int32)
    n = _match().orc_bf_crosscheck(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), mode,
                                   _p(oq, C.c_int), _p(ot, C.c_int), _p(od, C.c_int))
    keep = np.zeros(cap, np.uint8)
    md = C.c_int(0)
    kept = _match().orc_bf_filter(_p(od, C.c_int), n, C.byref(md), _p(keep, C.c_uint8))
    return dict(q=oq[:n].copy(), t=ot[:n].copy(), dist=od[:n].copy(), keep=keep[:n].copy(),
                n_kept=int(kept), min_dist=int(md.value))


def knn2(q, t):
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    idx = np.zeros((max(1, len(q)), 2), np.int32)
    dist = np.zeros((max(1, len(q)), 2), np.int32)
    _match().orc_knn2(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), _p(idx, C.c_int),
                      _p(dist, C.c_int))
    return idx[:len(q)], dist[:len(q)]


def set_num_threads(n):
    """Thread count of the OpenMP loops (sweep, batched BA); torchrun exports OMP_NUM_THREADS=1."""
    _match().orc_set_num_threads(int(n))
    return int(_match().orc_get_max_threads())


def sweep(bank, pair_a, pair_b):
    """bank [n_kf, n_desc, 32] -> (kept, matches, min_dist) per pair; OpenMP over pairs."""
    bank = _u8(bank)
    n_desc = bank.shape[1]
    pa, pb = _i32(pair_a), _i32(pair_b)
    n = len(pa)
    kept = np.zeros(max(1, n), np.int32)
    mt = np.zeros(max(1, n), np.int32)
    md = np.zeros(max(1, n), np.int32)
    _match().orc_sweep(_p(bank, C.c_uint8), n_desc, _p(pa, C.c_int), _p(pb, C.c_int), n,
                       _p(kept, C.c_int), _p(mt, C.c_int), _p(md, C.c_int))
    return kept[:n], mt[:n], md[:n]


def features_in_area(fr, x, y, r, min_level=-1, max_level=-1):
    out = np.zeros(max(1, fr["n_kp"]), np.int32)
    n = _match().orc_features_in_area(
        fr["n_kp"], _p(fr["kp_x"], C.c_float), _p(fr["kp_y"], C.c_float),
        _p(fr["kp_octave"], C.c_int), C.c_float(fr["min_x"]), C.c_float(fr["max_x"]),
        C.c_float(fr["min_y"]), C.c_float(fr["max_y"]), C.c_float(x), C.c_float(y), C.c_float(r),
        int(min_level), int(max_level), _p(out, C.c_int))
    return out[:n].copy()


def _frame_args(fr):
    return [fr["n_kp"], _p(fr["kp_x"], C.c_float), _p(fr["kp_y"], C.c_float),
            _p(fr["kp_octave"], C.c_int)]


def search_proj_points(fr, pts, th):
    """fr / pts: dicts of contiguous arrays as produced by lorb_slam_b200.synth."""
    n_kp, n_pts = fr["n_kp"], pts["n_pts"]
    kfp = np.full(max(1, n_pts), -1, np.int32)
    pfk = np.full(max(1, n_kp), -1, np.int32)
    nc = C.c_longlong(0)
    n = _match().orc_search_proj_points(
        *_frame_args(fr), _p(fr["kp_uright"], C.c_float), _p(fr["desc"], C.c_uint8),
        _p(fr["kp_claim_obs"], C.c_int), C.c_float(fr["min_x"]), C.c_float(fr["max_x"]),
        C.c_float(fr["min_y"]), C.c_float(fr["max_y"]), _p(fr["scale_factors"], C.c_float),
        n_pts, _p(pts["proj_x"], C.c_float), _p(pts["proj_y"], C.c_float),
        _p(pts["proj_xr"], C.c_float), _p(pts["level"], C.c_int), _p(pts["view_cos"], C.c_float),
        _p(pts["active"], C.c_uint8), _p(pts["mp_desc"], C.c_uint8), _p(pts["mp_nobs"], C.c_int),
        C.c_float(th), _p(kfp, C.c_int), _p(pfk, C.c_int), C.byref(nc))
    return dict(kp_for_point=kfp[:n_pts], point_for_kp=pfk[:n_kp], n_matches=int(n),
                n_candidates=int(nc.value))


def search_proj_frame(cur, last, th):
    n_kp, n_last = cur["n_kp"], last["n_last"]
    kfi = np.full(max(1, n_last), -1, np.int32)
    sfk = np.full(max(1, n_kp), -1, np.int32)
    nc = C.c_longlong(0)
    K = last["K"]
    n = _match().orc_search_proj_frame(
        *_frame_args(cur), _p(cur["kp_angle"], C.c_float), _p(cur["kp_uright"], C.c_float),
        _p(cur["desc"], C.c_uint8), _p(cur["kp_claim_obs"], C.c_int), C.c_float(cur["min_x"]),
        C.c_float(cur["max_x"]), C.c_float(cur["min_y"]), C.c_float(cur["max_y"]),
        _p(cur["scale_factors"], C.c_float), _p(last["tcw_cur"], C.c_float),
        _p(last["tcw_last"], C.c_float), C.c_float(K["fx"]), C.c_float(K["fy"]),
        C.c_float(K["cx"]), C.c_float(K["cy"]), C.c_float(K["mbf"]), C.c_float(K["mb"]), n_last,
        _p(last["valid"], C.c_uint8), _p(last["xw"], C.c_float), _p(last["octave"], C.c_int),
        _p(last["angle"], C.c_float), _p(last["mp_desc"], C.c_uint8), _p(last["mp_nobs"], C.c_int),
        C.c_float(th), _p(kfi, C.c_int), _p(sfk, C.c_int), C.byref(nc))
    return dict(kp_for_item=kfi[:n_last], state_for_kp=sfk[:n_kp], n_matches=int(n),
                n_candidates=int(nc.value))


def frustum_project(fp):
    """fp: dict from lorb_slam_b200.synth.make_frustum_points."""
    n = fp["n"]
    K = fp["K"]
    out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.full(n, -7.0, np.float32),
               proj_y=np.full(n, -7.0, np.float32), proj_xr=np.full(n, -7.0, np.float32),
               level=np.full(n, -7, np.int32), view_cos=np.full(n, -7.0, np.float32))
    _match().orc_frustum_project(
        _p(fp["tcw"], C.c_float), _p(fp["ow"], C.c_float), C.c_float(K["fx"]), C.c_float(K["fy"]),
        C.c_float(K["cx"]), C.c_float(K["cy"]), C.c_float(K["mbf"]), C.c_float(fp["min_x"]),
        C.c_float(fp["max_x"]), C.c_float(fp["min_y"]), C.c_float(fp["max_y"]), n,
        _p(fp["xw"], C.c_float), _p(fp["normal"], C.c_float), _p(fp["min_dist"], C.c_float),
        _p(fp["max_dist"], C.c_float), C.c_float(fp["cos_limit"]), C.c_float(fp["log_sf"]),
        int(fp["n_levels"]), _p(out["in_view"], C.c_uint8), _p(out["proj_x"], C.c_float),
        _p(out["proj_y"], C.c_float), _p(out["proj_xr"], C.c_float), _p(out["level"], C.c_int),
        _p(out["view_cos"], C.c_float))
    return out


def compute_descriptors(offsets, desc):
    offsets, desc = _i32(offsets), _u8(desc).reshape(-1, 32)
    n = len(offsets) - 1
    best, med = np.zeros(n, np.int32), np.zeros(n, np.int32)
    for k in range(n):
        m = C.c_int(0)
        sl = np.ascontiguousarray(desc[offsets[k]:offsets[k + 1]])
        best[k] = _match().orc_compute_descriptor(_p(sl, C.c_uint8) if len(sl) else None, len(sl),
                                                  C.byref(m))
        med[k] = m.value
    return best, med


def _pyr_args(pyr):
    """list of 2-D uint8 arrays -> (keepalive, w[], h[], step[], uint8** )"""
    lv = [np.ascontiguousarray(a, dtype=np.uint8) for a in pyr]
    w = np.array([a.shape[1] for a in lv], np.int32)
    h = np.array([a.shape[0] for a in lv], np.int32)
    st = np.array([a.strides[0] for a in lv], np.int32)
    ptrs = (C.POINTER(C.c_uint8) * len(lv))(*[_p(a, C.c_uint8) for a in lv])
    return lv, w, h, st, ptrs


def stereo_matches(st):
    """st: dict from lorb_slam_b200.synth.make_stereo_pair -> dict(uright, depth, n_matched)."""
    kl, lw, lh, ls, lp = _pyr_args(st["pyr_left"])
    kr, rw, rh, rs, rp = _pyr_args(st["pyr_right"])
    n = st["n_left"]
    ur, dp = np.zeros(max(1, n), np.float32), np.zeros(max(1, n), np.float32)
    _match().orc_stereo_matches.restype = C.c_int
    k = _match().orc_stereo_matches(
        int(st["n_levels"]), _p(lw, C.c_int), _p(lh, C.c_int), _p(ls, C.c_int), lp, _p(rw, C.c_int),
        _p(rh, C.c_int), _p(rs, C.c_int), rp, _p(_f32(st["scale_factors"]), C.c_float),
        _p(_f32(st["inv_scale_factors"]), C.c_float), C.c_float(st["mbf"]), C.c_float(st["mb"]), n,
        _p(st["lx"], C.c_float), _p(st["ly"], C.c_float), _p(st["loct"], C.c_int),
        _p(st["ldesc"], C.c_uint8), st["n_right"], _p(st["rx"], C.c_float), _p(st["ry"], C.c_float),
        _p(st["roct"], C.c_int), _p(st["rdesc"], C.c_uint8), _p(ur, C.c_float), _p(dp, C.c_float))
    return dict(uright=ur[:n], depth=dp[:n], n_matched=int(k))


def orb_describe(oi, pattern, umax):
    """oi: dict from synth.make_orb_inputs; pattern [512,2] int32, umax [16] int32 (the tables of
    the reference's ORBextractor constructor) -> (angle[n], desc[n,32])."""
    kr, w, h, sr, pr = _pyr_args(oi["pyr_raw"])
    kb, _, _, sb, pb = _pyr_args(oi["pyr_blur"])
    n = oi["n_kp"]
    pattern, umax = _i32(pattern).reshape(-1), _i32(umax)
    ang, desc = np.zeros(max(1, n), np.float32), np.zeros((max(1, n), 32), np.uint8)
    _match().orc_orb_describe(int(oi["n_levels"]), _p(w, C.c_int), _p(h, C.c_int), _p(sr, C.c_int), pr,
                              _p(sb, C.c_int), pb, n, _p(oi["kx"], C.c_float), _p(oi["ky"], C.c_float),
                              _p(oi["klevel"], C.c_int), _p(pattern, C.c_int), _p(umax, C.c_int),
                              _p(ang, C.c_float), _p(desc, C.c_uint8))
    return ang[:n], desc[:n]


def libm_sincosf(x, restated=False):
    """sinf/cosf of the process's libm (or of the host instantiation of the device restatement)."""
    x = _f32(x)
    s, c = np.zeros(len(x), np.float32), np.zeros(len(x), np.float32)
    f = _orb().orc_restated_sincosf if restated else _orb().orc_libm_sincosf
    f(len(x), _p(x, C.c_float), _p(s, C.c_float), _p(c, C.c_float))
    return s, c


def sincosf_mismatches(lo_bits, hi_bits, stride):
    return int(_orb().orc_sincosf_mismatches(lo_bits, hi_bits, stride))


def resize_linear(img, dw, dh):
    """cv::resize(img, (dw, dh), INTER_LINEAR) restated (oracle/orb_ref.cpp)."""
    img = _u8(img)
    out = np.zeros((dh, dw), np.uint8)
    _orb().orc_resize_linear_u8(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0],
                                _p(out, C.c_uint8), dw, dh, dw)
    return out


def gaussian7(img):
    """cv::GaussianBlur(img, (7, 7), 2, 2, BORDER_REFLECT_101) restated."""
    img = _u8(img)
    out = np.zeros_like(img)
    _orb().orc_gaussian7_u8(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0],
                            _p(out, C.c_uint8), img.shape[1])
    return out


def fast9_nms(img, threshold):
    """cv::FAST(img, kps, threshold, true) restated -> (x[n], y[n], score[n]) in OpenCV's order."""
    img = _u8(img)
    cap = img.size
    x, y, s = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n = _orb().orc_fast9_nms(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], int(threshold),
                             _p(x, C.c_int), _p(y, C.c_int), _p(s, C.c_int), cap)
    return x[:n], y[:n], s[:n]


def orb_level_candidates(img, ini_th=20, min_th=7):
    """vToDistributeKeys of one pyramid level (src/ORBextractor.cpp:808-862) -> (x, y, response)."""
    img = _u8(img)
    cap = img.size // 4 + 1024
    x, y, r = np.zeros(cap, np.float32), np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    n = _orb().orc_orb_level_candidates(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], ini_th,
                                        min_th, _p(x, C.c_float), _p(y, C.c_float), _p(r, C.c_float), cap)
    assert 0 <= n <= cap
    return x[:n], y[:n], r[:n]


def orb_pyramid(img, sizes):
    """ComputePyramid (:1157-1184): level l = resize(level l-1) to sizes[l] = (w, h)."""
    pyr = [_u8(img)]
    for (w, h) in sizes[1:]:
        pyr.append(resize_linear(pyr[-1], int(w), int(h)))
    return pyr


def orb_distribute(x, y, response, min_x, max_x, min_y, max_y, n_features):
    """ORBextractor::DistributeOctTree restated (oracle/orb_quadtree_ref.h) -> chosen indices, in the
    reference's output order."""
    x, y, r = _f32(x), _f32(y), _f32(response)
    out = np.zeros(max(1, len(x)), np.int32)
    n = _orb().orc_orb_distribute(len(x), _p(x, C.c_float), _p(y, C.c_float), _p(r, C.c_float), min_x, max_x, min_y,
                                  max_y, n_features, _p(out, C.c_int))
    return out[:n]


def fast_atan2(y, x):
    f = _match().orc_fast_atan2_export
    f.restype = C.c_float
    f.argtypes = [C.c_float, C.c_float]
    return float(f(y, x))


def project_rt(tcw, xw):
    tcw, xw = _f32(tcw).reshape(16), _f32(xw).reshape(3)
    out = np.zeros(3, np.float32)
    _match().orc_project_rt(_p(tcw, C.c_float), _p(xw, C.c_float), _p(out, C.c_float))
    return out


def tlc(tcw_cur, tcw_last):
    a, b = _f32(tcw_cur).reshape(16), _f32(tcw_last).reshape(16)
    out = np.zeros(3, np.float32)
    _match().orc_tlc(_p(a, C.c_float), _p(b, C.c_float), _p(out, C.c_float))
    return out


def three_maxima(hist):
    h = _i32(hist)
    i1, i2, i3 = C.c_int(), C.c_int(), C.c_int()
    _match().orc_three_maxima(_p(h, C.c_int), len(h), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


# ------------------------------------------------------------------------ BA

class BAOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("jacobi_scaling", C.c_int),
                ("max_consecutive_invalid_steps", C.c_int), ("reserved0", C.c_int),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double)]


class BASummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("final_radius", C.c_double), ("final_gradient_max_norm", C.c_double),
                ("iterations", C.c_int), ("num_successful_steps", C.c_int),
                ("num_unsuccessful_steps", C.c_int), ("termination", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def ba_options(**kw):
    o = BAOptions()
    _ba().orc_ba_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def ba_pose_only(xw, uv, K, rt, opt=None):
    xw, uv, K = _f32(xw).reshape(-1, 3), _f32(uv).reshape(-1, 2), _f32(K).reshape(4)
    rt = _f64(rt).reshape(6).copy()
    opt = opt or ba_options()
    s = BASummary()
    rc = _ba().orc_ba_pose_only(len(xw), _p(xw, C.c_float), _p(uv, C.c_float), _p(K, C.c_float),
                                _p(rt, C.c_double), C.byref(opt), C.byref(s))
    assert rc == 0
    return rt, s.as_dict()


def ba_local(pb, opt=None):
    """pb: dict(cams[C,6], pts[P,3], obs_cam, obs_pt, obs_uv, fix_pt, fix_uv, fix_rt, K)."""
    cams, pts = _f64(pb["cams"]).copy(), _f64(pb["pts"]).copy()
    oc, op, ouv = _i32(pb["obs_cam"]), _i32(pb["obs_pt"]), _f32(pb["obs_uv"])
    fp, fuv, frt = _i32(pb.get("fix_pt", [])), _f32(pb.get("fix_uv", [])), _f32(pb.get("fix_rt", []))
    K = _f32(pb["K"]).reshape(4)
    opt = opt or ba_options()
    s = BASummary()
    rc = _ba().orc_ba_local(len(cams), _p(cams, C.c_double), len(pts), _p(pts, C.c_double), len(oc),
                            _p(oc, C.c_int), _p(op, C.c_int), _p(ouv, C.c_float), len(fp),
                            _p(fp, C.c_int), _p(fuv, C.c_float), _p(frt, C.c_float),
                            _p(K, C.c_float), C.byref(opt), C.byref(s))
    assert rc == 0
    return cams, pts, s.as_dict()


def ba_local_cost(pb, cams=None, pts=None):
    cams = _f64(pb["cams"] if cams is None else cams)
    pts = _f64(pb["pts"] if pts is None else pts)
    oc, op, ouv = _i32(pb["obs_cam"]), _i32(pb["obs_pt"]), _f32(pb["obs_uv"])
    fp, fuv, frt = _i32(pb.get("fix_pt", [])), _f32(pb.get("fix_uv", [])), _f32(pb.get("fix_rt", []))
    K = _f32(pb["K"]).reshape(4)
    return float(_ba().orc_ba_local_cost(
        len(cams), _p(cams, C.c_double), len(pts), _p(pts, C.c_double), len(oc), _p(oc, C.c_int),
        _p(op, C.c_int), _p(ouv, C.c_float), len(fp), _p(fp, C.c_int), _p(fuv, C.c_float),
        _p(frt, C.c_float), _p(K, C.c_float)))


def ba_residual_jac(kind, cam, pt, uv, K):
    cam, pt = _f64(cam).reshape(6), _f64(pt).reshape(3)
    uv, K = _f32(uv).reshape(2), _f32(K).reshape(4)
    r, Jc, Jp = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 3))
    _ba().orc_ba_residual_jac(int(kind), _p(cam, C.c_double), _p(pt, C.c_double), _p(uv, C.c_float),
                              _p(K, C.c_float), _p(r, C.c_double), _p(Jc, C.c_double),
                              _p(Jp, C.c_double))
    return r, Jc, Jp


def ba_local_batched(cam_off, cams, pt_off, pts, obs_off, obs_cam, obs_pt, obs_uv, K, opt=None):
    cam_off, pt_off, obs_off = _i32(cam_off), _i32(pt_off), _i32(obs_off)
    cams, pts = _f64(cams).copy(), _f64(pts).copy()
    oc, op, ouv = _i32(obs_cam), _i32(obs_pt), _f32(obs_uv)
    K = _f32(K).reshape(4)
    nw = len(cam_off) - 1
    opt = opt or ba_options()
    sums = (BASummary * nw)()
    _ba().orc_ba_local_batched(nw, _p(cam_off, C.c_int), _p(cams, C.c_double), _p(pt_off, C.c_int),
                               _p(pts, C.c_double), _p(obs_off, C.c_int), _p(oc, C.c_int),
                               _p(op, C.c_int), _p(ouv, C.c_float), _p(K, C.c_float), C.byref(opt),
                               sums)
    return cams, pts, [s.as_dict() for s in sums]
