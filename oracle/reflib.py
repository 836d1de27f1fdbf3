"""ctypes front-end of oracle/_ref/libref.so: the reference's own hot-path sources
(/root/reference/src/{matcher,frame,map_point,map,bundle_adjust}.cpp) compiled unmodified
against the OpenCV / Ceres stand-ins of oracle/refshim/, behind oracle/ref_harness.cpp.

TEST INFRASTRUCTURE ONLY (tests/, the golden-vector generator, bench.py's CPU legs).
/root/reference exists only in the build container: `available()` is False on the GPU box unless
the prebuilt library travelled with the snapshot.  Same argument dictionaries as oracle/ref.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import ref as _orc
from .ref import BAOptions, BASummary, _f32, _f64, _i32, _p, _u8, ba_options  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_ROOT = os.environ.get("LORB_REFERENCE_ROOT", "/root/reference")
_lib = None


def can_build():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def build(force=False):
    """make -C oracle ref: compiles the reference sources where they lie; output only in oracle/_ref/."""
    if not can_build():
        return False
    srcs = [os.path.join(_HERE, "ref_harness.cpp"), os.path.join(_HERE, "refshim", "lorb_cvshim.hpp"),
            os.path.join(_HERE, "refshim", "lorb_ceresshim.hpp"), os.path.join(_HERE, "Makefile"),
            os.path.join(_HERE, "ref_orb_harness.cpp"), os.path.join(_HERE, "orb_ref.cpp"),
            os.path.join(_HERE, "orb_quadtree_ref.h")]
    if force or not os.path.exists(_LIB) or max(map(os.path.getmtime, srcs)) > os.path.getmtime(_LIB):
        subprocess.run(["make", "-C", _HERE, "ref", "REFERENCE=" + REFERENCE_ROOT], check=True,
                       capture_output=True)
    return True


def available():
    return os.path.exists(_LIB) or can_build()


def lib():
    global _lib
    if _lib is None:
        build()
        if not os.path.exists(_LIB):
            raise RuntimeError("oracle/_ref/libref.so is absent and /root/reference is not here to build it")
        _lib = C.CDLL(_LIB)
        _lib.ref_radius_by_viewing_cos.restype = C.c_float
        _lib.ref_radius_by_viewing_cos.argtypes = [C.c_float]
    return _lib


# ------------------------------------------------------------ stand-in arithmetic (cv2 pins)

def cv_rt(R, x, t):
    R, x, t = _f32(R).reshape(9), _f32(x).reshape(3), _f32(t).reshape(3)
    out = np.zeros(3, np.float32)
    lib().ref_cv_rt(_p(R, C.c_float), _p(x, C.c_float), _p(t, C.c_float), _p(out, C.c_float))
    return out


def cv_mul4(T, x):
    T, x = _f32(T).reshape(16), _f32(x).reshape(4)
    out = np.zeros(4, np.float32)
    lib().ref_cv_mul4(_p(T, C.c_float), _p(x, C.c_float), _p(out, C.c_float))
    return out


def cv_inv4(T):
    T = _f32(T).reshape(16)
    out = np.zeros(16, np.float32)
    lib().ref_cv_inv4(_p(T, C.c_float), _p(out, C.c_float))
    return out.reshape(4, 4)


def cv_rodrigues(r):
    r = _f32(r).reshape(3)
    out = np.zeros(9, np.float32)
    lib().ref_cv_rodrigues(_p(r, C.c_float), _p(out, C.c_float))
    return out.reshape(3, 3)


def cv_bfmatch(q, t):
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    cap = max(1, len(q))
    oq, ot, od = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n = lib().ref_cv_bfmatch(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), _p(oq, C.c_int),
                             _p(ot, C.c_int), _p(od, C.c_int))
    return oq[:n].copy(), ot[:n].copy(), od[:n].copy()


# ------------------------------------------------------------------------- matching

def descriptor_distance(a, b):
    a, b = _u8(a).reshape(32), _u8(b).reshape(32)
    return int(lib().ref_descriptor_distance(_p(a, C.c_uint8), _p(b, C.c_uint8)))


def three_maxima(hist):
    h = _i32(hist)
    i1, i2, i3 = C.c_int(), C.c_int(), C.c_int()
    lib().ref_three_maxima(_p(h, C.c_int), len(h), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


def radius_by_viewing_cos(c):
    return float(lib().ref_radius_by_viewing_cos(C.c_float(c)))


def constants():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().ref_constants(C.byref(a), C.byref(b), C.byref(c))
    return dict(TH_LOW=a.value, TH_HIGH=b.value, HISTO_LENGTH=c.value)


def features_in_area(fr, x, y, r, min_level=-1, max_level=-1):
    out = np.zeros(max(1, fr["n_kp"]), np.int32)
    n = lib().ref_features_in_area(
        fr["n_kp"], _p(fr["kp_x"], C.c_float), _p(fr["kp_y"], C.c_float), _p(fr["kp_octave"], C.c_int),
        C.c_float(fr["min_x"]), C.c_float(fr["max_x"]), C.c_float(fr["min_y"]), C.c_float(fr["max_y"]),
        C.c_float(x), C.c_float(y), C.c_float(r), int(min_level), int(max_level), _p(out, C.c_int))
    return out[:n].copy()


def search_bf(q, t, present=None, use_set=False):
    """Matcher::SearchByProjection(curr, prev) / SearchLocalPoints -> (n_kept, assign[n_q])."""
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    pres = None if present is None else _u8(present)
    out = np.full(max(1, len(q)), -1, np.int32)
    n = lib().ref_search_bf(int(use_set), len(q), _p(q, C.c_uint8), len(t), _p(t, C.c_uint8),
                            _p(pres, C.c_uint8) if pres is not None else None, _p(out, C.c_int))
    return int(n), out[:len(q)].copy()


class Sweep:
    """Keyframe-pair sweep through Matcher::SearchByProjection(curr, prev); OpenMP over pairs."""

    def __init__(self, bank):
        self.bank = _u8(bank)
        lib().ref_sweep_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().ref_sweep_create(_p(self.bank, C.c_uint8), self.bank.shape[0], self.bank.shape[1]))

    def run(self, pair_a, pair_b):
        pa, pb = _i32(pair_a), _i32(pair_b)
        kept = np.zeros(max(1, len(pa)), np.int32)
        lib().ref_sweep_run(self.h, _p(pa, C.c_int), _p(pb, C.c_int), len(pa), _p(kept, C.c_int))
        return kept[:len(pa)]

    def close(self):
        if self.h:
            lib().ref_sweep_destroy(self.h)
            self.h = None


def search_proj_points(fr, pts, th):
    n_kp, n_pts = fr["n_kp"], pts["n_pts"]
    pfk = np.full(max(1, n_kp), -1, np.int32)
    n = lib().ref_search_proj_points(
        n_kp, _p(fr["kp_x"], C.c_float), _p(fr["kp_y"], C.c_float), _p(fr["kp_octave"], C.c_int),
        _p(fr["kp_uright"], C.c_float), _p(fr["desc"], C.c_uint8), _p(fr["kp_claim_obs"], C.c_int),
        C.c_float(fr["min_x"]), C.c_float(fr["max_x"]), C.c_float(fr["min_y"]), C.c_float(fr["max_y"]),
        _p(fr["scale_factors"], C.c_float), len(fr["scale_factors"]), n_pts,
        _p(pts["proj_x"], C.c_float), _p(pts["proj_y"], C.c_float), _p(pts["proj_xr"], C.c_float),
        _p(pts["level"], C.c_int), _p(pts["view_cos"], C.c_float), _p(pts["active"], C.c_uint8),
        _p(pts["mp_desc"], C.c_uint8), _p(pts["mp_nobs"], C.c_int), C.c_float(th), _p(pfk, C.c_int))
    return dict(point_for_kp=pfk[:n_kp], n_matches=int(n))


def search_proj_frame(cur, last, th):
    n_kp, n_last = cur["n_kp"], last["n_last"]
    sfk = np.full(max(1, n_kp), -1, np.int32)
    K = last["K"]
    mb = C.c_float(0)
    n = lib().ref_search_proj_frame(
        n_kp, _p(cur["kp_x"], C.c_float), _p(cur["kp_y"], C.c_float), _p(cur["kp_octave"], C.c_int),
        _p(cur["kp_angle"], C.c_float), _p(cur["kp_uright"], C.c_float), _p(cur["desc"], C.c_uint8),
        _p(cur["kp_claim_obs"], C.c_int), C.c_float(cur["min_x"]), C.c_float(cur["max_x"]),
        C.c_float(cur["min_y"]), C.c_float(cur["max_y"]), _p(cur["scale_factors"], C.c_float),
        len(cur["scale_factors"]), _p(last["tcw_cur"], C.c_float), _p(last["tcw_last"], C.c_float),
        C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]), C.c_float(K["cy"]),
        C.c_float(K["mbf"]), n_last, _p(last["valid"], C.c_uint8), _p(last["xw"], C.c_float),
        _p(last["octave"], C.c_int), _p(last["angle"], C.c_float), _p(last["mp_desc"], C.c_uint8),
        _p(last["mp_nobs"], C.c_int), C.c_float(th), _p(sfk, C.c_int), C.byref(mb))
    return dict(state_for_kp=sfk[:n_kp], n_matches=int(n), mb=float(mb.value))


def final_state_from_oracle(state_for_kp, kp_claim_obs):
    """Map the oracle's per-keypoint state (-2 = NULLed by the rotation check, whatever it held)
    to what the reference's Frame::mvpMapPoints can show at exit (see ref_harness.cpp)."""
    s = np.asarray(state_for_kp).copy()
    nulled = s == -2
    s[nulled & (np.asarray(kp_claim_obs) < 0)] = -1
    return s


def frustum_project(fp, scale_factor=np.float32(1.2)):
    """-> (outputs like oracle.ref.frustum_project, ow, log_sf) with the reference's own camera
    centre (mTcw.inv()) and log scale factor."""
    n = fp["n"]
    K = fp["K"]
    out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32),
               proj_y=np.zeros(n, np.float32), proj_xr=np.zeros(n, np.float32),
               level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32))
    ow = np.zeros(3, np.float32)
    lsf = C.c_float(0)
    lib().ref_frustum_project(
        _p(fp["tcw"], C.c_float), C.c_float(K["fx"]), C.c_float(K["fy"]), C.c_float(K["cx"]),
        C.c_float(K["cy"]), C.c_float(K["mbf"]), C.c_float(fp["min_x"]), C.c_float(fp["max_x"]),
        C.c_float(fp["min_y"]), C.c_float(fp["max_y"]), n, _p(fp["xw"], C.c_float),
        _p(fp["normal"], C.c_float), _p(fp["min_dist"], C.c_float), _p(fp["max_dist"], C.c_float),
        C.c_float(fp["cos_limit"]), C.c_float(scale_factor), int(fp["n_levels"]),
        _p(out["in_view"], C.c_uint8), _p(out["proj_x"], C.c_float), _p(out["proj_y"], C.c_float),
        _p(out["proj_xr"], C.c_float), _p(out["level"], C.c_int), _p(out["view_cos"], C.c_float),
        _p(ow, C.c_float), C.byref(lsf))
    return out, ow, float(lsf.value)


def stereo_matches(st):
    """Frame::ComputeStereoMatches of the compiled reference on the given pyramids."""
    kl, lw, lh, ls, lp = _orc._pyr_args(st["pyr_left"])
    kr, rw, rh, rs, rp = _orc._pyr_args(st["pyr_right"])
    n = st["n_left"]
    ur, dp = np.zeros(max(1, n), np.float32), np.zeros(max(1, n), np.float32)
    mb = C.c_float(0)
    k = lib().ref_stereo_matches(
        int(st["n_levels"]), _p(lw, C.c_int), _p(lh, C.c_int), _p(ls, C.c_int), lp, _p(rw, C.c_int),
        _p(rh, C.c_int), _p(rs, C.c_int), rp, _p(_f32(st["scale_factors"]), C.c_float),
        _p(_f32(st["inv_scale_factors"]), C.c_float), C.c_float(st["fx"]), C.c_float(st["mbf"]), n,
        _p(st["lx"], C.c_float), _p(st["ly"], C.c_float), _p(st["loct"], C.c_int),
        _p(st["ldesc"], C.c_uint8), st["n_right"], _p(st["rx"], C.c_float), _p(st["ry"], C.c_float),
        _p(st["roct"], C.c_int), _p(st["rdesc"], C.c_uint8), _p(ur, C.c_float), _p(dp, C.c_float),
        C.byref(mb))
    return dict(uright=ur[:n], depth=dp[:n], n_matched=int(k), mb=float(mb.value))


def orb_describe(oi):
    """IC_Angle + computeOrbDescriptor of the compiled reference (src/ORBextractor.cpp:79-149)
    -> (angle[n], desc[n,32], pattern[512,2], umax[16]); the two tables are the ones its
    ORBextractor constructor builds."""
    kr, w, h, sr, pr = _orc._pyr_args(oi["pyr_raw"])
    kb, _, _, sb, pb = _orc._pyr_args(oi["pyr_blur"])
    n = oi["n_kp"]
    ang, desc = np.zeros(max(1, n), np.float32), np.zeros((max(1, n), 32), np.uint8)
    umax, pattern = np.zeros(16, np.int32), np.zeros(1024, np.int32)
    lib().ref_orb_describe(int(oi["n_levels"]), _p(w, C.c_int), _p(h, C.c_int), _p(sr, C.c_int), pr,
                           _p(sb, C.c_int), pb, n, _p(oi["kx"], C.c_float), _p(oi["ky"], C.c_float),
                           _p(oi["klevel"], C.c_int), _p(ang, C.c_float), _p(desc, C.c_uint8),
                           _p(umax, C.c_int), _p(pattern, C.c_int))
    return ang[:n], desc[:n], pattern.reshape(512, 2), umax


def orb_extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """The reference's whole ORBextractor::operator() (src/ORBextractor.cpp:1087-1151) on one 8-bit
    image, run under the bump arena of ref_orb_harness.cpp (quadtree ties = creation order)."""
    img = np.ascontiguousarray(img, np.uint8)
    cap = 4 * nfeatures + 1024  # wide levels with small budgets return more than their share
    kx, ky = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    ko, ka = np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    kr, ks = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    desc, npl = np.zeros((cap, 32), np.uint8), np.zeros(nlevels, np.int32)
    n = lib().ref_orb_extract(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], nfeatures,
                              C.c_float(scale_factor), nlevels, ini_th, min_th, cap, _p(kx, C.c_float),
                              _p(ky, C.c_float), _p(ko, C.c_int), _p(ka, C.c_float), _p(kr, C.c_float),
                              _p(ks, C.c_float), _p(desc, C.c_uint8), _p(npl, C.c_int))
    assert 0 <= n <= cap
    return dict(n=n, x=kx[:n], y=ky[:n], octave=ko[:n], angle=ka[:n], response=kr[:n], size=ks[:n],
                desc=desc[:n], n_per_level=npl)


def distribute_octree(x, y, resp, min_x, max_x, min_y, max_y, n_features):
    """ORBextractor::DistributeOctTree of the compiled reference (bump arena) -> (x, y, response)."""
    x, y, resp = _f32(x), _f32(y), _f32(resp)
    n = len(x)
    ox, oy, orr = np.zeros(n + 8, np.float32), np.zeros(n + 8, np.float32), np.zeros(n + 8, np.float32)
    k = lib().ref_distribute_octree(n, _p(x, C.c_float), _p(y, C.c_float), _p(resp, C.c_float), min_x, max_x, min_y,
                                    max_y, n_features, _p(ox, C.c_float), _p(oy, C.c_float), _p(orr, C.c_float))
    assert 0 <= k <= n
    return ox[:k], oy[:k], orr[:k]


def compute_descriptor(desc):
    desc = _u8(desc).reshape(-1, 32)
    return int(lib().ref_compute_descriptor(_p(desc, C.c_uint8) if len(desc) else None, len(desc)))


# ------------------------------------------------------------------------------- BA

def ba_pose_only(xw, uv, K, rt, opt=None):
    xw, uv, K = _f32(xw).reshape(-1, 3), _f32(uv).reshape(-1, 2), _f32(K).reshape(4)
    rt_in = _f32(rt).reshape(6)
    rt32, rt64, tcw = np.zeros(6, np.float32), np.zeros(6), np.zeros(16, np.float32)
    s = BASummary()
    rc = lib().ref_ba_pose_only(len(xw), _p(xw, C.c_float), _p(uv, C.c_float), _p(K, C.c_float),
                                _p(rt_in, C.c_float), C.byref(opt) if opt is not None else None,
                                _p(rt32, C.c_float), _p(rt64, C.c_double), _p(tcw, C.c_float), C.byref(s))
    assert rc == 0
    return dict(rt_f32=rt32, rt_f64=rt64, tcw=tcw.reshape(4, 4), summary=s.as_dict())


def ba_local(pb, opt=None):
    cams, pts = _f32(pb["cams"]), _f32(pb["pts"])
    assert np.array_equal(cams.astype(np.float64), _f64(pb["cams"])), "inputs must be float-representable"
    oc, op, ouv = _i32(pb["obs_cam"]), _i32(pb["obs_pt"]), _f32(pb["obs_uv"])
    fp, fuv, frt = _i32(pb.get("fix_pt", [])), _f32(pb.get("fix_uv", [])), _f32(pb.get("fix_rt", []))
    K = _f32(pb["K"]).reshape(4)
    c32, c64 = np.zeros_like(cams), np.zeros(cams.shape)
    p32, p64 = np.zeros_like(pts), np.zeros(pts.shape)
    s = BASummary()
    rc = lib().ref_ba_local(len(cams), _p(cams, C.c_float), len(pts), _p(pts, C.c_float), len(oc),
                            _p(oc, C.c_int), _p(op, C.c_int), _p(ouv, C.c_float), len(fp),
                            _p(fp, C.c_int), _p(fuv, C.c_float), _p(frt, C.c_float), _p(K, C.c_float),
                            C.byref(opt) if opt is not None else None, _p(c32, C.c_float),
                            _p(c64, C.c_double), _p(p32, C.c_float), _p(p64, C.c_double), C.byref(s))
    return dict(cams_f32=c32, cams_f64=c64, pts_f32=p32, pts_f64=p64, summary=s.as_dict(), rc=int(rc))
