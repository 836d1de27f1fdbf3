"""GPU parity: Frame::ComputeStereoMatches (reference src/frame.cpp:125-333)
through the C ABI vs the line-by-line oracle (oracle/match_ref.c::orc_stereo_matches, itself
pinned bit-exactly to the compiled reference in tests/test_ref_golden_cpu.py).  Bar: bit-exact
mvuRight / mvDepth and match count."""
import numpy as np
import pytest

from lorb_slam_b200 import synth
from oracle import ref

pytestmark = pytest.mark.gpu


def _cmp(ctx, st):
    r, o = ctx.stereo_matches(st), ref.stereo_matches(st)
    assert r["n_matched"] == o["n_matched"]
    assert np.array_equal(r["uright"], o["uright"])
    assert np.array_equal(r["depth"], o["depth"])
    return o


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("n", [500, 2000])
def test_stereo_random(ctx, seed, n):
    o = _cmp(ctx, synth.make_stereo_pair(n, 10 + seed))
    assert o["n_matched"] > n // 4


def test_stereo_vga_1000_features(ctx):
    """BASELINE config 1 shape: 1000 features on a 640x480 stereo pair."""
    o = _cmp(ctx, synth.make_stereo_pair(1000, 0))
    assert (o["depth"][o["uright"] >= 0] > 0).all()


def test_stereo_edge_cases(ctx):
    st = synth.make_stereo_pair(800, 21)
    # keypoints whose patch leaves the level image (the reference would throw): unmatched in both
    st["lx"][:30] = np.linspace(0.0, 12.0, 30).astype(np.float32)
    st["ly"][30:60] = np.linspace(470.0, 479.9, 30).astype(np.float32)
    st["rx"][:40] = np.linspace(0.0, 25.0, 40).astype(np.float32)
    _cmp(ctx, st)
    # no right keypoints at all
    st2 = dict(st)
    st2["n_right"] = 0
    for k in ("rx", "ry"):
        st2[k] = np.zeros(0, np.float32)
    st2["roct"] = np.zeros(0, np.int32)
    st2["rdesc"] = np.zeros((0, 32), np.uint8)
    r = ctx.stereo_matches(st2)
    assert r["n_matched"] == 0 and (r["uright"] == -1).all() and (r["depth"] == -1).all()
    # identical descriptors everywhere: ties resolved by the lowest right index
    st3 = synth.make_stereo_pair(600, 22)
    st3["ldesc"][:] = st3["ldesc"][0]
    st3["rdesc"][:] = st3["ldesc"][0]
    _cmp(ctx, st3)
    # a single pyramid level, every keypoint on it
    st4 = synth.make_stereo_pair(700, 23, n_levels=1)
    _cmp(ctx, st4)


def test_stereo_strided_pyramid(ctx):
    """mvImagePyramid levels are views into bordered buffers (step > width)."""
    st = synth.make_stereo_pair(900, 24)
    tight = ctx.stereo_matches(st)
    st2 = dict(st)
    for side in ("pyr_left", "pyr_right"):
        lv = []
        for a in st[side]:
            big = np.zeros((a.shape[0] + 38, a.shape[1] + 38 + 7), np.uint8)
            big[19:19 + a.shape[0], 19:19 + a.shape[1]] = a
            lv.append(big[19:19 + a.shape[0], 19:19 + a.shape[1]])
        st2[side] = lv
    # capi / oracle front-ends make views contiguous; pass the strided views through the raw ABI
    import ctypes as C
    from lorb_slam_b200 import capi

    def view(pyr):
        w = np.array([a.shape[1] for a in pyr], np.int32)
        h = np.array([a.shape[0] for a in pyr], np.int32)
        s = np.array([a.strides[0] for a in pyr], np.int32)
        ptrs = (C.c_void_p * len(pyr))(*[a.ctypes.data for a in pyr])
        return capi.PyramidView(len(pyr), capi._ptr(w), capi._ptr(h), capi._ptr(s), C.cast(ptrs, C.c_void_p)), (w, h, s, ptrs)

    vl, kl = view(st2["pyr_left"])
    vr, kr = view(st2["pyr_right"])
    n = st["n_left"]
    ur, dp, nm = np.zeros(n, np.float32), np.zeros(n, np.float32), C.c_int(0)
    a = [np.ascontiguousarray(st[k]) for k in ("scale_factors", "inv_scale_factors", "lx", "ly", "loct", "ldesc",
                                               "rx", "ry", "roct", "rdesc")]
    capi._check(ctx._lib.lorb_stereo_matches(
        ctx._h, C.byref(vl), C.byref(vr), int(st["n_levels"]), capi._ptr(a[0]), capi._ptr(a[1]),
        C.c_float(st["mbf"]), C.c_float(st["mb"]), n, capi._ptr(a[2]), capi._ptr(a[3]), capi._ptr(a[4]),
        capi._ptr(a[5]), int(st["n_right"]), capi._ptr(a[6]), capi._ptr(a[7]), capi._ptr(a[8]), capi._ptr(a[9]),
        capi._ptr(ur), capi._ptr(dp), C.byref(nm)))
    assert nm.value == tight["n_matched"]
    assert np.array_equal(ur, tight["uright"]) and np.array_equal(dp, tight["depth"])
