"""CPU side of the ORBextractor stages: the restated oracle against the
compiled reference's outputs, and the pins of the arithmetic the device restates (libm sinf/cosf,
cv::fastAtan2, the umax table)."""
import ctypes as C

import numpy as np
import pytest

import orb_checks as OC
import ref_cases as RC
from lorb_slam_b200 import capi
from oracle import ref


@pytest.mark.parametrize("c", RC.ORB_DESCRIBE, ids=[c[0] for c in RC.ORB_DESCRIBE])
def test_oracle_orb_describe(c):
    OC.check_orb_describe(lambda oi, pat: ref.orb_describe(oi, pat, OC.golden()["orb/umax"]), c)


def test_fast_atan2_restatement_matches_cv2():
    g = OC.golden()
    got = np.array([ref.fast_atan2(float(y), float(x)) for y, x in zip(g["atan2/y"], g["atan2/x"])], np.float32)
    assert np.array_equal(got.view(np.uint32), g["atan2/deg"].view(np.uint32))


def test_sincosf_restatement_matches_libm():
    """Host instantiation of lorb_slam_b200/csrc/libm_sincosf.cuh == this process's libm, on every
    7th float of [0, 7] and its negative (155 M arguments; the full range was run once: 0 differ).
    The descriptor stage only uses [0, 2*pi]."""
    hi = int(np.float32(7.0).view(np.uint32))
    assert ref.sincosf_mismatches(0, hi, 7) == 0
    x = np.linspace(0, 119.9, 200001).astype(np.float32)  # the rest of the fast-reduction range
    a, b = ref.libm_sincosf(x), ref.libm_sincosf(x, restated=True)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_libm_is_not_correctly_rounded():
    """Why the restatement exists: rounding the double result is NOT what the reference gets."""
    x = np.linspace(0.001, 6.28, 200000).astype(np.float32)
    s, _ = ref.libm_sincosf(x)
    assert 0 < (s != np.sin(x.astype(np.float64)).astype(np.float32)).sum() < 0.02 * len(x)


def test_umax_table():
    umax = np.zeros(16, np.int32)
    lib = capi.load_library()
    lib.lorb_orb_umax(umax.ctypes.data_as(C.c_void_p))
    assert np.array_equal(umax, OC.golden()["orb/umax"])
