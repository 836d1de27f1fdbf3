"""GPU parity against the REFERENCE'S OWN compiled code: the CUDA path, called through the C ABI,
must reproduce tests/golden/ref_golden.npz (outputs of /root/reference's matcher.cpp / frame.cpp /
map_point.cpp / bundle_adjust.cpp built into oracle/_ref; see tests/golden/make_ref_golden.py).
Matching: bit-exact.  BA: rtol 1e-6 / atol 1e-8 on poses and points, identical LM bookkeeping."""
import pytest

import ref_cases as RC
import ref_checks as CK
from lorb_slam_b200 import capi

pytestmark = pytest.mark.gpu


class _Impl:
    """capi.Context with the method names ref_checks expects."""

    def __init__(self, ctx):
        self._c = ctx

    def __getattr__(self, name):
        return getattr(self._c, name)

    def bf_crosscheck(self, q, t):
        return self._c.match_bf_crosscheck(q, t)


@pytest.fixture(scope="module")
def impl(ctx):
    return _Impl(ctx)


@pytest.mark.parametrize("c", RC.PROJ_POINTS, ids=[c[0] for c in RC.PROJ_POINTS])
def test_cuda_proj_points(impl, c):
    CK.check_proj_points(impl, c)


@pytest.mark.parametrize("c", RC.PROJ_FRAME, ids=[c[0] for c in RC.PROJ_FRAME])
def test_cuda_proj_frame(impl, c):
    CK.check_proj_frame(impl, c)


@pytest.mark.parametrize("c", RC.FRUSTUM, ids=[c[0] for c in RC.FRUSTUM])
def test_cuda_frustum(impl, c):
    CK.check_frustum(impl, c)


@pytest.mark.parametrize("c", RC.SEARCH_BF, ids=[c[0] for c in RC.SEARCH_BF])
def test_cuda_search_bf(impl, c):
    CK.check_search_bf(impl, c)


@pytest.mark.parametrize("c", RC.STEREO, ids=[c[0] for c in RC.STEREO])
def test_cuda_stereo(impl, c):
    CK.check_stereo(impl, c)


def test_cuda_compute_descriptor(impl):
    CK.check_compute_descriptor(impl)


@pytest.mark.parametrize("c", RC.POSE_ONLY, ids=[c[0] for c in RC.POSE_ONLY])
def test_cuda_pose_only(impl, c):
    CK.check_pose_only(impl, c)


@pytest.mark.parametrize("c", RC.BA_LOCAL, ids=[c[0] for c in RC.BA_LOCAL])
def test_cuda_ba_local(impl, c):
    CK.check_ba_local(impl, c, capi.ba_options)
