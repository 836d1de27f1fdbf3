"""CPU: the restated oracle (oracle/match_ref.c, oracle/ba_ref.cpp) against the outputs of the
REFERENCE'S OWN compiled code (tests/golden/ref_golden.npz, made from oracle/_ref by
tests/golden/make_ref_golden.py).  This is what pins the oracle; the GPU twin of this file is
tests/test_ref_golden_gpu.py."""
import numpy as np
import pytest

import ref_cases as RC
import ref_checks as CK
from oracle import ref


@pytest.mark.parametrize("c", RC.PROJ_POINTS, ids=[c[0] for c in RC.PROJ_POINTS])
def test_oracle_proj_points(c):
    CK.check_proj_points(ref, c)


@pytest.mark.parametrize("c", RC.PROJ_FRAME, ids=[c[0] for c in RC.PROJ_FRAME])
def test_oracle_proj_frame(c):
    CK.check_proj_frame(ref, c)


@pytest.mark.parametrize("c", RC.FRUSTUM, ids=[c[0] for c in RC.FRUSTUM])
def test_oracle_frustum(c):
    CK.check_frustum(ref, c)


@pytest.mark.parametrize("c", RC.SEARCH_BF, ids=[c[0] for c in RC.SEARCH_BF])
def test_oracle_search_bf(c):
    CK.check_search_bf(ref, c)


@pytest.mark.parametrize("c", RC.STEREO, ids=[c[0] for c in RC.STEREO])
def test_oracle_stereo(c):
    CK.check_stereo(ref, c)


def test_oracle_compute_descriptor():
    CK.check_compute_descriptor(ref)


@pytest.mark.parametrize("c", RC.POSE_ONLY, ids=[c[0] for c in RC.POSE_ONLY])
def test_oracle_pose_only(c):
    CK.check_pose_only(ref, c)


@pytest.mark.parametrize("c", RC.BA_LOCAL, ids=[c[0] for c in RC.BA_LOCAL])
def test_oracle_ba_local(c):
    CK.check_ba_local(ref, c, ref.ba_options)
