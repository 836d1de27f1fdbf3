"""Seeded inputs shared by tests/golden/make_ref_golden.py (which runs the compiled reference,
oracle/_ref, on them in the build container) and by the parity tests (which regenerate the same
inputs and compare the oracle and the CUDA path with the committed outputs).

Every case is rebuilt from lorb_slam_b200.synth with fixed seeds; `digest()` of the inputs is stored
next to the outputs so a drift of the generator (numpy version, edited synth) fails loudly instead
of comparing against stale vectors.
"""
import hashlib

import numpy as np

from lorb_slam_b200 import synth


def digest(*objs):
    h = hashlib.sha256()

    def feed(o):
        if isinstance(o, dict):
            for k in sorted(o):
                h.update(str(k).encode())
                feed(o[k])
        elif isinstance(o, (list, tuple)):
            for x in o:
                feed(x)
        elif isinstance(o, np.ndarray):
            h.update(str(o.dtype).encode() + str(o.shape).encode())
            h.update(np.ascontiguousarray(o).tobytes())
        else:
            h.update(repr(o).encode())

    for o in objs:
        feed(o)
    return h.hexdigest()[:16]


# ---- a6: Matcher::SearchByProjection(F, set, th); BASELINE config 2 shapes
PROJ_POINTS = [
    # name, seed, n_kp, n_pts, th, stereo, nobs, claimed_frac, inactive_frac
    ("cfg2_th1_mono", 0, 2000, 5000, 1.0, False, 1, 0.0, 0.0),
    ("cfg2_th15_mono", 0, 2000, 5000, 15.0, False, 1, 0.0, 0.0),
    ("cfg2_th1_stereo_claims", 1, 2000, 5000, 1.0, True, (0, 1, 2), 0.2, 0.1),
    ("cfg2_th3_stereo_claims", 2, 2000, 5000, 3.0, True, (0, 1, 2), 0.2, 0.1),
    ("cfg2_th15_stereo_unprotected", 3, 2000, 5000, 15.0, True, 0, 0.1, 0.05),
    ("small_th1", 4, 300, 700, 1.0, False, (0, 1), 0.3, 0.0),
]


def proj_points_case(c):
    _, seed, n_kp, n_pts, th, stereo, nobs, cf, inf = c
    fr = synth.make_frame(n_kp, seed, stereo=stereo, claimed_frac=cf)
    pts = synth.make_proj_points(fr, n_pts, seed, nobs=nobs, inactive_frac=inf)
    return fr, pts, th


# ---- a5: Matcher::SearchByProjection(Cur, Last, th)
PROJ_FRAME = [
    # name, seed, n_kp, motion, th, claim_frac
    ("forward_th15", 0, 2000, "forward", 15.0, 0.0),
    ("forward_th30", 0, 2000, "forward", 30.0, 0.2),
    ("backward_th15", 1, 2000, "backward", 15.0, 0.2),
    ("still_th15", 2, 2000, "still", 15.0, 0.2),
    ("still_th7", 3, 2000, "still", 7.0, 0.0),
    ("small_forward_th30", 4, 400, "forward", 30.0, 0.3),
]


def proj_frame_case(c):
    _, seed, n_kp, motion, th, cf = c
    cur, last = synth.make_frame_pair(n_kp, seed, motion=motion)
    if cf > 0:
        rng = np.random.default_rng(seed + 5)
        cur["kp_claim_obs"] = np.where(rng.random(n_kp) < cf, rng.integers(0, 3, n_kp), -1).astype(np.int32)
    return cur, last, th


# ---- 8(f) rank 1: Frame::IsInFrustum + MapPoint::PredictScale
FRUSTUM = [("n5000_s0", 0, 5000), ("n5000_s1", 1, 5000), ("n20000_s2", 2, 20000)]


def frustum_case(c):
    return synth.make_frustum_points(c[2], c[1])


# ---- a3 / a4: the two brute-force entry points (post-filter and assignment are the reference's)
SEARCH_BF = [
    # name, seed, n_q, n_t, kind, present_frac, use_set
    ("uniform_1000", 0, 1000, 1000, "uniform", 1.0, False),
    ("noisy_1000_holes", 1, 1000, 1000, "noisy", 0.8, False),
    ("ties_300x500_set", 2, 300, 500, "ties", 0.9, True),
    ("noisy_700x400_set", 3, 700, 400, "noisy", 1.0, True),
]


def search_bf_case(c):
    _, seed, n_q, n_t, kind, pf, use_set = c
    rng = np.random.default_rng(seed + 31337)
    if kind == "uniform":
        q, t = synth.descriptors_uniform(n_q, rng), synth.descriptors_uniform(n_t, rng)
    elif kind == "ties":
        q, t = synth.descriptors_tie_stress(n_q, rng), synth.descriptors_tie_stress(n_t, rng)
    else:
        t = synth.descriptors_uniform(n_t, rng)
        q = synth.descriptors_noisy_copy(t[rng.integers(0, n_t, n_q)], rng, 0.05)
    present = (rng.random(n_t) < pf).astype(np.uint8)
    present[0] = 1
    return q, t, present, use_set


# ---- 8(f) rank 2: Frame::ComputeStereoMatches
STEREO = [("n1000_s0", 0, 1000), ("n2000_s1", 1, 2000), ("n2000_s2", 2, 2000), ("n300_s3", 3, 300)]


def stereo_case(c):
    return synth.make_stereo_pair(c[2], c[1])


# ---- 8(f) rank 4: MapPoint::ComputeDescriptor
def compute_descriptor_case(seed=0, n_points=200):
    rng = np.random.default_rng(seed + 999)
    offs, descs = [0], []
    for _ in range(n_points):
        m = int(rng.integers(1, 25))
        base = rng.integers(0, 256, (1, 32), dtype=np.uint8)
        flips = rng.random((m, 256)) < rng.choice([0.02, 0.1, 0.3])
        d = np.packbits(np.unpackbits(np.repeat(base, m, 0), axis=1) ^ flips.astype(np.uint8), axis=1)
        descs.append(d)
        offs.append(offs[-1] + m)
    return np.asarray(offs, np.int32), np.concatenate(descs, 0)


# ---- a10: BA::ProjectPoseOptimization
POSE_ONLY = [("n300_s0", 0, 300), ("n300_s1", 1, 300), ("n1000_s2", 2, 1000), ("n50_s3", 3, 50)]


def pose_only_case(c):
    po = synth.make_pose_only(c[1], n=c[2])
    po["rt32"] = po["rt"].astype(np.float32)  # what Frame::mRvec / mTvec hold
    return po


# ---- a11: BA::LocalPoseOptimization (sizes the dense stand-in solver finishes in seconds)
BA_LOCAL = [
    # name, seed, C, P, fixed_frac, obs_per_point, max_iter (None = Ceres default 50)
    ("c4_p120", 0, 4, 120, 0.0, 4, None),
    ("c6_p200_fixed10", 1, 6, 200, 0.1, 4, None),
    ("c5_p150_fixed20", 2, 5, 150, 0.2, 4, None),
    ("c10_p300", 3, 10, 300, 0.0, 6, None),
    ("c10_p300_iter2", 3, 10, 300, 0.0, 6, 2),
]


def ba_local_case(c):
    _, seed, C, P, ff, k, _ = c
    return synth.make_ba_problem(seed, C=C, P=P, fixed_frac=ff, obs_per_point=(k,))


def ba_case_options(c, ba_options):
    return None if c[-1] is None else ba_options(max_num_iterations=c[-1])


# ---- 8(f) rank 5 (descriptor side): IC_Angle + computeOrbDescriptor of ORBextractor.cpp
ORB_DESCRIBE = [("n1000_s0", 0, 1000), ("n2000_s1", 1, 2000), ("n3000_s2", 2, 3000), ("n64_s3", 3, 64)]


def orb_describe_case(c):
    return synth.make_orb_inputs(c[2], c[1])


# ---- 8(f) rank 5 (whole extractor): ORBextractor::operator(); BASELINE config 0 shapes
ORB_EXTRACT = [
    # name, seed, width, height, nfeatures, warped
    ("640x480_n1000_s0", 0, 640, 480, 1000, False),
    ("640x480_n1000_s0_warp", 0, 640, 480, 1000, True),
    ("752x480_n2000_s1", 1, 752, 480, 2000, False),
    ("320x240_n500_s2", 2, 320, 240, 500, False),
    ("640x480_n300_s3", 3, 640, 480, 300, False),
    # levels 6 and 7 are smaller than one 30 px cell: part of the pyramid, no keypoints (the reference
    # computes nCols = 0 there and loops over no cells)
    ("161x241_n500_s4", 4, 161, 241, 500, False),
]


def orb_extract_case(c):
    img = synth.make_orb_image(c[1], c[2], c[3])
    return synth.warp_orb_image(img) if c[5] else img


# ---- ORBextractor::DistributeOctTree alone (host code): random candidate sets
QUADTREE = [
    # name, seed, n_keys, width, height (of the bordered region), n_features
    ("lvl0_like", 0, 2800, 608, 448, 217), ("dense_small_budget", 1, 3000, 608, 448, 40),
    ("sparse", 2, 150, 608, 448, 217), ("kitti_like", 3, 4000, 1209, 344, 400),
    ("duplicates", 4, 1500, 300, 200, 500), ("one_key", 5, 1, 608, 448, 100),
    ("budget_exceeds_keys", 6, 300, 400, 300, 2000), ("small_level", 7, 400, 147, 102, 61),
]


def quadtree_case(c):
    _, seed, n, w, h, nf = c
    rng = np.random.default_rng(seed + 8800)
    span = 40 if c[0] == "duplicates" else w - 6  # many keys on the same pixel
    x = rng.integers(0, span, n).astype(np.float32)
    y = rng.integers(0, min(span, h - 6), n).astype(np.float32)
    r = rng.integers(7, 80, n).astype(np.float32)
    return x, y, r, 16, 16 + w, 16, 16 + h, nf
