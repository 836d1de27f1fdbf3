"""CPU: the mapper's hand-off index (lorb_host::LocalMapIndex) on its own -- host
logic only, no device call.  The flat local-BA problem it assembles for a keyframe must be exactly what
the reference's BA::LocalPoseOptimization walks produce (src/bundle_adjust.cpp:207-303): window = current
+ non-bad covisible frames, points in first-occurrence order, per point its observers in Frame* order,
in-window ones as PoseMPCost records and the others as MPCost records with their fixed float pose."""
import ctypes as C

import numpy as np
import pytest

from lorb_slam_b200 import synth
from lorb_slam_b200.host import build_host


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def _scenario(seed, n_kf, P, null_slots, n_cov):
    pb = synth.make_ba_problem(seed, C=n_kf, P=P, obs_per_point=(3, 4, 5))
    rng = np.random.default_rng(seed)
    slots = [[] for _ in range(n_kf)]
    for c, p, uv in zip(pb["obs_cam"], pb["obs_pt"], pb["obs_uv"]):
        slots[c].append((int(p), uv))
    for k in range(n_kf):
        slots[k] += [(-1, np.zeros(2, np.float32))] * null_slots
        rng.shuffle(slots[k])
    cov = [[j for j in range(k - 1, max(-1, k - 1 - n_cov), -1)] for k in range(n_kf)]
    return pb, slots, cov


def _expected(pb, slots, cov, k_ba, bad_kf, bad_pt):
    window = [k_ba] + [f for f in cov[k_ba] if not bad_kf[f]]
    widx = {}
    for i, f in enumerate(window):
        widx.setdefault(f, i)
    order, seen = [], set()
    for f in window:
        for p, _ in slots[f]:
            if p >= 0 and p not in seen and not bad_pt[p]:
                seen.add(p)
                order.append(p)
    oc, op, ouv, fp, fuv, frt = [], [], [], [], [], []
    cams = pb["cams"].astype(np.float32)
    for i, p in enumerate(order):
        for f in range(len(slots)):  # frames live in one array: address order == index order
            if bad_kf[f]:
                continue
            for q, uv in slots[f]:
                if q != p:
                    continue
                if f in widx:
                    oc.append(widx[f]); op.append(i); ouv.append(uv)
                else:
                    fp.append(i); fuv.append(uv); frt.append(cams[f])
    return window, order, oc, op, np.array(ouv, np.float32).reshape(-1, 2), fp, \
        np.array(fuv, np.float32).reshape(-1, 2), np.array(frt, np.float32).reshape(-1, 6)


@pytest.mark.parametrize("seed,n_kf,P,n_cov,k_ba,bad", [(1, 6, 300, 3, 5, False), (2, 9, 500, 4, 6, True),
                                                        (3, 3, 60, 2, 2, False), (4, 8, 400, 7, 7, True)])
def test_assemble_matches_the_reference_walk(seed, n_kf, P, n_cov, k_ba, bad):
    H = C.CDLL(build_host.build())
    pb, slots, cov = _scenario(seed, n_kf, P, 4, n_cov)
    rng = np.random.default_rng(100 + seed)
    bad_kf = np.zeros(n_kf, np.uint8)
    bad_pt = np.zeros(P, np.uint8)
    if bad:
        bad_kf[rng.integers(0, k_ba)] = 1          # one culled keyframe (not the current one)
        bad_pt[rng.integers(0, P, P // 20)] = 1    # some culled map points
    kf_rt = pb["cams"].astype(np.float32)
    pts = pb["pts"].astype(np.float32)
    kf_off = np.cumsum([0] + [len(x) for x in slots]).astype(np.int32)
    slot_pt = np.array([p for x in slots for p, _ in x], np.int32)
    slot_uv = np.array([uv for x in slots for _, uv in x], np.float32)
    cov_off = np.cumsum([0] + [len(x) for x in cov]).astype(np.int32)
    cov_idx = np.array([j for x in cov for j in x] + [0], np.int32)
    cap = len(slot_pt) + 16
    ow, opn = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    ooc, oop, ofp = (np.zeros(cap, np.int32) for _ in range(3))
    oouv, ofuv, ofrt = np.zeros((cap, 2), np.float32), np.zeros((cap, 2), np.float32), np.zeros((cap, 6), np.float32)
    sizes = np.zeros(7, np.int32)
    rc = H.harness_local_map_assemble(n_kf, _p(kf_rt), _p(kf_off), _p(slot_pt), _p(slot_uv), _p(cov_off),
                                      _p(cov_idx), P, _p(pts), _p(bad_kf), _p(bad_pt), k_ba, cap, _p(ow),
                                      _p(opn), _p(ooc), _p(oop), _p(oouv), _p(ofp), _p(ofuv), _p(ofrt), _p(sizes))
    assert rc == 0
    Cw, Pw, O, F = (int(x) for x in sizes[:4])
    window, order, oc, op, ouv, fp, fuv, frt = _expected(pb, slots, cov, k_ba, bad_kf, bad_pt)
    assert ow[:Cw].tolist() == window and opn[:Pw].tolist() == order
    assert ooc[:O].tolist() == oc and oop[:O].tolist() == op and np.array_equal(oouv[:O], ouv)
    assert ofp[:F].tolist() == fp and np.array_equal(ofuv[:F], fuv) and np.array_equal(ofrt[:F], frt)
    # the index itself: every keyframe and every map point held by a slot known; bad map points gain
    # no observation at insert time (src/local_mapping.cpp:59-60); one entry per (point, frame)
    n_obs = sum(1 for x in slots for p, _ in x if p >= 0 and not bad_pt[p])
    n_pts = len({p for x in slots for p, _ in x if p >= 0})
    assert sizes[4] == n_kf and sizes[5] == n_pts and sizes[6] == n_obs
