"""Generate tests/golden/ref_golden.npz and tests/golden/cvshim_golden.npz.

Run in the BUILD container (needs /root/reference to compile oracle/_ref, and cv2):
    python tests/golden/make_ref_golden.py

ref_golden.npz     outputs of the REFERENCE'S OWN code -- src/matcher.cpp, src/frame.cpp,
                   src/map_point.cpp, src/bundle_adjust.cpp compiled unmodified into
                   oracle/_ref/libref.so (oracle/Makefile `ref`, oracle/ref_harness.cpp) -- on the
                   seeded inputs of tests/ref_cases.py.  The matching outputs are what the
                   reference leaves in Frame::mvpMapPoints and returns; the BA outputs are what
                   its functors + problem assembly produce through the Ceres stand-in
                   (oracle/refshim/lorb_ceresshim.hpp: NOT the real Ceres).
cvshim_golden.npz  cv2 4.13 results for the OpenCV arithmetic the stand-in restates
                   (cv::Rodrigues, 4x4 float cv::invert, R*x+t gemm), so the stand-in itself is
                   pinned wherever the build container's libref.so is rebuilt.
The committed .npz files are what the tests read; the GPU box needs neither cv2 nor the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_cases as RC  # noqa: E402
from oracle import reflib as R  # noqa: E402


def summary_vec(s):
    return np.array([s["initial_cost"], s["final_cost"], s["final_radius"], s["final_gradient_max_norm"],
                     s["iterations"], s["num_successful_steps"], s["num_unsuccessful_steps"],
                     s["termination"]], np.float64)


def make_ref():
    out = {}
    for c in RC.PROJ_POINTS:
        fr, pts, th = RC.proj_points_case(c)
        r = R.search_proj_points(fr, pts, th)
        k = "pp/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(fr, pts, th))
        out[k + "/point_for_kp"] = r["point_for_kp"].astype(np.int32)
        out[k + "/n_matches"] = np.array(r["n_matches"])
    for c in RC.PROJ_FRAME:
        cur, last, th = RC.proj_frame_case(c)
        r = R.search_proj_frame(cur, last, th)
        k = "pf/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(cur, last, th))
        out[k + "/state_for_kp"] = r["state_for_kp"].astype(np.int32)
        out[k + "/n_matches"] = np.array(r["n_matches"])
    for c in RC.FRUSTUM:
        fp = RC.frustum_case(c)
        o, ow, lsf = R.frustum_project(fp)
        k = "fr/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(fp))
        out[k + "/ow"] = ow
        out[k + "/log_sf"] = np.array(lsf, np.float32)
        for f in o:
            out[k + "/" + f] = o[f]
    for c in RC.SEARCH_BF:
        q, t, present, use_set = RC.search_bf_case(c)
        n, assign = R.search_bf(q, t, present, use_set)
        k = "bf/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(q, t, present, use_set))
        out[k + "/assign"] = assign
        out[k + "/n_kept"] = np.array(n)
    for c in RC.STEREO:
        st = RC.stereo_case(c)
        r = R.stereo_matches(st)
        k = "st/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(st))
        out[k + "/uright"] = r["uright"]
        out[k + "/depth"] = r["depth"]
        out[k + "/n_matched"] = np.array(r["n_matched"])
    offs, desc = RC.compute_descriptor_case()
    chosen = np.array([R.compute_descriptor(desc[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)], np.int32)
    out["cd/digest"] = np.array(RC.digest(offs, desc))
    # the reference keeps the chosen descriptor, not its index: store the bytes
    out["cd/chosen_desc"] = np.stack([desc[offs[i] + chosen[i]] for i in range(len(chosen))])
    for c in RC.POSE_ONLY:
        po = RC.pose_only_case(c)
        r = R.ba_pose_only(po["xw"], po["uv"], po["K"], po["rt32"])
        k = "po/" + c[0]
        out[k + "/digest"] = np.array(RC.digest(po["xw"], po["uv"], po["K"], po["rt32"]))
        out[k + "/rt_f64"] = r["rt_f64"]
        out[k + "/rt_f32"] = r["rt_f32"]
        out[k + "/tcw"] = r["tcw"]
        out[k + "/summary"] = summary_vec(r["summary"])
    for c in RC.BA_LOCAL:
        pb = RC.ba_local_case(c)
        r = R.ba_local(pb, RC.ba_case_options(c, R.ba_options))
        assert r["rc"] == 0
        k = "bl/" + c[0]
        out[k + "/digest"] = np.array(RC.digest({x: pb[x] for x in ("cams", "pts", "obs_cam", "obs_pt", "obs_uv", "fix_pt", "fix_uv", "fix_rt", "K")}))
        for f in ("cams_f64", "pts_f64", "cams_f32", "pts_f32"):
            out[k + "/" + f] = r[f]
        out[k + "/summary"] = summary_vec(r["summary"])
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("ref_golden.npz:", len(out), "arrays")


def make_cvshim():
    import cv2
    rng = np.random.default_rng(20260)
    n = 1000
    rv = (rng.normal(0, 1, (n, 3)) * rng.choice([1e-9, 1e-4, 0.01, 0.3, 1.5, 3.0], (n, 1))).astype(np.float32)
    rv[0] = 0
    Rm = np.stack([cv2.Rodrigues(r.reshape(3, 1))[0] for r in rv]).astype(np.float32)
    T = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    T[:, :3, :3] = Rm
    T[:, :3, 3] = rng.normal(0, 3, (n, 3)).astype(np.float32)
    T[n // 2:] = rng.normal(0, 2, (n - n // 2, 4, 4)).astype(np.float32)  # general matrices too
    Ti = np.stack([cv2.invert(t)[1] for t in T]).astype(np.float32)
    A = rng.normal(0, 1, (n, 3, 3)).astype(np.float32)
    x = rng.normal(0, 5, (n, 3)).astype(np.float32)
    t = rng.normal(0, 2, (n, 3)).astype(np.float32)
    y = np.stack([cv2.gemm(A[i], x[i].reshape(3, 1), 1.0, t[i].reshape(3, 1), 1.0).ravel() for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "cvshim_golden.npz"), rvec=rv, rmat=Rm, T=T, Tinv=Ti, A=A, x=x,
                        t=t, y=y.astype(np.float32), cv2_version=np.array(cv2.__version__))
    print("cvshim_golden.npz: cv2", cv2.__version__)


if __name__ == "__main__":
    make_cvshim()
    make_ref()
