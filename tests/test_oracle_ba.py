"""CPU: the BA oracle (Ceres-contract restatement, parity vs Ceres unpinned)
cross-checked against finite differences and scipy.optimize.least_squares."""
import numpy as np
import pytest

from lorb_slam_b200 import synth
from oracle import ref


def test_jet_jacobians_against_finite_differences():
    rng = np.random.default_rng(1)
    worst = 0.0
    for i in range(60):
        cam = np.concatenate([rng.normal(0, 10 ** rng.uniform(-3, 0.6), 3), rng.normal(0, 1, 3)])
        pt = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(4, 20)])
        uv = np.array([300.0, 200.0], np.float32)
        kind = 0 if i % 2 == 0 else 2
        r, Jc, Jp = ref.ba_residual_jac(kind, cam, pt, uv, synth.K_DEFAULT)
        x = np.concatenate([cam, pt])
        J = np.concatenate([Jc, Jp], 1)
        ncol = 9 if kind == 0 else 6  # PoseCost has a constant (float) point
        for k in range(ncol):
            h = 1e-6 * max(1.0, abs(x[k]))
            xp, xm = x.copy(), x.copy()
            xp[k] += h
            xm[k] -= h
            fp = ref.ba_residual_jac(kind, xp[:6], xp[6:], uv, synth.K_DEFAULT)[0]
            fm = ref.ba_residual_jac(kind, xm[:6], xm[6:], uv, synth.K_DEFAULT)[0]
            fd = (fp - fm) / (2 * h)
            worst = max(worst, float(np.max(np.abs(fd - J[:, k]) / (np.abs(J[:, k]) + 1e-3))))
    assert worst < 5e-5


def test_small_angle_branch_is_first_order():
    """|w|^2 <= DBL_EPSILON: R X = X + w x X (ceres::AngleAxisRotatePoint)."""
    cam = np.array([1e-9, -2e-9, 3e-9, 0.1, -0.2, 0.3])
    pt = np.array([0.5, -0.4, 6.0])
    uv = np.zeros(2, np.float32)
    r, Jc, Jp = ref.ba_residual_jac(0, cam, pt, uv, synth.K_DEFAULT)
    q = pt + np.cross(cam[:3], pt) + cam[3:]
    fx, fy, cx, cy = (float(v) for v in synth.K_DEFAULT)
    np.testing.assert_allclose(r, [q[0] / q[2] * fx + cx, q[1] / q[2] * fy + cy], rtol=1e-14)


def test_pose_cost_uses_fx_for_v():
    """reference src/bundle_adjust.cpp:51."""
    K = np.array([400.0, 777.0, 320.0, 240.0], np.float32)
    cam = np.array([0.01, 0.02, -0.01, 0.1, 0.2, 0.3])
    pt = np.array([0.5, -0.75, 5.0])  # float-representable: PoseCost holds the point as cv::Point3f
    r_pose, _, _ = ref.ba_residual_jac(2, cam, pt, np.zeros(2, np.float32), K)
    r_mp, _, _ = ref.ba_residual_jac(0, cam, pt, np.zeros(2, np.float32), K)
    assert abs(r_pose[0] - r_mp[0]) < 1e-12
    assert abs((r_pose[1] - 240.0) * 777.0 / 400.0 - (r_mp[1] - 240.0)) < 1e-9


def test_lm_converges_and_matches_scipy_minimum():
    from scipy.optimize import least_squares
    pb = synth.make_ba_problem(11, C=4, P=60, obs_per_point=(3, 4), fixed_frac=0.2)
    cams, pts, s = ref.ba_local(pb, ref.ba_options(max_num_iterations=50))
    assert s["termination"] in (1, 2, 3) and s["final_cost"] < 0.2 * s["initial_cost"]
    C, P = len(pb["cams"]), len(pb["pts"])

    def fun(x):
        c, p = x[:6 * C].reshape(C, 6), x[6 * C:].reshape(P, 3)
        # residual vector with the same cost: evaluate through the oracle per block
        res = []
        for i in range(pb["O"]):
            r, _, _ = ref.ba_residual_jac(0, c[pb["obs_cam"][i]], p[pb["obs_pt"][i]],
                                          pb["obs_uv"][i], pb["K"])
            res.extend(r)
        for i in range(pb["F"]):
            r, _, _ = ref.ba_residual_jac(0, pb["fix_rt"][i].astype(np.float64), p[pb["fix_pt"][i]],
                                          pb["fix_uv"][i], pb["K"])
            res.extend(r)
        return np.asarray(res)

    x0 = np.concatenate([cams.ravel(), pts.ravel()])
    assert abs(0.5 * np.sum(fun(x0) ** 2) - s["final_cost"]) < 1e-9 * s["final_cost"]
    sol = least_squares(fun, x0, method="trf", x_scale="jac", xtol=1e-14, ftol=1e-14, gtol=1e-12,
                        max_nfev=30)
    # the oracle's answer is (to its function tolerance) the minimum scipy finds from there
    assert sol.cost <= s["final_cost"] * (1 + 1e-12)
    assert (s["final_cost"] - sol.cost) / s["final_cost"] < 1e-5


def test_pose_only_recovers_pose():
    po = synth.make_pose_only(3, 400, pixel_noise=0.2)
    rt, s = ref.ba_pose_only(po["xw"], po["uv"], po["K"], po["rt"])
    assert s["final_cost"] < s["initial_cost"] * 0.05
    np.testing.assert_allclose(rt, po["rt_true"], atol=5e-3)


def test_rejected_steps_shrink_radius():
    pb = synth.make_ba_problem(4, C=6, P=400, pose_noise=(0.08, 0.4), point_noise=0.5)
    _, _, s = ref.ba_local(pb, ref.ba_options(max_num_iterations=25))
    assert s["num_successful_steps"] >= 3 and s["final_cost"] < s["initial_cost"]
    _, _, s0 = ref.ba_local(pb, ref.ba_options(max_num_iterations=0))
    assert s0["iterations"] == 0 and s0["final_cost"] == s0["initial_cost"]
