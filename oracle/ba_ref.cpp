/*
 * oracle/ba_ref.cpp — CPU fp64 restatement of the reference's bundle adjustment.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may call, link or
 * import this file; it is the checker for the CUDA solver (tests/,
 * __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
 *
 * PARITY UNPINNED.  The reference delegates all BA arithmetic to Ceres Solver
 * (CMakeLists.txt:19 `find_package(Ceres REQUIRED)`, version not pinned, not
 * vendored, absent from this image).  What the reference itself contributes is
 *   - three autodiff cost functors  src/bundle_adjust.cpp:22-64 (PoseCost),
 *     :68-113 (MPCost), :116-151 (PoseMPCost),
 *   - Solver::Options with only linear_solver_type = DENSE_SCHUR (:189-191, :308-310),
 *   - the problem assembly (:173-186, :270-303).
 * This file restates those functors with forward-mode Jets (the arithmetic
 * ceres::AutoDiffCostFunction performs) and restates Ceres' published
 * algorithm for everything else: trust-region Levenberg-Marquardt with Jacobi
 * column scaling, clamped LM diagonal, dense Schur elimination of the point
 * blocks, dense Cholesky, step-quality based radius update and Ceres' default
 * tolerances.  No Ceres golden vectors exist anywhere in
 * the reference; the restatement is cross-checked at convergence
 * against scipy.optimize.least_squares (tests/test_oracle_ba.py).
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/lorb_cuda.h"

namespace {

/* ---------------------------------------------------------------- Jets
 * value + N partial derivatives, with Ceres' jet.h operator definitions
 * (product rule; quotient as f.a * (1/g.a); sqrt/sin/cos chain rules). */
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) {
    for (int i = 0; i < N; i++) v[i] = 0;
  }
  explicit Jet(double x) : a(x) {
    for (int i = 0; i < N; i++) v[i] = 0;
  }
  Jet(double x, int k) : a(x) {
    for (int i = 0; i < N; i++) v[i] = 0;
    v[k] = 1.0;
  }
};
template <int N>
Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h;
  h.a = f.a + g.a;
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i];
  return h;
}
template <int N>
Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h;
  h.a = f.a - g.a;
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i];
  return h;
}
template <int N>
Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h;
  h.a = f.a * g.a;
  for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
template <int N>
Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h;
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  h.a = f_a_by_g_a;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
template <int N>
Jet<N> sqrt(const Jet<N>& f) {
  Jet<N> h;
  const double tmp = std::sqrt(f.a);
  const double two_a_inverse = 1.0 / (2.0 * tmp);
  h.a = tmp;
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
template <int N>
Jet<N> cos(const Jet<N>& f) {
  Jet<N> h;
  h.a = std::cos(f.a);
  const double s = -std::sin(f.a);
  for (int i = 0; i < N; i++) h.v[i] = s * f.v[i];
  return h;
}
template <int N>
Jet<N> sin(const Jet<N>& f) {
  Jet<N> h;
  h.a = std::sin(f.a);
  const double c = std::cos(f.a);
  for (int i = 0; i < N; i++) h.v[i] = c * f.v[i];
  return h;
}
template <int N>
bool operator>(const Jet<N>& f, const Jet<N>& g) {
  return f.a > g.a;
}

/* plain-double versions so the same templates evaluate cost only */
inline double make_T(double x, const double*) { return x; }
template <int N>
Jet<N> make_T(double x, const Jet<N>*) {
  return Jet<N>(x);
}
using std::cos;
using std::sin;
using std::sqrt;

/* ceres::AngleAxisRotatePoint (ceres/rotation.h) as called at
 * src/bundle_adjust.cpp:44, :96, :135. */
template <typename T>
void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  const T theta2 =
      angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
  if (theta2 > make_T(std::numeric_limits<double>::epsilon(), (T*)0)) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = make_T(1.0, (T*)0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse,
                    angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp =
        (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (make_T(1.0, (T*)0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1],
                             angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

/* The projection shared by the three functors: u = x/z*fx + cx, v = y/z*fv + cy
 * (src/bundle_adjust.cpp:50-54, :100-104, :139-143).  fv is fy for MPCost and
 * PoseMPCost and — faithfully to :51 — fx for PoseCost. */
template <typename T>
void reproject(const T rvec[3], const T tvec[3], const T X[3], float fu, float fv, float cx,
               float cy, float ou, float ov, T residual[2]) {
  T q[3];
  AngleAxisRotatePoint(rvec, X, q);
  q[0] = q[0] + tvec[0];
  q[1] = q[1] + tvec[1];
  q[2] = q[2] + tvec[2];
  const T u = q[0] / q[2] * make_T((double)fu, (T*)0) + make_T((double)cx, (T*)0);
  const T v = q[1] / q[2] * make_T((double)fv, (T*)0) + make_T((double)cy, (T*)0);
  residual[0] = u - make_T((double)ou, (T*)0);
  residual[1] = v - make_T((double)ov, (T*)0);
}

enum ResKind { RES_POSE_MP = 0, RES_MP_FIXED = 1, RES_POSE_ONLY = 2 };

struct Residual {
  int kind;
  int cam;      /* window camera index (POSE_MP, POSE_ONLY) */
  int pt;       /* point index (POSE_MP, MP_FIXED); index into const_xw for POSE_ONLY */
  float u, v;   /* observation */
  float frt[6]; /* fixed float pose for MP_FIXED (mRvec/mTvec.at<float>, :87,:98) */
};

struct Problem {
  int C = 0, P = 0;
  std::vector<double> cams; /* C x 6 */
  std::vector<double> pts;  /* P x 3 */
  std::vector<float> const_xw; /* constant points of POSE_ONLY residuals */
  std::vector<Residual> res;
  float fx, fy, cx, cy;
  /* per point residual lists (CSR) */
  std::vector<int> pt_start, pt_res;
  std::vector<int> nopt_res; /* residuals without a variable point (POSE_ONLY) */
};

struct Lin {                 /* linearisation at x */
  std::vector<double> r;     /* 2 per residual */
  std::vector<double> Jc;    /* 12 per residual: d r / d cam (2x6 row-major) */
  std::vector<double> Jp;    /* 6 per residual:  d r / d point (2x3 row-major) */
};

/* Evaluate residual i at (cams, pts); optionally Jacobian blocks via Jets. */
void eval_residual(const Problem& pb, const double* cams, const double* pts, int i, double r[2],
                   double* Jc, double* Jp) {
  const Residual& R = pb.res[i];
  if (!Jc && !Jp) {
    double rv[3], tv[3], X[3];
    if (R.kind == RES_MP_FIXED) {
      for (int k = 0; k < 3; k++) {
        rv[k] = (double)R.frt[k];
        tv[k] = (double)R.frt[3 + k];
        X[k] = pts[3 * R.pt + k];
      }
      reproject<double>(rv, tv, X, pb.fx, pb.fy, pb.cx, pb.cy, R.u, R.v, r);
    } else if (R.kind == RES_POSE_MP) {
      for (int k = 0; k < 3; k++) {
        rv[k] = cams[6 * R.cam + k];
        tv[k] = cams[6 * R.cam + 3 + k];
        X[k] = pts[3 * R.pt + k];
      }
      reproject<double>(rv, tv, X, pb.fx, pb.fy, pb.cx, pb.cy, R.u, R.v, r);
    } else {
      for (int k = 0; k < 3; k++) {
        rv[k] = cams[6 * R.cam + k];
        tv[k] = cams[6 * R.cam + 3 + k];
        X[k] = (double)pb.const_xw[3 * R.pt + k];
      }
      reproject<double>(rv, tv, X, pb.fx, pb.fx /* :51 */, pb.cx, pb.cy, R.u, R.v, r);
    }
    return;
  }
  /* 9 independent variables: (pose 0..5, point 6..8) */
  typedef Jet<9> J;
  J rv[3], tv[3], X[3], res[2];
  if (R.kind == RES_MP_FIXED) {
    for (int k = 0; k < 3; k++) {
      rv[k] = J((double)R.frt[k]);
      tv[k] = J((double)R.frt[3 + k]);
      X[k] = J(pts[3 * R.pt + k], 6 + k);
    }
    reproject<J>(rv, tv, X, pb.fx, pb.fy, pb.cx, pb.cy, R.u, R.v, res);
  } else if (R.kind == RES_POSE_MP) {
    for (int k = 0; k < 3; k++) {
      rv[k] = J(cams[6 * R.cam + k], k);
      tv[k] = J(cams[6 * R.cam + 3 + k], 3 + k);
      X[k] = J(pts[3 * R.pt + k], 6 + k);
    }
    reproject<J>(rv, tv, X, pb.fx, pb.fy, pb.cx, pb.cy, R.u, R.v, res);
  } else {
    for (int k = 0; k < 3; k++) {
      rv[k] = J(cams[6 * R.cam + k], k);
      tv[k] = J(cams[6 * R.cam + 3 + k], 3 + k);
      X[k] = J((double)pb.const_xw[3 * R.pt + k]);
    }
    reproject<J>(rv, tv, X, pb.fx, pb.fx /* :51 */, pb.cx, pb.cy, R.u, R.v, res);
  }
  for (int a = 0; a < 2; a++) {
    r[a] = res[a].a;
    for (int k = 0; k < 6; k++) Jc[6 * a + k] = res[a].v[k];
    for (int k = 0; k < 3; k++) Jp[3 * a + k] = res[a].v[6 + k];
  }
}

double eval_cost(const Problem& pb, const double* cams, const double* pts) {
  double c = 0;
  for (size_t i = 0; i < pb.res.size(); i++) {
    double r[2];
    eval_residual(pb, cams, pts, (int)i, r, nullptr, nullptr);
    c += r[0] * r[0] + r[1] * r[1];
  }
  return 0.5 * c;
}

double linearize(const Problem& pb, Lin& L) {
  const size_t n = pb.res.size();
  L.r.assign(2 * n, 0.0);
  L.Jc.assign(12 * n, 0.0);
  L.Jp.assign(6 * n, 0.0);
  double c = 0;
  for (size_t i = 0; i < n; i++) {
    eval_residual(pb, pb.cams.data(), pb.pts.data(), (int)i, &L.r[2 * i], &L.Jc[12 * i],
                  &L.Jp[6 * i]);
    c += L.r[2 * i] * L.r[2 * i] + L.r[2 * i + 1] * L.r[2 * i + 1];
  }
  return 0.5 * c;
}

inline bool has_cam(const Residual& R) { return R.kind != RES_MP_FIXED; }
inline bool has_pt(const Residual& R) { return R.kind != RES_POSE_ONLY; }

/* in-place dense Cholesky A = L L^T (lower), returns false on non-positive pivot */
bool cholesky(std::vector<double>& A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
void chol_solve(const std::vector<double>& Lm, int n, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= Lm[(size_t)i * n + k] * b[k];
    b[i] = s / Lm[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= Lm[(size_t)k * n + i] * b[k];
    b[i] = s / Lm[(size_t)i * n + i];
  }
}

/* symmetric 3x3 inverse via Cholesky; false if not positive definite */
bool inv3_spd(const double H[9], double Hi[9]) {
  std::vector<double> A(H, H + 9);
  if (!cholesky(A, 3)) return false;
  for (int c = 0; c < 3; c++) {
    double e[3] = {0, 0, 0};
    e[c] = 1.0;
    chol_solve(A, 3, e);
    for (int r = 0; r < 3; r++) Hi[3 * r + c] = e[r];
  }
  return true;
}

/* The whole trust-region loop (Ceres TrustRegionMinimizer +
 * LevenbergMarquardtStrategy + SchurComplementSolver(DENSE_SCHUR)). */
int solve_lm(Problem& pb, const lorb_ba_options& opt, lorb_ba_summary* sum) {
  const int C = pb.C, P = pb.P, nc = 6 * C, np = 3 * P;
  const size_t nres = pb.res.size();
  /* residual lists per point */
  pb.pt_start.assign(P + 1, 0);
  pb.nopt_res.clear();
  for (size_t i = 0; i < nres; i++) {
    if (has_pt(pb.res[i]))
      pb.pt_start[pb.res[i].pt + 1]++;
    else
      pb.nopt_res.push_back((int)i);
  }
  for (int p = 0; p < P; p++) pb.pt_start[p + 1] += pb.pt_start[p];
  pb.pt_res.assign(pb.pt_start[P], 0);
  {
    std::vector<int> fill(pb.pt_start.begin(), pb.pt_start.end() - 1);
    for (size_t i = 0; i < nres; i++)
      if (has_pt(pb.res[i])) pb.pt_res[fill[pb.res[i].pt]++] = (int)i;
  }

  Lin L;
  double cost = linearize(pb, L);
  std::vector<double> scale_c(nc, 1.0), scale_p(np, 1.0);
  std::vector<double> grad_c(nc), grad_p(np);
  auto gradient_max = [&]() {
    std::fill(grad_c.begin(), grad_c.end(), 0.0);
    std::fill(grad_p.begin(), grad_p.end(), 0.0);
    for (size_t i = 0; i < nres; i++) {
      const Residual& R = pb.res[i];
      for (int a = 0; a < 2; a++) {
        if (has_cam(R))
          for (int k = 0; k < 6; k++) grad_c[6 * R.cam + k] += L.Jc[12 * i + 6 * a + k] * L.r[2 * i + a];
        if (has_pt(R))
          for (int k = 0; k < 3; k++) grad_p[3 * R.pt + k] += L.Jp[6 * i + 3 * a + k] * L.r[2 * i + a];
      }
    }
    double m = 0;
    for (double g : grad_c) m = std::fmax(m, std::fabs(g));
    for (double g : grad_p) m = std::fmax(m, std::fabs(g));
    return m;
  };
  auto col_sqnorms = [&](std::vector<double>& dc, std::vector<double>& dp) {
    dc.assign(nc, 0.0);
    dp.assign(np, 0.0);
    for (size_t i = 0; i < nres; i++) {
      const Residual& R = pb.res[i];
      for (int a = 0; a < 2; a++) {
        if (has_cam(R))
          for (int k = 0; k < 6; k++) {
            double j = L.Jc[12 * i + 6 * a + k];
            dc[6 * R.cam + k] += j * j;
          }
        if (has_pt(R))
          for (int k = 0; k < 3; k++) {
            double j = L.Jp[6 * i + 3 * a + k];
            dp[3 * R.pt + k] += j * j;
          }
      }
    }
  };
  auto scale_columns = [&]() {
    for (size_t i = 0; i < nres; i++) {
      const Residual& R = pb.res[i];
      for (int a = 0; a < 2; a++) {
        if (has_cam(R))
          for (int k = 0; k < 6; k++) L.Jc[12 * i + 6 * a + k] *= scale_c[6 * R.cam + k];
        if (has_pt(R))
          for (int k = 0; k < 3; k++) L.Jp[6 * i + 3 * a + k] *= scale_p[3 * R.pt + k];
      }
    }
  };
  auto x_norm_of = [&]() {
    double s = 0;
    for (double v : pb.cams) s += v * v;
    for (double v : pb.pts) s += v * v;
    return std::sqrt(s);
  };

  sum->initial_cost = cost;
  sum->num_successful_steps = 0;
  sum->num_unsuccessful_steps = 0;
  double gmax = gradient_max(); /* on the unscaled Jacobian */
  if (opt.jacobi_scaling) {
    std::vector<double> dc, dp;
    col_sqnorms(dc, dp);
    for (int k = 0; k < nc; k++) scale_c[k] = 1.0 / (1.0 + std::sqrt(dc[k]));
    for (int k = 0; k < np; k++) scale_p[k] = 1.0 / (1.0 + std::sqrt(dp[k]));
    scale_columns();
  }
  double x_norm = x_norm_of();
  double radius = opt.initial_trust_region_radius;
  double decrease_factor = 2.0;
  int iteration = 0, invalid_run = 0;
  int termination = LORB_BA_NO_CONVERGENCE;
  if (gmax <= opt.gradient_tolerance) termination = LORB_BA_CONV_GRADIENT;

  std::vector<double> S((size_t)nc * nc), rhs(nc), yc(nc), yp(np), Hpp_inv((size_t)9 * P),
      gp(np);
  std::vector<double> dcq, dpq;
  std::vector<double> cand_c(nc), cand_p(np);

  while (termination == LORB_BA_NO_CONVERGENCE) {
    if (iteration >= opt.max_num_iterations) break;
    if (radius < opt.min_trust_region_radius) {
      termination = LORB_BA_CONV_RADIUS;
      break;
    }
    iteration++;
    /* ---- LevenbergMarquardtStrategy::ComputeStep */
    col_sqnorms(dcq, dpq); /* of the scaled Jacobian */
    for (double& d : dcq) d = std::fmin(std::fmax(d, opt.min_lm_diagonal), opt.max_lm_diagonal) / radius;
    for (double& d : dpq) d = std::fmin(std::fmax(d, opt.min_lm_diagonal), opt.max_lm_diagonal) / radius;
    /* ---- SchurComplementSolver: S = F'F + Df^2 - F'E (E'E + De^2)^-1 E'F */
    std::fill(S.begin(), S.end(), 0.0);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    for (int k = 0; k < nc; k++) S[(size_t)k * nc + k] = dcq[k];
    auto add_cam_block = [&](int i) {
      const Residual& R = pb.res[i];
      const double* J = &L.Jc[12 * (size_t)i];
      const double* r = &L.r[2 * (size_t)i];
      for (int a = 0; a < 6; a++) {
        for (int b = 0; b < 6; b++)
          S[(size_t)(6 * R.cam + a) * nc + 6 * R.cam + b] += J[a] * J[b] + J[6 + a] * J[6 + b];
        rhs[6 * R.cam + a] += J[a] * r[0] + J[6 + a] * r[1];
      }
    };
    for (int i : pb.nopt_res) add_cam_block(i);
    bool ok = true;
    for (int p = 0; p < P && ok; p++) {
      double H[9] = {dpq[3 * p], 0, 0, 0, dpq[3 * p + 1], 0, 0, 0, dpq[3 * p + 2]};
      double g[3] = {0, 0, 0};
      for (int e = pb.pt_start[p]; e < pb.pt_start[p + 1]; e++) {
        const int i = pb.pt_res[e];
        const double* J = &L.Jp[6 * (size_t)i];
        const double* r = &L.r[2 * (size_t)i];
        for (int a = 0; a < 3; a++) {
          for (int b = 0; b < 3; b++) H[3 * a + b] += J[a] * J[b] + J[3 + a] * J[3 + b];
          g[a] += J[a] * r[0] + J[3 + a] * r[1];
        }
        if (has_cam(pb.res[i])) add_cam_block(i);
      }
      double* Hi = &Hpp_inv[(size_t)9 * p];
      if (!inv3_spd(H, Hi)) {
        ok = false;
        break;
      }
      for (int a = 0; a < 3; a++) gp[3 * p + a] = g[a];
      /* W_i = Jc_i^T Jp_i (6x3);  Y_i = W_i Hinv */
      for (int e = pb.pt_start[p]; e < pb.pt_start[p + 1]; e++) {
        const int i = pb.pt_res[e];
        if (!has_cam(pb.res[i])) continue;
        const int ci = pb.res[i].cam;
        double Wi[18], Yi[18];
        const double* Jc = &L.Jc[12 * (size_t)i];
        const double* Jp = &L.Jp[6 * (size_t)i];
        for (int a = 0; a < 6; a++)
          for (int b = 0; b < 3; b++) Wi[3 * a + b] = Jc[a] * Jp[b] + Jc[6 + a] * Jp[3 + b];
        for (int a = 0; a < 6; a++)
          for (int b = 0; b < 3; b++)
            Yi[3 * a + b] = Wi[3 * a] * Hi[b] + Wi[3 * a + 1] * Hi[3 + b] + Wi[3 * a + 2] * Hi[6 + b];
        for (int a = 0; a < 6; a++)
          rhs[6 * ci + a] -= Yi[3 * a] * g[0] + Yi[3 * a + 1] * g[1] + Yi[3 * a + 2] * g[2];
        for (int f = pb.pt_start[p]; f < pb.pt_start[p + 1]; f++) {
          const int j = pb.pt_res[f];
          if (!has_cam(pb.res[j])) continue;
          const int cj = pb.res[j].cam;
          const double* Jc2 = &L.Jc[12 * (size_t)j];
          const double* Jp2 = &L.Jp[6 * (size_t)j];
          for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) {
              /* (Y_i W_j^T)[a][b] = sum_k Yi[a][k] * Wj[b][k] */
              double s = 0;
              for (int k = 0; k < 3; k++)
                s += Yi[3 * a + k] * (Jc2[b] * Jp2[k] + Jc2[6 + b] * Jp2[3 + k]);
              S[(size_t)(6 * ci + a) * nc + 6 * cj + b] -= s;
            }
        }
      }
    }
    std::vector<double> Sf = S;
    if (ok && nc > 0) ok = cholesky(Sf, nc);
    bool step_valid = ok;
    double model_cost_change = 0;
    if (ok) {
      yc = rhs;
      if (nc > 0) chol_solve(Sf, nc, yc.data());
      for (int p = 0; p < P; p++) {
        double b[3] = {gp[3 * p], gp[3 * p + 1], gp[3 * p + 2]};
        for (int e = pb.pt_start[p]; e < pb.pt_start[p + 1]; e++) {
          const int i = pb.pt_res[e];
          if (!has_cam(pb.res[i])) continue;
          const int ci = pb.res[i].cam;
          const double* Jc = &L.Jc[12 * (size_t)i];
          const double* Jp = &L.Jp[6 * (size_t)i];
          /* b -= W_i^T yc_ci  = Jp^T (Jc yc) */
          double t0 = 0, t1 = 0;
          for (int a = 0; a < 6; a++) {
            t0 += Jc[a] * yc[6 * ci + a];
            t1 += Jc[6 + a] * yc[6 * ci + a];
          }
          for (int k = 0; k < 3; k++) b[k] -= Jp[k] * t0 + Jp[3 + k] * t1;
        }
        const double* Hi = &Hpp_inv[(size_t)9 * p];
        for (int a = 0; a < 3; a++) yp[3 * p + a] = Hi[3 * a] * b[0] + Hi[3 * a + 1] * b[1] + Hi[3 * a + 2] * b[2];
      }
      /* step = -y ; model_cost_change = -(J step).(r + J step / 2) */
      for (int k = 0; k < nc; k++)
        if (!std::isfinite(yc[k])) step_valid = false;
      for (int k = 0; k < np; k++)
        if (!std::isfinite(yp[k])) step_valid = false;
      if (step_valid) {
        double acc = 0;
        for (size_t i = 0; i < nres; i++) {
          const Residual& R = pb.res[i];
          for (int a = 0; a < 2; a++) {
            double m = 0;
            if (has_cam(R))
              for (int k = 0; k < 6; k++) m -= L.Jc[12 * i + 6 * a + k] * yc[6 * R.cam + k];
            if (has_pt(R))
              for (int k = 0; k < 3; k++) m -= L.Jp[6 * i + 3 * a + k] * yp[3 * R.pt + k];
            acc += m * (L.r[2 * i + a] + m / 2.0);
          }
        }
        model_cost_change = -acc;
        step_valid = model_cost_change > 0.0;
      }
    }
    if (!step_valid) {
      /* HandleInvalidStep */
      invalid_run++;
      sum->num_unsuccessful_steps++;
      if (invalid_run >= opt.max_consecutive_invalid_steps) {
        termination = LORB_BA_FAILURE;
        break;
      }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    invalid_run = 0;
    /* delta = step .* scaling ; candidate = x + delta */
    double step_sq = 0;
    for (int k = 0; k < nc; k++) {
      double d = -yc[k] * scale_c[k];
      cand_c[k] = pb.cams[k] + d;
      step_sq += d * d;
    }
    for (int k = 0; k < np; k++) {
      double d = -yp[k] * scale_p[k];
      cand_p[k] = pb.pts[k] + d;
      step_sq += d * d;
    }
    const double cand_cost = eval_cost(pb, cand_c.data(), cand_p.data());
    const double step_norm = std::sqrt(step_sq);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      termination = LORB_BA_CONV_PARAMETER;
      break;
    }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * cost) {
      termination = LORB_BA_CONV_FUNCTION;
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {
      /* HandleSuccessfulStep */
      pb.cams.assign(cand_c.begin(), cand_c.end());
      pb.pts.assign(cand_p.begin(), cand_p.end());
      x_norm = x_norm_of();
      cost = linearize(pb, L);
      gmax = gradient_max();
      if (opt.jacobi_scaling) scale_columns();
      sum->num_successful_steps++;
      if (gmax <= opt.gradient_tolerance) {
        termination = LORB_BA_CONV_GRADIENT;
        break;
      }
      const double t = 2.0 * relative_decrease - 1.0;
      radius = radius / std::fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = std::fmin(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
    } else {
      sum->num_unsuccessful_steps++;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
    }
  }
  sum->final_cost = cost;
  sum->final_radius = radius;
  sum->final_gradient_max_norm = gmax;
  sum->iterations = iteration;
  sum->termination = termination;
  return 0;
}

}  // namespace

extern "C" {

void orc_ba_default_options(lorb_ba_options* o) {
  o->max_num_iterations = 50;
  o->jacobi_scaling = 1;
  o->max_consecutive_invalid_steps = 5;
  o->reserved0 = 0;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
}

/* BA::ProjectPoseOptimization, src/bundle_adjust.cpp:158-202. */
int orc_ba_pose_only(int n, const float* xw, const float* uv, const float* K, double* rt,
                     const lorb_ba_options* opt, lorb_ba_summary* summary) {
  Problem pb;
  pb.C = 1;
  pb.P = 0;
  pb.cams.assign(rt, rt + 6);
  pb.const_xw.assign(xw, xw + 3 * (size_t)n);
  pb.fx = K[0];
  pb.fy = K[1];
  pb.cx = K[2];
  pb.cy = K[3];
  pb.res.resize(n);
  for (int i = 0; i < n; i++) {
    Residual& R = pb.res[i];
    R.kind = RES_POSE_ONLY;
    R.cam = 0;
    R.pt = i;
    R.u = uv[2 * i];
    R.v = uv[2 * i + 1];
  }
  int rc = solve_lm(pb, *opt, summary);
  for (int k = 0; k < 6; k++) rt[k] = pb.cams[k];
  return rc;
}

/* BA::LocalPoseOptimization, src/bundle_adjust.cpp:207-330. */
int orc_ba_local(int C, double* cams, int P, double* pts, int O, const int* obs_cam,
                 const int* obs_pt, const float* obs_uv, int F, const int* fix_pt,
                 const float* fix_uv, const float* fix_rt, const float* K,
                 const lorb_ba_options* opt, lorb_ba_summary* summary) {
  Problem pb;
  pb.C = C;
  pb.P = P;
  pb.cams.assign(cams, cams + 6 * (size_t)C);
  pb.pts.assign(pts, pts + 3 * (size_t)P);
  pb.fx = K[0];
  pb.fy = K[1];
  pb.cx = K[2];
  pb.cy = K[3];
  pb.res.resize((size_t)O + F);
  for (int i = 0; i < O; i++) {
    Residual& R = pb.res[i];
    R.kind = RES_POSE_MP;
    R.cam = obs_cam[i];
    R.pt = obs_pt[i];
    R.u = obs_uv[2 * i];
    R.v = obs_uv[2 * i + 1];
    if (R.cam < 0 || R.cam >= C || R.pt < 0 || R.pt >= P) return -1;
  }
  for (int i = 0; i < F; i++) {
    Residual& R = pb.res[(size_t)O + i];
    R.kind = RES_MP_FIXED;
    R.cam = -1;
    R.pt = fix_pt[i];
    R.u = fix_uv[2 * i];
    R.v = fix_uv[2 * i + 1];
    for (int k = 0; k < 6; k++) R.frt[k] = fix_rt[6 * i + k];
    if (R.pt < 0 || R.pt >= P) return -1;
  }
  int rc = solve_lm(pb, *opt, summary);
  memcpy(cams, pb.cams.data(), sizeof(double) * 6 * (size_t)C);
  memcpy(pts, pb.pts.data(), sizeof(double) * 3 * (size_t)P);
  return rc;
}

/* Residuals + Jet Jacobians of one PoseMPCost block, for Jacobian tests:
 * J_cam[2x6], J_pt[2x3] row-major.  kind: 0 PoseMP, 2 PoseOnly (fx quirk). */
void orc_ba_residual_jac(int kind, const double* cam, const double* pt, const float* uv,
                         const float* K, double* r, double* Jc, double* Jp) {
  Problem pb;
  pb.C = 1;
  pb.P = 1;
  pb.cams.assign(cam, cam + 6);
  pb.pts.assign(pt, pt + 3);
  pb.const_xw = {(float)pt[0], (float)pt[1], (float)pt[2]};
  pb.fx = K[0];
  pb.fy = K[1];
  pb.cx = K[2];
  pb.cy = K[3];
  Residual R;
  R.kind = kind;
  R.cam = 0;
  R.pt = 0;
  R.u = uv[0];
  R.v = uv[1];
  pb.res.push_back(R);
  eval_residual(pb, pb.cams.data(), pb.pts.data(), 0, r, Jc, Jp);
}

/* Cost only (1/2 sum r^2) of a local-BA state, used by the scipy cross-check. */
double orc_ba_local_cost(int C, const double* cams, int P, const double* pts, int O,
                         const int* obs_cam, const int* obs_pt, const float* obs_uv, int F,
                         const int* fix_pt, const float* fix_uv, const float* fix_rt,
                         const float* K) {
  lorb_ba_options o;
  orc_ba_default_options(&o);
  o.max_num_iterations = 0;
  lorb_ba_summary s;
  std::vector<double> c(cams, cams + 6 * (size_t)C), p(pts, pts + 3 * (size_t)P);
  orc_ba_local(C, c.data(), P, p.data(), O, obs_cam, obs_pt, obs_uv, F, fix_pt, fix_uv, fix_rt, K,
               &o, &s);
  return s.initial_cost;
}

/* OpenMP loop over independent windows (CPU baseline of BASELINE config 4). */
int orc_ba_local_batched(int n_windows, const int* cam_off, double* cams, const int* pt_off,
                         double* pts, const int* obs_off, const int* obs_cam, const int* obs_pt,
                         const float* obs_uv, const float* K, const lorb_ba_options* opt,
                         lorb_ba_summary* summaries) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int w = 0; w < n_windows; w++) {
    orc_ba_local(cam_off[w + 1] - cam_off[w], cams + 6 * (size_t)cam_off[w],
                 pt_off[w + 1] - pt_off[w], pts + 3 * (size_t)pt_off[w],
                 obs_off[w + 1] - obs_off[w], obs_cam + obs_off[w], obs_pt + obs_off[w],
                 obs_uv + 2 * (size_t)obs_off[w], 0, nullptr, nullptr, nullptr, K, opt,
                 &summaries[w]);
  }
  return 0;
}
}
