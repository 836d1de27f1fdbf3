// Stand-in for the slice of Ceres Solver's public API that the reference's
// src/bundle_adjust.cpp uses (AutoDiffCostFunction, Problem::AddResidualBlock, Solver::Options
// with DENSE_SCHUR, Solve, AngleAxisRotatePoint), so that file compiles UNMODIFIED into
// oracle/_ref/libref.so in an image without Ceres (reference CMakeLists.txt:19 names no
// version and vendors nothing).
//
// TEST INFRASTRUCTURE ONLY.  It pins what the reference's own code contributes -- the three
// cost functors (incl. PoseCost's fx-for-v), the window / point / observation assembly, the
// parameter layout, the float read-in and write-back -- and it is a second, independently
// written statement of Ceres' published trust-region Levenberg-Marquardt loop:
// generic residual blocks, forward-mode Jets, DENSE normal equations factorised by Cholesky
// (mathematically the same step as DENSE_SCHUR; different arithmetic order), against which the
// structured oracle oracle/ba_ref.cpp and the CUDA solver are compared.  It is NOT Ceres: the LM
// trajectory stays "unpinned against the real library" (DESIGN.md section 3).
#ifndef LORB_ORACLE_CERESSHIM_HPP
#define LORB_ORACLE_CERESSHIM_HPP
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <limits>
#include <map>
#include <vector>

namespace ceres {

// ------------------------------------------------------------------- Jet
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() { for (int i = 0; i < N; i++) v[i] = T(); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; i++) v[i] = T(); }  // NOLINT
  template <typename S>
  explicit Jet(const S& s) : a(T(s)) { for (int i = 0; i < N; i++) v[i] = T(); }
  Jet& operator+=(const Jet& g) { a += g.a; for (int i = 0; i < N; i++) v[i] += g.v[i]; return *this; }
  Jet& operator-=(const Jet& g) { a -= g.a; for (int i = 0; i < N; i++) v[i] -= g.v[i]; return *this; }
  Jet& operator*=(const Jet& g) { *this = *this * g; return *this; }
  Jet& operator/=(const Jet& g) { *this = *this / g; return *this; }
};
template <typename T, int N>
inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h;
}
template <typename T, int N>
inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h;
}
template <typename T, int N>
inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // Ceres: g_a_inverse = 1/g.a; f_a_by_g_a = f.a * g_a_inverse; v = (f.v - f_a_by_g_a * g.v) * g_a_inverse
  Jet<T, N> h;
  const T gi = T(1.0) / g.a;
  const T fg = f.a * gi;
  h.a = fg;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a += s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a += s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a -= s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h = -f; h.a += s; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) {
  Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; i++) h.v[i] = f.v[i] * s; return h;
}
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, T s) { return f * (T(1.0) / s); }
template <typename T, int N> inline Jet<T, N> operator/(T s, const Jet<T, N>& f) { return Jet<T, N>(s) / f; }
template <typename T, int N> inline bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
template <typename T, int N> inline bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
template <typename T, int N> inline bool operator>=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a >= g.a; }
template <typename T, int N> inline bool operator<=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a <= g.a; }
template <typename T, int N>
inline Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::sqrt(f.a); const T d = T(1.0) / (T(2.0) * h.a);
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h;
}
template <typename T, int N>
inline Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::cos(f.a); const T d = -std::sin(f.a);
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h;
}
template <typename T, int N>
inline Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::sin(f.a); const T d = std::cos(f.a);
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h;
}
using std::cos;
using std::sin;
using std::sqrt;

// ceres/rotation.h: AngleAxisRotatePoint (Rodrigues' formula away from zero, first-order
// Taylor `pt + w x pt` when theta^2 <= epsilon)
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  const T theta2 = angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// --------------------------------------------------------- cost functions
class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int> sizes_;
  int num_residuals_ = 0;
};
class LossFunction;

// Jacobians are row-major [num_residuals x block size] per parameter block, as in Ceres.
template <typename Functor, int kNumResiduals, int N0, int N1 = 0, int N2 = 0>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) {
    num_residuals_ = kNumResiduals;
    sizes_.push_back(N0);
    if (N1) sizes_.push_back(N1);
    if (N2) sizes_.push_back(N2);
  }
  ~AutoDiffCostFunction() override { delete f_; }
  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    if (!jacobians) return call(*f_, p, residuals);
    enum { N = N0 + N1 + N2 };
    typedef Jet<double, N> J;
    J x[N];
    int k = 0;
    const int sz[3] = {N0, N1, N2};
    for (int b = 0; b < 3; b++)
      for (int i = 0; i < sz[b]; i++, k++) {
        x[k] = J(p[b][i]);
        x[k].v[k] = 1.0;
      }
    const J* px[3] = {x, x + N0, x + N0 + N1};
    J r[kNumResiduals];
    if (!call(*f_, px, r)) return false;
    for (int i = 0; i < kNumResiduals; i++) residuals[i] = r[i].a;
    int off = 0;
    for (int b = 0; b < 3; b++) {
      if (sz[b] && jacobians[b])
        for (int i = 0; i < kNumResiduals; i++)
          for (int c = 0; c < sz[b]; c++) jacobians[b][i * sz[b] + c] = r[i].v[off + c];
      off += sz[b];
    }
    return true;
  }

 private:
  template <typename T>
  static bool call(const Functor& f, T const* const* p, T* r) {
    return call_n(f, p, r, std::integral_constant<int, (N1 ? 1 : 0) + (N2 ? 1 : 0)>());
  }
  template <typename T>
  static bool call_n(const Functor& f, T const* const* p, T* r, std::integral_constant<int, 0>) { return f(p[0], r); }
  template <typename T>
  static bool call_n(const Functor& f, T const* const* p, T* r, std::integral_constant<int, 1>) { return f(p[0], p[1], r); }
  template <typename T>
  static bool call_n(const Functor& f, T const* const* p, T* r, std::integral_constant<int, 2>) { return f(p[0], p[1], p[2], r); }
  Functor* f_;
};

// ---------------------------------------------------------------- problem
class Problem {
 public:
  struct Block { const CostFunction* cost; std::vector<double*> params; };
  ~Problem() { for (auto& b : blocks_) delete b.cost; }
  void AddResidualBlock(CostFunction* c, LossFunction*, double* x0) { add(c, {x0}); }
  void AddResidualBlock(CostFunction* c, LossFunction*, double* x0, double* x1) { add(c, {x0, x1}); }
  void AddResidualBlock(CostFunction* c, LossFunction*, double* x0, double* x1, double* x2) { add(c, {x0, x1, x2}); }
  // parameter blocks in order of first appearance
  std::vector<Block> blocks_;
  std::vector<double*> param_ptr_;
  std::vector<int> param_size_, param_off_;
  std::map<double*, int> index_;
  int num_params_ = 0, num_res_ = 0;

 private:
  void add(CostFunction* c, std::vector<double*> ps) {
    for (size_t i = 0; i < ps.size(); i++)
      if (!index_.count(ps[i])) {
        index_[ps[i]] = (int)param_ptr_.size();
        param_ptr_.push_back(ps[i]);
        param_size_.push_back(c->parameter_block_sizes()[i]);
        param_off_.push_back(num_params_);
        num_params_ += c->parameter_block_sizes()[i];
      }
    num_res_ += c->num_residuals();
    blocks_.push_back(Block{c, ps});
  }
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

class Solver {
 public:
  struct Options {  // Ceres' documented defaults
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    bool jacobi_scaling = true;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;
  };
  struct Summary {
    double initial_cost = 0, final_cost = 0, final_radius = 0, final_gradient_max_norm = 0;
    int iterations = 0, num_successful_steps = 0, num_unsuccessful_steps = 0;
    int lorb_termination = 0;  // LORB_BA_* code of include/lorb_cuda.h, for the comparison harness
    TerminationType termination_type = NO_CONVERGENCE;
  };
};

// What the last Solve() saw, for the comparison harness (oracle/ref_harness.cpp): the doubles the
// reference keeps in local arrays are otherwise lost when it rounds them to float.
struct LorbLastSolve {
  std::vector<double*> ptr;
  std::vector<int> size;
  std::vector<double> value;  // concatenated, in `ptr` order
  Solver::Summary summary;
  bool have_override = false;
  Solver::Options override_options;
};
inline LorbLastSolve& lorb_last_solve() {
  static thread_local LorbLastSolve s;
  return s;
}

namespace shim_detail {
struct Lin {
  std::vector<double> r, J;  // J dense row-major [num_res x num_params]
};
inline double evaluate(const Problem& pb, const std::vector<double>& x, Lin* L) {
  const int n = pb.num_params_, m = pb.num_res_;
  std::vector<double> r(m, 0.0);
  if (L) L->J.assign((size_t)m * n, 0.0);
  int row = 0;
  double jbuf[3][64];
  for (const auto& b : pb.blocks_) {
    const double* p[3] = {nullptr, nullptr, nullptr};
    double* jac[3] = {nullptr, nullptr, nullptr};
    int off[3] = {0, 0, 0}, sz[3] = {0, 0, 0};
    for (size_t i = 0; i < b.params.size(); i++) {
      const int id = pb.index_.at(b.params[i]);
      off[i] = pb.param_off_[id];
      sz[i] = pb.param_size_[id];
      p[i] = &x[off[i]];
      jac[i] = jbuf[i];
    }
    const int nr = b.cost->num_residuals();
    b.cost->Evaluate(p, &r[row], L ? jac : nullptr);
    if (L)
      for (size_t i = 0; i < b.params.size(); i++)
        for (int a = 0; a < nr; a++)
          for (int c = 0; c < sz[i]; c++) L->J[(size_t)(row + a) * n + off[i] + c] += jac[i][a * sz[i] + c];
    row += nr;
  }
  double s = 0;
  for (double v : r) s += v * v;
  if (L) L->r = r;
  return 0.5 * s;
}
inline bool cholesky(std::vector<double>& A, int n) {  // lower, in place
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
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
inline void chol_solve(const std::vector<double>& L, int n, std::vector<double>& b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}
}  // namespace shim_detail

// TrustRegionMinimizer + LevenbergMarquardtStrategy as published (Ceres 1.x / 2.0 loop order):
// Jacobi scaling from the initial Jacobian, D^2 = clamp(diag(J'J), min, max) / radius on the scaled
// Jacobian, step from (J'J + D^2) y = J'r, delta = -y .* scale, model decrease -m.(r + m/2) with
// m = -J y, parameter then function tolerance on the candidate BEFORE the acceptance test, step
// quality rho, radius update radius / max(1/3, 1 - (2 rho - 1)^3) or halving with doubling factor.
inline void Solve(const Solver::Options& options_in, Problem* problem, Solver::Summary* summary) {
  using namespace shim_detail;
  LorbLastSolve& last = lorb_last_solve();
  const Solver::Options opt = last.have_override ? last.override_options : options_in;
  Problem& pb = *problem;
  const int n = pb.num_params_, m = pb.num_res_;
  std::vector<double> x(n);
  for (size_t b = 0; b < pb.param_ptr_.size(); b++)
    for (int i = 0; i < pb.param_size_[b]; i++) x[pb.param_off_[b] + i] = pb.param_ptr_[b][i];

  enum { T_NO = 0, T_FUNCTION = 1, T_GRADIENT = 2, T_PARAMETER = 3, T_RADIUS = 4, T_FAILURE = 5 };  // = LORB_BA_*
  Lin L;
  double cost = evaluate(pb, x, &L);
  Solver::Summary S;
  S.initial_cost = cost;
  std::vector<double> scale(n, 1.0), g(n);
  auto gradient_max = [&]() {
    double mx = 0;
    for (int c = 0; c < n; c++) {
      double s = 0;
      for (int r = 0; r < m; r++) s += L.J[(size_t)r * n + c] * L.r[r];
      g[c] = s;
      mx = std::fmax(mx, std::fabs(s));
    }
    return mx;
  };
  auto scale_columns = [&]() {
    for (int r = 0; r < m; r++)
      for (int c = 0; c < n; c++) L.J[(size_t)r * n + c] *= scale[c];
  };
  auto x_norm_of = [&]() { double s = 0; for (double v : x) s += v * v; return std::sqrt(s); };
  double gmax = gradient_max();  // on the unscaled Jacobian
  if (opt.jacobi_scaling) {
    for (int c = 0; c < n; c++) {
      double s = 0;
      for (int r = 0; r < m; r++) s += L.J[(size_t)r * n + c] * L.J[(size_t)r * n + c];
      scale[c] = 1.0 / (1.0 + std::sqrt(s));
    }
    scale_columns();
  }
  double x_norm = x_norm_of();
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  int iteration = 0, invalid_run = 0, term = T_NO;
  if (gmax <= opt.gradient_tolerance) term = T_GRADIENT;
  std::vector<double> H((size_t)n * n), y(n), cand(n);
  while (term == T_NO) {
    if (iteration >= opt.max_num_iterations) break;
    if (radius < opt.min_trust_region_radius) { term = T_RADIUS; break; }
    iteration++;
    // normal equations of the scaled Jacobian, damped diagonal
    for (int a = 0; a < n; a++)
      for (int b = a; b < n; b++) H[(size_t)a * n + b] = 0.0;
    for (int r = 0; r < m; r++) {
      const double* Jr = &L.J[(size_t)r * n];
      for (int a = 0; a < n; a++) {
        if (Jr[a] == 0.0) continue;
        for (int b = a; b < n; b++) H[(size_t)a * n + b] += Jr[a] * Jr[b];
      }
    }
    for (int a = 0; a < n; a++) {
      for (int b = 0; b < a; b++) H[(size_t)a * n + b] = H[(size_t)b * n + a];
      const double d = std::fmin(std::fmax(H[(size_t)a * n + a], opt.min_lm_diagonal), opt.max_lm_diagonal) / radius;
      H[(size_t)a * n + a] += d;
    }
    for (int c = 0; c < n; c++) {
      double s = 0;
      for (int r = 0; r < m; r++) s += L.J[(size_t)r * n + c] * L.r[r];
      y[c] = s;
    }
    bool step_valid = cholesky(H, n);
    double model_cost_change = 0;
    if (step_valid) {
      chol_solve(H, n, y);
      for (int c = 0; c < n; c++)
        if (!std::isfinite(y[c])) step_valid = false;
    }
    if (step_valid) {
      double acc = 0;
      for (int r = 0; r < m; r++) {
        double mm = 0;
        for (int c = 0; c < n; c++) mm -= L.J[(size_t)r * n + c] * y[c];
        acc += mm * (L.r[r] + mm / 2.0);
      }
      model_cost_change = -acc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      invalid_run++;
      S.num_unsuccessful_steps++;
      if (invalid_run >= opt.max_num_consecutive_invalid_steps) { term = T_FAILURE; break; }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    invalid_run = 0;
    double step_sq = 0;
    for (int c = 0; c < n; c++) {
      const double d = -y[c] * scale[c];
      cand[c] = x[c] + d;
      step_sq += d * d;
    }
    const double cand_cost = evaluate(pb, cand, nullptr);
    if (std::sqrt(step_sq) <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { term = T_PARAMETER; break; }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * cost) { term = T_FUNCTION; break; }
    const double rho = cost_change / model_cost_change;
    if (rho > opt.min_relative_decrease) {
      x = cand;
      x_norm = x_norm_of();
      cost = evaluate(pb, x, &L);
      gmax = gradient_max();
      if (opt.jacobi_scaling) scale_columns();
      S.num_successful_steps++;
      if (gmax <= opt.gradient_tolerance) { term = T_GRADIENT; break; }
      const double t = 2.0 * rho - 1.0;
      radius = std::fmin(opt.max_trust_region_radius, radius / std::fmax(1.0 / 3.0, 1.0 - t * t * t));
      decrease_factor = 2.0;
    } else {
      S.num_unsuccessful_steps++;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
    }
  }
  S.final_cost = cost;
  S.final_radius = radius;
  S.final_gradient_max_norm = gmax;
  S.iterations = iteration;
  S.lorb_termination = term;
  S.termination_type = term == T_NO ? NO_CONVERGENCE : (term == T_FAILURE ? FAILURE : CONVERGENCE);
  for (size_t b = 0; b < pb.param_ptr_.size(); b++)
    for (int i = 0; i < pb.param_size_[b]; i++) pb.param_ptr_[b][i] = x[pb.param_off_[b] + i];
  last.ptr = pb.param_ptr_;
  last.size = pb.param_size_;
  last.value = x;
  last.summary = S;
  if (summary) *summary = S;
}

}  // namespace ceres
#endif
