// ba_local.cu — local bundle adjustment on sm_100a (general multi-CTA path).
//
// Replaces the Ceres solve inside BA::LocalPoseOptimization (reference
// src/bundle_adjust.cpp:207-330): PoseMPCost / MPCost residuals, trust-region
// Levenberg-Marquardt with Jacobi scaling, DENSE_SCHUR.
//
// Data layout in HBM: observations sorted by point (CSR
// pt_ptr), 12 B each (camera index, u, v); parameters fp64 SoA-by-block
// (cams[C][6], pts[P][3]) double-buffered (current / candidate); per camera a
// 36-double rotation block (R, dR/dw) rebuilt per linearisation.
// Off-diagonal blocks W = Jc^T Jp are never stored: the Schur products are
// formed from recomputed Jacobians (the 48 B/obs/iteration design of
// SURVEY §8(d)).
//
// One LM attempt = build (Jacobians, H_pp, g_p, H_cc, g_c, Schur products)
//                -> [allreduce when sharded] -> finish (damping, mirror, rhs)
//                -> dense Cholesky + solve -> candidate cameras
//                -> back-substitution + model decrease + candidate cost
//                -> [allreduce] -> control (accept / reject, radius update).
// The trust-region state lives on the device; the host only enqueues.
#include <algorithm>
#include <new>
#include <vector>

#include "ba_math.cuh"
#include "dist.cuh"

namespace lorb {

constexpr int BA_THREADS = 256;
constexpr int HCC = 21;  // upper triangle of a 6x6 block

struct BADev {
  int C, P, n;  // n = 6C
  double* cams[2];
  double* pts[2];
  double* camrot[2];     // C x 36
  const int* pt_ptr;     // P+1
  const int* obs_cam;    // >=0 window camera; <0: fixed observation -1-f
  const float2* obs_uv;
  const double* fixrt;   // F x 12: R (9) + t (3) of out-of-window observers
  Intr K;
  double* scale_c;       // 6C
  double* scale_p;       // 3P
  double* lin;           // [S n*n | Hcc 21C | gc 6C | rhs_corr 6C | tail 4]
  double* S;
  double* Hcc;
  double* gc;
  double* rhs_corr;
  double* tail;          // acc_cur2, acc_xcur2, bad, spare
  double* rhs;           // n (becomes y_c after the solve)
  double* pt_hinv;       // P x 6
  double* pt_gp;         // P x 3
  LMState* st;
};

__device__ __forceinline__ int upper_idx(int a, int b) {  // a <= b, 6x6
  return a * 6 - (a * (a - 1)) / 2 + (b - a);
}

__global__ void ba_camrot_kernel(const BADev* __restrict__ probs, int which /*0 = current, 1 = candidate*/, int force) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done && !force) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const int buf = st->cur ^ which;
  cam_rotation(p.cams[buf] + 6 * c, p.camrot[buf] + CAMROT * c, true);
}

// Evaluate observation `o` of point X with scaled Jacobians.
__device__ __forceinline__ int eval_obs(const BADev& p, const double* __restrict__ cams,
                                        const double* __restrict__ camrot, int o, const double X[3],
                                        const double sp[3], double r[2], double Jc[12],
                                        double Jp[6]) {
  const int cam = p.obs_cam[o];
  const float2 uv = p.obs_uv[o];
  if (cam >= 0) {
    const double* t = cams + 6 * cam + 3;
    const double tt[3] = {t[0], t[1], t[2]};
    obs_eval<true, true>(camrot + CAMROT * cam, tt, X, p.K, (double)uv.x, (double)uv.y, r, Jc, Jp);
    const double* sc = p.scale_c + 6 * cam;
#pragma unroll
    for (int a = 0; a < 6; a++) {
      Jc[a] *= sc[a];
      Jc[6 + a] *= sc[a];
    }
  } else {
    const double* f = p.fixrt + 12 * (size_t)(-1 - cam);
    const double tt[3] = {f[9], f[10], f[11]};
    obs_eval<false, true>(f, tt, X, p.K, (double)uv.x, (double)uv.y, r, Jc, Jp);
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    Jp[a] *= sp[a];
    Jp[3 + a] *= sp[a];
  }
  return cam;
}

__device__ __forceinline__ double group_sum8(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 2, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 1, 8);
  return v;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
  return t;  // valid on thread 0
}

__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t = fmax(t, red[w]);
  return t;
}

// Build pass.  Eight lanes per point, one observation per lane per round.
// FULL = false: initial pass (scales are 1): cost, H_cc / g_c, point column
// norms -> scale_p, |x|^2, gradient max.  FULL = true: one LM attempt.
// SMALL = true: windows with few cameras (n = 6C <= 96).  All blocks of such a
// window hit the same few hundred H_cc / S addresses, and same-address fp64
// atomics serialise in L2 (profiles/ba_launches_r01.md), so every CTA first
// accumulates into a private copy of `lin` in shared memory and flushes it once.
template <bool FULL, bool SMALL>
__global__ void __launch_bounds__(BA_THREADS)
    ba_build_kernel(const BADev* __restrict__ probs, lorb_ba_options opt, int force) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done && !force) return;
  __shared__ double Wsm[BA_THREADS][18];
  __shared__ int Csm[BA_THREADS];
  __shared__ double red[BA_THREADS / 32];
  extern __shared__ double slin[];  // SMALL: [S n*n | Hcc 21C | gc 6C | rhs_corr 6C]
  const int lin_n = p.n * p.n + (HCC + 12) * p.C;
  if (SMALL) {
    for (int i = threadIdx.x; i < lin_n; i += BA_THREADS) slin[i] = 0.0;
    __syncthreads();
  }
  double* const accS = SMALL ? slin : p.S;
  double* const accH = SMALL ? slin + (size_t)p.n * p.n : p.Hcc;
  double* const accG = SMALL ? slin + (size_t)p.n * p.n + HCC * p.C : p.gc;
  double* const accR = SMALL ? slin + (size_t)p.n * p.n + (HCC + 6) * p.C : p.rhs_corr;
  const int cur = st->cur;
  const double* cams = p.cams[cur];
  const double* pts = p.pts[cur];
  const double* camrot = p.camrot[cur];
  const double radius = st->radius;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gl = lane & 7, gw = lane >> 3;
  const int n = p.n;
  double cost_acc = 0, gmax_acc = 0, xn_acc = 0, bad_acc = 0;
  for (int base = (blockIdx.x * (BA_THREADS / 32) + warp) * 4; base < p.P;
       base += gridDim.x * (BA_THREADS / 8)) {
    const int pt = base + gw;
    const bool pv = pt < p.P;
    int s = 0, e = 0;
    double X[3] = {0, 0, 0}, sp[3] = {1, 1, 1};
    if (pv) {
      s = p.pt_ptr[pt];
      e = p.pt_ptr[pt + 1];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        X[a] = pts[3 * (size_t)pt + a];
        sp[a] = p.scale_p[3 * (size_t)pt + a];
      }
    }
    const int rounds = __reduce_max_sync(0xffffffffu, (e - s + 7) >> 3);
    // ---- phase 1: H_pp, g_p, H_cc, g_c, cost
    double h[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    double r[2], Jc[12], Jp[6];
    int cam = -1;
    bool has = false;
    for (int rd = 0; rd < rounds; rd++) {
      const int o = s + rd * 8 + gl;
      has = o < e;
      if (has) {
        cam = eval_obs(p, cams, camrot, o, X, sp, r, Jc, Jp);
        cost_acc += r[0] * r[0] + r[1] * r[1];
        h[0] += Jp[0] * Jp[0] + Jp[3] * Jp[3];
        h[1] += Jp[0] * Jp[1] + Jp[3] * Jp[4];
        h[2] += Jp[0] * Jp[2] + Jp[3] * Jp[5];
        h[3] += Jp[1] * Jp[1] + Jp[4] * Jp[4];
        h[4] += Jp[1] * Jp[2] + Jp[4] * Jp[5];
        h[5] += Jp[2] * Jp[2] + Jp[5] * Jp[5];
#pragma unroll
        for (int a = 0; a < 3; a++) g[a] += Jp[a] * r[0] + Jp[3 + a] * r[1];
        if (cam >= 0) {
          double* H = accH + HCC * (size_t)cam;
          int k = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = a; b < 6; b++) atomicAdd(&H[k++], Jc[a] * Jc[b] + Jc[6 + a] * Jc[6 + b]);
            atomicAdd(&accG[6 * cam + a], Jc[a] * r[0] + Jc[6 + a] * r[1]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) h[a] = group_sum8(h[a]);
#pragma unroll
    for (int a = 0; a < 3; a++) g[a] = group_sum8(g[a]);
    if (!FULL) {
      if (pv && gl == 0) {
        if (opt.jacobi_scaling && force == 1) {
          p.scale_p[3 * (size_t)pt + 0] = 1.0 / (1.0 + sqrt(h[0]));
          p.scale_p[3 * (size_t)pt + 1] = 1.0 / (1.0 + sqrt(h[3]));
          p.scale_p[3 * (size_t)pt + 2] = 1.0 / (1.0 + sqrt(h[5]));
        }
        xn_acc += X[0] * X[0] + X[1] * X[1] + X[2] * X[2];
        // gradient of the unscaled problem (this pass may run with scales already set)
        gmax_acc = fmax(gmax_acc, fmax(fabs(g[0] / sp[0]), fmax(fabs(g[1] / sp[1]), fabs(g[2] / sp[2]))));
      }
      continue;
    }
    if (pv && gl == 0)
      gmax_acc = fmax(gmax_acc, fmax(fabs(g[0] / sp[0]), fmax(fabs(g[1] / sp[1]), fabs(g[2] / sp[2]))));
    // LM diagonal on the scaled columns, clamped (LevenbergMarquardtStrategy)
    double hd[6] = {h[0], h[1], h[2], h[3], h[4], h[5]}, hi[6];
    hd[0] += clamp_diag(h[0], opt.min_lm_diagonal, opt.max_lm_diagonal) / radius;
    hd[3] += clamp_diag(h[3], opt.min_lm_diagonal, opt.max_lm_diagonal) / radius;
    hd[5] += clamp_diag(h[5], opt.min_lm_diagonal, opt.max_lm_diagonal) / radius;
    const bool ok = inv3_spd(hd, hi);
    if (pv && gl == 0) {
      if (!ok) bad_acc += 1.0;
#pragma unroll
      for (int a = 0; a < 6; a++) p.pt_hinv[6 * (size_t)pt + a] = ok ? hi[a] : 0.0;
#pragma unroll
      for (int a = 0; a < 3; a++) p.pt_gp[3 * (size_t)pt + a] = g[a];
    }
    if (!ok) {
#pragma unroll
      for (int a = 0; a < 6; a++) hi[a] = 0.0;
    }
    // ---- phase 2: Schur products  S -= W_i Hinv W_j^T ,  rhs_corr += W_i Hinv g_p
    for (int ri = 0; ri < rounds; ri++) {
      const int oi = s + ri * 8 + gl;
      const bool has_i = oi < e;
      int ci = -1;
      double Wi[18], Yi[18];
      if (rounds > 1 || ri > 0) {
        if (has_i) ci = eval_obs(p, cams, camrot, oi, X, sp, r, Jc, Jp);
      } else {
        ci = has ? cam : -1;  // single round: Jacobians of phase 1 are still live
      }
      const bool act_i = has_i && ci >= 0;
      if (act_i) {
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) Wi[3 * a + b] = Jc[a] * Jp[b] + Jc[6 + a] * Jp[3 + b];
#pragma unroll
        for (int a = 0; a < 6; a++) {
          const double w0 = Wi[3 * a], w1 = Wi[3 * a + 1], w2 = Wi[3 * a + 2];
          Yi[3 * a + 0] = w0 * hi[0] + w1 * hi[1] + w2 * hi[2];
          Yi[3 * a + 1] = w0 * hi[1] + w1 * hi[3] + w2 * hi[4];
          Yi[3 * a + 2] = w0 * hi[2] + w1 * hi[4] + w2 * hi[5];
          atomicAdd(&accR[6 * ci + a],
                    Yi[3 * a] * g[0] + Yi[3 * a + 1] * g[1] + Yi[3 * a + 2] * g[2]);
        }
      }
      for (int rj = 0; rj < rounds; rj++) {
        if (rj == ri) {
          Csm[tid] = act_i ? ci : -1;
          if (act_i) {
#pragma unroll
            for (int a = 0; a < 18; a++) Wsm[tid][a] = Wi[a];
          }
        } else {
          const int oj = s + rj * 8 + gl;
          int cj = -1;
          if (oj < e) {
            double r2[2], Jc2[12], Jp2[6];
            cj = eval_obs(p, cams, camrot, oj, X, sp, r2, Jc2, Jp2);
            if (cj >= 0) {
#pragma unroll
              for (int a = 0; a < 6; a++)
#pragma unroll
                for (int b = 0; b < 3; b++)
                  Wsm[tid][3 * a + b] = Jc2[a] * Jp2[b] + Jc2[6 + a] * Jp2[3 + b];
            }
          }
          Csm[tid] = cj;
        }
        __syncwarp();
        for (int j = 0; j < 8; j++) {
          const int src = (tid & ~7) + j;
          const int cj = Csm[src];
          // upper block triangle only (finish mirrors); equal cameras keep both orders
          if (act_i && cj >= ci) {
            const double* Wj = Wsm[src];
            double* Sblk = accS + (size_t)(6 * ci) * n + 6 * cj;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
              for (int b = 0; b < 6; b++) {
                const double v = Yi[3 * a] * Wj[3 * b] + Yi[3 * a + 1] * Wj[3 * b + 1] +
                                 Yi[3 * a + 2] * Wj[3 * b + 2];
                atomicAdd(&Sblk[(size_t)a * n + b], -v);
              }
          }
        }
        __syncwarp();
      }
    }
  }
  if (SMALL) {
    __syncthreads();
    for (int i = tid; i < lin_n; i += BA_THREADS) {
      const double v = slin[i];
      if (v != 0.0) atomicAdd(&p.lin[i], v);
    }
  }
  double t = block_sum(cost_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&p.tail[0], t);
  t = block_sum(xn_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&p.tail[1], t);
  t = block_sum(bad_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&p.tail[2], t);
  t = block_max(gmax_acc, red);
  if (tid == 0 && t > 0.0) atomic_max_nonneg(&st->acc_gmax, t);
}

// After the initial pass: camera column scales, |x|, cost, gradient test.
__global__ void ba_init_finish_kernel(const BADev* __restrict__ probs, lorb_ba_options opt) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  __shared__ double red[32];
  double xn = 0, gm = 0;
  for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
    const int c = i / 6, a = i % 6;
    const double d = p.Hcc[HCC * (size_t)c + upper_idx(a, a)];
    gm = fmax(gm, fabs(p.gc[i]));
    if (opt.jacobi_scaling) p.scale_c[i] = 1.0 / (1.0 + sqrt(d));
    const double x = p.cams[st->cur][i];
    xn += x * x;
  }
  const double xs = block_sum(xn, red);
  const double gmx = block_max(gm, red);
  if (threadIdx.x == 0) {
    st->x_norm = sqrt(xs + p.tail[1]);
    st->cost = 0.5 * p.tail[0];
    st->initial_cost = st->cost;
    st->gmax = fmax(gmx, st->acc_gmax);
    st->radius = opt.initial_trust_region_radius;
    st->decrease_factor = 2.0;
    st->iteration = st->n_success = st->n_fail = st->invalid_run = 0;
    st->termination = LORB_BA_NO_CONVERGENCE;
    st->done = 0;
    st->check_gradient = 0;
    st->solve_ok = 1;
    if (st->gmax <= opt.gradient_tolerance) {
      st->termination = LORB_BA_CONV_GRADIENT;
      st->done = 1;
    }
    if (opt.max_num_iterations <= 0) st->done = 1;
    lm_zero_acc(st);
  }
}

// Gradient-tolerance test of the point accepted by the previous attempt
// (Ceres checks it right after HandleSuccessfulStep; here the gradient only
// exists after the next build pass).  One CTA.
__global__ void ba_gradcheck_kernel(const BADev* __restrict__ probs, lorb_ba_options opt, int force) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done && !force) return;
  __shared__ double red[32];
  double gm = 0;
  for (int i = threadIdx.x; i < p.n; i += blockDim.x) gm = fmax(gm, fabs(p.gc[i] / p.scale_c[i]));
  const double gmx = block_max(gm, red);
  if (threadIdx.x == 0 && st->check_gradient) {
    st->gmax = fmax(gmx, st->acc_gmax);
    st->check_gradient = 0;
    if (st->gmax <= opt.gradient_tolerance) {
      st->termination = LORB_BA_CONV_GRADIENT;
      st->done = 1;
    }
  }
}

// S <- blockdiag(H_cc + D_c^2) + (Schur part, upper blocks mirrored); rhs = g_c - corr.
__global__ void ba_finish_kernel(const BADev* __restrict__ probs, lorb_ba_options opt) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  const int n = p.n;
  const double radius = st->radius;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)n * n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    const int ci = i / 6, cj = j / 6;
    if (ci == cj) {
      const int a = i % 6, b = j % 6;
      double v = p.Hcc[HCC * (size_t)ci + (a <= b ? upper_idx(a, b) : upper_idx(b, a))];
      if (a == b) v += clamp_diag(v, opt.min_lm_diagonal, opt.max_lm_diagonal) / radius;
      p.S[idx] += v;
    } else if (ci > cj) {
      p.S[idx] = p.S[(size_t)j * n + i];  // mirror the upper block triangle
    }
    if (j == 0) p.rhs[i] = p.gc[i] - p.rhs_corr[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) st->solve_ok = (p.tail[2] == 0.0) ? 1 : 0;
}

// ---- dense Cholesky (lower, in place) + solve.  Small systems: one CTA, matrix in smem.
__global__ void __launch_bounds__(256) ba_chol_small_kernel(const BADev* __restrict__ probs) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  extern __shared__ double A[];  // n*n + n
  __shared__ int s_ok;
  const int n = p.n, tid = threadIdx.x;
  double* b = A + (size_t)n * n;
  for (int i = tid; i < n * n; i += blockDim.x) A[i] = p.S[i];
  for (int i = tid; i < n; i += blockDim.x) b[i] = p.rhs[i];
  if (tid == 0) s_ok = 1;
  __syncthreads();
  for (int j = 0; j < n; j++) {
    if (tid == 0) {
      const double d = A[j * n + j];
      if (!(d > 0.0)) s_ok = 0;
      A[j * n + j] = sqrt(d > 0.0 ? d : 1.0);
    }
    __syncthreads();
    const double dj = A[j * n + j];
    for (int i = j + 1 + tid; i < n; i += blockDim.x) A[i * n + j] /= dj;
    __syncthreads();
    // trailing update of the lower triangle
    const int m = n - j - 1;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) A[i * n + k] -= A[i * n + j] * A[k * n + j];
    }
    __syncthreads();
  }
  // forward / backward substitution by warp 0
  if (tid < 32) {
    for (int i = 0; i < n; i++) {
      double sacc = 0;
      for (int k = tid; k < i; k += 32) sacc += A[i * n + k] * b[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (tid == 0) b[i] = (b[i] - sacc) / A[i * n + i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; i--) {
      double sacc = 0;
      for (int k = i + 1 + tid; k < n; k += 32) sacc += A[k * n + i] * b[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (tid == 0) b[i] = (b[i] - sacc) / A[i * n + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) p.rhs[i] = b[i];
  if (tid == 0 && !s_ok) st->solve_ok = 0;
}

// Small windows: one CTA does everything between the build pass and the
// back-substitution: gradient-tolerance test, assembly of the damped reduced
// camera system in shared memory, Cholesky, both triangular solves, candidate
// cameras and their rotation blocks.  L is kept transposed in the (unused) upper
// triangle so a column step needs a single barrier.
__global__ void __launch_bounds__(256)
    ba_solve_small_kernel(const BADev* __restrict__ probs, lorb_ba_options opt) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  extern __shared__ double A[];  // n*n | b[n] | diag[n]
  __shared__ double red[8];
  __shared__ int s_ok, s_done;
  const int n = p.n, tid = threadIdx.x;
  double* b = A + (size_t)n * n;
  double* dg = b + n;
  // gradient tolerance of the point accepted by the previous attempt
  double gm = 0;
  for (int i = tid; i < n; i += blockDim.x) gm = fmax(gm, fabs(p.gc[i] / p.scale_c[i]));
  const double gmx = block_max(gm, red);
  if (tid == 0) {
    s_done = 0;
    s_ok = (p.tail[2] == 0.0) ? 1 : 0;
    if (st->check_gradient) {
      st->gmax = fmax(gmx, st->acc_gmax);
      st->check_gradient = 0;
      if (st->gmax <= opt.gradient_tolerance) {
        st->termination = LORB_BA_CONV_GRADIENT;
        st->done = 1;
        s_done = 1;
      }
    }
  }
  __syncthreads();
  if (s_done) return;
  const double radius = st->radius;
  for (int idx = tid; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx % n;
    const int ci = i / 6, cj = j / 6;
    double v;
    if (ci == cj) {
      const int a = i % 6, c = j % 6;
      const double h = p.Hcc[HCC * (size_t)ci + (a <= c ? upper_idx(a, c) : upper_idx(c, a))];
      v = p.S[idx] + h;
      if (a == c) v += clamp_diag(h, opt.min_lm_diagonal, opt.max_lm_diagonal) / radius;
    } else {
      v = ci < cj ? p.S[idx] : p.S[(size_t)j * n + i];
    }
    A[idx] = v;
  }
  for (int i = tid; i < n; i += blockDim.x) b[i] = p.gc[i] - p.rhs_corr[i];
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    if (!(d > 0.0)) {
      if (tid == 0) s_ok = 0;
      d = 1.0;
    }
    const double inv_d = 1.0 / d, inv_s = 1.0 / sqrt(d);
    if (tid == 0) dg[j] = sqrt(d);
    for (int i = j + 1 + tid; i < n; i += 256) A[j * n + i] = A[i * n + j] * inv_s;  // L_ij at [j][i]
    for (int i = j + 1 + ty; i < n; i += 16) {
      const double aij = A[i * n + j] * inv_d;
      for (int k = j + 1 + tx; k <= i; k += 16) A[i * n + k] -= aij * A[k * n + j];
    }
    __syncthreads();
  }
  if (tid < 32) {  // L y = b ; L^T x = y   (L_ik = A[k][i] for k < i)
    for (int i = 0; i < n; i++) {
      double sacc = 0;
      for (int k = tid; k < i; k += 32) sacc += A[k * n + i] * b[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (tid == 0) b[i] = (b[i] - sacc) / dg[i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; i--) {
      double sacc = 0;
      for (int k = i + 1 + tid; k < n; k += 32) sacc += A[i * n + k] * b[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (tid == 0) b[i] = (b[i] - sacc) / dg[i];
      __syncwarp();
    }
  }
  __syncthreads();
  // candidate cameras + rotation blocks
  const int cur = st->cur;
  double step2 = 0, xc2 = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    p.rhs[i] = b[i];
    const double dlt = -b[i] * p.scale_c[i];
    const double x = p.cams[cur][i] + dlt;
    p.cams[cur ^ 1][i] = x;
    step2 += dlt * dlt;
    xc2 += x * x;
  }
  const double s2 = block_sum(step2, red);
  const double x2 = block_sum(xc2, red);
  if (tid == 0) {
    p.tail[3] = s2;
    p.tail[4] = x2;
    st->solve_ok = s_ok;
  }
  __syncthreads();
  for (int c = tid; c < p.C; c += blockDim.x)
    cam_rotation(p.cams[cur ^ 1] + 6 * c, p.camrot[cur ^ 1] + CAMROT * c, true);
}

// Large systems: right-looking blocked Cholesky in global memory, NB = 32.
constexpr int NB = 32;

// Panel: every CTA factors the diagonal block in smem; CTA b>0 then solves its
// row block  L_ik = A_ik L_kk^-T.
__global__ void __launch_bounds__(NB * NB / 4)
    ba_chol_panel_kernel(const BADev* __restrict__ probs, int kb) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  __shared__ double D[NB][NB + 1];
  __shared__ double Ablk[NB][NB + 1];
  __shared__ int s_ok;
  const int n = p.n, tid = threadIdx.x;
  const int k0 = kb * NB, kn = min(NB, n - k0);
  if (k0 >= n || (kb + (int)blockIdx.x) * NB >= n) return;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    D[i][j] = (i < kn && j < kn) ? p.S[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
  }
  if (tid == 0) s_ok = 1;
  __syncthreads();
  if (tid < 32) {  // one warp, lane = row
    for (int j = 0; j < kn; j++) {
      double d = D[j][j];
      if (!(d > 0.0)) {
        if (tid == 0) s_ok = 0;
        d = 1.0;
      }
      const double dj = sqrt(d);
      __syncwarp();
      if (tid == j) D[j][j] = dj;
      if (tid > j && tid < kn) D[tid][j] /= dj;
      __syncwarp();
      if (tid > j && tid < kn)
        for (int k = j + 1; k <= tid; k++) D[tid][k] -= D[tid][j] * D[k][j];
      __syncwarp();
    }
  }
  __syncthreads();
  const int ib = kb + blockIdx.x;
  const int i0 = ib * NB, in = min(NB, n - i0);
  if (blockIdx.x == 0) {
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int i = e / NB, j = e % NB;
      if (i < kn && j <= i) p.S[(size_t)(k0 + i) * n + k0 + j] = D[i][j];
    }
    if (tid == 0 && !s_ok) st->solve_ok = 0;
    return;
  }
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    Ablk[i][j] = (i < in && j < kn) ? p.S[(size_t)(i0 + i) * n + k0 + j] : 0.0;
  }
  __syncthreads();
  if (tid < in) {  // x L^T = a, one row per thread
    for (int c = 0; c < kn; c++) {
      double v = Ablk[tid][c];
      for (int m = 0; m < c; m++) v -= Ablk[tid][m] * D[c][m];
      Ablk[tid][c] = v / D[c][c];
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    if (i < in && j < kn) p.S[(size_t)(i0 + i) * n + k0 + j] = Ablk[i][j];
  }
}

// Trailing update A_ij -= L_ik L_jk^T for block pairs i >= j > kb (lower triangle).
__global__ void __launch_bounds__(256) ba_chol_update_kernel(const BADev* __restrict__ probs, int kb, int nblk) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  __shared__ double Li[NB][NB + 1];
  __shared__ double Lj[NB][NB + 1];
  const int n = p.n, tid = threadIdx.x;
  nblk = (n + NB - 1) / NB;
  // decode (bi, bj) from the linear block id over the lower triangle of the trailing part
  const int m = nblk - kb - 1;
  if (m <= 0) return;
  int t = blockIdx.x, bi = 0;
  while (t >= bi + 1) {
    t -= bi + 1;
    bi++;
  }
  const int bj = t;
  if (bi >= m) return;
  const int i0 = (kb + 1 + bi) * NB, j0 = (kb + 1 + bj) * NB, k0 = kb * NB;
  const int in = min(NB, n - i0), jn = min(NB, n - j0), kn = min(NB, n - k0);
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, k = e % NB;
    Li[i][k] = (i < in && k < kn) ? p.S[(size_t)(i0 + i) * n + k0 + k] : 0.0;
    Lj[i][k] = (i < jn && k < kn) ? p.S[(size_t)(j0 + i) * n + k0 + k] : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    if (i < in && j < jn) {
      double acc = 0;
#pragma unroll 8
      for (int k = 0; k < NB; k++) acc += Li[i][k] * Lj[j][k];
      p.S[(size_t)(i0 + i) * n + j0 + j] -= acc;
    }
  }
}

// Triangular solves with the factor in global memory; one CTA, vector in smem.
__global__ void __launch_bounds__(1024) ba_chol_solve_kernel(const BADev* __restrict__ probs) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  extern __shared__ double y[];
  const int n = p.n, tid = threadIdx.x;
  const int nblk = (n + NB - 1) / NB;
  for (int i = tid; i < n; i += blockDim.x) y[i] = p.rhs[i];
  __syncthreads();
  for (int kb = 0; kb < nblk; kb++) {  // L y = b
    const int k0 = kb * NB, kn = min(NB, n - k0);
    if (tid < 32) {
      for (int j = 0; j < kn; j++) {
        if (tid == j) y[k0 + j] /= p.S[(size_t)(k0 + j) * n + k0 + j];
        __syncwarp();
        if (tid > j && tid < kn) y[k0 + tid] -= p.S[(size_t)(k0 + tid) * n + k0 + j] * y[k0 + j];
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = k0 + kn + tid; i < n; i += blockDim.x) {
      const double* Lrow = p.S + (size_t)i * n + k0;
      double acc = 0;
      for (int m = 0; m < kn; m++) acc += Lrow[m] * y[k0 + m];
      y[i] -= acc;
    }
    __syncthreads();
  }
  for (int kb = nblk - 1; kb >= 0; kb--) {  // L^T x = y
    const int k0 = kb * NB, kn = min(NB, n - k0);
    if (tid < 32) {
      for (int j = kn - 1; j >= 0; j--) {
        if (tid == j) y[k0 + j] /= p.S[(size_t)(k0 + j) * n + k0 + j];
        __syncwarp();
        if (tid < j) y[k0 + tid] -= p.S[(size_t)(k0 + j) * n + k0 + tid] * y[k0 + j];
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = tid; i < k0; i += blockDim.x) {
      double acc = 0;
      for (int m = 0; m < kn; m++) acc += p.S[(size_t)(k0 + m) * n + i] * y[k0 + m];
      y[i] -= acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += blockDim.x) p.rhs[i] = y[i];
}

// Candidate cameras x_c - y_c * scale_c (+ their rotation blocks).
__global__ void ba_candcam_kernel(const BADev* __restrict__ probs) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const int cur = st->cur;
  double step2 = 0, xc2 = 0;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const double d = -p.rhs[6 * c + a] * p.scale_c[6 * c + a];
    const double x = p.cams[cur][6 * c + a] + d;
    p.cams[cur ^ 1][6 * c + a] = x;
    step2 += d * d;
    xc2 += x * x;
  }
  cam_rotation(p.cams[cur ^ 1] + 6 * c, p.camrot[cur ^ 1] + CAMROT * c, true);
  // camera parts are replicated on every rank: kept apart from the all-reduced sums
  atomicAdd(&p.tail[3], step2);
  atomicAdd(&p.tail[4], xc2);
}

// Back-substitution, model decrease, candidate points and candidate cost.
__device__ __forceinline__ void lm_control(const BADev& p, const lorb_ba_options& opt, int* n_active) {
  // volatile copy: accumulators were written by other CTAs through L2 atomics
  LMState loc;
  {
    const volatile long long* src = reinterpret_cast<const volatile long long*>(p.st);
    long long* dst = reinterpret_cast<long long*>(&loc);
    for (int i = 0; i < (int)(sizeof(LMState) / 8); i++) dst[i] = src[i];
  }
  const volatile double* tail = p.tail;
  loc.acc_step2 += tail[3];
  loc.acc_xcand2 += tail[4];
  const int acc = lm_decide(&loc, opt);
  if (acc) loc.cur ^= 1;
  lm_zero_acc(&loc);
  loc.blocks_done = 0;
  if (!loc.done) atomicAdd(n_active, 1);
  *p.st = loc;
}

__global__ void __launch_bounds__(BA_THREADS)
    ba_backsub_kernel(const BADev* __restrict__ probs, lorb_ba_options opt, int fuse_control,
                      int* __restrict__ n_active) {
  const BADev p = probs[blockIdx.y];
  LMState* st = p.st;
  if (st->done) return;
  __shared__ double red[BA_THREADS / 32];
  const int cur = st->cur;
  const double* cams = p.cams[cur];
  const double* pts = p.pts[cur];
  const double* camrot = p.camrot[cur];
  const double* cams_c = p.cams[cur ^ 1];
  const double* camrot_c = p.camrot[cur ^ 1];
  double* pts_c = p.pts[cur ^ 1];
  const double* yc = p.rhs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gl = lane & 7, gw = lane >> 3;
  double model_acc = 0, cost_acc = 0, step_acc = 0, xn_acc = 0;
  for (int base = (blockIdx.x * (BA_THREADS / 32) + warp) * 4; base < p.P;
       base += gridDim.x * (BA_THREADS / 8)) {
    const int pt = base + gw;
    const bool pv = pt < p.P;
    int s = 0, e = 0;
    double X[3] = {0, 0, 0}, sp[3] = {1, 1, 1};
    if (pv) {
      s = p.pt_ptr[pt];
      e = p.pt_ptr[pt + 1];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        X[a] = pts[3 * (size_t)pt + a];
        sp[a] = p.scale_p[3 * (size_t)pt + a];
      }
    }
    const int rounds = __reduce_max_sync(0xffffffffu, (e - s + 7) >> 3);
    double bsum[3] = {0, 0, 0};
    double r[2], Jc[12], Jp[6];
    int cam = -1;
    bool has = false;
    double jcy[2] = {0, 0};
    for (int rd = 0; rd < rounds; rd++) {
      const int o = s + rd * 8 + gl;
      has = o < e;
      if (has) {
        cam = eval_obs(p, cams, camrot, o, X, sp, r, Jc, Jp);
        jcy[0] = jcy[1] = 0;
        if (cam >= 0) {
#pragma unroll
          for (int a = 0; a < 6; a++) {
            jcy[0] += Jc[a] * yc[6 * cam + a];
            jcy[1] += Jc[6 + a] * yc[6 * cam + a];
          }
#pragma unroll
          for (int a = 0; a < 3; a++) bsum[a] += Jp[a] * jcy[0] + Jp[3 + a] * jcy[1];
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) bsum[a] = group_sum8(bsum[a]);
    double yp[3] = {0, 0, 0}, Xc[3] = {0, 0, 0};
    if (pv) {
      const double* hi = p.pt_hinv + 6 * (size_t)pt;
      const double* gp = p.pt_gp + 3 * (size_t)pt;
      const double b0 = gp[0] - bsum[0], b1 = gp[1] - bsum[1], b2 = gp[2] - bsum[2];
      yp[0] = hi[0] * b0 + hi[1] * b1 + hi[2] * b2;
      yp[1] = hi[1] * b0 + hi[3] * b1 + hi[4] * b2;
      yp[2] = hi[2] * b0 + hi[4] * b1 + hi[5] * b2;
#pragma unroll
      for (int a = 0; a < 3; a++) Xc[a] = X[a] - yp[a] * sp[a];
      if (gl == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const double d = yp[a] * sp[a];
          pts_c[3 * (size_t)pt + a] = Xc[a];
          step_acc += d * d;
          xn_acc += Xc[a] * Xc[a];
        }
      }
    }
    // model residual m = J*step = -(Jc yc + Jp yp); candidate cost
    for (int rd = 0; rd < rounds; rd++) {
      const int o = s + rd * 8 + gl;
      if (o < e) {
        if (rounds > 1) {
          cam = eval_obs(p, cams, camrot, o, X, sp, r, Jc, Jp);
          jcy[0] = jcy[1] = 0;
          if (cam >= 0) {
#pragma unroll
            for (int a = 0; a < 6; a++) {
              jcy[0] += Jc[a] * yc[6 * cam + a];
              jcy[1] += Jc[6 + a] * yc[6 * cam + a];
            }
          }
        }
        const double m0 = -(jcy[0] + Jp[0] * yp[0] + Jp[1] * yp[1] + Jp[2] * yp[2]);
        const double m1 = -(jcy[1] + Jp[3] * yp[0] + Jp[4] * yp[1] + Jp[5] * yp[2]);
        model_acc += m0 * (r[0] + 0.5 * m0) + m1 * (r[1] + 0.5 * m1);
        const float2 uv = p.obs_uv[o];
        if (cam >= 0) {
          const double* t = cams_c + 6 * cam + 3;
          const double tt[3] = {t[0], t[1], t[2]};
          cost_acc += obs_cost(camrot_c + CAMROT * cam, tt, Xc, p.K, (double)uv.x, (double)uv.y);
        } else {
          const double* f = p.fixrt + 12 * (size_t)(-1 - cam);
          const double tt[3] = {f[9], f[10], f[11]};
          cost_acc += obs_cost(f, tt, Xc, p.K, (double)uv.x, (double)uv.y);
        }
      }
    }
  }
  double t = block_sum(model_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&st->acc_model, t);
  t = block_sum(cost_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&st->acc_cost2, t);
  t = block_sum(step_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&st->acc_step2, t);
  t = block_sum(xn_acc, red);
  if (tid == 0 && t != 0.0) atomicAdd(&st->acc_xcand2, t);
  if (fuse_control && tid == 0) {
    __threadfence();
    const int prev = atomicAdd(&st->blocks_done, 1);
    if (prev == (int)gridDim.x - 1) {
      __threadfence();
      lm_control(p, opt, n_active);
    }
  }
}

__global__ void ba_control_kernel(const BADev* __restrict__ probs, lorb_ba_options opt, int* __restrict__ n_active) {
  const BADev p = probs[blockIdx.x];
  LMState* st = p.st;
  if (st->done) return;
  lm_control(p, opt, n_active);
}

}  // namespace lorb

using namespace lorb;


// ------------------------------------------------------------------ host side
// A lorb_ba_problem is a batch of nw >= 1 independent windows resident in HBM
// (nw = 1 for lorb_ba_local / the sharded large solve).
struct lorb_ba_problem {
  lorb_ctx* ctx = nullptr;
  int nw = 0;
  std::vector<lorb::BADev> h_dev;     // host copies of the per-window descriptors
  std::vector<int> h_cam_off, h_pt_off;
  lorb::Buf params, topo, work, descs, hstate, counter;
  double *cams0 = nullptr, *pts0 = nullptr;  // initial parameters of all windows
  size_t cam_doubles = 0, pt_doubles = 0;
  int maxC = 0, maxP = 0;
  size_t lin_doubles_max = 0, lin_bytes_total = 0;
  double* lin_base = nullptr;
  lorb::LMState* d_states = nullptr;
};

namespace lorb {

static size_t al(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

struct WindowSpec {  // host view of one window's inputs
  int C, P, O, F;
  const double* cams;
  const double* pts;
  const int* obs_cam;
  const int* obs_pt;
  const float* obs_uv;
  const int* fix_pt;
  const float* fix_uv;
  const float* fix_rt;
};

static void host_fix_rotation(const float* rt, double* R) {
  // float angle-axis widened to double (T(mRvec.at<float>(i)), reference
  // src/bundle_adjust.cpp:87), Rodrigues / small-angle as ceres::AngleAxisRotatePoint
  const double w0 = rt[0], w1 = rt[1], w2 = rt[2];
  const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
  if (th2 > DBL_EPSILON) {
    const double th = sqrt(th2), s = sin(th), cth = cos(th), ith = 1.0 / th;
    const double k0 = w0 * ith, k1 = w1 * ith, k2 = w2 * ith, omc = 1.0 - cth;
    R[0] = cth + omc * k0 * k0; R[1] = -s * k2 + omc * k0 * k1; R[2] = s * k1 + omc * k0 * k2;
    R[3] = s * k2 + omc * k1 * k0; R[4] = cth + omc * k1 * k1; R[5] = -s * k0 + omc * k1 * k2;
    R[6] = -s * k1 + omc * k2 * k0; R[7] = s * k0 + omc * k2 * k1; R[8] = cth + omc * k2 * k2;
  } else {
    R[0] = 1; R[1] = -w2; R[2] = w1; R[3] = w2; R[4] = 1; R[5] = -w0; R[6] = -w1; R[7] = w0; R[8] = 1;
  }
  R[9] = rt[3];
  R[10] = rt[4];
  R[11] = rt[5];
}

static int problem_build(lorb_ba_problem* pb, lorb_ctx* c, const std::vector<WindowSpec>& ws,
                         const float* K) {
  pb->ctx = c;
  const int nw = (int)ws.size();
  pb->nw = nw;
  pb->h_dev.resize(nw);
  pb->h_cam_off.assign(nw + 1, 0);
  pb->h_pt_off.assign(nw + 1, 0);
  size_t tot_obs = 0, tot_fix = 0, tot_ptr = 0, tot_lin = 0, tot_rot = 0;
  for (int w = 0; w < nw; w++) {
    const WindowSpec& W = ws[w];
    LORB_REQUIRE(W.C > 0 && W.P >= 0 && W.O >= 0 && W.F >= 0, "window sizes");
    pb->h_cam_off[w + 1] = pb->h_cam_off[w] + W.C;
    pb->h_pt_off[w + 1] = pb->h_pt_off[w] + W.P;
    pb->maxC = std::max(pb->maxC, W.C);
    pb->maxP = std::max(pb->maxP, W.P);
    const int n = 6 * W.C;
    const size_t lin = (size_t)n * n + (size_t)HCC * W.C + 12 * (size_t)W.C + 8;
    pb->lin_doubles_max = std::max(pb->lin_doubles_max, lin);
    tot_lin += lin;
    tot_obs += (size_t)W.O + W.F;
    tot_fix += W.F;
    tot_ptr += (size_t)W.P + 1;
    tot_rot += (size_t)W.C * CAMROT;
  }
  const size_t totC = pb->h_cam_off[nw], totP = pb->h_pt_off[nw];
  pb->cam_doubles = totC * 6;
  pb->pt_doubles = totP * 3;
  // ---- host staging of the topology (CSR by point per window)
  std::vector<int> h_ptr(tot_ptr), h_cam(std::max<size_t>(tot_obs, 1));
  std::vector<float2> h_uv(std::max<size_t>(tot_obs, 1));
  std::vector<double> h_fix(std::max<size_t>(tot_fix, 1) * 12);
  std::vector<double> h_cams(pb->cam_doubles), h_pts(std::max<size_t>(pb->pt_doubles, 1));
  std::vector<size_t> o_ptr(nw), o_obs(nw), o_fix(nw);
  {
    size_t a = 0, b = 0, f = 0;
    for (int w = 0; w < nw; w++) {
      const WindowSpec& W = ws[w];
      o_ptr[w] = a;
      o_obs[w] = b;
      o_fix[w] = f;
      int* ptr = &h_ptr[a];
      for (int i = 0; i <= W.P; i++) ptr[i] = 0;
      for (int i = 0; i < W.O; i++) {
        LORB_REQUIRE(W.obs_pt[i] >= 0 && W.obs_pt[i] < W.P && W.obs_cam[i] >= 0 && W.obs_cam[i] < W.C,
                     "observation index out of range");
        ptr[W.obs_pt[i] + 1]++;
      }
      for (int i = 0; i < W.F; i++) {
        LORB_REQUIRE(W.fix_pt[i] >= 0 && W.fix_pt[i] < W.P, "fixed observation point out of range");
        ptr[W.fix_pt[i] + 1]++;
      }
      for (int i = 0; i < W.P; i++) ptr[i + 1] += ptr[i];
      std::vector<int> fill(ptr, ptr + W.P);
      for (int i = 0; i < W.O; i++) {
        const int d = fill[W.obs_pt[i]]++;
        h_cam[b + d] = W.obs_cam[i];
        h_uv[b + d] = make_float2(W.obs_uv[2 * i], W.obs_uv[2 * i + 1]);
      }
      for (int i = 0; i < W.F; i++) {
        const int d = fill[W.fix_pt[i]]++;
        h_cam[b + d] = -1 - i;
        h_uv[b + d] = make_float2(W.fix_uv[2 * i], W.fix_uv[2 * i + 1]);
        host_fix_rotation(W.fix_rt + 6 * (size_t)i, &h_fix[12 * (f + i)]);
      }
      memcpy(&h_cams[6 * (size_t)pb->h_cam_off[w]], W.cams, (size_t)W.C * 48);
      if (W.P) memcpy(&h_pts[3 * (size_t)pb->h_pt_off[w]], W.pts, (size_t)W.P * 24);
      a += (size_t)W.P + 1;
      b += (size_t)W.O + W.F;
      f += W.F;
    }
  }
  // ---- device layout
  const size_t cb = al(pb->cam_doubles * 8), pbts = al(std::max<size_t>(pb->pt_doubles, 1) * 8);
  LORB_TRY(pb->params.reserve(3 * cb + 3 * pbts));
  uint8_t* q = pb->params.as<uint8_t>();
  double* d_cams[2] = {(double*)q, (double*)(q + cb)};
  pb->cams0 = (double*)(q + 2 * cb);
  double* d_pts[2] = {(double*)(q + 3 * cb), (double*)(q + 3 * cb + pbts)};
  pb->pts0 = (double*)(q + 3 * cb + 2 * pbts);
  const size_t t_ptr = al(tot_ptr * 4), t_cam = al(std::max<size_t>(tot_obs, 1) * 4),
               t_uv = al(std::max<size_t>(tot_obs, 1) * 8), t_fix = al(std::max<size_t>(tot_fix, 1) * 96);
  LORB_TRY(pb->topo.reserve(t_ptr + t_cam + t_uv + t_fix));
  uint8_t* t = pb->topo.as<uint8_t>();
  int* d_ptr = (int*)t;
  int* d_ocam = (int*)(t + t_ptr);
  float2* d_uv = (float2*)(t + t_ptr + t_cam);
  double* d_fix = (double*)(t + t_ptr + t_cam + t_uv);
  const size_t w_rot = al(tot_rot * 8), w_sc = cb, w_sp = pbts, w_lin = al(tot_lin * 8), w_rhs = cb,
               w_hinv = al(std::max<size_t>(totP, 1) * 48), w_gp = pbts,
               w_st = al(sizeof(LMState) * (size_t)nw);
  LORB_TRY(pb->work.reserve(2 * w_rot + w_sc + w_sp + w_lin + w_rhs + w_hinv + w_gp + w_st));
  uint8_t* wk = pb->work.as<uint8_t>();
  double* d_rot[2] = {(double*)wk, (double*)(wk + w_rot)};
  wk += 2 * w_rot;
  double* d_sc = (double*)wk;   wk += w_sc;
  double* d_sp = (double*)wk;   wk += w_sp;
  double* d_lin = (double*)wk;  wk += w_lin;
  double* d_rhs = (double*)wk;  wk += w_rhs;
  double* d_hinv = (double*)wk; wk += w_hinv;
  double* d_gp = (double*)wk;   wk += w_gp;
  pb->d_states = (LMState*)wk;
  pb->lin_base = d_lin;
  pb->lin_bytes_total = tot_lin * 8;
  {
    size_t lin_off = 0, rot_off = 0;
    for (int w = 0; w < nw; w++) {
      const WindowSpec& W = ws[w];
      BADev& d = pb->h_dev[w];
      const size_t co = 6 * (size_t)pb->h_cam_off[w], po = 3 * (size_t)pb->h_pt_off[w];
      d.C = W.C;
      d.P = W.P;
      d.n = 6 * W.C;
      for (int b = 0; b < 2; b++) {
        d.cams[b] = d_cams[b] + co;
        d.pts[b] = d_pts[b] + po;
        d.camrot[b] = d_rot[b] + rot_off;
      }
      d.pt_ptr = d_ptr + o_ptr[w];
      d.obs_cam = d_ocam + o_obs[w];
      d.obs_uv = d_uv + o_obs[w];
      d.fixrt = d_fix + 12 * o_fix[w];
      d.K.fu = (double)K[0];
      d.K.fv = (double)K[1];
      d.K.cx = (double)K[2];
      d.K.cy = (double)K[3];
      d.scale_c = d_sc + co;
      d.scale_p = d_sp + po;
      d.lin = d_lin + lin_off;
      d.S = d.lin;
      d.Hcc = d.S + (size_t)d.n * d.n;
      d.gc = d.Hcc + (size_t)HCC * W.C;
      d.rhs_corr = d.gc + 6 * (size_t)W.C;
      d.tail = d.rhs_corr + 6 * (size_t)W.C;
      d.rhs = d_rhs + co;
      d.pt_hinv = d_hinv + 2 * po;
      d.pt_gp = d_gp + po;
      d.st = pb->d_states + w;
      lin_off += (size_t)d.n * d.n + (size_t)HCC * W.C + 12 * (size_t)W.C + 8;
      rot_off += (size_t)W.C * CAMROT;
    }
  }
  LORB_TRY(pb->descs.reserve(sizeof(BADev) * (size_t)nw));
  LORB_TRY(pb->counter.reserve(256));
  pb->hstate.pinned = true;
  LORB_TRY(pb->hstate.reserve(sizeof(LMState) * (size_t)nw + 64));
  cudaStream_t s = c->stream;
  LORB_CUDA_TRY(cudaMemcpyAsync(pb->descs.p, pb->h_dev.data(), sizeof(BADev) * (size_t)nw, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaMemcpyAsync(pb->cams0, h_cams.data(), pb->cam_doubles * 8, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaMemcpyAsync(pb->pts0, h_pts.data(), pb->pt_doubles * 8, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaMemcpyAsync(d_ptr, h_ptr.data(), tot_ptr * 4, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaMemcpyAsync(d_ocam, h_cam.data(), tot_obs * 4, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaMemcpyAsync(d_uv, h_uv.data(), tot_obs * 8, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaMemcpyAsync(d_fix, h_fix.data(), tot_fix * 96, cudaMemcpyHostToDevice, s));
  LORB_CUDA_TRY(cudaStreamSynchronize(s));  // host staging vectors go out of scope
  return LORB_OK;
}

__global__ void fill_ones_kernel(double* a, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a[i] = 1.0;
}

static int problem_reset(lorb_ba_problem* pb) {
  lorb_ctx* c = pb->ctx;
  const BADev& d0 = pb->h_dev[0];
  LORB_CUDA_TRY(cudaMemcpyAsync(d0.cams[0], pb->cams0, pb->cam_doubles * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (pb->pt_doubles)
    LORB_CUDA_TRY(cudaMemcpyAsync(d0.pts[0], pb->pts0, pb->pt_doubles * 8, cudaMemcpyDeviceToDevice, c->stream));
  return LORB_OK;
}

static int run_cholesky(lorb_ba_problem* pb) {
  lorb_ctx* c = pb->ctx;
  const BADev* dp = pb->descs.as<BADev>();
  const int n = 6 * pb->maxC, nw = pb->nw;
  const size_t small_bytes = ((size_t)n * n + n) * 8;
  if (small_bytes <= 160 * 1024) {
    LORB_CUDA_TRY(cudaFuncSetAttribute(ba_chol_small_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_bytes));
    LORB_LAUNCH(c, ba_chol_small_kernel, dim3(1, nw), 256, small_bytes, dp);
    return LORB_OK;
  }
  const int nblk = (n + NB - 1) / NB;
  for (int kb = 0; kb < nblk; kb++) {
    LORB_LAUNCH(c, ba_chol_panel_kernel, dim3(nblk - kb, nw), NB * NB / 4, 0, dp, kb);
    const int m = nblk - kb - 1;
    if (m > 0) LORB_LAUNCH(c, ba_chol_update_kernel, dim3(m * (m + 1) / 2, nw), 256, 0, dp, kb, nblk);
  }
  LORB_REQUIRE((size_t)n * 8 <= 200 * 1024, "reduced camera system too large for the solve kernel");
  LORB_CUDA_TRY(cudaFuncSetAttribute(ba_chol_solve_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8));
  LORB_LAUNCH(c, ba_chol_solve_kernel, dim3(1, nw), 1024, (size_t)n * 8, dp);
  return LORB_OK;
}

static int problem_solve(lorb_ba_problem* pb, const lorb_ba_options* optp, int sharded,
                         lorb_ba_summary* sums) {
  lorb_ctx* c = pb->ctx;
  const BADev* dp = pb->descs.as<BADev>();
  const BADev& d0 = pb->h_dev[0];
  const int nw = pb->nw;
  lorb_ba_options opt = *optp;
  LORB_REQUIRE(!sharded || (dist_ready(c) && nw == 1), "sharded solve needs lorb_dist_init and one window");
  const int gx_pts = std::max(1, std::min((pb->maxP + 31) / 32, std::max(1, c->sm_count * 8 / nw)));
  const dim3 grid_pts(gx_pts, nw);
  const dim3 grid_cam((pb->maxC + 127) / 128, nw);
  const int nmax = 6 * pb->maxC;
  const dim3 grid_fin(std::max(1, std::min(c->sm_count * 2 / std::min(nw, c->sm_count) + 1, (nmax * nmax + 255) / 256)), nw);
  cudaStream_t s = c->stream;
  int* d_active = pb->counter.as<int>();
  int* h_active = reinterpret_cast<int*>(pb->hstate.as<uint8_t>() + sizeof(LMState) * (size_t)nw);
  // small windows: shared-memory privatised accumulation + one fused solve kernel
  const bool small = nmax <= 96;
  const size_t smem_lin = small ? ((size_t)nmax * nmax + (size_t)(HCC + 12) * pb->maxC) * 8 : 0;
  const size_t smem_solve = ((size_t)nmax * nmax + 2 * (size_t)nmax) * 8;
  if (small) {
    LORB_CUDA_TRY(cudaFuncSetAttribute(ba_build_kernel<true, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lin));
    LORB_CUDA_TRY(cudaFuncSetAttribute(ba_build_kernel<false, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lin));
    LORB_CUDA_TRY(cudaFuncSetAttribute(ba_solve_small_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
  }
  auto launch_build = [&](bool full, int force) -> int {
    if (small) {
      if (full)
        LORB_LAUNCH(c, (ba_build_kernel<true, true>), grid_pts, BA_THREADS, smem_lin, dp, opt, force);
      else
        LORB_LAUNCH(c, (ba_build_kernel<false, true>), grid_pts, BA_THREADS, smem_lin, dp, opt, force);
    } else {
      if (full)
        LORB_LAUNCH(c, (ba_build_kernel<true, false>), grid_pts, BA_THREADS, 0, dp, opt, force);
      else
        LORB_LAUNCH(c, (ba_build_kernel<false, false>), grid_pts, BA_THREADS, 0, dp, opt, force);
    }
    return LORB_OK;
  };
  // ---- initial evaluation (iteration 0)
  LORB_CUDA_TRY(cudaMemsetAsync(pb->d_states, 0, sizeof(LMState) * (size_t)nw, s));
  LORB_CUDA_TRY(cudaMemsetAsync(pb->lin_base, 0, pb->lin_bytes_total, s));
  LORB_LAUNCH(c, fill_ones_kernel, 128, 256, 0, d0.scale_c, pb->cam_doubles);
  LORB_LAUNCH(c, fill_ones_kernel, 512, 256, 0, d0.scale_p, pb->pt_doubles);
  LORB_LAUNCH(c, ba_camrot_kernel, grid_cam, 128, 0, dp, 0, 1);
  LORB_TRY(launch_build(false, 1));
  if (sharded) {
    LORB_TRY(dist_allreduce_sum(c, d0.Hcc, (size_t)HCC * d0.C + 12 * (size_t)d0.C + 8));
    LORB_TRY(dist_allreduce_max_u64(c, &d0.st->acc_gmax, 1));
  }
  LORB_LAUNCH(c, ba_init_finish_kernel, dim3(1, nw), 1024, 0, dp, opt);
  // ---- LM attempts; the device decides, the host polls a counter every few attempts
  const int poll_every = 4;
  const int fuse_control = sharded ? 0 : 1;
  for (int it = 0; it < opt.max_num_iterations; it++) {
    LORB_CUDA_TRY(cudaMemsetAsync(pb->lin_base, 0, pb->lin_bytes_total, s));
    LORB_CUDA_TRY(cudaMemsetAsync(d_active, 0, 4, s));
    LORB_TRY(launch_build(true, 0));
    if (sharded) {
      LORB_TRY(dist_allreduce_sum(c, d0.lin, pb->lin_doubles_max));
      LORB_TRY(dist_allreduce_max_u64(c, &d0.st->acc_gmax, 1));
    }
    if (small) {
      LORB_LAUNCH(c, ba_solve_small_kernel, dim3(1, nw), 256, smem_solve, dp, opt);
    } else {
      LORB_LAUNCH(c, ba_gradcheck_kernel, dim3(1, nw), 256, 0, dp, opt, 0);
      LORB_LAUNCH(c, ba_finish_kernel, grid_fin, 256, 0, dp, opt);
      LORB_TRY(run_cholesky(pb));
      LORB_LAUNCH(c, ba_candcam_kernel, grid_cam, 128, 0, dp);
    }
    LORB_LAUNCH(c, ba_backsub_kernel, grid_pts, BA_THREADS, 0, dp, opt, fuse_control, d_active);
    if (sharded) {
      LORB_TRY(dist_allreduce_sum(c, &d0.st->acc_cost2, 4));
      LORB_LAUNCH(c, ba_control_kernel, nw, 1, 0, dp, opt, d_active);
    }
    if ((it + 1) % poll_every == 0 && it + 1 < opt.max_num_iterations) {
      LORB_CUDA_TRY(cudaMemcpyAsync(h_active, d_active, 4, cudaMemcpyDeviceToHost, s));
      LORB_CUDA_TRY(cudaStreamSynchronize(s));
      if (*h_active == 0) break;
    }
  }
  // gradient at the final point of windows whose last attempt moved it (Ceres tests the
  // gradient tolerance on acceptance); windows without a pending check ignore the pass
  LMState* hs = pb->hstate.as<LMState>();
  LORB_CUDA_TRY(cudaMemcpyAsync(hs, pb->d_states, sizeof(LMState) * (size_t)nw, cudaMemcpyDeviceToHost, s));
  LORB_CUDA_TRY(cudaStreamSynchronize(s));
  bool pending = false;
  for (int w = 0; w < nw; w++) pending |= hs[w].check_gradient != 0;
  if (pending) {
    LORB_CUDA_TRY(cudaMemsetAsync(pb->lin_base, 0, pb->lin_bytes_total, s));
    LORB_LAUNCH(c, ba_camrot_kernel, grid_cam, 128, 0, dp, 0, 2);
    LORB_TRY(launch_build(false, 2));
    if (sharded) {
      LORB_TRY(dist_allreduce_sum(c, d0.Hcc, (size_t)HCC * d0.C + 12 * (size_t)d0.C + 8));
      LORB_TRY(dist_allreduce_max_u64(c, &d0.st->acc_gmax, 1));
    }
    LORB_LAUNCH(c, ba_gradcheck_kernel, dim3(1, nw), 256, 0, dp, opt, 2);
    LORB_CUDA_TRY(cudaMemcpyAsync(hs, pb->d_states, sizeof(LMState) * (size_t)nw, cudaMemcpyDeviceToHost, s));
    LORB_CUDA_TRY(cudaStreamSynchronize(s));
  }
  // make buffer 0 the current one so download / the next solve find the result there
  for (int w = 0; w < nw; w++) {
    if (hs[w].cur == 1) {
      const BADev& d = pb->h_dev[w];
      LORB_CUDA_TRY(cudaMemcpyAsync(d.cams[0], d.cams[1], (size_t)d.n * 8, cudaMemcpyDeviceToDevice, s));
      if (d.P) LORB_CUDA_TRY(cudaMemcpyAsync(d.pts[0], d.pts[1], (size_t)d.P * 24, cudaMemcpyDeviceToDevice, s));
    }
    if (sums) {
      lorb_ba_summary& sum = sums[w];
      sum.initial_cost = hs[w].initial_cost;
      sum.final_cost = hs[w].cost;
      sum.final_radius = hs[w].radius;
      sum.final_gradient_max_norm = hs[w].gmax;
      sum.iterations = hs[w].iteration;
      sum.num_successful_steps = hs[w].n_success;
      sum.num_unsuccessful_steps = hs[w].n_fail;
      sum.termination = hs[w].termination;
    }
  }
  LORB_CUDA_TRY(cudaStreamSynchronize(s));
  return LORB_OK;
}

static void problem_free(lorb_ba_problem* pb) {
  if (!pb) return;
  if (pb->ctx) {
    cudaSetDevice(pb->ctx->device);
    cudaStreamSynchronize(pb->ctx->stream);
  }
  pb->params.release();
  pb->topo.release();
  pb->work.release();
  pb->descs.release();
  pb->hstate.release();
  pb->counter.release();
  delete pb;
}

}  // namespace lorb

extern "C" {

int lorb_ba_problem_create(lorb_ctx* c, int C, const double* cams, int P, const double* pts, int O,
                           const int* obs_cam, const int* obs_pt, const float* obs_uv, int F,
                           const int* fix_pt, const float* fix_uv, const float* fix_rt,
                           const float* K, lorb_ba_problem** out) {
  LORB_REQUIRE(c && out && K, "ctx / out / K");
  LORB_REQUIRE(C > 0 && P >= 0 && O >= 0 && F >= 0, "sizes");
  LORB_REQUIRE(cams && (P == 0 || pts), "parameters");
  LORB_REQUIRE(O == 0 || (obs_cam && obs_pt && obs_uv), "observations");
  LORB_REQUIRE(F == 0 || (fix_pt && fix_uv && fix_rt), "fixed observations");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  lorb_ba_problem* pb = new (std::nothrow) lorb_ba_problem();
  if (!pb) return LORB_ERR_NOMEM;
  std::vector<WindowSpec> ws(1);
  ws[0] = WindowSpec{C, P, O, F, cams, pts, obs_cam, obs_pt, obs_uv, fix_pt, fix_uv, fix_rt};
  int rc = problem_build(pb, c, ws, K);
  if (rc == LORB_OK) rc = problem_reset(pb);
  if (rc != LORB_OK) {
    problem_free(pb);
    return rc;
  }
  *out = pb;
  return LORB_OK;
}

int lorb_ba_problem_reset(lorb_ba_problem* pb) {
  LORB_REQUIRE(pb, "problem");
  LORB_CUDA_TRY(cudaSetDevice(pb->ctx->device));
  return problem_reset(pb);
}

int lorb_ba_problem_solve(lorb_ba_problem* pb, const lorb_ba_options* opt, int sharded,
                          lorb_ba_summary* summary) {
  LORB_REQUIRE(pb && opt, "problem / options");
  LORB_CUDA_TRY(cudaSetDevice(pb->ctx->device));
  return problem_solve(pb, opt, sharded, summary);
}

int lorb_ba_problem_download(lorb_ba_problem* pb, double* cams, double* pts) {
  LORB_REQUIRE(pb, "problem");
  lorb_ctx* c = pb->ctx;
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  const BADev& d0 = pb->h_dev[0];
  if (cams)
    LORB_CUDA_TRY(cudaMemcpyAsync(cams, d0.cams[0], pb->cam_doubles * 8, cudaMemcpyDeviceToHost, c->stream));
  if (pts && pb->pt_doubles)
    LORB_CUDA_TRY(cudaMemcpyAsync(pts, d0.pts[0], pb->pt_doubles * 8, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return LORB_OK;
}

int lorb_ba_problem_destroy(lorb_ba_problem* pb) {
  problem_free(pb);
  return LORB_OK;
}

int lorb_ba_local(lorb_ctx* c, int C, double* cams, int P, double* pts, int O, const int* obs_cam,
                  const int* obs_pt, const float* obs_uv, int F, const int* fix_pt,
                  const float* fix_uv, const float* fix_rt, const float* K,
                  const lorb_ba_options* opt, lorb_ba_summary* summary) {
  LORB_REQUIRE(opt, "options");
  lorb_ba_problem* pb = nullptr;
  LORB_TRY(lorb_ba_problem_create(c, C, cams, P, pts, O, obs_cam, obs_pt, obs_uv, F, fix_pt, fix_uv,
                                  fix_rt, K, &pb));
  int rc = lorb_ba_problem_solve(pb, opt, 0, summary);
  if (rc == LORB_OK) rc = lorb_ba_problem_download(pb, cams, pts);
  problem_free(pb);
  return rc;
}

int lorb_ba_local_batched(lorb_ctx* c, int n_windows, const int* cam_off, double* cams,
                          const int* pt_off, double* pts, const int* obs_off, const int* obs_cam,
                          const int* obs_pt, const float* obs_uv, const int* fix_off,
                          const int* fix_pt, const float* fix_uv, const float* fix_rt,
                          const float* K, const lorb_ba_options* opt, lorb_ba_summary* summaries) {
  LORB_REQUIRE(c && opt && K, "ctx / options / K");
  LORB_REQUIRE(n_windows > 0 && cam_off && pt_off && obs_off && cams, "window offsets");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  std::vector<WindowSpec> ws((size_t)n_windows);
  for (int w = 0; w < n_windows; w++) {
    WindowSpec& W = ws[w];
    W.C = cam_off[w + 1] - cam_off[w];
    W.P = pt_off[w + 1] - pt_off[w];
    W.O = obs_off[w + 1] - obs_off[w];
    W.F = fix_off ? fix_off[w + 1] - fix_off[w] : 0;
    W.cams = cams + 6 * (size_t)cam_off[w];
    W.pts = pts ? pts + 3 * (size_t)pt_off[w] : nullptr;
    W.obs_cam = obs_cam + obs_off[w];
    W.obs_pt = obs_pt + obs_off[w];
    W.obs_uv = obs_uv + 2 * (size_t)obs_off[w];
    W.fix_pt = fix_off ? fix_pt + fix_off[w] : nullptr;
    W.fix_uv = fix_off ? fix_uv + 2 * (size_t)fix_off[w] : nullptr;
    W.fix_rt = fix_off ? fix_rt + 6 * (size_t)fix_off[w] : nullptr;
  }
  lorb_ba_problem* pb = new (std::nothrow) lorb_ba_problem();
  if (!pb) return LORB_ERR_NOMEM;
  int rc = problem_build(pb, c, ws, K);
  if (rc == LORB_OK) rc = problem_reset(pb);
  if (rc == LORB_OK) rc = problem_solve(pb, opt, 0, summaries);
  if (rc == LORB_OK) rc = lorb_ba_problem_download(pb, cams, pts);
  problem_free(pb);
  return rc;
}

}  // extern "C"
