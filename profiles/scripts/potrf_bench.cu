// Standalone timing of the warp-level 32x32 POTRF variants considered for the dense factorisations
// (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/potrf_bench potrf_bench.cu).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 32;

__device__ __forceinline__ double rsqrt_fast(double d) {
  // float seed + two Newton steps in double (relative error ~1e-16 after the second)
  double y = (double)rsqrtf((float)d);
  const double hd = 0.5 * d;
  y = y * (1.5 - hd * y * y);
  y = y * (1.5 - hd * y * y);
  return y;
}

template <int VAR>
__device__ __forceinline__ void potrf32(double (*A)[NB + 1], double* idg, int lane) {
  double a[NB];
#pragma unroll
  for (int c = 0; c < NB; c++) a[c] = A[lane][c];
#pragma unroll
  for (int c = 0; c < NB; c++) {
    double d = __shfl_sync(0xffffffffu, a[c], c);
    if (VAR == 0 && !(d > 0.0)) d = 1.0;
    if (VAR != 0) d = d > 0.0 ? d : 1.0;
    const double is = (VAR >= 2) ? rsqrt_fast(d) : rsqrt(d);
    const double l = a[c] * is;
    a[c] = l;
    if (lane == c) idg[c] = is;
    if (VAR >= 3) {
      // the next pivot does not need a shuffle: lane c+1 holds both factors
      if (c + 1 < NB) a[c + 1] -= (lane == c + 1) ? l * l : 0.0;
#pragma unroll
      for (int q = c + 1; q < NB; q++) {
        const double lq = __shfl_sync(0xffffffffu, l, q);
        if (q == c + 1) { if (lane != q) a[q] -= l * lq; } else a[q] -= l * lq;
      }
    } else {
#pragma unroll
      for (int q = c + 1; q < NB; q++) {
        const double lq = __shfl_sync(0xffffffffu, l, q);
        a[q] -= l * lq;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NB; c++)
    if (lane > c) A[c][lane] = a[c];
}

__device__ __forceinline__ double rcp_fast(double d) {
  // float seed + two Newton steps in double
  double y = (double)__frcp_rn((float)d);
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  return y;
}

// Two pivots per step through the 2x2 block [d1 e; e d2]: one reciprocal (of its determinant) on the
// dependency chain, the Cholesky columns of the pair computed off the chain.  VAR 4: __drcp_rn,
// VAR 5: float-seeded reciprocal and rsqrt.
template <int VAR>
__device__ __forceinline__ void potrf32_pair(double (*A)[NB + 1], double* idg, int lane) {
  double a[NB];
#pragma unroll
  for (int c = 0; c < NB; c++) a[c] = A[lane][c];
  double my_is = 1.0;
#pragma unroll
  for (int c = 0; c < NB; c += 2) {
    double d1 = __shfl_sync(0xffffffffu, a[c], c);
    double e = __shfl_sync(0xffffffffu, a[c], c + 1);
    double d2 = __shfl_sync(0xffffffffu, a[c + 1], c + 1);
    double det = d1 * d2 - e * e;
    if (!(d1 > 0.0) || !(det > 0.0)) {
      d1 = d2 = det = 1.0;
      e = 0.0;
    }
    const double rdet = VAR == 4 ? __drcp_rn(det) : rcp_fast(det);
    const double x = a[c], y = a[c + 1];
    const double u = (x * d2 - y * e) * rdet;
    const double v = (y * d1 - x * e) * rdet;
#pragma unroll
    for (int q = c + 2; q < NB; q++) {
      const double xq = __shfl_sync(0xffffffffu, x, q);
      const double yq = __shfl_sync(0xffffffffu, y, q);
      a[q] = fma(-v, yq, fma(-u, xq, a[q]));
    }
    const double is1 = VAR == 4 ? rsqrt(d1) : rsqrt_fast(d1);
    const double s2 = det * (is1 * is1);
    const double is2 = VAR == 4 ? rsqrt(s2) : rsqrt_fast(s2);
    const double l1 = x * is1;
    a[c] = l1;
    a[c + 1] = (y - l1 * (e * is1)) * is2;
    if (lane == c) my_is = is1;
    if (lane == c + 1) my_is = is2;
  }
  idg[lane] = my_is;
#pragma unroll
  for (int c = 0; c < NB; c++)
    if (lane > c) A[c][lane] = a[c];
}

template <int VAR>
__global__ void bench(const double* in, double* out, long long* cyc) {
  __shared__ double A[NB][NB + 1];
  __shared__ double idg[NB];
  const int lane = threadIdx.x;
  for (int c = 0; c < NB; c++) A[lane][c] = in[lane * NB + c];
  __syncwarp();
  const long long t0 = clock64();
  if (VAR >= 4) potrf32_pair<VAR>(A, idg, lane); else potrf32<VAR>(A, idg, lane);
  __syncwarp();
  const long long t1 = clock64();
  if (lane == 0) cyc[VAR] = t1 - t0;
  for (int c = 0; c < NB; c++) out[lane * NB + c] = (c < lane) ? A[c][lane] : (c == lane ? 1.0 / idg[c] : 0.0);
}

int main() {
  double h[NB * NB], L[NB * NB];
  for (int i = 0; i < NB; i++)
    for (int j = 0; j < NB; j++) h[i * NB + j] = (i == j ? NB + 1.0 : 0.0) + 1.0 / (1.0 + i + j);
  double *din, *dout;
  long long* dc;
  cudaMalloc(&din, sizeof(h));
  cudaMalloc(&dout, sizeof(h));
  cudaMalloc(&dc, 64);
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) {
    bench<0><<<1, 32>>>(din, dout, dc);
    bench<1><<<1, 32>>>(din, dout, dc);
    bench<2><<<1, 32>>>(din, dout, dc);
    bench<3><<<1, 32>>>(din, dout, dc);
    bench<4><<<1, 32>>>(din, dout, dc);
    bench<5><<<1, 32>>>(din, dout, dc);
  }
  cudaMemcpy(L, dout, sizeof(h), cudaMemcpyDeviceToHost);
  long long c[6];
  cudaMemcpy(c, dc, 48, cudaMemcpyDeviceToHost);
  double err = 0;
  for (int i = 0; i < NB; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0;
      for (int k = 0; k <= j; k++) s += L[i * NB + k] * L[j * NB + k];
      err = fmax(err, fabs(s - h[i * NB + j]));
    }
  printf("potrf32 cycles: rsqrt+branch %lld | select %lld | fast rsqrt %lld | + local next pivot %lld | 2x2 pivots drcp %lld | 2x2 fast %lld ; |LL^T - A| = %.3g (last variant)\n",
         c[0], c[1], c[2], c[3], c[4], c[5], err);
  return 0;
}
