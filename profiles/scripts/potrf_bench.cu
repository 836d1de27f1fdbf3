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

template <int VAR>
__global__ void bench(const double* in, double* out, long long* cyc) {
  __shared__ double A[NB][NB + 1];
  __shared__ double idg[NB];
  const int lane = threadIdx.x;
  for (int c = 0; c < NB; c++) A[lane][c] = in[lane * NB + c];
  __syncwarp();
  const long long t0 = clock64();
  potrf32<VAR>(A, idg, lane);
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
  }
  cudaMemcpy(L, dout, sizeof(h), cudaMemcpyDeviceToHost);
  long long c[4];
  cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost);
  double err = 0;
  for (int i = 0; i < NB; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0;
      for (int k = 0; k <= j; k++) s += L[i * NB + k] * L[j * NB + k];
      err = fmax(err, fabs(s - h[i * NB + j]));
    }
  printf("potrf32 cycles: rsqrt+branch %lld | select %lld | fast rsqrt %lld | + local next pivot %lld ; |LL^T - A| = %.3g (last variant)\n",
         c[0], c[1], c[2], c[3], err);
  return 0;
}
