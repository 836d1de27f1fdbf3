// ctx.cu — context lifecycle, error text, scratch buffers, integer-pipe
// micro-benchmark (roofline denominator for the matching kernels).
#include <stdarg.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace lorb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int Buf::reserve(size_t bytes) {
  if (bytes <= cap) return LORB_OK;
  size_t want = bytes + bytes / 4 + 256;
  void* np = nullptr;
  cudaError_t e = pinned ? cudaMallocHost(&np, want) : cudaMalloc(&np, want);
  if (e != cudaSuccess) {
    set_error("allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? LORB_ERR_NOMEM : LORB_ERR_CUDA;
  }
  release();
  p = np;
  cap = want;
  return LORB_OK;
}

void Buf::release() {
  if (p) {
    if (pinned)
      cudaFreeHost(p);
    else
      cudaFree(p);
  }
  p = nullptr;
  cap = 0;
}

void dist_destroy(lorb_ctx* ctx);  // dist.cu
void ba_cache_free(lorb_ctx* ctx);  // ba_local.cu
void orb_graph_free(lorb_ctx* ctx); // orb.cu

// ---------------------------------------------------------------- microbench
// kind 0: 8 x (XOR, POPC, add) per 256-bit pair; kind 1: carry-save body.
// Operands live in registers and change every iteration so nothing folds.
template <bool CSA>
__global__ void __launch_bounds__(256) popc_bench_kernel(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t q[4][8], t[8];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int i = 0; i < 8; i++) q[k][i] = seed * (threadIdx.x + 1) + 0x9E3779B9u * (8 * k + i + blockIdx.x);
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = seed ^ (0x85EBCA6Bu * (i + 1));
  uint32_t best[4] = {KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE};
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = t[i] * 1664525u + 1013904223u;  // 8 IMAD per 4 pairs
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t d = hamming256<CSA>(q[k], t);
      best[k] = min(best[k], make_key(d, (uint32_t)it));
    }
  }
  uint32_t r = min(min(best[0], best[1]), min(best[2], best[3]));
  if (r == 0x12345u) sink[0] = r;  // practically never; keeps the loop alive
}

// kind 0: 8 independent DFMA chains per thread; kind 1: 8 independent DMMA accumulators per warp.
template <int KIND>
__global__ void __launch_bounds__(256) fp64_bench_kernel(double* sink, int iters, double seed) {
  double acc[8][2];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    acc[k][0] = seed * (threadIdx.x + k);
    acc[k][1] = seed * (blockIdx.x + k + 1);
  }
  const double a = 1.0 + seed * 1e-9 * threadIdx.x, b = 1.0 - seed * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (KIND == 0) {
        acc[k][0] = fma(acc[k][0], a, b);
        acc[k][1] = fma(acc[k][1], b, a);
      } else {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[k][0]), "+d"(acc[k][1])
                     : "d"(a), "d"(b));
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) r += acc[k][0] + acc[k][1];
  if (r == 0.12345) sink[0] = r;
}

struct Prof {
  struct Ev { cudaEvent_t a, b; int slot; };
  std::vector<Ev> ev;
  size_t used = 0;
  double total[LORB_PROF_SLOTS] = {};
  long long count[LORB_PROF_SLOTS] = {};
};

void prof_begin(lorb_ctx* c, int slot) {
  if (!c->prof_on) return;
  Prof* P = static_cast<Prof*>(c->prof);
  if (P->used == P->ev.size()) {
    Prof::Ev e;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    P->ev.push_back(e);
  }
  P->ev[P->used].slot = slot;
  cudaEventRecord(P->ev[P->used].a, c->stream);
}

void prof_end(lorb_ctx* c, int slot) {
  if (!c->prof_on) return;
  Prof* P = static_cast<Prof*>(c->prof);
  if (P->used >= P->ev.size() || P->ev[P->used].slot != slot) return;
  cudaEventRecord(P->ev[P->used].b, c->stream);
  P->used++;
}

static void prof_collect(lorb_ctx* c) {
  Prof* P = static_cast<Prof*>(c->prof);
  if (!P) return;
  cudaStreamSynchronize(c->stream);
  for (size_t i = 0; i < P->used; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, P->ev[i].a, P->ev[i].b) == cudaSuccess) {
      P->total[P->ev[i].slot] += ms;
      P->count[P->ev[i].slot]++;
    }
  }
  P->used = 0;
}

}  // namespace lorb

using namespace lorb;

extern "C" {

int lorb_ctx_profile(lorb_ctx* c, int enable) {
  LORB_REQUIRE(c, "ctx");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  if (!c->prof) c->prof = new (std::nothrow) Prof();
  LORB_REQUIRE(c->prof, "profile state");
  Prof* P = static_cast<Prof*>(c->prof);
  prof_collect(c);
  if (enable) {
    for (int s = 0; s < LORB_PROF_SLOTS; s++) {
      P->total[s] = 0;
      P->count[s] = 0;
    }
  }
  c->prof_on = enable ? 1 : 0;
  return LORB_OK;
}

int lorb_ctx_profile_read(lorb_ctx* c, int slot, double* total_ms, long long* count) {
  LORB_REQUIRE(c && c->prof && total_ms && count, "ctx / outputs (call lorb_ctx_profile first)");
  LORB_REQUIRE(slot >= 0 && slot < LORB_PROF_SLOTS, "slot");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  prof_collect(c);
  Prof* P = static_cast<Prof*>(c->prof);
  *total_ms = P->total[slot];
  *count = P->count[slot];
  return LORB_OK;
}

int lorb_microbench_fp64(lorb_ctx* c, int kind, int iters, double* flop_per_s) {
  LORB_REQUIRE(c && flop_per_s, "ctx/out");
  LORB_REQUIRE(iters > 0 && (kind == 0 || kind == 1), "iters / kind");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  LORB_TRY(dev_reserve(c, 0, 256));
  const int grid = c->sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  LORB_CUDA_TRY(cudaEventCreate(&e0));
  LORB_CUDA_TRY(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    LORB_CUDA_TRY(cudaEventRecord(e0, c->stream));
    if (kind == 0)
      LORB_LAUNCH(c, fp64_bench_kernel<0>, grid, block, 0, c->d[0].as<double>(), iters, 1.000001);
    else
      LORB_LAUNCH(c, fp64_bench_kernel<1>, grid, block, 0, c->d[0].as<double>(), iters, 1.000001);
    LORB_CUDA_TRY(cudaEventRecord(e1, c->stream));
    LORB_CUDA_TRY(cudaEventSynchronize(e1));
  }
  float ms = 0;
  LORB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // kind 0: 16 DFMA (2 flop) per thread and iteration; kind 1: 8 DMMA (512 flop) per warp and iteration
  const double per_iter = kind == 0 ? (double)grid * block * 16.0 * 2.0 : (double)grid * (block / 32) * 8.0 * 512.0;
  *flop_per_s = per_iter * iters / (ms * 1e-3);
  return LORB_OK;
}

const char* lorb_last_error(void) { return g_err; }
const char* lorb_version(void) { return "lorb-b200 0.1 (sm_100a)"; }

int lorb_ctx_create(int device, lorb_ctx** out) {
  LORB_REQUIRE(out != nullptr, "out");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no usable CUDA device (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return LORB_ERR_CUDA;
  }
  LORB_REQUIRE(device >= 0 && device < n, "device ordinal out of range");
  LORB_CUDA_TRY(cudaSetDevice(device));
  lorb_ctx* c = new (std::nothrow) lorb_ctx();
  if (!c) return LORB_ERR_NOMEM;
  c->device = device;
  cudaDeviceProp prop;
  LORB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
              prop.minor);
    delete c;
    return LORB_ERR_CUDA;
  }
  LORB_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& b : c->h) b.pinned = true;
  *out = c;
  return LORB_OK;
}

int lorb_ctx_destroy(lorb_ctx* c) {
  if (!c) return LORB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  dist_destroy(c);
  ba_cache_free(c);
  orb_graph_free(c);
  for (auto& b : c->d) b.release();
  for (auto& b : c->h) b.release();
  c->bank.release();
  c->plan_pairs.release();
  c->plan_out.release();
  c->tc_img.release();
  c->tc_units.release();
  c->tc_keys.release();
  if (c->prof) {
    Prof* P = static_cast<Prof*>(c->prof);
    for (auto& e : P->ev) {
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
    delete P;
  }
  cudaStreamDestroy(c->stream);
  delete c;
  return LORB_OK;
}

int lorb_ctx_sync(lorb_ctx* c) {
  LORB_REQUIRE(c, "ctx");
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return LORB_OK;
}

void* lorb_ctx_stream(lorb_ctx* c) { return c ? (void*)c->stream : nullptr; }
long long lorb_ctx_launch_count(lorb_ctx* c) { return c ? c->launches : 0; }

void lorb_ba_default_options(lorb_ba_options* o) {
  if (!o) return;
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

int lorb_microbench_popc(lorb_ctx* c, int kind, int iters, double* words_per_s) {
  LORB_REQUIRE(c && words_per_s, "ctx/out");
  LORB_REQUIRE(iters > 0, "iters");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  LORB_TRY(dev_reserve(c, 0, 256));
  const int grid = c->sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  LORB_CUDA_TRY(cudaEventCreate(&e0));
  LORB_CUDA_TRY(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    LORB_CUDA_TRY(cudaEventRecord(e0, c->stream));
    if (kind == 0)
      LORB_LAUNCH(c, popc_bench_kernel<false>, grid, block, 0, c->d[0].as<uint32_t>(), iters, 12345u);
    else
      LORB_LAUNCH(c, popc_bench_kernel<true>, grid, block, 0, c->d[0].as<uint32_t>(), iters, 12345u);
    LORB_CUDA_TRY(cudaEventRecord(e1, c->stream));
    LORB_CUDA_TRY(cudaEventSynchronize(e1));
  }
  float ms = 0;
  LORB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double pairs = (double)grid * block * 4.0 * iters;
  *words_per_s = pairs * 8.0 / (ms * 1e-3);
  return LORB_OK;
}

}  // extern "C"
