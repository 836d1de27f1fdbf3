// orb.cu — the descriptor side of ORBextractor (reference src/ORBextractor.cpp; SURVEY 8(f)
// rank 5): intensity-centroid orientation (IC_Angle :79-106) and the steered BRIEF descriptor
// (computeOrbDescriptor :110-149) for keypoints given in the coordinates of their pyramid level.
//
// One warp per keypoint.  Orientation: lane r owns patch row v = r - 15 (31 rows of the circular
// patch, half width umax[|v|]); the two moments are integer sums, so any summation order gives
// the reference's bits; cv::fastAtan2 is restated in fp32 with the reference's operation order
// and no contraction.  Descriptor: lane b owns output byte b = 8 comparisons = 16 rotated pattern
// points; the rotation uses sinf/cosf with the bits of the host's libm (libm_sincosf.cuh) and
// separate fp32 multiply / add (the reference is compiled without FMA), each coordinate rounded
// half-to-even as cvRound does.  The pattern (512 points, int8 pairs) sits in shared memory.
#include "common.cuh"
#include "libm_sincosf.cuh"

namespace lorb {

constexpr int ORB_MAX_LEVELS = 16;
constexpr int ORB_HALF_PATCH = 15;  // HALF_PATCH_SIZE (:75)
constexpr int ORB_EDGE = 19;        // EDGE_THRESHOLD (:76)

struct OrbLevelsDev {
  const uint8_t* raw[ORB_MAX_LEVELS];   // tightly packed rows, stride = w
  const uint8_t* blur[ORB_MAX_LEVELS];
  int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
};

// cv::fastAtan2 (degrees, [0, 360)), scalar form: 7th-order odd polynomial on min/max, fp32,
// evaluated as (((p7*c2 + p5)*c2 + p3)*c2 + p1)*c without contraction.
__device__ __forceinline__ float fast_atan2_cv(float y, float x) {
  const float scale = (float)(180 / 3.1415926535897932384626433832795);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  const bool xs = ax >= ay;
  const float c = __fdiv_rn(xs ? ay : ax, __fadd_rn(xs ? ax : ay, eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fadd_rn(__fmul_rn(p7, c2), p5);
  a = __fadd_rn(__fmul_rn(a, c2), p3);
  a = __fadd_rn(__fmul_rn(a, c2), p1);
  a = __fmul_rn(a, c);
  if (!xs) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

struct OrbTables {
  char2 pattern[512];
  int umax[ORB_HALF_PATCH + 1];
};

__global__ void __launch_bounds__(256)
    orb_describe_kernel(OrbLevelsDev L, int n_kp, const float* __restrict__ kx, const float* __restrict__ ky,
                        const int* __restrict__ klevel, const OrbTables* __restrict__ tab,
                        float* __restrict__ out_angle, uint32_t* __restrict__ out_desc, int with_angle_in) {
  __shared__ char2 s_pat[512];
  __shared__ int s_umax[ORB_HALF_PATCH + 1];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_pat[i] = tab->pattern[i];
  if (threadIdx.x <= ORB_HALF_PATCH) s_umax[threadIdx.x] = tab->umax[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_kp) return;
  const int l = klevel[i];
  const int step = L.w[l];
  const int px = __float2int_rn(kx[i]), py = __float2int_rn(ky[i]);  // cvRound(pt.x), cvRound(pt.y)
  const size_t centre = (size_t)py * step + px;

  float angle;
  if (with_angle_in) {
    angle = out_angle[i];
  } else {
    // IC_Angle: m_10 = sum u*I, m_01 = sum v*I over the circular patch
    int m01 = 0, m10 = 0;
    if (lane <= 2 * ORB_HALF_PATCH) {
      const int v = lane - ORB_HALF_PATCH;
      const int d = s_umax[v < 0 ? -v : v];
      const uint8_t* row = L.raw[l] + centre + (ptrdiff_t)v * step;
      int s = 0;
      for (int u = -d; u <= d; ++u) {
        const int val = row[u];
        s += val;
        m10 += u * val;
      }
      m01 = v * s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    angle = fast_atan2_cv((float)m01, (float)m10);
    if (lane == 0) out_angle[i] = angle;
  }

  // computeOrbDescriptor
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  const float rad = __fmul_rn(angle, factorPI);
  const float a = lorb_libm::cosf_libm(rad), b = lorb_libm::sinf_libm(rad);
  const uint8_t* img = L.blur[l] + centre;
  uint32_t val = 0;
#pragma unroll
  for (int bit = 0; bit < 8; bit++) {
    const char2 p0 = s_pat[lane * 16 + 2 * bit], p1 = s_pat[lane * 16 + 2 * bit + 1];
    const float x0 = (float)p0.x, y0 = (float)p0.y, x1 = (float)p1.x, y1 = (float)p1.y;
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = img[r0 * step + c0], t1 = img[r1 * step + c1];
    val |= (uint32_t)(t0 < t1) << bit;
  }
  // byte `lane` of the 32-byte row: gather 4 lanes into one 32-bit store
  const uint32_t b1 = __shfl_down_sync(0xffffffffu, val, 1), b2 = __shfl_down_sync(0xffffffffu, val, 2),
                 b3 = __shfl_down_sync(0xffffffffu, val, 3);
  if ((lane & 3) == 0) out_desc[(size_t)i * 8 + (lane >> 2)] = val | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// Arithmetic pins (tests only): the device's restated libm sinf/cosf and cv::fastAtan2.
__global__ void orb_selftest_kernel(int n, const float* __restrict__ a, const float* __restrict__ b,
                                    float* __restrict__ o_sin, float* __restrict__ o_cos,
                                    float* __restrict__ o_atan2) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    o_sin[i] = lorb_libm::sinf_libm(a[i]);
    o_cos[i] = lorb_libm::cosf_libm(a[i]);
    o_atan2[i] = fast_atan2_cv(a[i], b[i]);
  }
}

struct OPacker {
  size_t off = 0;
  size_t add(size_t bytes) {
    const size_t o = off;
    off = (off + bytes + 255) & ~(size_t)255;
    return o;
  }
};

static int check_pyr(const lorb_pyramid_view* p, int n_levels) {
  LORB_REQUIRE(p && p->n_levels == n_levels && p->width && p->height && p->step && p->data, "pyramid view");
  for (int l = 0; l < n_levels; l++)
    LORB_REQUIRE(p->width[l] > 0 && p->height[l] > 0 && p->step[l] >= p->width[l] && p->data[l], "pyramid level");
  return LORB_OK;
}

static void pack_level(uint8_t* dst, const lorb_pyramid_view* p, int l) {
  const int w = p->width[l], h = p->height[l], st = p->step[l];
  if (st == w) {
    memcpy(dst, p->data[l], (size_t)w * h);
  } else {
    for (int r = 0; r < h; r++) memcpy(dst + (size_t)r * w, p->data[l] + (size_t)r * st, w);
  }
}

// umax of the ORBextractor constructor (:469-482): half widths of the rows of a radius-15 disc,
// the lower octant rounded from sqrt, the upper one filled so that the disc is symmetric.
static void make_umax(int* umax) {
  const int hp = ORB_HALF_PATCH;
  const int vmax = (int)floor(hp * sqrt(2.f) / 2 + 1), vmin = (int)ceil(hp * sqrt(2.f) / 2);
  for (int v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt((double)hp * hp - (double)v * v));
  for (int v = hp, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

}  // namespace lorb

using namespace lorb;

extern "C" {

void lorb_orb_umax(int* umax16) { make_umax(umax16); }

int lorb_orb_describe(lorb_ctx* c, const lorb_pyramid_view* raw, const lorb_pyramid_view* blurred, int n_levels,
                      const int* pattern, int n_kp, const float* kx, const float* ky, const int* klevel,
                      const float* angle_in, float* out_angle, uint8_t* out_desc) {
  LORB_REQUIRE(c, "ctx");
  LORB_REQUIRE(n_levels > 0 && n_levels <= ORB_MAX_LEVELS, "levels");
  LORB_TRY(check_pyr(blurred, n_levels));
  if (!angle_in) LORB_TRY(check_pyr(raw, n_levels));
  LORB_REQUIRE(pattern, "pattern");
  LORB_REQUIRE(n_kp >= 0, "keypoint count");
  if (n_kp == 0) return LORB_OK;
  LORB_REQUIRE(kx && ky && klevel && out_desc && (angle_in || out_angle), "keypoint arrays");
  for (int k = 0; k < 512; k++)  // a rotated pattern point must stay inside the EDGE_THRESHOLD margin
    LORB_REQUIRE(pattern[2 * k] * pattern[2 * k] + pattern[2 * k + 1] * pattern[2 * k + 1] < ORB_EDGE * ORB_EDGE,
                 "pattern radius");
  for (int l = 0; l < n_levels; l++)
    if (!angle_in)
      LORB_REQUIRE(raw->width[l] == blurred->width[l] && raw->height[l] == blurred->height[l], "pyramid shapes");
  for (int i = 0; i < n_kp; i++) {
    LORB_REQUIRE(klevel[i] >= 0 && klevel[i] < n_levels, "keypoint level");
    const int l = klevel[i];
    const long px = lrintf(kx[i]), py = lrintf(ky[i]);
    // the extractor only produces keypoints EDGE_THRESHOLD inside their level (:76, :820-823)
    LORB_REQUIRE(px >= ORB_EDGE && py >= ORB_EDGE && px < blurred->width[l] - ORB_EDGE &&
                     py < blurred->height[l] - ORB_EDGE,
                 "keypoint closer than EDGE_THRESHOLD to the border of its level");
  }
  LORB_CUDA_TRY(cudaSetDevice(c->device));

  OPacker in;
  size_t o_raw[ORB_MAX_LEVELS], o_blur[ORB_MAX_LEVELS];
  for (int l = 0; l < n_levels; l++) {
    o_raw[l] = angle_in ? 0 : in.add((size_t)raw->width[l] * raw->height[l]);
    o_blur[l] = in.add((size_t)blurred->width[l] * blurred->height[l]);
  }
  const size_t i_kx = in.add((size_t)n_kp * 4), i_ky = in.add((size_t)n_kp * 4), i_kl = in.add((size_t)n_kp * 4),
               i_tab = in.add(sizeof(OrbTables)), i_ang = in.add((size_t)n_kp * 4);
  OPacker out;
  const size_t o_ang = out.add((size_t)n_kp * 4), o_desc = out.add((size_t)n_kp * 32);
  LORB_TRY(pin_reserve(c, 0, in.off));
  LORB_TRY(pin_reserve(c, 1, out.off));
  LORB_TRY(dev_reserve(c, 0, in.off));
  LORB_TRY(dev_reserve(c, 2, out.off));
  uint8_t* h = c->h[0].as<uint8_t>();
  for (int l = 0; l < n_levels; l++) {
    if (!angle_in) pack_level(h + o_raw[l], raw, l);
    pack_level(h + o_blur[l], blurred, l);
  }
  memcpy(h + i_kx, kx, (size_t)n_kp * 4);
  memcpy(h + i_ky, ky, (size_t)n_kp * 4);
  memcpy(h + i_kl, klevel, (size_t)n_kp * 4);
  OrbTables* t = (OrbTables*)(h + i_tab);
  for (int k = 0; k < 512; k++) t->pattern[k] = make_char2((signed char)pattern[2 * k], (signed char)pattern[2 * k + 1]);
  make_umax(t->umax);
  uint8_t* d = c->d[0].as<uint8_t>();
  uint8_t* dout = c->d[2].as<uint8_t>();
  LORB_CUDA_TRY(cudaMemcpyAsync(d, h, i_ang, cudaMemcpyHostToDevice, c->stream));
  if (angle_in) {
    memcpy(c->h[1].as<uint8_t>() + o_ang, angle_in, (size_t)n_kp * 4);
    LORB_CUDA_TRY(cudaMemcpyAsync(dout + o_ang, c->h[1].as<uint8_t>() + o_ang, (size_t)n_kp * 4,
                                  cudaMemcpyHostToDevice, c->stream));
  }
  OrbLevelsDev L;
  for (int l = 0; l < n_levels; l++) {
    L.raw[l] = angle_in ? nullptr : d + o_raw[l];
    L.blur[l] = d + o_blur[l];
    L.w[l] = blurred->width[l];
    L.h[l] = blurred->height[l];
  }
  const int warps_per_cta = 8;
  LORB_LAUNCH(c, orb_describe_kernel, (n_kp + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, L, n_kp,
              (const float*)(d + i_kx), (const float*)(d + i_ky), (const int*)(d + i_kl),
              (const OrbTables*)(d + i_tab), (float*)(dout + o_ang), (uint32_t*)(dout + o_desc), angle_in ? 1 : 0);
  uint8_t* ho = c->h[1].as<uint8_t>();
  LORB_CUDA_TRY(cudaMemcpyAsync(ho, dout, out.off, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (out_angle) memcpy(out_angle, ho + o_ang, (size_t)n_kp * 4);
  memcpy(out_desc, ho + o_desc, (size_t)n_kp * 32);
  return LORB_OK;
}

int lorb_orb_selftest(lorb_ctx* c, int n, const float* a, const float* b, float* out_sin, float* out_cos,
                      float* out_atan2) {
  LORB_REQUIRE(c && n >= 0, "ctx / n");
  if (n == 0) return LORB_OK;
  LORB_REQUIRE(a && b && out_sin && out_cos && out_atan2, "arrays");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * 4;
  LORB_TRY(dev_reserve(c, 0, 2 * bytes));
  LORB_TRY(dev_reserve(c, 2, 3 * bytes));
  float* din = c->d[0].as<float>();
  float* dout = c->d[2].as<float>();
  LORB_CUDA_TRY(cudaMemcpyAsync(din, a, bytes, cudaMemcpyHostToDevice, c->stream));
  LORB_CUDA_TRY(cudaMemcpyAsync(din + n, b, bytes, cudaMemcpyHostToDevice, c->stream));
  LORB_LAUNCH(c, orb_selftest_kernel, c->sm_count * 8, 256, 0, n, din, din + n, dout, dout + n, dout + 2 * (size_t)n);
  LORB_CUDA_TRY(cudaMemcpyAsync(out_sin, dout, bytes, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaMemcpyAsync(out_cos, dout + n, bytes, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaMemcpyAsync(out_atan2, dout + 2 * (size_t)n, bytes, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return LORB_OK;
}

}  // extern "C"
