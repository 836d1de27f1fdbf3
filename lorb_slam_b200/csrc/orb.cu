// orb.cu — ORBextractor (reference src/ORBextractor.cpp): pyramid, FAST per
// cell, quadtree selection (orb_quadtree_gpu.cuh), blur, orientation and descriptors, one CUDA graph
// per frame (see OrbPipeline::launch).  First the descriptor side: intensity-centroid orientation
// (IC_Angle :79-106) and the steered BRIEF descriptor (computeOrbDescriptor :110-149) for keypoints
// given in the coordinates of their pyramid level.
//
// One warp per keypoint.  Orientation: lane r owns patch row v = r - 15 (31 rows of the circular
// patch, half width umax[|v|]); the two moments are integer sums, so any summation order gives
// the reference's bits; cv::fastAtan2 is restated in fp32 with the reference's operation order
// and no contraction.  Descriptor: lane b owns output byte b = 8 comparisons = 16 rotated pattern
// points; the rotation uses sinf/cosf with the bits of the host's libm (libm_sincosf.cuh) and
// separate fp32 multiply / add (the reference is compiled without FMA), each coordinate rounded
// half-to-even as cvRound does.  The pattern (512 points, int8 pairs) sits in shared memory.
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <utility>
#include <vector>

#include "common.cuh"
#include "libm_sincosf.cuh"
#include "orb_quadtree_gpu.cuh"
#include "stereo_dev.cuh"

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


// IC_Angle (:79-106) by one warp: m_10 = sum u*I, m_01 = sum v*I over the radius-15 disc.
__device__ __forceinline__ float orb_ic_angle(const uint8_t* __restrict__ centre, int step, int lane,
                                              const int* s_umax) {
  int m01 = 0, m10 = 0;
  if (lane <= 2 * ORB_HALF_PATCH) {
    const int v = lane - ORB_HALF_PATCH;
    const int d = s_umax[v < 0 ? -v : v];
    const uint8_t* row = centre + (ptrdiff_t)v * step;
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
  return fast_atan2_cv((float)m01, (float)m10);
}

// computeOrbDescriptor (:110-149) by one warp: lane b computes byte b; lanes with (lane & 3) == 0
// return the 32-bit word made of their own and the next three lanes' bytes.
__device__ __forceinline__ uint32_t orb_brief_word(const uint8_t* __restrict__ img, int step, float angle, int lane,
                                                   const char2* s_pat) {
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  const float rad = __fmul_rn(angle, factorPI);
  const float a = lorb_libm::cosf_libm(rad), b = lorb_libm::sinf_libm(rad);
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
  const uint32_t b1 = __shfl_down_sync(0xffffffffu, val, 1), b2 = __shfl_down_sync(0xffffffffu, val, 2),
                 b3 = __shfl_down_sync(0xffffffffu, val, 3);
  return val | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

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
    angle = orb_ic_angle(L.raw[l] + centre, step, lane, s_umax);
    if (lane == 0) out_angle[i] = angle;
  }
  const uint32_t word = orb_brief_word(L.blur[l] + centre, step, angle, lane, s_pat);
  if ((lane & 3) == 0) out_desc[(size_t)i * 8 + (lane >> 2)] = word;
}

// The same two stages for the keypoints the device quadtree selected (orb_quadtree_gpu.cuh): warp i
// finds its level from the per-level counts, its keypoint from the level's chosen list, and writes
// the finished KeyPoint (level-0 coordinates, octave, angle, response) and descriptor row both to
// mapped pinned host memory (the caller's result) and to device arrays (the stereo matcher's input).
struct OrbSelDev {
  OrbLevelsDev LV;
  int n_levels;
  const uint32_t* keys[ORB_MAX_LEVELS];  // packed candidates of the level, candidate order
  const int* chosen[ORB_MAX_LEVELS];
  const int* count[ORB_MAX_LEVELS];
  float scale[ORB_MAX_LEVELS];
  const OrbTables* tab;
  int cap;  // capacity of the output arrays
  float *h_x, *h_y, *h_angle, *h_resp;
  int* h_oct;
  uint32_t* h_desc;
  float *d_x, *d_y;
  int* d_oct;
  uint32_t* d_desc;
};

__global__ void __launch_bounds__(256) orb_describe_selected_kernel(OrbSelDev Q) {
  __shared__ char2 s_pat[512];
  __shared__ int s_umax[ORB_HALF_PATCH + 1];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_pat[i] = Q.tab->pattern[i];
  if (threadIdx.x <= ORB_HALF_PATCH) s_umax[threadIdx.x] = Q.tab->umax[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= Q.cap) return;
  int l = 0, off = 0;
  for (; l < Q.n_levels; l++) {
    const int n = max(*Q.count[l], 0);
    if (i < off + n) break;
    off += n;
  }
  if (l == Q.n_levels) return;
  const uint32_t e = Q.keys[l][Q.chosen[l][i - off]];
  // pt = candidate + minBorder (:873-874), in level coordinates
  const float x = __fadd_rn((float)(e & 0xfff), (float)(ORB_EDGE - 3));
  const float y = __fadd_rn((float)((e >> 12) & 0xfff), (float)(ORB_EDGE - 3));
  const int step = Q.LV.w[l];
  const size_t centre = (size_t)__float2int_rn(y) * step + __float2int_rn(x);
  const float angle = orb_ic_angle(Q.LV.raw[l] + centre, step, lane, s_umax);
  const uint32_t word = orb_brief_word(Q.LV.blur[l] + centre, step, angle, lane, s_pat);
  if ((lane & 3) == 0) {
    Q.h_desc[(size_t)i * 8 + (lane >> 2)] = word;
    Q.d_desc[(size_t)i * 8 + (lane >> 2)] = word;
  }
  if (lane == 0) {
    // keypoint->pt *= scale for level != 0 (:1142-1147)
    const float sx = l ? __fmul_rn(x, Q.scale[l]) : x, sy = l ? __fmul_rn(y, Q.scale[l]) : y;
    Q.h_x[i] = sx;
    Q.h_y[i] = sy;
    Q.h_oct[i] = l;
    Q.h_angle[i] = angle;
    Q.h_resp[i] = (float)(e >> 24);
    Q.d_x[i] = sx;
    Q.d_y[i] = sy;
    Q.d_oct[i] = l;
  }
}

// Arithmetic pins (tests only): the device's restated libm sinf/cosf and cv::fastAtan2.
__global__ void orb_selftest_kernel(int n, const float* __restrict__ a, const float* __restrict__ b,
                                    float* __restrict__ o_sin, float* __restrict__ o_cos,
                                    float* __restrict__ o_atan2) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bool in_domain = fabsf(a[i]) < 120.0f;  // outside it the restatement traps by design
    o_sin[i] = in_domain ? lorb_libm::sinf_libm(a[i]) : nanf("");
    o_cos[i] = in_domain ? lorb_libm::cosf_libm(a[i]) : nanf("");
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


// =====================================================================================
// The detection side: pyramid, FAST per cell, blur (ORBextractor::operator() :1087-1151).
// =====================================================================================
//
// ComputePyramid (:1157-1184): level l = cv::resize(level l-1, INTER_LINEAR).  8-bit INTER_LINEAR
// is OpenCV's 11-bit fixed-point bilinear: coefficients round(2048*(1-f)), round(2048*f) from a
// float fraction f of (d+0.5)*scale-0.5, horizontal pass in int, vertical pass
// (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2.  One thread per output pixel; the source
// level (<= 300 KB) is read through L1/L2.
__global__ void __launch_bounds__(256)
    orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh,
                      double scale_x, double scale_y) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
  if (dx >= dw || dy >= dh) return;
  float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, scale_x), 0.5);
  int sx = __float2int_rd(fx);
  fx = __fsub_rn(fx, (float)sx);
  if (sx < 0) fx = 0.f, sx = 0;
  if (sx >= sw - 1) fx = 0.f, sx = sw - 1;
  float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale_y), 0.5);
  const int sy = __float2int_rd(fy);
  fy = __fsub_rn(fy, (float)sy);
  const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
  const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.f));
  const int sx1 = min(sx + 1, sw - 1);
  const uint8_t* r0 = src + (size_t)min(max(sy, 0), sh - 1) * sw;
  const uint8_t* r1 = src + (size_t)min(max(sy + 1, 0), sh - 1) * sw;
  const int h0 = r0[sx] * a0 + r0[sx1] * a1, h1 = r1[sx] * a0 + r1[sx1] * a1;
  dst[(size_t)dy * dw + dx] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

// Everything the multi-level kernels need to find their level from blockIdx.x.
struct OrbPlanDev {
  int n_levels;
  int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
  uint8_t* raw[ORB_MAX_LEVELS];
  uint8_t* blur[ORB_MAX_LEVELS];
  int tile_start[ORB_MAX_LEVELS + 1];  // blur tiles
  int tiles_x[ORB_MAX_LEVELS];
  int cell_start[ORB_MAX_LEVELS + 1];  // FAST cells
  int n_cols[ORB_MAX_LEVELS], w_cell[ORB_MAX_LEVELS], h_cell[ORB_MAX_LEVELS];
  int slot_cap;                        // candidate slots per cell
  int ini_th, min_th;
};

// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of an 8-bit image in OpenCV 4's fixed-point
// form: taps {18,34,48,56,48,34,18}/256 in both directions, exact integer accumulation, one
// rounding (+2^15)>>16.  64x16 output tile per CTA, separable through shared memory.
constexpr int BLUR_TX = 64, BLUR_TY = 16;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) orb_blur_kernel(OrbPlanDev P) {
  __shared__ uint8_t s_in[BLUR_TY + 6][BLUR_TX + 6 + 2];
  __shared__ uint16_t s_h[BLUR_TY + 6][BLUR_TX];
  int l = 0;
  while (l + 1 < P.n_levels && (int)blockIdx.x >= P.tile_start[l + 1]) ++l;
  const int t = blockIdx.x - P.tile_start[l];
  const int w = P.w[l], h = P.h[l];
  const int x0 = (t % P.tiles_x[l]) * BLUR_TX, y0 = (t / P.tiles_x[l]) * BLUR_TY;
  const uint8_t* __restrict__ src = P.raw[l];
  for (int p = threadIdx.x; p < (BLUR_TY + 6) * (BLUR_TX + 6); p += blockDim.x) {
    const int ty = p / (BLUR_TX + 6), tx = p % (BLUR_TX + 6);
    s_in[ty][tx] = src[(size_t)reflect101(y0 + ty - 3, h) * w + reflect101(x0 + tx - 3, w)];
  }
  __syncthreads();
  for (int p = threadIdx.x; p < (BLUR_TY + 6) * BLUR_TX; p += blockDim.x) {
    const int ty = p / BLUR_TX, tx = p % BLUR_TX;
    const uint8_t* r = &s_in[ty][tx];
    s_h[ty][tx] = (uint16_t)(18 * (r[0] + r[6]) + 34 * (r[1] + r[5]) + 48 * (r[2] + r[4]) + 56 * r[3]);
  }
  __syncthreads();
  uint8_t* __restrict__ dst = P.blur[l];
  for (int p = threadIdx.x; p < BLUR_TY * BLUR_TX; p += blockDim.x) {
    const int ty = p / BLUR_TX, tx = p % BLUR_TX;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) continue;
    const uint32_t v = 18u * (s_h[ty][tx] + s_h[ty + 6][tx]) + 34u * (s_h[ty + 1][tx] + s_h[ty + 5][tx]) +
                       48u * (s_h[ty + 2][tx] + s_h[ty + 4][tx]) + 56u * s_h[ty + 3][tx];
    dst[(size_t)y * w + x] = (uint8_t)((v + (1u << 15)) >> 16);
  }
}

// ComputeKeyPointsOctTree's detection loop (:799-880): the level is cut into ~30x30 cells, each
// cell image (cell + 6 px, clipped to 16 px inside the level) goes through cv::FAST(iniThFAST,
// nonmax) and, if that finds nothing, cv::FAST(minThFAST, nonmax).
//
// FAST-9/16 restated: with A = the largest, over the 16 arcs of 9 ring pixels and both
// polarities, of the smallest |centre - ring| on the arc, a pixel is a corner at threshold T iff
// A > T and its OpenCV score (cornerScore<16>) is A - 1, independent of T.  cv::FAST only scores
// pixels 3 px inside the image it is given and suppresses against the scores of that image, so a
// keypoint of a cell is: score >= T, strictly greater than its 8 neighbours' scores, all taken
// inside the cell interior.  One score map per cell therefore serves both thresholds.
//
// One CTA per cell: cell image in shared memory, A by a doubling sliding minimum over the ring,
// suppression flags, block vote for "iniThFAST found something", ordered (row-major, as cv::FAST
// emits) compaction into the cell's slot.  Candidates are packed score<<24 | y<<12 | x with x, y
// relative to the level's 16-px margin (what the reference hands to DistributeOctTree).  The slots
// and the per-cell counts stay in device memory: the quadtree kernel (orb_quadtree_gpu.cuh) flattens
// them into the reference's candidate order (cell row, cell column, row-major inside the cell).
constexpr int CELL_MAX = 68;  // cell image side: wCell + 6 < 60 + 6

__device__ __forceinline__ int arc9_min_max(const int (&d)[16]) {
  int m2[16], m4[16], best = -256;
#pragma unroll
  for (int k = 0; k < 16; k++) m2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) best = max(best, min(min(m4[k], m4[(k + 4) & 15]), d[(k + 8) & 15]));
  return best;
}

__global__ void __launch_bounds__(256)
    orb_fast_cells_kernel(OrbPlanDev P, int cell_base, uint32_t* __restrict__ slots, int* __restrict__ cell_count) {
  __shared__ uint8_t s_img[CELL_MAX * CELL_MAX];
  __shared__ int16_t s_score[CELL_MAX * CELL_MAX];
  __shared__ uint8_t s_keep[CELL_MAX * CELL_MAX];
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int cell = cell_base + blockIdx.x;  // cells are numbered level after level
  int l = 0;
  while (l + 1 < P.n_levels && cell >= P.cell_start[l + 1]) ++l;
  const int c = cell - P.cell_start[l];
  const int ci = c / P.n_cols[l], cj = c % P.n_cols[l];
  const int w = P.w[l], h = P.h[l];
  const int min_b = ORB_EDGE - 3, max_bx = w - ORB_EDGE + 3, max_by = h - ORB_EDGE + 3;
  const int ini_y = min_b + ci * P.h_cell[l], ini_x = min_b + cj * P.w_cell[l];
  const int max_y = min(ini_y + P.h_cell[l] + 6, max_by), max_x = min(ini_x + P.w_cell[l] + 6, max_bx);
  const int sw = max_x - ini_x, sh = max_y - ini_y;
  // the reference skips these cells (:832-833, :841-842); cv::FAST returns nothing below 7 px
  const bool active = !(ini_y >= max_by - 3 || ini_x >= max_bx - 6 || sw < 7 || sh < 7);
  if (threadIdx.x == 0) s_base = 0;
  if (active) {
  const uint8_t* __restrict__ img = P.raw[l];
  for (int p = threadIdx.x; p < sw * sh; p += blockDim.x) {
    const int y = p / sw, x = p - y * sw;
    s_img[y * CELL_MAX + x] = img[(size_t)(ini_y + y) * w + ini_x + x];
    s_score[y * CELL_MAX + x] = 0;
    s_keep[y * CELL_MAX + x] = 0;
  }
  __syncthreads();
  const int iw = sw - 6, ih = sh - 6, n_in = iw * ih;
  const int th_lo = min(P.ini_th, P.min_th);
  for (int q = threadIdx.x; q < n_in; q += blockDim.x) {
    const int y = q / iw + 3, x = q % iw + 3;
    const uint8_t* p = &s_img[y * CELL_MAX + x];
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * CELL_MAX];
    d[1] = v - p[3 * CELL_MAX + 1];
    d[2] = v - p[2 * CELL_MAX + 2];
    d[3] = v - p[CELL_MAX + 3];
    d[4] = v - p[3];
    d[5] = v - p[-CELL_MAX + 3];
    d[6] = v - p[-2 * CELL_MAX + 2];
    d[7] = v - p[-3 * CELL_MAX + 1];
    d[8] = v - p[-3 * CELL_MAX];
    d[9] = v - p[-3 * CELL_MAX - 1];
    d[10] = v - p[-2 * CELL_MAX - 2];
    d[11] = v - p[-CELL_MAX - 3];
    d[12] = v - p[-3];
    d[13] = v - p[CELL_MAX - 3];
    d[14] = v - p[2 * CELL_MAX - 2];
    d[15] = v - p[3 * CELL_MAX - 1];
    const int dark = arc9_min_max(d);  // ring darker than the centre
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = -d[k];
    const int a = max(dark, arc9_min_max(d));
    if (a > th_lo) s_score[y * CELL_MAX + x] = (int16_t)(a - 1);
  }
  __syncthreads();
  int any_ini = 0;
  for (int q = threadIdx.x; q < n_in; q += blockDim.x) {
    const int y = q / iw + 3, x = q % iw + 3;
    const int16_t* s = &s_score[y * CELL_MAX + x];
    const int v = s[0];
    const bool keep = v > 0 && v > s[-1] && v > s[1] && v > s[-CELL_MAX - 1] && v > s[-CELL_MAX] &&
                      v > s[-CELL_MAX + 1] && v > s[CELL_MAX - 1] && v > s[CELL_MAX] && v > s[CELL_MAX + 1];
    s_keep[y * CELL_MAX + x] = keep;
    any_ini |= keep && v >= P.ini_th;
  }
  const int th = __syncthreads_or(any_ini) ? P.ini_th : P.min_th;
  uint32_t* __restrict__ slot = slots + (size_t)cell * P.slot_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q0 = 0; q0 < n_in; q0 += blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool keep = false;
    int x = 0, y = 0, v = 0;
    if (q < n_in) {
      y = q / iw + 3;
      x = q % iw + 3;
      v = s_score[y * CELL_MAX + x];
      keep = s_keep[y * CELL_MAX + x] && v >= th;
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int off = s_base + __popc(ballot & ((1u << lane) - 1));
    for (int k = 0; k < warp; k++) off += s_warp[k];
    if (keep && off < P.slot_cap)
      slot[off] = ((uint32_t)v << 24) | ((uint32_t)(ini_y + y - min_b) << 12) | (uint32_t)(ini_x + x - min_b);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < 8; k++) tot += s_warp[k];
      s_base += tot;
    }
    __syncthreads();
  }
  }  // active
  if (threadIdx.x == 0) cell_count[cell] = active ? s_base : 0;
}

// ------------------------------------------------------------------ host: level plan + quadtree
struct OrbLevelPlan {
  int n_levels;
  float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS];
  int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
  int n_features[ORB_MAX_LEVELS];
  int n_cols[ORB_MAX_LEVELS], n_rows[ORB_MAX_LEVELS], w_cell[ORB_MAX_LEVELS], h_cell[ORB_MAX_LEVELS];
};

// The ORBextractor constructor (:412-461) and the level sizes of ComputePyramid (:1161-1163), in
// the reference's float arithmetic.
static int make_level_plan(const lorb_orb_params* p, int width, int height, OrbLevelPlan* L) {
  LORB_REQUIRE(p->nlevels >= 1 && p->nlevels <= ORB_MAX_LEVELS, "nlevels");
  LORB_REQUIRE(p->scale_factor > 1.0f, "scale_factor");
  LORB_REQUIRE(p->nfeatures >= 1, "nfeatures");
  LORB_REQUIRE(p->min_th_fast >= 1 && p->ini_th_fast >= 1 && p->min_th_fast <= 254 && p->ini_th_fast <= 254,
               "FAST thresholds");
  const int n = L->n_levels = p->nlevels;
  L->scale[0] = 1.0f;
  for (int i = 1; i < n; i++) L->scale[i] = L->scale[i - 1] * p->scale_factor;
  for (int i = 0; i < n; i++) L->inv_scale[i] = 1.0f / L->scale[i];
  const float factor = 1.0f / p->scale_factor;
  float per_scale = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
  int sum = 0;
  for (int l = 0; l < n - 1; l++) {
    L->n_features[l] = (int)lrintf(per_scale);
    sum += L->n_features[l];
    per_scale *= factor;
  }
  L->n_features[n - 1] = std::max(p->nfeatures - sum, 0);
  for (int l = 0; l < n; l++) {
    L->w[l] = (int)lrintf((float)width * L->inv_scale[l]);
    L->h[l] = (int)lrintf((float)height * L->inv_scale[l]);
    const float bw = (float)(L->w[l] - 2 * ORB_EDGE + 6), bh = (float)(L->h[l] - 2 * ORB_EDGE + 6);
    L->n_cols[l] = (int)(bw / 30.f);  // W = 30 (:803)
    L->n_rows[l] = (int)(bh / 30.f);
    LORB_REQUIRE(L->w[l] >= 7 && L->h[l] >= 7, "pyramid level smaller than the 7 x 7 blur kernel");
    if (bw < 30.f || bh < 30.f) {
      // Below one cell the reference computes nCols (or nRows) = 0, divides by it (:821-822) and then
      // loops over no cells: the level is still part of the pyramid but yields no keypoints.  Same here.
      L->n_cols[l] = L->n_rows[l] = 0;
      L->w_cell[l] = L->h_cell[l] = 1;
      continue;
    }
    L->w_cell[l] = (int)ceilf(bw / L->n_cols[l]);
    L->h_cell[l] = (int)ceilf(bh / L->n_rows[l]);
    LORB_REQUIRE(L->w[l] - 2 * ORB_EDGE + 6 < 4096 && L->h[l] - 2 * ORB_EDGE + 6 < 4096, "image larger than 4096 px");
  }
  return LORB_OK;
}

struct OrbGraphKey {  // everything the captured chain depends on
  OrbPlanDev P;
  const void* stage;
  int width, height, with_blur;
  int n_features[ORB_MAX_LEVELS];
  uint64_t pattern_hash;
};

struct OrbGraph {
  OrbGraphKey key;
  cudaGraphExec_t exec = nullptr;
  int n_kernels = 0;
};

void orb_graph_free(lorb_ctx* c) {
  for (auto& g : c->orb_graph)
    if (g) {
      OrbGraph* G = static_cast<OrbGraph*>(g);
      if (G->exec) cudaGraphExecDestroy(G->exec);
      delete G;
      g = nullptr;
    }
  for (auto& e : c->orb_ev)
    if (e) {
      cudaEventDestroy(e);
      e = nullptr;
    }
  if (c->orb_stream2) {
    cudaStreamDestroy(c->orb_stream2);
    c->orb_stream2 = nullptr;
  }
  if (c->orb_stream3) {
    cudaStreamDestroy(c->orb_stream3);
    c->orb_stream3 = nullptr;
  }
  if (c->orb_join) {
    cudaEventDestroy(c->orb_join);
    c->orb_join = nullptr;
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

// Shared body of lorb_orb_extract / lorb_orb_stages.
namespace {

struct ExtractOut {
  // stages (all optional)
  uint8_t* const* raw_levels = nullptr;   // [n_levels] host buffers w*h
  uint8_t* const* blur_levels = nullptr;  // [n_levels]
  int cand_cap = 0;
  float *cand_x = nullptr, *cand_y = nullptr, *cand_resp = nullptr;
  int* cand_level_start = nullptr;  // [n_levels + 1]
  // final keypoints
  int cap = 0;
  float *kx = nullptr, *ky = nullptr, *kangle = nullptr, *kresp = nullptr, *ksize = nullptr;
  int* koct = nullptr;
  uint8_t* desc = nullptr;
  int* n_out = nullptr;
  int* n_per_level = nullptr;
  int* level_w = nullptr;
  int* level_h = nullptr;
};

// LORB_ORB_TRACE=1: host timeline of a call on stderr (microseconds since the call started)
struct OrbTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  OrbTrace() : on(getenv("LORB_ORB_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what, int a = -1) const {
    if (!on) return;
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[orb %8.1f us] %s %d\n", us, what, a);
  }
};

// One image going through the extractor.  Two jobs (left, right) share the context.
struct OrbJob {
  const uint8_t* image = nullptr;
  int step = 0;
  ExtractOut O;
  int slot = 0;  // which cached graph / buffer half of the context this job uses
  // layout: offsets into the job's device area (d[3]) and pinned area (h[2])
  size_t dev_base = 0, pin_base = 0;
  size_t o_raw[ORB_MAX_LEVELS], o_blur[ORB_MAX_LEVELS], o_tab = 0, o_slots = 0, o_cnt = 0;
  size_t o_keys[ORB_MAX_LEVELS], o_qscr[ORB_MAX_LEVELS], o_chosen[ORB_MAX_LEVELS], o_count = 0;
  size_t o_dx = 0, o_dy = 0, o_doct = 0, o_ddesc = 0;
  size_t h_img = 0, h_count = 0, h_x = 0, h_y = 0, h_oct = 0, h_ang = 0, h_resp = 0, h_desc = 0;
  OrbPlanDev P;
  int n_total = 0;
  int n_level[ORB_MAX_LEVELS];
  // device views of the selected keypoints (valid after the chain has run)
  const float *d_sx = nullptr, *d_sy = nullptr;  // level-0 coordinates (KeyPoint::pt)
  const int* d_lvl = nullptr;
  const uint32_t* d_desc = nullptr;
};

struct OrbPipeline {
  OrbTrace tr;
  lorb_ctx* c;
  OrbLevelPlan L;
  int nl = 0, width = 0, height = 0, n_cells = 0, n_tiles = 0, slot_cap = 0, key_cap = 0;
  bool want_desc = false;
  const int* pattern = nullptr;
  int ini_th = 0, min_th = 0;
  int n_ini[ORB_MAX_LEVELS], out_cap[ORB_MAX_LEVELS], node_cap[ORB_MAX_LEVELS], lvl_key_cap[ORB_MAX_LEVELS];
  size_t dev_bytes = 0, pin_bytes = 0;

  int plan(lorb_ctx* ctx, int w, int h, const lorb_orb_params* prm, const int* pat, bool desc, int cap) {
    c = ctx;
    width = w;
    height = h;
    pattern = pat;
    want_desc = desc;
    LORB_TRY(make_level_plan(prm, w, h, &L));
    nl = L.n_levels;
    ini_th = prm->ini_th_fast;
    min_th = prm->min_th_fast;
    if (want_desc) {
      LORB_REQUIRE(pattern, "pattern");
      for (int k = 0; k < 512; k++)
        LORB_REQUIRE(pattern[2 * k] * pattern[2 * k] + pattern[2 * k + 1] * pattern[2 * k + 1] < ORB_EDGE * ORB_EDGE,
                     "pattern radius");
    }
    n_cells = n_tiles = slot_cap = key_cap = 0;
    for (int l = 0; l < nl; l++) {
      n_tiles += ((L.w[l] + BLUR_TX - 1) / BLUR_TX) * ((L.h[l] + BLUR_TY - 1) / BLUR_TY);
      n_cells += L.n_cols[l] * L.n_rows[l];
      LORB_REQUIRE(L.w_cell[l] + 6 <= CELL_MAX && L.h_cell[l] + 6 <= CELL_MAX, "cell size");
      // suppression leaves no two adjacent survivors: at most ceil(w/2)*ceil(h/2) per interior
      slot_cap = std::max(slot_cap, ((L.w_cell[l] + 1) / 2) * ((L.h_cell[l] + 1) / 2));
    }
    for (int l = 0; l < nl; l++) {
      const int bw = L.w[l] - 2 * ORB_EDGE + 6, bh = L.h[l] - 2 * ORB_EDGE + 6;
      n_ini[l] = L.n_cols[l] > 0 ? std::max(1, (int)roundf((float)bw / bh)) : 1;  // (a level without cells has no quadtree)
      out_cap[l] = qt_out_cap(L.n_features[l], n_ini[l]);
      lvl_key_cap[l] = L.n_cols[l] * L.n_rows[l] * slot_cap;
      node_cap[l] = std::max(qt_node_cap(lvl_key_cap[l], L.n_features[l], n_ini[l]), L.n_cols[l] * L.n_rows[l] + 64);
      key_cap += out_cap[l];
    }
    (void)cap;
    return LORB_OK;
  }

  // Offsets of job j inside the shared device / pinned buffers.
  void layout(OrbJob* J, int j) {
    OPacker dv;
    for (int l = 0; l < nl; l++) J->o_raw[l] = dv.add((size_t)L.w[l] * L.h[l]);
    for (int l = 0; l < nl; l++) J->o_blur[l] = dv.add((size_t)L.w[l] * L.h[l]);
    J->o_tab = dv.add(sizeof(OrbTables));
    J->o_slots = dv.add((size_t)n_cells * slot_cap * 4);
    J->o_cnt = dv.add((size_t)n_cells * 4);
    for (int l = 0; l < nl; l++) {
      J->o_keys[l] = dv.add((size_t)lvl_key_cap[l] * 4);
      J->o_qscr[l] = dv.add(qt_scratch_ints(lvl_key_cap[l], node_cap[l]) * 4);
      J->o_chosen[l] = dv.add((size_t)out_cap[l] * 4);
    }
    J->o_count = dv.add(ORB_MAX_LEVELS * 4);
    J->o_dx = dv.add((size_t)key_cap * 4);
    J->o_dy = dv.add((size_t)key_cap * 4);
    J->o_doct = dv.add((size_t)key_cap * 4);
    J->o_ddesc = dv.add((size_t)key_cap * 32);
    dev_bytes = dv.off;
    J->dev_base = (size_t)j * dev_bytes;
    OPacker hs;
    J->h_img = hs.add((size_t)width * height);
    J->h_count = hs.add(ORB_MAX_LEVELS * 4);
    J->h_x = hs.add((size_t)key_cap * 4);
    J->h_y = hs.add((size_t)key_cap * 4);
    J->h_oct = hs.add((size_t)key_cap * 4);
    J->h_ang = hs.add((size_t)key_cap * 4);
    J->h_resp = hs.add((size_t)key_cap * 4);
    J->h_desc = hs.add((size_t)key_cap * 32);
    pin_bytes = hs.off;
    J->pin_base = (size_t)j * pin_bytes;
    J->slot = j;
  }

  int reserve(int n_jobs) {
    LORB_TRY(dev_reserve(c, 3, dev_bytes * n_jobs));
    LORB_TRY(pin_reserve(c, 2, pin_bytes * n_jobs));
    if (!c->orb_stream2) {
      LORB_CUDA_TRY(cudaStreamCreateWithFlags(&c->orb_stream2, cudaStreamNonBlocking));
      for (auto& e : c->orb_ev) LORB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      LORB_CUDA_TRY(cudaStreamCreateWithFlags(&c->orb_stream3, cudaStreamNonBlocking));
      LORB_CUDA_TRY(cudaEventCreateWithFlags(&c->orb_join, cudaEventDisableTiming));
    }
    return LORB_OK;
  }

  // One frame = ONE CUDA-graph launch.  The graph (captured once per frame size, parameters and
  // buffers) holds the whole extractor:
  //     H2D frame
  //     FAST level 0 -----------------> quadtree level 0 -> blur of all levels   (side branch;
  //     resize x7 -> FAST levels 1..7 -> quadtree levels 1..7                     the blur waits for the pyramid)
  //     (join) orientation + descriptors of the selected keypoints, results written to pinned memory
  // Issued one by one these 13 operations cost the host more time than the GPU needs to run them,
  // and a quadtree on the host (80 us for level 0, sequential) would sit in the middle of the
  // chain; as a graph the host stages the frame, launches, and waits once.
  int launch(OrbJob* J, cudaStream_t run_on = nullptr) {
    if (!run_on) run_on = c->stream;
    OrbPlanDev& P = J->P;
    memset(&P, 0, sizeof(P));
    P.n_levels = nl;
    P.ini_th = ini_th;
    P.min_th = min_th;
    int cells = 0, tiles = 0;
    uint8_t* d = c->d[3].as<uint8_t>() + J->dev_base;
    uint8_t* hp = c->h[2].as<uint8_t>() + J->pin_base;
    for (int l = 0; l < nl; l++) {
      P.w[l] = L.w[l];
      P.h[l] = L.h[l];
      P.raw[l] = d + J->o_raw[l];
      P.blur[l] = d + J->o_blur[l];
      P.tile_start[l] = tiles;
      P.tiles_x[l] = (L.w[l] + BLUR_TX - 1) / BLUR_TX;
      tiles += P.tiles_x[l] * ((L.h[l] + BLUR_TY - 1) / BLUR_TY);
      P.cell_start[l] = cells;
      P.n_cols[l] = L.n_cols[l];
      P.w_cell[l] = L.w_cell[l];
      P.h_cell[l] = L.h_cell[l];
      cells += L.n_cols[l] * L.n_rows[l];
    }
    P.tile_start[nl] = tiles;
    P.cell_start[nl] = cells;
    P.slot_cap = slot_cap;

    OrbGraph*& G = reinterpret_cast<OrbGraph*&>(c->orb_graph[J->slot]);
    OrbGraphKey key;
    memset(&key, 0, sizeof(key));
    key.P = P;
    key.stage = hp;
    key.width = width;
    key.height = height;
    key.with_blur = 1;
    for (int l = 0; l < nl; l++) key.n_features[l] = L.n_features[l];
    uint64_t hsh = 1469598103934665603ull;
    if (pattern)
      for (int q = 0; q < 1024; q++) hsh = (hsh ^ (uint32_t)pattern[q]) * 1099511628211ull;
    key.pattern_hash = hsh;
    if (!G || memcmp(&G->key, &key, sizeof(key)) != 0) {
      if (G) {
        cudaGraphExecDestroy(G->exec);
        delete G;
        G = nullptr;
      }
      LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
      // descriptor tables live in the job's device area
      OrbTables t;
      memset(&t, 0, sizeof(t));
      if (pattern)
        for (int q = 0; q < 512; q++)
          t.pattern[q] = make_char2((signed char)pattern[2 * q], (signed char)pattern[2 * q + 1]);
      make_umax(t.umax);
      LORB_CUDA_TRY(cudaMemcpyAsync(d + J->o_tab, &t, sizeof(t), cudaMemcpyHostToDevice, c->stream));
      LORB_CUDA_TRY(cudaMemsetAsync(d + J->o_count, 0, ORB_MAX_LEVELS * 4, c->stream));
      LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));  // `t` is on this stack frame
      const long long launches_before = c->launches;
      LORB_CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = capture_chain(J, d, hp);
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
      if (rc != LORB_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      LORB_CUDA_TRY(ce);
      OrbGraph* ng = new OrbGraph();
      ng->key = key;
      ng->n_kernels = (int)(c->launches - launches_before);
      c->launches = launches_before;  // captured, not launched
      ce = cudaGraphInstantiate(&ng->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) {
        delete ng;
        LORB_CUDA_TRY(ce);
      }
      G = ng;
    }
    if (J->step == width) {
      memcpy(hp + J->h_img, J->image, (size_t)width * height);
    } else {
      for (int r = 0; r < height; r++) memcpy(hp + J->h_img + (size_t)r * width, J->image + (size_t)r * J->step, width);
    }
    tr.mark("frame staged");
    LORB_CUDA_TRY(cudaGraphLaunch(G->exec, run_on));
    c->launches += G->n_kernels;
    tr.mark("graph launched");
    J->d_sx = (const float*)(d + J->o_dx);
    J->d_sy = (const float*)(d + J->o_dy);
    J->d_lvl = (const int*)(d + J->o_doct);
    J->d_desc = (const uint32_t*)(d + J->o_ddesc);
    return LORB_OK;
  }

  QtLevel qt_level(const OrbJob* J, uint8_t* d, uint8_t* hp, int l) const {
    QtLevel q;
    memset(&q, 0, sizeof(q));
    q.width = L.w[l] - 2 * ORB_EDGE + 6;
    q.height = L.h[l] - 2 * ORB_EDGE + 6;
    q.n_features = L.n_features[l];
    q.scratch = (int*)(d + J->o_qscr[l]);
    q.node_cap = node_cap[l];
    q.slots = (const uint32_t*)(d + J->o_slots) + (size_t)J->P.cell_start[l] * slot_cap;
    q.cell_cnt = (const int*)(d + J->o_cnt) + J->P.cell_start[l];
    q.n_cells = J->P.cell_start[l + 1] - J->P.cell_start[l];
    q.slot_cap = slot_cap;
    q.keys_flat = (uint32_t*)(d + J->o_keys[l]);
    q.key_cap = lvl_key_cap[l];
    q.out_index = (int*)(d + J->o_chosen[l]);
    q.out_count = (int*)(d + J->o_count) + l;
    q.out_count_host = (int*)(hp + J->h_count) + l;
    return q;
  }

  int capture_chain(OrbJob* J, uint8_t* d, uint8_t* hp) {
    const OrbPlanDev& P = J->P;
    cudaStream_t s = c->stream, s2 = c->orb_stream2;
    LORB_CUDA_TRY(cudaMemcpyAsync(d + J->o_raw[0], hp + J->h_img, (size_t)width * height, cudaMemcpyHostToDevice, s));
    uint32_t* slots = (uint32_t*)(d + J->o_slots);
    int* cnt = (int*)(d + J->o_cnt);
    // level 0 holds 40 % of the candidates and has the longest quadtree: its FAST runs first and its
    // quadtree on a side branch, under the pyramid chain and the other levels
    if (P.cell_start[1] > 0) LORB_LAUNCH(c, orb_fast_cells_kernel, P.cell_start[1], 256, 0, P, 0, slots, cnt);
    LORB_CUDA_TRY(cudaEventRecord(c->orb_ev[0], s));
    LORB_CUDA_TRY(cudaStreamWaitEvent(s2, c->orb_ev[0], 0));
    {
      QtArgs A;
      memset(&A, 0, sizeof(A));
      A.lv[0] = qt_level(J, d, hp, 0);
      orb_quadtree_kernel<<<1, QT_THREADS, 0, s2>>>(A);
      c->launches++;
      LORB_CUDA_TRY(cudaGetLastError());
    }
    for (int l = 1; l < nl; l++) {
      // scale = 1 / (dsize / ssize) in double, as cv::resize derives it from the two sizes
      const double sx = 1. / ((double)L.w[l] / L.w[l - 1]), sy = 1. / ((double)L.h[l] / L.h[l - 1]);
      const dim3 blk(32, 8), grd((L.w[l] + 31) / 32, (L.h[l] + 7) / 8);
      LORB_LAUNCH(c, orb_resize_kernel, grd, blk, 0, P.raw[l - 1], L.w[l - 1], L.h[l - 1], P.raw[l], L.w[l], L.h[l],
                  sx, sy);
    }
    // the blur only needs the pyramid: it joins the side branch behind the level-0 quadtree and runs
    // under the detection of levels 1..7
    LORB_CUDA_TRY(cudaEventRecord(c->orb_ev[2], s));
    LORB_CUDA_TRY(cudaStreamWaitEvent(s2, c->orb_ev[2], 0));
    orb_blur_kernel<<<n_tiles, 256, 0, s2>>>(P);
    c->launches++;
    LORB_CUDA_TRY(cudaGetLastError());
    LORB_CUDA_TRY(cudaEventRecord(c->orb_ev[1], s2));
    if (nl > 1) {
      if (P.cell_start[nl] > P.cell_start[1])
        LORB_LAUNCH(c, orb_fast_cells_kernel, P.cell_start[nl] - P.cell_start[1], 256, 0, P, P.cell_start[1], slots, cnt);
      QtArgs A;
      memset(&A, 0, sizeof(A));
      for (int l = 1; l < nl; l++) A.lv[l - 1] = qt_level(J, d, hp, l);
      LORB_LAUNCH(c, orb_quadtree_kernel, nl - 1, QT_THREADS, 0, A);
    }
    LORB_CUDA_TRY(cudaStreamWaitEvent(s, c->orb_ev[1], 0));  // join the side branch
    OrbSelDev Q;
    memset(&Q, 0, sizeof(Q));
    for (int l = 0; l < nl; l++) {
      Q.LV.raw[l] = P.raw[l];
      Q.LV.blur[l] = P.blur[l];
      Q.LV.w[l] = L.w[l];
      Q.LV.h[l] = L.h[l];
      Q.keys[l] = (const uint32_t*)(d + J->o_keys[l]);
      Q.chosen[l] = (const int*)(d + J->o_chosen[l]);
      Q.count[l] = (const int*)(d + J->o_count) + l;
      Q.scale[l] = L.scale[l];
    }
    Q.n_levels = nl;
    Q.tab = (const OrbTables*)(d + J->o_tab);
    Q.cap = key_cap;
    Q.h_x = (float*)(hp + J->h_x);
    Q.h_y = (float*)(hp + J->h_y);
    Q.h_oct = (int*)(hp + J->h_oct);
    Q.h_angle = (float*)(hp + J->h_ang);
    Q.h_resp = (float*)(hp + J->h_resp);
    Q.h_desc = (uint32_t*)(hp + J->h_desc);
    Q.d_x = (float*)(d + J->o_dx);
    Q.d_y = (float*)(d + J->o_dy);
    Q.d_oct = (int*)(d + J->o_doct);
    Q.d_desc = (uint32_t*)(d + J->o_ddesc);
    LORB_LAUNCH(c, orb_describe_selected_kernel, (key_cap + 7) / 8, 256, 0, Q);
    return LORB_OK;
  }

  int queue_level_copies(OrbJob* J) {
    const ExtractOut& O = J->O;
    if (O.raw_levels || O.blur_levels)
      for (int l = 0; l < nl; l++) {
        if (O.raw_levels)
          LORB_CUDA_TRY(cudaMemcpyAsync(O.raw_levels[l], J->P.raw[l], (size_t)L.w[l] * L.h[l], cudaMemcpyDeviceToHost,
                                        c->stream));
        if (O.blur_levels)
          LORB_CUDA_TRY(cudaMemcpyAsync(O.blur_levels[l], J->P.blur[l], (size_t)L.w[l] * L.h[l],
                                        cudaMemcpyDeviceToHost, c->stream));
      }
    return LORB_OK;
  }

  // After the stream has been synchronised: counts, then the results into the caller's arrays.
  int collect(OrbJob* J) {
    const ExtractOut& O = J->O;
    const uint8_t* hp = c->h[2].as<uint8_t>() + J->pin_base;
    const int* cnt = (const int*)(hp + J->h_count);
    J->n_total = 0;
    for (int l = 0; l < nl; l++) {
      LORB_REQUIRE(cnt[l] >= 0 && cnt[l] <= out_cap[l], "device quadtree ran out of node slots (internal)");
      J->n_level[l] = cnt[l];
      J->n_total += cnt[l];
    }
    if (O.cand_level_start) {  // vToDistributeKeys (:813), for lorb_orb_stages: fetch the slots
      std::vector<uint32_t> slots((size_t)n_cells * slot_cap);
      std::vector<int> ccnt(n_cells);
      const uint8_t* d = c->d[3].as<uint8_t>() + J->dev_base;
      LORB_CUDA_TRY(cudaMemcpy(slots.data(), d + J->o_slots, slots.size() * 4, cudaMemcpyDeviceToHost));
      LORB_CUDA_TRY(cudaMemcpy(ccnt.data(), d + J->o_cnt, ccnt.size() * 4, cudaMemcpyDeviceToHost));
      int total = 0;
      for (int l = 0; l < nl; l++) {
        O.cand_level_start[l] = total;
        for (int k = J->P.cell_start[l]; k < J->P.cell_start[l + 1]; k++) {
          LORB_REQUIRE(ccnt[k] >= 0 && ccnt[k] <= slot_cap, "candidate slot overflow (internal)");
          LORB_REQUIRE(total + ccnt[k] <= O.cand_cap, "candidate capacity");
          for (int e = 0; e < ccnt[k]; e++) {
            const uint32_t v = slots[(size_t)k * slot_cap + e];
            O.cand_x[total] = (float)(v & 0xfff);
            O.cand_y[total] = (float)((v >> 12) & 0xfff);
            O.cand_resp[total] = (float)(v >> 24);
            total++;
          }
        }
      }
      O.cand_level_start[nl] = total;
    }
    if (O.n_out) {
      *O.n_out = J->n_total;
      LORB_REQUIRE(J->n_total <= O.cap, "keypoint capacity (nfeatures + a few: the quadtree stops at >= N per level)");
    }
    if (O.kx || O.desc) {
      const int n = J->n_total;
      if (O.kx) memcpy(O.kx, hp + J->h_x, (size_t)n * 4);
      if (O.ky) memcpy(O.ky, hp + J->h_y, (size_t)n * 4);
      if (O.koct) memcpy(O.koct, hp + J->h_oct, (size_t)n * 4);
      if (O.kangle) memcpy(O.kangle, hp + J->h_ang, (size_t)n * 4);
      if (O.kresp) memcpy(O.kresp, hp + J->h_resp, (size_t)n * 4);
      if (O.desc) memcpy(O.desc, hp + J->h_desc, (size_t)n * 32);
      if (O.ksize) {
        int k = 0;
        for (int l = 0; l < nl; l++) {
          const float patch = (float)(int)(31 * L.scale[l]);  // scaledPatchSize = PATCH_SIZE*mvScaleFactor (:868)
          for (int i = 0; i < J->n_level[l]; i++) O.ksize[k++] = patch;
        }
      }
    }
    if (O.n_per_level)
      for (int l = 0; l < nl; l++) O.n_per_level[l] = J->n_level[l];
    if (O.level_w)
      for (int l = 0; l < nl; l++) O.level_w[l] = L.w[l], O.level_h[l] = L.h[l];
    return LORB_OK;
  }
};

int orb_run(lorb_ctx* c, const uint8_t* image, int width, int height, int step, const lorb_orb_params* prm,
            const int* pattern, const ExtractOut& O) {
  LORB_REQUIRE(c && image && prm, "ctx / image / params");
  LORB_REQUIRE(width > 0 && height > 0 && step >= width, "image shape");
  OrbPipeline pl;
  LORB_TRY(pl.plan(c, width, height, prm, pattern, O.desc != nullptr, O.cap));
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  OrbJob J;
  J.image = image;
  J.step = step;
  J.O = O;
  pl.layout(&J, 0);
  LORB_TRY(pl.reserve(1));
  LORB_TRY(pl.launch(&J));
  LORB_TRY(pl.queue_level_copies(&J));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  pl.tr.mark("stream drained");
  LORB_TRY(pl.collect(&J));
  pl.tr.mark("results scattered");
  return LORB_OK;
}

}  // namespace

// ORBextractor::DistributeOctTree (:554-797) alone, on the device (orb_quadtree_gpu.cuh): one CTA.
// Coordinates must be integers in [0, 4096), responses integers in [0, 256) (what the FAST kernel
// produces).
int lorb_orb_distribute(lorb_ctx* c, int n_keys, const float* x, const float* y, const float* response, int min_x,
                            int max_x, int min_y, int max_y, int n_features, int* out_index, int* n_out) {
  LORB_REQUIRE(c && n_keys >= 0 && n_out && (n_keys == 0 || (x && y && response && out_index)), "arguments");
  LORB_REQUIRE(max_x > min_x && max_y > min_y && n_features >= 0, "bounds");
  *n_out = 0;
  if (n_keys == 0) return LORB_OK;
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  const int n_ini = std::max(1, (int)roundf((float)(max_x - min_x) / (max_y - min_y)));
  const int node_cap = qt_node_cap(n_keys, n_features, n_ini);
  OPacker pk;
  const size_t o_keys = pk.add((size_t)n_keys * 4), o_scr = pk.add(qt_scratch_ints(n_keys, node_cap) * 4),
               o_out = pk.add((size_t)qt_out_cap(n_features, n_ini) * 4), o_cnt = pk.add(4);
  LORB_TRY(dev_reserve(c, 5, pk.off));
  LORB_TRY(pin_reserve(c, 4, std::max((size_t)n_keys * 4, (size_t)qt_out_cap(n_features, n_ini) * 4 + 512)));
  uint32_t* hk = c->h[4].as<uint32_t>();
  for (int i = 0; i < n_keys; i++) {
    LORB_REQUIRE(x[i] >= 0 && x[i] < 4096 && y[i] >= 0 && y[i] < 4096 && response[i] >= 0 && response[i] < 256 &&
                     x[i] == (float)(int)x[i] && y[i] == (float)(int)y[i] && response[i] == (float)(int)response[i],
                 "keys must be integer pixels below 4096 with integer responses below 256");
    hk[i] = ((uint32_t)response[i] << 24) | ((uint32_t)y[i] << 12) | (uint32_t)x[i];
  }
  uint8_t* d = c->d[5].as<uint8_t>();
  LORB_CUDA_TRY(cudaMemcpyAsync(d + o_keys, hk, (size_t)n_keys * 4, cudaMemcpyHostToDevice, c->stream));
  QtArgs A;
  memset(&A, 0, sizeof(A));
  QtLevel& L = A.lv[0];
  L.keys = (const uint32_t*)(d + o_keys);
  L.n_keys = n_keys;
  L.key_cap = n_keys;
  L.width = max_x - min_x;
  L.height = max_y - min_y;
  L.n_features = n_features;
  L.scratch = (int*)(d + o_scr);
  L.node_cap = node_cap;
  L.out_index = (int*)(d + o_out);
  L.out_count = (int*)(d + o_cnt);
  LORB_LAUNCH(c, orb_quadtree_kernel, 1, QT_THREADS, 0, A);
  int* ho = c->h[4].as<int>();
  LORB_CUDA_TRY(cudaMemcpyAsync(ho, d + o_out, o_cnt + 4 - o_out, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int n = *(const int*)((const uint8_t*)ho + (o_cnt - o_out));
  LORB_REQUIRE(n >= 0 && n <= qt_out_cap(n_features, n_ini), "device quadtree ran out of node slots (internal)");
  memcpy(out_index, ho, (size_t)n * 4);
  *n_out = n;
  return LORB_OK;
}

// Upper bound of the number of keypoints ORBextractor::operator() can return for this geometry:
// per level max(N_l + 3, 4 * nIni_l) -- the careful phase stops at the first split that reaches
// N_l, but the first quadtree pass is unconditional and may already yield 4 nodes per initial
// column (wide levels with small budgets return more than nfeatures in total).
int lorb_orb_max_keypoints(const lorb_orb_params* prm, int width, int height, int* max_keypoints) {
  LORB_REQUIRE(prm && max_keypoints, "arguments");
  OrbLevelPlan L;
  LORB_TRY(make_level_plan(prm, width, height, &L));
  int total = 0;
  for (int l = 0; l < L.n_levels; l++) {
    const int bw = L.w[l] - 2 * ORB_EDGE + 6, bh = L.h[l] - 2 * ORB_EDGE + 6;
    if (L.n_cols[l] > 0) total += qt_out_cap(L.n_features[l], std::max(1, (int)roundf((float)bw / bh)));
  }
  *max_keypoints = total;
  return LORB_OK;
}

int lorb_orb_level_sizes(const lorb_orb_params* prm, int width, int height, int* level_w, int* level_h,
                         int* n_features_per_level, float* scale_factors) {
  LORB_REQUIRE(prm && level_w && level_h, "arguments");
  OrbLevelPlan L;
  LORB_TRY(make_level_plan(prm, width, height, &L));
  for (int l = 0; l < L.n_levels; l++) {
    level_w[l] = L.w[l];
    level_h[l] = L.h[l];
    if (n_features_per_level) n_features_per_level[l] = L.n_features[l];
    if (scale_factors) scale_factors[l] = L.scale[l];
  }
  return LORB_OK;
}

int lorb_orb_extract(lorb_ctx* c, const uint8_t* image, int width, int height, int step,
                     const lorb_orb_params* prm, const int* pattern, int cap, float* kp_x, float* kp_y,
                     int* kp_octave, float* kp_angle, float* kp_response, float* kp_size, uint8_t* desc,
                     int* n_out, uint8_t* const* raw_levels) {
  LORB_REQUIRE(cap >= 0 && kp_x && kp_y && kp_octave && kp_angle && desc && n_out, "output arrays");
  ExtractOut O;
  O.raw_levels = raw_levels;
  O.cap = cap;
  O.kx = kp_x;
  O.ky = kp_y;
  O.koct = kp_octave;
  O.kangle = kp_angle;
  O.kresp = kp_response;
  O.ksize = kp_size;
  O.desc = desc;
  O.n_out = n_out;
  return orb_run(c, image, width, height, step, prm, pattern, O);
}

// The device work of Frame::Frame(imgLeft, imgRight, camera) (reference src/frame.cpp:17-68):
// ORBextractor on both images (:42-46, two host threads in the reference) and
// ComputeStereoMatches (:49).  Both extractions share the stream: the right image's pyramid / FAST
// run while the host distributes the left image's keypoints; the stereo matcher then reads the
// pyramids, keypoints and descriptors where they already are -- on the device.
int lorb_stereo_frame(lorb_ctx* c, const uint8_t* left, const uint8_t* right, int width, int height,
                      int step_left, int step_right, const lorb_orb_params* prm, const int* pattern, float mbf,
                      float mb, int cap, lorb_orb_keypoints* out_left, lorb_orb_keypoints* out_right,
                      float* out_uright, float* out_depth, int* n_matched) {
  LORB_REQUIRE(c && left && right && prm && out_left && out_right, "ctx / images / params / outputs");
  LORB_REQUIRE(width > 0 && height > 0 && step_left >= width && step_right >= width, "image shape");
  LORB_REQUIRE(mb > 0.0f, "baseline");
  LORB_REQUIRE(out_uright && out_depth, "stereo outputs");
  lorb_orb_keypoints* outs[2] = {out_left, out_right};
  for (int j = 0; j < 2; j++)
    LORB_REQUIRE(outs[j]->x && outs[j]->y && outs[j]->octave && outs[j]->angle && outs[j]->desc, "keypoint arrays");
  OrbPipeline pl;
  LORB_TRY(pl.plan(c, width, height, prm, pattern, true, cap));
  LORB_REQUIRE(pl.nl <= STEREO_MAX_LEVELS, "levels");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  OrbJob J[2];
  for (int j = 0; j < 2; j++) {
    J[j].image = j ? right : left;
    J[j].step = j ? step_right : step_left;
    ExtractOut& O = J[j].O;
    O.cap = cap;
    O.kx = outs[j]->x;
    O.ky = outs[j]->y;
    O.koct = outs[j]->octave;
    O.kangle = outs[j]->angle;
    O.kresp = outs[j]->response;
    O.ksize = outs[j]->size;
    O.desc = outs[j]->desc;
    O.n_out = &outs[j]->n;
    O.raw_levels = outs[j]->raw_levels;
    pl.layout(&J[j], j);
  }
  LORB_TRY(pl.reserve(2));
  // the two frames are independent: their graphs run concurrently on two streams (the reference
  // uses two host threads); the host needs the two keypoint counts to size the stereo launch, so it
  // waits once here and once at the end
  LORB_TRY(pl.launch(&J[0]));
  LORB_TRY(pl.launch(&J[1], c->orb_stream3));
  LORB_CUDA_TRY(cudaEventRecord(c->orb_join, c->orb_stream3));
  LORB_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->orb_join, 0));
  LORB_TRY(pl.queue_level_copies(&J[0]));
  LORB_TRY(pl.queue_level_copies(&J[1]));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  LORB_TRY(pl.collect(&J[0]));
  LORB_TRY(pl.collect(&J[1]));
  const int n_left = J[0].n_total, n_right = J[1].n_total;
  if (n_matched) *n_matched = 0;
  size_t o_ur = 0, o_dp = 0, o_n = 0;
  if (n_left > 0) {
    LORB_REQUIRE((unsigned)n_right < KEY_IDX_MASK, "keypoint counts");
    OPacker so;
    o_ur = so.add((size_t)n_left * 4);
    o_dp = so.add((size_t)n_left * 4);
    o_n = so.add(4);
    const size_t o_sad = so.add((size_t)n_left * 4);
    LORB_TRY(dev_reserve(c, 4, so.off));
    LORB_TRY(pin_reserve(c, 3, so.off));
    StereoDev S;
    memset(&S, 0, sizeof(S));
    for (int l = 0; l < pl.nl; l++) {
      S.left.lvl[l] = J[0].P.raw[l];
      S.right.lvl[l] = J[1].P.raw[l];
      S.left.w[l] = S.right.w[l] = pl.L.w[l];
      S.left.h[l] = S.right.h[l] = pl.L.h[l];
      S.sf[l] = pl.L.scale[l];
      S.inv_sf[l] = pl.L.inv_scale[l];
    }
    S.n_left = n_left;
    S.n_right = n_right;
    S.n_levels = pl.nl;
    S.n_rows = pl.L.h[0];  // nRows = mvImagePyramid[0].rows (src/frame.cpp:132)
    S.lx = J[0].d_sx;
    S.ly = J[0].d_sy;
    S.loct = J[0].d_lvl;
    S.ldesc = reinterpret_cast<const uint4*>(J[0].d_desc);
    S.rx = J[1].d_sx;
    S.ry = J[1].d_sy;
    S.roct = J[1].d_lvl;
    S.rdesc = reinterpret_cast<const uint4*>(J[1].d_desc);
    S.mbf = mbf;
    S.mb = mb;
    uint8_t* ds = c->d[4].as<uint8_t>();
    LORB_TRY(stereo_launch(c, S, (float*)(ds + o_ur), (float*)(ds + o_dp), (int*)(ds + o_sad), (int*)(ds + o_n)));
    LORB_CUDA_TRY(cudaMemcpyAsync(c->h[3].as<uint8_t>(), ds, o_sad, cudaMemcpyDeviceToHost, c->stream));
    LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    const uint8_t* hs = c->h[3].as<uint8_t>();
    memcpy(out_uright, hs + o_ur, (size_t)n_left * 4);
    memcpy(out_depth, hs + o_dp, (size_t)n_left * 4);
    if (n_matched) *n_matched = *(const int*)(hs + o_n);
  }
  return LORB_OK;
}

int lorb_orb_stages(lorb_ctx* c, const uint8_t* image, int width, int height, int step,
                    const lorb_orb_params* prm, uint8_t* const* raw_levels, uint8_t* const* blur_levels,
                    int cand_cap, float* cand_x, float* cand_y, float* cand_response, int* cand_level_start) {
  ExtractOut O;
  O.raw_levels = raw_levels;
  O.blur_levels = blur_levels;
  if (cand_level_start) {
    LORB_REQUIRE(cand_cap >= 0 && cand_x && cand_y && cand_response, "candidate arrays");
    O.cand_cap = cand_cap;
    O.cand_x = cand_x;
    O.cand_y = cand_y;
    O.cand_resp = cand_response;
    O.cand_level_start = cand_level_start;
  }
  return orb_run(c, image, width, height, step, prm, nullptr, O);
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
