// match_stereo.cu — Frame::ComputeStereoMatches (reference src/frame.cpp:125-333) on sm_100a:
// the other caller of Matcher::DescriptorDistance.  Runs once per frame.
//
// Reference algorithm, per left keypoint iL:
//   1. candidates = right keypoints whose row band [floor(y-r), ceil(y+r)], r = 2*scale[octave],
//      contains row (int)vL (:141-160), octave within +-1 (:209), uR in [uL-maxD, uL] (:214);
//      best = first strict minimum of the Hamming distance, starting from TH_HIGH (:190-226);
//   2. if best < (TH_HIGH+TH_LOW)/2: 11x11 SAD of centre-normalised patches at the keypoint's
//      pyramid level over 11 horizontal shifts (:233-279), parabola fit (:287-296), disparity
//      checks (:303-317) -> mvuRight, mvDepth, and the SAD kept for step 3;
//   3. matches whose SAD is >= 1.5*1.4*median(SAD) are removed (:323-337).
//
// GPU mapping: one warp per left keypoint.  Lanes stride over the right keypoints (row-band,
// octave and disparity gates, then the 256-bit Hamming distance against the left descriptor
// held in registers); best = min of (dist << 20 | iR), i.e. the reference's first-strict-minimum
// in ascending iR order.  The SAD is integer-exact: patch values are u8 - u8 centre, so
// sum |a - b| <= 121 * 510 fits an int and equals the reference's float/double L1 norm bit for
// bit; lanes stride over the 121 pixels with 11 accumulators (one per shift) and the warp reduces
// them with shuffles.  Lane 0 finishes the float post-processing with explicit _rn intrinsics
// (no FMA contraction; the reference is -O0 x86-64).  A second single-CTA kernel finds the
// (n/2)-th order statistic of the kept SADs by counting and applies the outlier threshold.
//
// Bounds: the reference relies on cv::Mat::rowRange/colRange assertions (an out-of-image patch
// throws); here such a keypoint is simply left unmatched (documented deviation: no exception).
#include "common.cuh"
#include "stereo_dev.cuh"

namespace lorb {

__global__ void __launch_bounds__(256)
    stereo_match_kernel(StereoDev S, float* __restrict__ out_uright, float* __restrict__ out_depth,
                        int* __restrict__ out_sad) {
  const int lane = threadIdx.x & 31;
  const int iL = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (iL >= S.n_left) return;
  if (lane == 0) {
    out_uright[iL] = -1.0f;  // :127-128
    out_depth[iL] = -1.0f;
    out_sad[iL] = -1;
  }
  const float uL = S.lx[iL], vL = S.ly[iL];
  const int levelL = S.loct[iL];
  const float minZ = S.mb, minD = 0.0f;
  const float maxD = __fdiv_rn(S.mbf, minZ);  // :164-166
  const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
  const int row = (int)vL;  // vRowIndices[vL] :185
  if (vL < 0.0f || row >= S.n_rows) return;
  if (maxU < 0.0f) return;  // :193

  uint32_t q[8];
  {
    const uint4 a = S.ldesc[2 * (size_t)iL], b = S.ldesc[2 * (size_t)iL + 1];
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
  }
  uint32_t best = make_key(LORB_TH_HIGH, 0);  // bestDist = TH_HIGH, bestIdxR = 0 (:196-197)
  for (int iR = lane; iR < S.n_right; iR += 32) {
    const float kpY = S.ry[iR];
    const int octR = S.roct[iR];
    const float r = __fmul_rn(2.0f, S.sf[octR]);                // :152
    const int maxr = (int)ceilf(__fadd_rn(kpY, r)), minr = (int)floorf(__fsub_rn(kpY, r));
    if (row < minr || row > maxr) continue;                     // :156-157
    if (octR < levelL - 1 || octR > levelL + 1) continue;       // :209
    const float uR = S.rx[iR];
    if (!(uR >= minU && uR <= maxU)) continue;                  // :214
    const uint4 a = S.rdesc[2 * (size_t)iR], b = S.rdesc[2 * (size_t)iR + 1];
    const uint32_t t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t d = hamming256_popc8(q, t);
    best = min(best, make_key(d, (uint32_t)iR));                // strict '<' in ascending iR (:219)
  }
  best = __reduce_min_sync(0xffffffffu, best);
  const int bestDist = (int)(best >> KEY_IDX_BITS);
  const int bestIdxR = (int)(best & KEY_IDX_MASK);
  const int thOrbDist = (LORB_TH_HIGH + LORB_TH_LOW) / 2;       // :130
  if (!(bestDist < thOrbDist)) return;                          // :231

  // ---- sub-pixel refinement by correlation (:233-279)
  const float uR0 = S.rx[bestIdxR];
  const float scaleFactor = S.inv_sf[levelL];
  const float scaleduL = roundf(__fmul_rn(uL, scaleFactor));
  const float scaledvL = roundf(__fmul_rn(vL, scaleFactor));
  const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
  const int wl = S.left.w[levelL], hl = S.left.h[levelL];
  const int wr = S.right.w[levelL], hr = S.right.h[levelL];
  const float iniu = scaleduR0 + (float)STEREO_L - (float)STEREO_W;        // :252
  const float endu = scaleduR0 + (float)STEREO_L + (float)STEREO_W + 1.0f;  // :253
  if (iniu < 0.0f || endu >= (float)wr) return;                             // :254
  const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
  // cv::Mat::rowRange / colRange would assert outside the image: leave such keypoints unmatched
  if (cv - STEREO_W < 0 || cv + STEREO_W >= hl || cv + STEREO_W >= hr) return;
  if (cu - STEREO_W < 0 || cu + STEREO_W >= wl) return;
  if (cr - STEREO_L - STEREO_W < 0 || cr + STEREO_L + STEREO_W >= wr) return;

  const uint8_t* IL = S.left.lvl[levelL];
  const uint8_t* IR = S.right.lvl[levelL];
  const int cL = IL[(size_t)cv * wl + cu];
  int acc[2 * STEREO_L + 1];
  int cR[2 * STEREO_L + 1];
#pragma unroll
  for (int s = 0; s <= 2 * STEREO_L; s++) {
    acc[s] = 0;
    cR[s] = IR[(size_t)cv * wr + cr + s - STEREO_L];
  }
  constexpr int PW = 2 * STEREO_W + 1;
  for (int p = lane; p < PW * PW; p += 32) {
    const int pr = p / PW, pc = p - pr * PW;
    const int a = (int)IL[(size_t)(cv - STEREO_W + pr) * wl + cu - STEREO_W + pc] - cL;
    const uint8_t* rowR = IR + (size_t)(cv - STEREO_W + pr) * wr + cr - STEREO_L - STEREO_W + pc;
#pragma unroll
    for (int s = 0; s <= 2 * STEREO_L; s++) acc[s] += abs(a - ((int)rowR[s] - cR[s]));
  }
#pragma unroll
  for (int s = 0; s <= 2 * STEREO_L; s++) acc[s] = __reduce_add_sync(0xffffffffu, acc[s]);
  if (lane != 0) return;

  int bestSad = 0x7fffffff, bestincR = 0;  // :243-244
#pragma unroll
  for (int s = 0; s <= 2 * STEREO_L; s++)
    if (acc[s] < bestSad) {  // float dist < int bestDist on exact integers (:266)
      bestSad = acc[s];
      bestincR = s - STEREO_L;
    }
  if (bestincR == -STEREO_L || bestincR == STEREO_L) return;  // :276
  float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
  for (int s = 1; s < 2 * STEREO_L; s++)
    if (s - STEREO_L == bestincR) {
      d1 = (float)acc[s - 1];
      d2 = (float)acc[s];
      d3 = (float)acc[s + 1];
    }
  // deltaR = (dist1-dist3)/(2.0f*(dist1+dist3-2.0f*dist2))   (:287)
  const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
  const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
  if (deltaR < -1.0f || deltaR > 1.0f) return;  // :290 (NaN passes, then fails the disparity test)
  float bestuR = __fmul_rn(S.sf[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestincR), deltaR));  // :297
  float disparity = __fsub_rn(uL, bestuR);                                                            // :300
  if (disparity >= minD && disparity < maxD) {  // :302
    if (disparity <= 0.0f) {
      disparity = 0.01f;                         // :306 (double literal stored to float)
      bestuR = (float)((double)uL - 0.01);      // :307
    }
    out_depth[iL] = __fdiv_rn(S.mbf, disparity);  // :311
    out_uright[iL] = bestuR;
    out_sad[iL] = bestSad;
  }
}

// :320-337: sort(vDistIdx); median = vDistIdx[size/2].first; thDist = 1.5f*1.4f*median; every
// match with SAD >= thDist is reset.  Sorting pairs (sad, iL) makes the (size/2)-th FIRST member
// the (size/2)-th order statistic of the SADs, found here by counting.
__global__ void __launch_bounds__(1024)
    stereo_finalize_kernel(int n_left, const int* __restrict__ sad, float* __restrict__ out_uright,
                           float* __restrict__ out_depth, int* __restrict__ n_matched) {
  __shared__ int s_n, s_median, s_kept;
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  if (tid == 0) { s_n = 0; s_median = -1; s_kept = 0; }
  __syncthreads();
  int local = 0;
  for (int i = tid; i < n_left; i += blockDim.x) local += sad[i] >= 0;
  local = __reduce_add_sync(0xffffffffu, local);
  if ((tid & 31) == 0) s_warp[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) n += s_warp[w];
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  if (n == 0) {
    if (tid == 0) *n_matched = 0;
    return;
  }
  const int k = n / 2;
  for (int i = tid; i < n_left; i += blockDim.x) {
    const int v = sad[i];
    if (v < 0) continue;
    int less = 0, leq = 0;
    for (int j = 0; j < n_left; j++) {
      const int u = sad[j];
      less += (u >= 0) & (u < v);
      leq += (u >= 0) & (u <= v);
    }
    if (less <= k && k < leq) s_median = v;  // every such i holds the same value
  }
  __syncthreads();
  const float median = (float)s_median;
  const float thDist = __fmul_rn(1.5f * 1.4f, median);  // 1.5f*1.4f folds in fp32 (:325)
  int kept = 0;
  for (int i = tid; i < n_left; i += blockDim.x) {
    const int v = sad[i];
    if (v < 0) continue;
    if ((float)v < thDist) {
      kept++;
    } else {  // :331-335
      out_uright[i] = -1.0f;
      out_depth[i] = -1.0f;
    }
  }
  atomicAdd(&s_kept, kept);
  __syncthreads();
  if (tid == 0) *n_matched = s_kept;
}

struct SPacker {
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

int stereo_launch(lorb_ctx* c, const StereoDev& S, float* d_uright, float* d_depth, int* d_sad, int* d_n_matched) {
  const int warps_per_cta = 8;
  LORB_LAUNCH(c, stereo_match_kernel, (S.n_left + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, S,
              d_uright, d_depth, d_sad);
  LORB_LAUNCH(c, stereo_finalize_kernel, 1, 1024, 0, S.n_left, (const int*)d_sad, d_uright, d_depth, d_n_matched);
  return LORB_OK;
}

}  // namespace lorb

using namespace lorb;

extern "C" {

int lorb_stereo_matches(lorb_ctx* c, const lorb_pyramid_view* left, const lorb_pyramid_view* right,
                        int n_levels, const float* scale_factors, const float* inv_scale_factors,
                        float mbf, float mb, int n_left, const float* lx, const float* ly,
                        const int* loct, const uint8_t* ldesc, int n_right, const float* rx,
                        const float* ry, const int* roct, const uint8_t* rdesc, float* out_uright,
                        float* out_depth, int* n_matched) {
  LORB_REQUIRE(c, "ctx");
  LORB_REQUIRE(n_levels > 0 && n_levels <= STEREO_MAX_LEVELS && scale_factors && inv_scale_factors, "levels");
  LORB_TRY(check_pyr(left, n_levels));
  LORB_TRY(check_pyr(right, n_levels));
  LORB_REQUIRE(n_left >= 0 && n_right >= 0 && (unsigned)n_right < KEY_IDX_MASK, "keypoint counts");
  LORB_REQUIRE(mb > 0.0f, "baseline");
  if (n_matched) *n_matched = 0;
  if (n_left == 0) return LORB_OK;
  LORB_REQUIRE(lx && ly && loct && ldesc && out_uright && out_depth, "left arrays");
  if (n_right > 0) LORB_REQUIRE(rx && ry && roct && rdesc, "right arrays");
  for (int i = 0; i < n_left; i++) LORB_REQUIRE(loct[i] >= 0 && loct[i] < n_levels, "left octave");
  for (int i = 0; i < n_right; i++) LORB_REQUIRE(roct[i] >= 0 && roct[i] < n_levels, "right octave");
  LORB_CUDA_TRY(cudaSetDevice(c->device));

  SPacker in, out;
  size_t o_img[2][STEREO_MAX_LEVELS];
  const lorb_pyramid_view* pv[2] = {left, right};
  for (int s = 0; s < 2; s++)
    for (int l = 0; l < n_levels; l++) o_img[s][l] = in.add((size_t)pv[s]->width[l] * pv[s]->height[l]);
  const size_t i_lx = in.add((size_t)n_left * 4), i_ly = in.add((size_t)n_left * 4),
               i_lo = in.add((size_t)n_left * 4), i_ld = in.add((size_t)n_left * 32),
               i_rx = in.add((size_t)n_right * 4), i_ry = in.add((size_t)n_right * 4),
               i_ro = in.add((size_t)n_right * 4), i_rd = in.add((size_t)n_right * 32);
  const size_t o_ur = out.add((size_t)n_left * 4), o_dp = out.add((size_t)n_left * 4), o_n = out.add(4);
  const size_t o_sad = out.add((size_t)n_left * 4);
  LORB_TRY(pin_reserve(c, 0, in.off));
  LORB_TRY(pin_reserve(c, 1, out.off));
  LORB_TRY(dev_reserve(c, 0, in.off));
  LORB_TRY(dev_reserve(c, 2, out.off));
  uint8_t* h = c->h[0].as<uint8_t>();
  for (int s = 0; s < 2; s++)
    for (int l = 0; l < n_levels; l++) {
      const int w = pv[s]->width[l], hh = pv[s]->height[l], st = pv[s]->step[l];
      uint8_t* dst = h + o_img[s][l];
      if (st == w) {
        memcpy(dst, pv[s]->data[l], (size_t)w * hh);
      } else {
        for (int r = 0; r < hh; r++) memcpy(dst + (size_t)r * w, pv[s]->data[l] + (size_t)r * st, w);
      }
    }
  memcpy(h + i_lx, lx, (size_t)n_left * 4);
  memcpy(h + i_ly, ly, (size_t)n_left * 4);
  memcpy(h + i_lo, loct, (size_t)n_left * 4);
  memcpy(h + i_ld, ldesc, (size_t)n_left * 32);
  if (n_right > 0) {
    memcpy(h + i_rx, rx, (size_t)n_right * 4);
    memcpy(h + i_ry, ry, (size_t)n_right * 4);
    memcpy(h + i_ro, roct, (size_t)n_right * 4);
    memcpy(h + i_rd, rdesc, (size_t)n_right * 32);
  }
  uint8_t* d = c->d[0].as<uint8_t>();
  uint8_t* dout = c->d[2].as<uint8_t>();
  LORB_CUDA_TRY(cudaMemcpyAsync(d, h, in.off, cudaMemcpyHostToDevice, c->stream));

  StereoDev S;
  for (int l = 0; l < n_levels; l++) {
    S.left.lvl[l] = d + o_img[0][l];
    S.left.w[l] = left->width[l];
    S.left.h[l] = left->height[l];
    S.right.lvl[l] = d + o_img[1][l];
    S.right.w[l] = right->width[l];
    S.right.h[l] = right->height[l];
    S.sf[l] = scale_factors[l];
    S.inv_sf[l] = inv_scale_factors[l];
  }
  S.n_left = n_left;
  S.n_right = n_right;
  S.n_levels = n_levels;
  S.n_rows = left->height[0];  // nRows = mvImagePyramid[0].rows (:132)
  S.lx = (const float*)(d + i_lx);
  S.ly = (const float*)(d + i_ly);
  S.loct = (const int*)(d + i_lo);
  S.ldesc = (const uint4*)(d + i_ld);
  S.rx = (const float*)(d + i_rx);
  S.ry = (const float*)(d + i_ry);
  S.roct = (const int*)(d + i_ro);
  S.rdesc = (const uint4*)(d + i_rd);
  S.mbf = mbf;
  S.mb = mb;
  LORB_TRY(stereo_launch(c, S, (float*)(dout + o_ur), (float*)(dout + o_dp), (int*)(dout + o_sad), (int*)(dout + o_n)));
  uint8_t* ho = c->h[1].as<uint8_t>();
  LORB_CUDA_TRY(cudaMemcpyAsync(ho, dout, o_sad, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  memcpy(out_uright, ho + o_ur, (size_t)n_left * 4);
  memcpy(out_depth, ho + o_dp, (size_t)n_left * 4);
  if (n_matched) *n_matched = *(const int*)(ho + o_n);
  return LORB_OK;
}

}  // extern "C"
