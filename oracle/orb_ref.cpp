// orb_ref.cpp — CPU oracle pieces for the ORBextractor stages.
// TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may
// load this; the product path (lorb_slam_b200/) never does.
//
// Part 1: pins of the device's sinf/cosf restatement.  computeOrbDescriptor (reference
// src/ORBextractor.cpp:116) calls the C library's cosf/sinf; lorb_slam_b200/csrc/libm_sincosf.cuh
// restates glibc's algorithm for the device.  Its host instantiation is compared here with the
// libm this process links, bit for bit.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../lorb_slam_b200/csrc/libm_sincosf.cuh"
#include "orb_quadtree_ref.h"

extern "C" {

void orc_libm_sincosf(int n, const float* x, float* s, float* c) {
  for (int i = 0; i < n; i++) {
    s[i] = sinf(x[i]);
    c[i] = cosf(x[i]);
  }
}

void orc_restated_sincosf(int n, const float* x, float* s, float* c) {
  for (int i = 0; i < n; i++) {
    s[i] = lorb_libm::sinf_libm(x[i]);
    c[i] = lorb_libm::cosf_libm(x[i]);
  }
}

// Number of floats with bit patterns lo, lo+stride, ... <= hi (and their negatives) on which the
// restatement and libm disagree in sinf or cosf.
long long orc_sincosf_mismatches(uint32_t lo, uint32_t hi, uint32_t stride) {
  long long bad = 0;
  const long long cnt = ((long long)hi - lo) / stride + 1;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (long long k = 0; k < cnt; k++) {
    const uint32_t bits = lo + (uint32_t)k * stride;
    for (int neg = 0; neg < 2; neg++) {
      const uint32_t b = bits | (neg ? 0x80000000u : 0u);
      float x;
      memcpy(&x, &b, 4);
      const float s = sinf(x), c = cosf(x), s2 = lorb_libm::sinf_libm(x), c2 = lorb_libm::cosf_libm(x);
      if (memcmp(&s, &s2, 4) || memcmp(&c, &c2, 4)) bad++;
    }
  }
  return bad;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Part 2: the OpenCV image operations ORBextractor calls (src/ORBextractor.cpp:851-858, 1132,
// 1172) restated for 8-bit single-channel images.  OpenCV is a third-party dependency that is not
// in /root/reference (CMakeLists.txt:10 asks for 3.1); these follow the published algorithms of
// OpenCV 4.x imgproc / features2d and are pinned bit-exactly against cv2 4.13
// (tests/golden/orb_golden.npz, tests/test_orb_cpu.py).  oracle/refshim's cv::resize /
// cv::GaussianBlur / cv::FAST stand-ins, which the compiled reference calls, are these functions.
#include <algorithm>
#include <vector>

namespace orc {

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR), CV_8UC1: 11-bit fixed-point coefficients from a
// float fraction, horizontal pass into 32-bit ints, vertical pass
// ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.
static inline short sat_short(float v) {
  const long r = lrintf(v);
  return (short)std::min<long>(std::max<long>(r, -32768), 32767);
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
  const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) fx = 0, sx = 0;
    if (sx >= sw - 1) fx = 0, sx = sw - 1;
    xofs[dx] = sx;
    ialpha[2 * dx] = sat_short((1.f - fx) * 2048);
    ialpha[2 * dx + 1] = sat_short(fx * 2048);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    yofs[dy] = sy;
    ibeta[2 * dy] = sat_short((1.f - fy) * 2048);
    ibeta[2 * dy + 1] = sat_short(fy * 2048);
  }
  for (int dy = 0; dy < dh; dy++) {
    const int sy0 = std::min(std::max(yofs[dy], 0), sh - 1), sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
    const uint8_t *r0 = src + (size_t)sy0 * sstep, *r1 = src + (size_t)sy1 * sstep;
    const int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
    for (int dx = 0; dx < dw; dx++) {
      const int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
      const int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
      const int h0 = r0[sx] * a0 + r0[sx1] * a1, h1 = r1[sx] * a0 + r1[sx1] * a1;
      dst[(size_t)dy * dstep + dx] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
    }
  }
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101), CV_8U: OpenCV 4's fixed-point
// path -- the 7-tap kernel in 8 fractional bits (error-diffused so that it sums to 256), exact
// integer accumulation over both passes, one rounding (+2^15) >> 16.
static const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};

static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

void gaussian7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
  std::vector<uint32_t> hor((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t s = 0;
      for (int k = 0; k < 7; k++) s += kGauss7[k] * src[(size_t)y * sstep + reflect101(x + k - 3, w)];
      hor[(size_t)y * w + x] = s;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t s = 0;
      for (int k = 0; k < 7; k++) s += kGauss7[k] * hor[(size_t)reflect101(y + k - 3, h) * w + x];
      dst[(size_t)y * dstep + x] = (uint8_t)((s + (1u << 15)) >> 16);
    }
}

// cv::FAST(img, keypoints, threshold, true) = FAST-9/16 with non-maximum suppression.
// A pixel at least 3 px inside the image is a corner at threshold T iff some 9 contiguous pixels of
// its 16-pixel Bresenham ring are all brighter than v+T or all darker than v-T; its score is the
// largest such T.  Equivalently, with A = max over the 16 arcs of 9 (both polarities) of the
// smallest |difference| on the arc: corner iff A > T, score = A - 1 (OpenCV's cornerScore<16>).
// With suppression, a corner survives iff its score is strictly greater than the scores of its 8
// neighbours (non-corners and the 3-px frame count as 0).  Output is row-major.
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int fast_arc_strength(const uint8_t* p, int step) {  // A of the comment above (<= 0: no arc)
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * step + kRingDx[k]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = 0;
  for (int s = 0; s < 16; s++) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; k++) {
      mn = std::min(mn, d[s + k]);
      mx = std::max(mx, d[s + k]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

int fast9_nms(const uint8_t* img, int w, int h, int step, int threshold, int* out_x, int* out_y, int* out_score,
              int cap) {
  if (w < 7 || h < 7) return 0;
  std::vector<int> score((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const int a = fast_arc_strength(img + (size_t)y * step + x, step);
      if (a > threshold) score[(size_t)y * w + x] = a - 1;
    }
  int n = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const int s = score[(size_t)y * w + x];
      if (s == 0 && !(threshold == 0 && fast_arc_strength(img + (size_t)y * step + x, step) > 0)) continue;
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; dy++)
        for (int dx = -1; dx <= 1; dx++)
          if ((dx || dy) && score[(size_t)(y + dy) * w + x + dx] >= s) {
            keep = false;
            break;
          }
      if (!keep) continue;
      if (n < cap) out_x[n] = x, out_y[n] = y, out_score[n] = s;
      n++;
    }
  return n;
}

// The detection loop of ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cpp:808-862) on
// one pyramid level: ~30x30 cells, cv::FAST(iniThFAST) on the cell image (cell + 6 px, clipped to
// the 16-px margin), cv::FAST(minThFAST) if that found nothing; keypoints relative to the margin.
int orb_level_candidates(const uint8_t* img, int cols, int rows, int step, int ini_th, int min_th, float* out_x,
                         float* out_y, float* out_resp, int cap) {
  const int EDGE = 19;
  const float W = 30;
  const int minBorderX = EDGE - 3, minBorderY = minBorderX, maxBorderX = cols - EDGE + 3, maxBorderY = rows - EDGE + 3;
  const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
  const int nCols = (int)(width / W), nRows = (int)(height / W);
  if (nCols < 1 || nRows < 1) return -1;
  const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
  std::vector<int> x(4096), y(4096), sc(4096);
  int n = 0;
  for (int i = 0; i < nRows; i++) {
    const float iniY = (float)(minBorderY + i * hCell);
    float maxY = iniY + hCell + 6;
    if (iniY >= maxBorderY - 3) continue;
    if (maxY > maxBorderY) maxY = (float)maxBorderY;
    for (int j = 0; j < nCols; j++) {
      const float iniX = (float)(minBorderX + j * wCell);
      float maxX = iniX + wCell + 6;
      if (iniX >= maxBorderX - 6) continue;
      if (maxX > maxBorderX) maxX = (float)maxBorderX;
      const uint8_t* sub = img + (size_t)(int)iniY * step + (int)iniX;
      const int sw = (int)maxX - (int)iniX, sh = (int)maxY - (int)iniY;
      int k = fast9_nms(sub, sw, sh, step, ini_th, x.data(), y.data(), sc.data(), 4096);
      if (k == 0) k = fast9_nms(sub, sw, sh, step, min_th, x.data(), y.data(), sc.data(), 4096);
      for (int q = 0; q < k; q++) {
        if (n < cap) {
          out_x[n] = (float)x[q] + j * wCell;
          out_y[n] = (float)y[q] + i * hCell;
          out_resp[n] = (float)sc[q];
        }
        n++;
      }
    }
  }
  return n;
}

}  // namespace orc

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
  orc::resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orc_gaussian7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
  orc::gaussian7_u8(src, w, h, sstep, dst, dstep);
}
int orc_orb_level_candidates(const uint8_t* img, int cols, int rows, int step, int ini_th, int min_th,
                             float* out_x, float* out_y, float* out_resp, int cap) {
  return orc::orb_level_candidates(img, cols, rows, step, ini_th, min_th, out_x, out_y, out_resp, cap);
}
int orc_fast9_nms(const uint8_t* img, int w, int h, int step, int threshold, int* out_x, int* out_y,
                  int* out_score, int cap) {
  return orc::fast9_nms(img, w, h, step, threshold, out_x, out_y, out_score, cap);
}

// Part 3: ORBextractor::DistributeOctTree (:554-797), see orb_quadtree_ref.h.  out_index receives the
// indices of the chosen keypoints in the reference's output order; returns their number.
int orc_orb_distribute(int n_keys, const float* x, const float* y, const float* response, int min_x, int max_x,
                       int min_y, int max_y, int n_features, int* out_index) {
  std::vector<orc::QKey> keys(n_keys);
  for (int i = 0; i < n_keys; i++) keys[i] = orc::QKey{x[i], y[i], response[i]};
  std::vector<int> chosen;
  orc::distribute_quadtree(keys, min_x, max_x, min_y, max_y, n_features, &chosen);
  for (size_t i = 0; i < chosen.size(); i++) out_index[i] = chosen[i];
  return (int)chosen.size();
}

}  // extern "C"
