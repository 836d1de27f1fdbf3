// Second translation unit of oracle/_ref/libref.so: the reference's src/ORBextractor.cpp, textually
// included (unmodified, from where it lies) so that its file-static functions IC_Angle and
// computeOrbDescriptor -- the orientation and steered-BRIEF stages, SURVEY 8(f) rank 5 -- can be
// called on given pyramid levels and keypoints.  TEST INFRASTRUCTURE ONLY.
//
// The image-processing calls of the rest of that file (cv::resize, copyMakeBorder, GaussianBlur,
// FAST) go to the cv2-pinned stand-ins of ref_harness.cpp / orb_ref.cpp, so the whole
// ORBextractor::operator() runs as well (ref_orb_extract).
//
// DistributeOctTree (:554-797) sorts (size, ExtractorNode*) pairs (:695) and splits the largest
// nodes first: among nodes of equal size the order is the order of their ADDRESSES, i.e. it depends
// on the allocator.  To make the reference's output a function of its input, this library replaces
// operator new/delete (for itself only: -Bsymbolic) and, while ref_orb_extract runs, serves every
// allocation from a bump arena that never reuses memory: addresses then grow with creation order,
// so "larger address" = "created later".  That is the rule the product's quadtree restates.
#include <cstdio>
#include <cstdlib>
#include <new>

namespace lorb_arena {
static thread_local char* base = nullptr;
static thread_local size_t used = 0, cap = 0;
static thread_local bool on = false;
inline void* take(size_t n) {
  n = (n + 15) & ~(size_t)15;
  if (!on || used + n > cap) return nullptr;
  void* p = base + used;
  used += n;
  return p;
}
inline bool owns(void* p) { return base && (char*)p >= base && (char*)p < base + cap; }
}  // namespace lorb_arena

void* operator new(size_t n) {
  if (void* p = lorb_arena::take(n)) return p;
  if (lorb_arena::on) {
    std::fprintf(stderr, "oracle/_ref: bump arena exhausted\n");
    std::abort();
  }
  void* p = std::malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
  if (!lorb_arena::owns(p)) std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

#include "ORBextractor.cpp"  // -I$(REFERENCE)/src

#include <cstdint>
#include <cstring>

using namespace Simple_ORB_SLAM;

namespace {
cv::Mat level_mat(int w, int h, int step, const uint8_t* data) {
  cv::Mat m(h, w, CV_8U);
  for (int r = 0; r < h; r++) std::memcpy(m.ptr(r), data + (size_t)r * step, (size_t)w);
  return m;
}
}  // namespace

extern "C" {

// For keypoint i at (kx, ky) in the coordinates of pyramid level klevel[i]:
//   out_angle[i] = IC_Angle(raw level image)            (src/ORBextractor.cpp:79-106)
//   out_desc[i]  = computeOrbDescriptor(blurred level)  (:110-149), with that angle
// Keypoints must lie >= 19 px inside their level (EDGE_THRESHOLD), as the extractor guarantees.
int ref_orb_describe(int n_levels, const int* w, const int* h, const int* step_raw, const uint8_t* const* raw,
                     const int* step_blur, const uint8_t* const* blurred, int n_kp, const float* kx,
                     const float* ky, const int* klevel, float* out_angle, uint8_t* out_desc, int* umax_out,
                     int* pattern_out) {
  ORBextractor ex(1000, 1.2f, n_levels, 20, 7);
  std::vector<cv::Mat> R, B;
  for (int l = 0; l < n_levels; l++) {
    R.push_back(level_mat(w[l], h[l], step_raw[l], raw[l]));
    B.push_back(level_mat(w[l], h[l], step_blur[l], blurred[l]));
  }
  for (int i = 0; i < n_kp; i++) {
    cv::KeyPoint kp;
    kp.pt = cv::Point2f(kx[i], ky[i]);
    kp.octave = klevel[i];
    kp.angle = IC_Angle(R[klevel[i]], kp.pt, ex.umax);
    out_angle[i] = kp.angle;
    computeOrbDescriptor(kp, B[klevel[i]], &ex.pattern[0], out_desc + 32 * (size_t)i);
  }
  if (umax_out)
    for (int v = 0; v <= HALF_PATCH_SIZE; v++) umax_out[v] = ex.umax[v];
  if (pattern_out)
    for (int k = 0; k < 512; k++) {
      pattern_out[2 * k] = ex.pattern[k].x;
      pattern_out[2 * k + 1] = ex.pattern[k].y;
    }
  return 0;
}

// ORBextractor::DistributeOctTree (:554-797, a protected member: the library is compiled with
// -fno-access-control) on given candidates, under the bump arena.  out_x/out_y/out_resp receive
// the chosen keypoints in the reference's output order; returns their number.
int ref_distribute_octree(int n_keys, const float* x, const float* y, const float* resp, int min_x, int max_x,
                          int min_y, int max_y, int n_features, float* out_x, float* out_y, float* out_resp) {
  const size_t arena = (size_t)1 << 28;
  char* mem = (char*)std::malloc(arena);
  if (!mem) return -1;
  int n = -1;
  lorb_arena::base = mem;
  lorb_arena::used = 0;
  lorb_arena::cap = arena;
  lorb_arena::on = true;
  {
    ORBextractor ex(n_features, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> in(n_keys);
    for (int i = 0; i < n_keys; i++) {
      in[i].pt = cv::Point2f(x[i], y[i]);
      in[i].response = resp[i];
    }
    std::vector<cv::KeyPoint> out = ex.DistributeOctTree(in, min_x, max_x, min_y, max_y, n_features, 0);
    n = (int)out.size();
    for (int i = 0; i < n; i++) {
      out_x[i] = out[i].pt.x;
      out_y[i] = out[i].pt.y;
      out_resp[i] = out[i].response;
    }
  }
  lorb_arena::on = false;
  lorb_arena::base = nullptr;
  lorb_arena::cap = 0;
  std::free(mem);
  return n;
}

// The whole ORBextractor::operator() (:1087-1151) on one 8-bit image, under the bump arena.
// Outputs in the reference's order (level by level): keypoint x, y (level-0 coordinates), octave,
// angle, response, size, and the descriptor rows.  Returns the number of keypoints (<= cap) or -1.
// out_level_xy (optional, [cap x 2]) receives the keypoints in level coordinates, re-derived as the
// integer pixel the reference scaled (pt / scale rounds back exactly: checked).
int ref_orb_extract(const uint8_t* img, int w, int h, int step, int nfeatures, float scale_factor, int nlevels,
                    int ini_th, int min_th, int cap, float* kx, float* ky, int* koct, float* kangle,
                    float* kresp, float* ksize, uint8_t* desc, int* n_per_level) {
  const size_t arena = (size_t)1 << 30;
  char* mem = (char*)std::malloc(arena);
  if (!mem) return -1;
  int n = -1;
  {
    cv::Mat image = level_mat(w, h, step, img);
    lorb_arena::base = mem;
    lorb_arena::used = 0;
    lorb_arena::cap = arena;
    lorb_arena::on = true;
    {
      ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
      std::vector<cv::KeyPoint> kps;
      cv::Mat d;
      ex(image, cv::Mat(), kps, d);
      n = (int)kps.size();
      if (n_per_level)
        for (int l = 0; l < nlevels; l++) n_per_level[l] = 0;
      for (int i = 0; i < n && i < cap; i++) {
        kx[i] = kps[i].pt.x;
        ky[i] = kps[i].pt.y;
        koct[i] = kps[i].octave;
        kangle[i] = kps[i].angle;
        kresp[i] = kps[i].response;
        ksize[i] = kps[i].size;
        std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
        if (n_per_level) n_per_level[kps[i].octave]++;
      }
    }
    lorb_arena::on = false;
  }
  lorb_arena::base = nullptr;
  lorb_arena::cap = 0;
  std::free(mem);
  return n;
}

}  // extern "C"
