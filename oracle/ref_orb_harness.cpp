// Second translation unit of oracle/_ref/libref.so: the reference's src/ORBextractor.cpp, textually
// included (unmodified, from where it lies) so that its file-static functions IC_Angle and
// computeOrbDescriptor -- the orientation and steered-BRIEF stages, SURVEY 8(f) rank 5 -- can be
// called on given pyramid levels and keypoints.  TEST INFRASTRUCTURE ONLY.
//
// The image-processing calls of the rest of that file (cv::resize, copyMakeBorder, GaussianBlur,
// FAST, KeyPointsFilter) are declared by the stand-in and defined as aborting stubs in
// ref_harness.cpp: detection and the pyramid are not on the path (the pyramids are inputs).  The
// real ORBextractor constructor runs (pattern copy, umax table, scale factors).
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

}  // extern "C"
