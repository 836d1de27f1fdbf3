// stereo_dev.cuh — device-side argument block of the stereo matcher (match_stereo.cu), shared with
// the frame front-end in orb.cu, which feeds it pyramids, keypoints and descriptors that are
// already resident on the device.
#pragma once
#include "common.cuh"

namespace lorb {

constexpr int STEREO_MAX_LEVELS = 16;
constexpr int STEREO_W = 5;  // patch half size  (:239)
constexpr int STEREO_L = 5;  // shift half range (:246)

struct PyrDev {
  const uint8_t* lvl[STEREO_MAX_LEVELS];  // tightly packed rows, stride = w
  int w[STEREO_MAX_LEVELS], h[STEREO_MAX_LEVELS];
};

struct StereoDev {
  PyrDev left, right;
  int n_left, n_right, n_levels, n_rows;
  const float *lx, *ly, *rx, *ry;
  const int *loct, *roct;
  const uint4 *ldesc, *rdesc;
  float sf[STEREO_MAX_LEVELS], inv_sf[STEREO_MAX_LEVELS];
  float mbf, mb;
};

// Launches stereo_match_kernel + stereo_finalize_kernel on the context's stream; every pointer
// (in S and the outputs) is a device pointer.  d_sad is scratch of n_left ints.
int stereo_launch(lorb_ctx* c, const StereoDev& S, float* d_uright, float* d_depth, int* d_sad, int* d_n_matched);

}  // namespace lorb
