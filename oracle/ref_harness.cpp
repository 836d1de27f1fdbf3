// oracle/_ref/libref.so = the REFERENCE'S OWN hot-path sources, compiled unmodified from
// /root/reference/src/{matcher,frame,map_point,map,bundle_adjust}.cpp, plus this harness.
//
// TEST INFRASTRUCTURE ONLY (never linked into or called by lorb_slam_b200/).
//
// The reference needs OpenCV-C++ and Ceres, which this image lacks; oracle/refshim/ provides
// stand-ins for exactly the types and functions those five files mention (see the headers there
// for what is pinned against cv2 and what is a restatement).  Everything else that runs below is
// the reference's code: Matcher::*, Frame::GetFeaturesInArea / AssignFeaturesToGrid / PosInGrid /
// ComputeImageBounds / IsInFrustum / SetPose, MapPoint::PredictScale / ComputeDescriptor,
// BA::ProjectPoseOptimization / LocalPoseOptimization with their three cost functors.
//
// This file only converts between the flat arrays of include/lorb_cuda.h and the reference's
// object graph (Frame*, MapPoint*, std::set<MapPoint*>, Camera*), built with -fno-access-control
// because Frame keeps its descriptors and grid private (include/frame.h:126-139).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <vector>

#include "frame.h"
#include "map.h"
#include "map_point.h"
#include "matcher.h"
#include "bundle_adjust.h"
#include "../include/lorb_cuda.h"

using namespace Simple_ORB_SLAM;

namespace {

// Camera has a single constructor that parses a YAML file (src/camera.cpp, not compiled here).
// All its members are PODs or cv::Mat stand-ins whose all-zero state is the empty matrix, so a
// zeroed block is a valid, never-destroyed Camera for the fields the hot path reads.
Camera* make_camera(float fx, float fy, float cx, float cy, float bf) {
  Camera* c = static_cast<Camera*>(std::calloc(1, sizeof(Camera)));
  c->fx = fx;
  c->fy = fy;
  c->cx = cx;
  c->cy = cy;
  c->bf = bf;
  return c;
}

cv::Mat mat4(const float* t16) {
  cv::Mat m(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m.at<float>(i, j) = t16[4 * i + j];
  return m;
}

cv::Mat desc_row(const uint8_t* d) {
  cv::Mat m(1, 32, CV_8U);
  std::memcpy(m.ptr<uint8_t>(), d, 32);
  return m;
}

void init_frame_scalars(Frame* F, Camera* cam) {
  F->mpCamera = cam;
  F->mnId = Frame::Idx++;
  F->mbBadFlag = false;
  F->mbFirstConnection = true;
  F->mpMap = nullptr;
  F->mpORBextractorLeft = F->mpORBextractorRight = nullptr;
  F->mTcw = cv::Mat::eye(4, 4, CV_32F);  // as the real constructor, src/frame.cpp:20
  F->mnMapPoints = 0;
}

// The current-frame side of every search: what Frame::Frame leaves behind after extraction
// (src/frame.cpp:17-62), with the keypoint arrays given instead of extracted.
Frame* make_search_frame(Camera* cam, int n_kp, const float* kx, const float* ky, const int* koct,
                         const float* kangle, const float* kuright, const uint8_t* kdesc,
                         float min_x, float max_x, float min_y, float max_y, const float* sf,
                         int n_levels) {
  Frame* F = new Frame();
  init_frame_scalars(F, cam);
  if (min_x == 0.0f && min_y == 0.0f && max_x == (float)(int)max_x && max_y == (float)(int)max_y) {
    cv::Mat img((int)max_y, (int)max_x, CV_8U);
    F->ComputeImageBounds(img);  // the reference's own bounds + grid-pitch arithmetic, :70-85
  } else {
    F->mbf = cam->bf;
    F->mb = F->mbf / cam->fx;
    F->cx = cam->cx; F->cy = cam->cy; F->fx = cam->fx; F->fy = cam->fy;
    F->mnMinX = min_x; F->mnMaxX = max_x; F->mnMinY = min_y; F->mnMaxY = max_y;
    F->mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (F->mnMaxX - F->mnMinX);
    F->mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (F->mnMaxY - F->mnMinY);
  }
  F->mnMapPoints = n_kp;
  F->mvKeys.resize(n_kp);
  for (int i = 0; i < n_kp; i++) {
    F->mvKeys[i].pt.x = kx[i];
    F->mvKeys[i].pt.y = ky[i];
    F->mvKeys[i].octave = koct[i];
    F->mvKeys[i].angle = kangle ? kangle[i] : 0.0f;
  }
  F->UndistortKeyPoints();  // mvKeysUn = mvKeys, :336-340
  F->mvuRight.assign(n_kp, -1.0f);
  if (kuright) F->mvuRight.assign(kuright, kuright + n_kp);
  F->mvDepth.assign(n_kp, -1.0f);
  F->mvpMapPoints = std::vector<MapPoint*>(n_kp, static_cast<MapPoint*>(NULL));
  F->mvbOutlier = std::vector<bool>(n_kp, false);
  F->mDescriptors = cv::Mat(n_kp, 32, CV_8U);
  if (kdesc)
    for (int i = 0; i < n_kp; i++) std::memcpy(F->mDescriptors.ptr<uint8_t>(i), kdesc + 32 * (size_t)i, 32);
  F->mnScaleLevels = n_levels;
  F->mvScaleFactors.assign(sf, sf + n_levels);
  F->mfScaleFactor = n_levels > 1 ? sf[1] : 1.2f;
  F->mfLogScaleFactor = log(F->mfScaleFactor);
  F->AssignFeaturesToGrid();  // :87-103
  return F;
}

// n MapPoints in one block, so that std::set<MapPoint*> iterates them in index order.
struct PointBlock {
  MapPoint* p = nullptr;
  int n = 0;
  Frame* birth = nullptr;
  PointBlock(int n_, Camera* cam) : n(n_) {
    birth = new Frame();
    init_frame_scalars(birth, cam);
    p = static_cast<MapPoint*>(::operator new(sizeof(MapPoint) * (size_t)(n > 0 ? n : 1)));
    for (int i = 0; i < n; i++) new (p + i) MapPoint(cv::Point3f(0, 0, 0), birth, nullptr);
  }
  ~PointBlock() {
    for (int i = 0; i < n; i++) p[i].~MapPoint();
    ::operator delete(p);
    delete birth;
  }
  int index_of(const MapPoint* q) const { return (q >= p && q < p + n) ? (int)(q - p) : -1; }
};

// keypoints of the frame that already hold a map point on entry (kclaim_obs[i] >= 0 = its mnObs)
void apply_claims(Frame* F, PointBlock& claims, const int* kclaim_obs) {
  for (int i = 0; i < claims.n; i++) {
    if (kclaim_obs[i] < 0) continue;
    claims.p[i].mnObs = (size_t)kclaim_obs[i];
    F->mvpMapPoints[i] = &claims.p[i];
  }
}

int final_state(const Frame* F, int kp, const PointBlock& items, const int* kclaim_obs) {
  const MapPoint* q = F->mvpMapPoints[kp];
  if (!q) return kclaim_obs[kp] >= 0 ? -2 : -1;  // -2: an entry claim was NULLed
  return items.index_of(q);                       // -1: still the entry claim
}

}  // namespace

extern "C" {

// ---- the stand-in's arithmetic, exposed so that tests can pin it against cv2 golden vectors
void ref_cv_rt(const float* R9, const float* x3, const float* t3, float* out3) {
  cv::Mat R(3, 3, CV_32F), x(3, 1, CV_32F), t(3, 1, CV_32F);
  for (int i = 0; i < 9; i++) R.at<float>(i / 3, i % 3) = R9[i];
  for (int i = 0; i < 3; i++) { x.at<float>(i) = x3[i]; t.at<float>(i) = t3[i]; }
  cv::Mat y = R * x + t;
  for (int i = 0; i < 3; i++) out3[i] = y.at<float>(i);
}
void ref_cv_mul4(const float* T16, const float* x4, float* out4) {
  cv::Mat x(4, 1, CV_32F);
  for (int i = 0; i < 4; i++) x.at<float>(i) = x4[i];
  cv::Mat y = mat4(T16) * x;
  for (int i = 0; i < 4; i++) out4[i] = y.at<float>(i);
}
void ref_cv_inv4(const float* T16, float* out16) {
  cv::Mat y = mat4(T16).inv();
  for (int i = 0; i < 16; i++) out16[i] = y.at<float>(i / 4, i % 4);
}
void ref_cv_rodrigues(const float* r3, float* R9) {
  cv::Mat r = (cv::Mat_<float>(3, 1) << r3[0], r3[1], r3[2]), R;
  cv::Rodrigues(r, R);
  for (int i = 0; i < 9; i++) R9[i] = R.at<float>(i / 3, i % 3);
}
int ref_cv_bfmatch(const uint8_t* q, int N, const uint8_t* t, int M, int* oq, int* ot, int* od) {
  cv::Mat Q(N, 32, CV_8U), T(M, 32, CV_8U);
  if (N) std::memcpy(Q.ptr<uint8_t>(), q, 32 * (size_t)N);
  if (M) std::memcpy(T.ptr<uint8_t>(), t, 32 * (size_t)M);
  std::vector<cv::DMatch> m;
  cv::BFMatcher(cv::NORM_HAMMING, true).match(Q, T, m);
  for (size_t i = 0; i < m.size(); i++) { oq[i] = m[i].queryIdx; ot[i] = m[i].trainIdx; od[i] = (int)m[i].distance; }
  return (int)m.size();
}

// ---- a2 / a8: the reference's scalar helpers
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  return Matcher::DescriptorDistance(desc_row(a), desc_row(b));
}
void ref_three_maxima(const int* counts, int L, int* i1, int* i2, int* i3) {
  std::vector<std::vector<int>> h(L);
  for (int i = 0; i < L; i++) h[i].assign(counts[i], 0);
  *i1 = *i2 = *i3 = -1;
  Matcher::ComputeThreeMaxima(h.data(), L, *i1, *i2, *i3);
}
float ref_radius_by_viewing_cos(float c) { return Matcher::RadiusByViewingCos(c); }
void ref_constants(int* th_low, int* th_high, int* histo) {
  *th_low = Matcher::TH_LOW; *th_high = Matcher::TH_HIGH; *histo = Matcher::HISTO_LENGTH;
}

// ---- a7: Frame::GetFeaturesInArea on the reference's own grid
int ref_features_in_area(int n_kp, const float* kx, const float* ky, const int* koct, float min_x,
                         float max_x, float min_y, float max_y, float x, float y, float r,
                         int minLevel, int maxLevel, int* out) {
  Camera* cam = make_camera(458, 458, 320, 240, 47.9f);
  const float sf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  Frame* F = make_search_frame(cam, n_kp, kx, ky, koct, nullptr, nullptr, nullptr, min_x, max_x, min_y, max_y, sf, 8);
  std::vector<size_t> v = F->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
  for (size_t i = 0; i < v.size(); i++) out[i] = (int)v[i];
  delete F;
  std::free(cam);
  return (int)v.size();
}

// ---- a3 / a4: Matcher::SearchByProjection(curr, prev) and SearchLocalPoints(curr, set).
// train_present[j] != 0 <=> prev->mvpMapPoints[j] is non-NULL (its descriptor = train_desc[j]).
// out_assign[i] = index j of the map point written to curr->mvpMapPoints[i], or -1.
int ref_search_bf(int use_set, int n_q, const uint8_t* q_desc, int n_t, const uint8_t* train_desc,
                  const uint8_t* train_present, int* out_assign) {
  Camera* cam = make_camera(458, 458, 320, 240, 47.9f);
  const float sf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  std::vector<float> z(n_q > n_t ? n_q : n_t, 1.0f);
  std::vector<int> zi(z.size(), 0);
  Frame* cur = make_search_frame(cam, n_q, z.data(), z.data(), zi.data(), nullptr, nullptr, q_desc, 0, 640, 0, 480, sf, 8);
  Frame* prev = make_search_frame(cam, n_t, z.data(), z.data(), zi.data(), nullptr, nullptr, nullptr, 0, 640, 0, 480, sf, 8);
  PointBlock pts(n_t, cam);
  std::set<MapPoint*> s;
  for (int j = 0; j < n_t; j++) {
    pts.p[j].mDescriptor = desc_row(train_desc + 32 * (size_t)j);
    if (train_present && !train_present[j]) continue;
    prev->mvpMapPoints[j] = &pts.p[j];
    s.insert(&pts.p[j]);
  }
  size_t n = use_set ? Matcher::SearchLocalPoints(cur, s) : Matcher::SearchByProjection(cur, prev);
  for (int i = 0; i < n_q; i++) out_assign[i] = cur->mvpMapPoints[i] ? pts.index_of(cur->mvpMapPoints[i]) : -1;
  delete cur;
  delete prev;
  std::free(cam);
  return (int)n;
}

// ---- the keyframe-pair sweep of BASELINE config 5 run through the reference's own entry point:
// for every pair (a, b), Matcher::SearchByProjection(curr = keyframe a, prev = keyframe b), i.e. its
// descriptor gathering (GetDescriptors clone, row-by-row push_back), the stand-in BFMatcher and its
// minDist / max(2*minDist, 30) filter.  Used by bench.py's reference arm (all host threads: OpenMP
// over pairs; the reference itself is single-threaded) and checked against the oracle's sweep.
struct RefSweep {
  Camera* cam = nullptr;
  int n_kf = 0, n_desc = 0;
  const uint8_t* bank = nullptr;
  std::vector<Frame*> prev;        // one per keyframe: every keypoint holds a map point
  std::vector<PointBlock*> pts;
};
void* ref_sweep_create(const uint8_t* bank, int n_kf, int n_desc) {
  RefSweep* S = new RefSweep();
  S->cam = make_camera(458, 458, 320, 240, 47.9f);
  S->n_kf = n_kf;
  S->n_desc = n_desc;
  S->bank = bank;
  const float sf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  std::vector<float> z(n_desc > 0 ? n_desc : 1, 1.0f);
  std::vector<int> zi(z.size(), 0);
  for (int k = 0; k < n_kf; k++) {
    const uint8_t* d = bank + (size_t)k * n_desc * 32;
    Frame* F = make_search_frame(S->cam, n_desc, z.data(), z.data(), zi.data(), nullptr, nullptr, d, 0, 640, 0, 480, sf, 8);
    PointBlock* P = new PointBlock(n_desc, S->cam);
    for (int j = 0; j < n_desc; j++) {
      P->p[j].mDescriptor = desc_row(d + 32 * (size_t)j);
      F->mvpMapPoints[j] = &P->p[j];
    }
    S->prev.push_back(F);
    S->pts.push_back(P);
  }
  return S;
}
void ref_sweep_run(void* h, const int* pair_a, const int* pair_b, int n_pairs, int* kept) {
  RefSweep* S = static_cast<RefSweep*>(h);
  const float sf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  std::vector<float> z(S->n_desc > 0 ? S->n_desc : 1, 1.0f);
  std::vector<int> zi(z.size(), 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < n_pairs; i++) {
    // a private current frame per pair: SearchByProjection writes curr->mvpMapPoints
    Frame* cur = make_search_frame(S->cam, S->n_desc, z.data(), z.data(), zi.data(), nullptr, nullptr,
                                   S->bank + (size_t)pair_a[i] * S->n_desc * 32, 0, 640, 0, 480, sf, 8);
    kept[i] = (int)Matcher::SearchByProjection(cur, S->prev[pair_b[i]]);
    delete cur;
  }
}
void ref_sweep_destroy(void* h) {
  RefSweep* S = static_cast<RefSweep*>(h);
  for (Frame* F : S->prev) delete F;
  for (PointBlock* P : S->pts) delete P;
  std::free(S->cam);
  delete S;
}

// ---- a6: Matcher::SearchByProjection(F, set, th).  Same arguments as orc_search_proj_points;
// out_point_for_kp = what F->mvpMapPoints holds at exit (-1 entry state, k = map point k).
int ref_search_proj_points(int n_kp, const float* kx, const float* ky, const int* koct,
                           const float* kuright, const uint8_t* kdesc, const int* kclaim_obs,
                           float min_x, float max_x, float min_y, float max_y,
                           const float* scale_factors, int n_levels, int n_pts, const float* proj_x,
                           const float* proj_y, const float* proj_xr, const int* level,
                           const float* view_cos, const uint8_t* active, const uint8_t* mp_desc,
                           const int* mp_nobs, float th, int* out_point_for_kp) {
  Camera* cam = make_camera(458, 458, 320, 240, 47.9f);
  Frame* F = make_search_frame(cam, n_kp, kx, ky, koct, nullptr, kuright, kdesc, min_x, max_x, min_y, max_y, scale_factors, n_levels);
  PointBlock claims(n_kp, cam), pts(n_pts, cam);
  apply_claims(F, claims, kclaim_obs);
  std::set<MapPoint*> s;
  for (int k = 0; k < n_pts; k++) {
    MapPoint& P = pts.p[k];
    P.mTrackProjX = proj_x[k];
    P.mTrackProjY = proj_y[k];
    P.mTrackProjXR = proj_xr[k];
    P.mnTrackScaleLevel = level[k];
    P.mTrackViewCos = view_cos[k];
    P.mbTrackInView = active[k] != 0;
    P.mnObs = (size_t)mp_nobs[k];
    P.mDescriptor = desc_row(mp_desc + 32 * (size_t)k);
    s.insert(&P);
  }
  size_t n = Matcher::SearchByProjection(F, s, th);
  for (int i = 0; i < n_kp; i++) out_point_for_kp[i] = final_state(F, i, pts, kclaim_obs);
  delete F;
  std::free(cam);
  return (int)n;
}

// ---- a5: Matcher::SearchByProjection(Cur, Last, th).  out_state_for_kp: i = Last item i,
// -1 = entry state, -2 = an entry claim NULLed by the rotation check.
int ref_search_proj_frame(int n_kp, const float* kx, const float* ky, const int* koct,
                          const float* kangle, const float* kuright, const uint8_t* kdesc,
                          const int* kclaim_obs, float min_x, float max_x, float min_y,
                          float max_y, const float* scale_factors, int n_levels, const float* tcw_cur,
                          const float* tcw_last, float fx, float fy, float cx, float cy, float mbf,
                          int n_last, const uint8_t* last_valid, const float* last_xw,
                          const int* last_octave, const float* last_angle, const uint8_t* mp_desc,
                          const int* mp_nobs, float th, int* out_state_for_kp, float* out_mb) {
  Camera* cam = make_camera(fx, fy, cx, cy, mbf);
  Frame* Cur = make_search_frame(cam, n_kp, kx, ky, koct, kangle, kuright, kdesc, min_x, max_x, min_y, max_y, scale_factors, n_levels);
  Cur->mTcw = mat4(tcw_cur);
  std::vector<float> z(n_last > 0 ? n_last : 1, 1.0f);
  Frame* Last = make_search_frame(cam, n_last, z.data(), z.data(), last_octave, last_angle, nullptr, nullptr, min_x, max_x, min_y, max_y, scale_factors, n_levels);
  Last->mTcw = mat4(tcw_last);
  PointBlock claims(n_kp, cam), pts(n_last, cam);
  apply_claims(Cur, claims, kclaim_obs);
  for (int i = 0; i < n_last; i++) {
    MapPoint& P = pts.p[i];
    P.SetWorldPos(cv::Point3f(last_xw[3 * i], last_xw[3 * i + 1], last_xw[3 * i + 2]));
    P.mnObs = (size_t)mp_nobs[i];
    P.mDescriptor = desc_row(mp_desc + 32 * (size_t)i);
    if (last_valid[i]) Last->mvpMapPoints[i] = &P;
  }
  if (out_mb) *out_mb = Cur->mb;
  size_t n = Matcher::SearchByProjection(Cur, Last, th);
  for (int i = 0; i < n_kp; i++) out_state_for_kp[i] = final_state(Cur, i, pts, kclaim_obs);
  delete Cur;
  delete Last;
  std::free(cam);
  return (int)n;
}

// ---- 8(f) rank 1: Frame::IsInFrustum + MapPoint::PredictScale.  The camera centre is NOT an
// input: the reference takes it from mTcw.inv() (src/frame.cpp:461-462); it is returned in ow_out.
void ref_frustum_project(const float* tcw, float fx, float fy, float cx, float cy, float mbf,
                         float min_x, float max_x, float min_y, float max_y, int n, const float* xw,
                         const float* normal, const float* min_dist, const float* max_dist,
                         float cos_limit, float scale_factor, int n_levels, uint8_t* in_view,
                         float* proj_x, float* proj_y, float* proj_xr, int* level, float* view_cos,
                         float* ow_out, float* log_sf_out) {
  Camera* cam = make_camera(fx, fy, cx, cy, mbf);
  std::vector<float> sf(n_levels, 1.0f);
  for (int l = 1; l < n_levels; l++) sf[l] = sf[l - 1] * scale_factor;
  Frame* F = make_search_frame(cam, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, min_x, max_x, min_y, max_y, sf.data(), n_levels);
  F->mfScaleFactor = scale_factor;
  F->mfLogScaleFactor = log(F->mfScaleFactor);  // src/frame.cpp:35
  F->mTcw = mat4(tcw);
  cv::Mat Twc = F->mTcw.inv();
  for (int i = 0; i < 3; i++) ow_out[i] = Twc.at<float>(i, 3);
  *log_sf_out = F->mfLogScaleFactor;
  PointBlock pts(n, cam);
  for (int i = 0; i < n; i++) {
    MapPoint& P = pts.p[i];
    P.SetWorldPos(cv::Point3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]));
    P.mNormalVector = (cv::Mat_<float>(3, 1) << normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
    P.mfMinDistance = min_dist[i];
    P.mfMaxDistance = max_dist[i];
    P.mTrackProjX = P.mTrackProjY = P.mTrackProjXR = P.mTrackViewCos = -7.0f;
    P.mnTrackScaleLevel = -7;
    const bool ok = F->IsInFrustum(&P, cos_limit);
    in_view[i] = ok ? 1 : 0;
    proj_x[i] = P.mTrackProjX;
    proj_y[i] = P.mTrackProjY;
    proj_xr[i] = P.mTrackProjXR;
    level[i] = P.mnTrackScaleLevel;
    view_cos[i] = P.mTrackViewCos;
  }
  delete F;
  std::free(cam);
}

// ---- 8(f) rank 2: Frame::ComputeStereoMatches (private; src/frame.cpp:125-333) on given pyramids.
// ORBextractor's constructor is not on the path (aborting stub): its two instances are zeroed
// blocks (an all-zero std::vector is the empty vector) that only carry mvImagePyramid.
int ref_stereo_matches(int n_levels, const int* lw, const int* lh, const int* lstep,
                       const uint8_t* const* pyr_left, const int* rw, const int* rh, const int* rstep,
                       const uint8_t* const* pyr_right, const float* sf, const float* inv_sf, float fx,
                       float mbf, int N, const float* lx, const float* ly, const int* loct,
                       const uint8_t* ldesc, int Nr, const float* rx, const float* ry, const int* roct,
                       const uint8_t* rdesc, float* out_uright, float* out_depth, float* out_mb) {
  Camera* cam = make_camera(fx, fx, lw[0] / 2.0f, lh[0] / 2.0f, mbf);
  Frame* F = make_search_frame(cam, N, lx, ly, loct, nullptr, nullptr, ldesc, 0, (float)lw[0], 0, (float)lh[0], sf, n_levels);
  F->mvInvScaleFactors.assign(inv_sf, inv_sf + n_levels);
  F->mvKeysRight.resize(Nr);
  for (int i = 0; i < Nr; i++) {
    F->mvKeysRight[i].pt.x = rx[i];
    F->mvKeysRight[i].pt.y = ry[i];
    F->mvKeysRight[i].octave = roct[i];
  }
  F->mDescriptorsRight = cv::Mat(Nr, 32, CV_8U);
  for (int i = 0; i < Nr; i++) std::memcpy(F->mDescriptorsRight.ptr<uint8_t>(i), rdesc + 32 * (size_t)i, 32);
  ORBextractor* ex[2];
  for (int s = 0; s < 2; s++) {
    ex[s] = static_cast<ORBextractor*>(std::calloc(1, sizeof(ORBextractor)));
    const int* w = s ? rw : lw; const int* h = s ? rh : lh; const int* st = s ? rstep : lstep;
    const uint8_t* const* img = s ? pyr_right : pyr_left;
    for (int l = 0; l < n_levels; l++) {
      cv::Mat m(h[l], w[l], CV_8U);
      for (int r = 0; r < h[l]; r++) std::memcpy(m.ptr<uint8_t>(r), img[l] + (size_t)r * st[l], w[l]);
      ex[s]->mvImagePyramid.push_back(m);
    }
  }
  F->mpORBextractorLeft = ex[0];
  F->mpORBextractorRight = ex[1];
  if (out_mb) *out_mb = F->mb;
  F->ComputeStereoMatches();
  int kept = 0;
  for (int i = 0; i < N; i++) {
    out_uright[i] = F->mvuRight[i];
    out_depth[i] = F->mvDepth[i];
    kept += F->mvuRight[i] != -1.0f || F->mvDepth[i] != -1.0f;
  }
  for (int s = 0; s < 2; s++) { ex[s]->mvImagePyramid.clear(); std::free(ex[s]); }
  delete F;
  std::free(cam);
  return kept;
}

// ---- 8(f) rank 4: MapPoint::ComputeDescriptor over m observing frames (one descriptor each).
// Returns the index of the chosen observation.
int ref_compute_descriptor(const uint8_t* desc, int m) {
  if (m <= 0) return -1;
  Camera* cam = make_camera(458, 458, 320, 240, 47.9f);
  Frame* frames = new Frame[m];  // one block: std::map<Frame*, size_t> iterates in index order
  PointBlock pt(1, cam);
  for (int i = 0; i < m; i++) {
    init_frame_scalars(&frames[i], cam);
    frames[i].mDescriptors = desc_row(desc + 32 * (size_t)i);
    frames[i].mnMapPoints = 1;
    pt.p[0].AddObservation(&frames[i], 0);
  }
  pt.p[0].ComputeDescriptor();
  cv::Mat d = pt.p[0].GetDescriptor();
  int chosen = -1;
  for (int i = 0; i < m && chosen < 0; i++)
    if (std::memcmp(d.ptr<uint8_t>(), desc + 32 * (size_t)i, 32) == 0) chosen = i;
  delete[] frames;
  std::free(cam);
  return chosen;
}

// ------------------------------------------------------------------------ BA
static void set_override(const lorb_ba_options* o) {
  ceres::LorbLastSolve& L = ceres::lorb_last_solve();
  L.have_override = o != nullptr;
  if (!o) return;
  ceres::Solver::Options c;
  c.linear_solver_type = ceres::DENSE_SCHUR;
  c.max_num_iterations = o->max_num_iterations;
  c.jacobi_scaling = o->jacobi_scaling != 0;
  c.max_num_consecutive_invalid_steps = o->max_consecutive_invalid_steps;
  c.function_tolerance = o->function_tolerance;
  c.gradient_tolerance = o->gradient_tolerance;
  c.parameter_tolerance = o->parameter_tolerance;
  c.initial_trust_region_radius = o->initial_trust_region_radius;
  c.max_trust_region_radius = o->max_trust_region_radius;
  c.min_trust_region_radius = o->min_trust_region_radius;
  c.min_relative_decrease = o->min_relative_decrease;
  c.min_lm_diagonal = o->min_lm_diagonal;
  c.max_lm_diagonal = o->max_lm_diagonal;
  L.override_options = c;
}
static void get_summary(lorb_ba_summary* s) {
  const ceres::Solver::Summary& c = ceres::lorb_last_solve().summary;
  s->initial_cost = c.initial_cost;
  s->final_cost = c.final_cost;
  s->final_radius = c.final_radius;
  s->final_gradient_max_norm = c.final_gradient_max_norm;
  s->iterations = c.iterations;
  s->num_successful_steps = c.num_successful_steps;
  s->num_unsuccessful_steps = c.num_unsuccessful_steps;
  s->termination = c.lorb_termination;
}

// a10: BA::ProjectPoseOptimization.  rt_in = (rvec, tvec) as the floats the Frame stores;
// rt_f32 = what the reference writes back (Frame::mRvec / mTvec), rt_f64 = the doubles its
// solver held before that rounding, tcw_out = the 4x4 pose Frame::SetPose derives.
int ref_ba_pose_only(int n, const float* xw, const float* uv, const float* K, const float* rt_in,
                     const lorb_ba_options* opt, float* rt_f32, double* rt_f64, float* tcw_out,
                     lorb_ba_summary* sum) {
  Camera* cam = make_camera(K[0], K[1], K[2], K[3], 47.9f);
  const float sf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  std::vector<float> kx(n > 0 ? n : 1), ky(n > 0 ? n : 1);
  std::vector<int> ko(n > 0 ? n : 1, 0);
  for (int i = 0; i < n; i++) { kx[i] = uv[2 * i]; ky[i] = uv[2 * i + 1]; }
  Frame* F = make_search_frame(cam, n, kx.data(), ky.data(), ko.data(), nullptr, nullptr, nullptr, 0, 640, 0, 480, sf, 8);
  F->mRvec = (cv::Mat_<float>(3, 1) << rt_in[0], rt_in[1], rt_in[2]);
  F->mTvec = (cv::Mat_<float>(3, 1) << rt_in[3], rt_in[4], rt_in[5]);
  PointBlock pts(n, cam);
  for (int i = 0; i < n; i++) {
    pts.p[i].SetWorldPos(cv::Point3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]));
    F->mvpMapPoints[i] = &pts.p[i];
  }
  set_override(opt);
  BA::ProjectPoseOptimization(F);
  set_override(nullptr);
  const ceres::LorbLastSolve& L = ceres::lorb_last_solve();
  for (int i = 0; i < 3; i++) {
    rt_f32[i] = F->mRvec.at<float>(i);
    rt_f32[3 + i] = F->mTvec.at<float>(i);
  }
  // blocks in order of first appearance: initialR then initialT (src/bundle_adjust.cpp:184)
  for (int i = 0; i < 6; i++) rt_f64[i] = L.value.size() == 6 ? L.value[i] : (double)rt_in[i];
  for (int i = 0; i < 16; i++) tcw_out[i] = F->mTcw.at<float>(i / 4, i % 4);
  get_summary(sum);
  delete F;
  std::free(cam);
  return 0;
}

// a11: BA::LocalPoseOptimization.  Window camera 0 is pCurrFrame, cameras 1..C-1 its covisible
// frames; every fixed observation gets its own out-of-window Frame with pose fix_rt.
// cams_in / pts_in are the floats the reference's objects store.
int ref_ba_local(int C, const float* cams_in, int P, const float* pts_in, int O, const int* obs_cam,
                 const int* obs_pt, const float* obs_uv, int F, const int* fix_pt, const float* fix_uv,
                 const float* fix_rt, const float* K, const lorb_ba_options* opt, float* cams_f32,
                 double* cams_f64, float* pts_f32, double* pts_f64, lorb_ba_summary* sum) {
  Camera* cam = make_camera(K[0], K[1], K[2], K[3], 47.9f);
  const int NF = C + F;
  Frame* frames = new Frame[NF > 0 ? NF : 1];
  PointBlock pts(P, cam);
  for (int j = 0; j < P; j++) pts.p[j].SetWorldPos(cv::Point3f(pts_in[3 * j], pts_in[3 * j + 1], pts_in[3 * j + 2]));
  for (int f = 0; f < NF; f++) {
    Frame& Fr = frames[f];
    init_frame_scalars(&Fr, cam);
    const float* rt = f < C ? cams_in + 6 * f : fix_rt + 6 * (size_t)(f - C);
    Fr.mRvec = (cv::Mat_<float>(3, 1) << rt[0], rt[1], rt[2]);
    Fr.mTvec = (cv::Mat_<float>(3, 1) << rt[3], rt[4], rt[5]);
  }
  auto add_obs = [&](Frame& Fr, int pt, float u, float v) {
    const size_t slot = Fr.mvpMapPoints.size();
    Fr.mvpMapPoints.push_back(&pts.p[pt]);
    Fr.mKps2d.push_back(cv::Point2f(u, v));  // what GetKps2d() returns (src/frame.cpp:628-631)
    cv::KeyPoint kp;
    kp.pt = cv::Point2f(u, v);
    Fr.mvKeys.push_back(kp);
    Fr.mvKeysUn.push_back(kp);
    Fr.mnMapPoints = Fr.mvpMapPoints.size();
    pts.p[pt].AddObservation(&Fr, slot);
  };
  for (int o = 0; o < O; o++) add_obs(frames[obs_cam[o]], obs_pt[o], obs_uv[2 * o], obs_uv[2 * o + 1]);
  for (int f = 0; f < F; f++) add_obs(frames[C + f], fix_pt[f], fix_uv[2 * f], fix_uv[2 * f + 1]);
  for (int c = 1; c < C; c++) frames[0].mvpOrderedKeyFrames.push_back(&frames[c]);

  // the order in which the reference will meet the points (src/bundle_adjust.cpp:224-241)
  std::vector<int> order;
  {
    std::vector<char> seen(P > 0 ? P : 1, 0);
    for (int c = 0; c < C; c++)
      for (MapPoint* q : frames[c].mvpMapPoints) {
        const int j = pts.index_of(q);
        if (!seen[j]) { seen[j] = 1; order.push_back(j); }
      }
  }
  set_override(opt);
  BA::LocalPoseOptimization(&frames[0]);
  set_override(nullptr);
  const ceres::LorbLastSolve& L = ceres::lorb_last_solve();
  // doubles: framesPose[][6] and mpsPose[][3] are two arrays, so blocks of one size sort by address
  std::map<double*, const double*> b6, b3;
  {
    size_t off = 0;
    for (size_t b = 0; b < L.ptr.size(); b++) {
      (L.size[b] == 6 ? b6 : b3)[L.ptr[b]] = &L.value[off];
      off += L.size[b];
    }
  }
  for (int c = 0; c < C; c++)
    for (int k = 0; k < 6; k++) cams_f64[6 * c + k] = (double)cams_in[6 * c + k];
  for (int j = 0; j < P; j++)
    for (int k = 0; k < 3; k++) pts_f64[3 * j + k] = (double)pts_in[3 * j + k];
  int rc = 0;
  if ((int)b6.size() == C) {  // every window camera has at least one residual
    int c = 0;
    for (auto& kv : b6) { for (int k = 0; k < 6; k++) cams_f64[6 * c + k] = kv.second[k]; c++; }
  } else {
    rc = 1;  // cannot map camera blocks back by address; floats below are still valid
  }
  if (b3.size() == order.size()) {
    size_t i = 0;
    for (auto& kv : b3) { for (int k = 0; k < 3; k++) pts_f64[3 * order[i] + k] = kv.second[k]; i++; }
  } else {
    rc |= 2;
  }
  for (int c = 0; c < C; c++)
    for (int k = 0; k < 3; k++) {
      cams_f32[6 * c + k] = frames[c].mRvec.at<float>(k);
      cams_f32[6 * c + 3 + k] = frames[c].mTvec.at<float>(k);
    }
  for (int j = 0; j < P; j++) {
    const cv::Point3f p = pts.p[j].GetPos();
    pts_f32[3 * j] = p.x; pts_f32[3 * j + 1] = p.y; pts_f32[3 * j + 2] = p.z;
  }
  get_summary(sum);
  delete[] frames;
  std::free(cam);
  return rc;
}

}  // extern "C"

// ---- symbols the compiled reference files mention but the hot path never reaches (the extractor,
// the GUI, the YAML camera).  ctypes loads with RTLD_NOW, so they need a definition; reaching one
// is a harness bug and aborts loudly.
#define LORB_OFF_PATH(what)                                                        \
  do {                                                                             \
    std::fprintf(stderr, "oracle/_ref: %s is not on the hot path\n", what);        \
    std::abort();                                                                  \
  } while (0)
namespace cv {
void imshow(const std::string&, const Mat&) { LORB_OFF_PATH("cv::imshow"); }
int waitKey(int) { LORB_OFF_PATH("cv::waitKey"); }
Ptr<ORB> ORB::create() { LORB_OFF_PATH("cv::ORB::create"); }
void Feature2D::detect(const Mat&, std::vector<KeyPoint>&) { LORB_OFF_PATH("cv::Feature2D::detect"); }
void Feature2D::compute(const Mat&, std::vector<KeyPoint>&, Mat&) { LORB_OFF_PATH("cv::Feature2D::compute"); }
// The four image operations of ORBextractor: the cv2-pinned restatements of oracle/orb_ref.cpp
// (linked into this library).  Only the forms ORBextractor.cpp uses are accepted.
}  // namespace cv
namespace orc {
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep);
void gaussian7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep);
int fast9_nms(const uint8_t* img, int w, int h, int step, int threshold, int* out_x, int* out_y, int* out_score,
              int cap);
}  // namespace orc
namespace cv {
void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation) {
  if (src.type() != CV_8U || fx != 0 || fy != 0 || interpolation != INTER_LINEAR) LORB_OFF_PATH("cv::resize form");
  if (dst.rows != dsize.height || dst.cols != dsize.width || dst.type() != CV_8U) dst.create(dsize.height, dsize.width, CV_8U);
  orc::resize_linear_u8(src.ptr(0), src.cols, src.rows, (int)src.step, dst.ptr(0), dst.cols, dst.rows, (int)dst.step);
}
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
  if (src.type() != CV_8U || (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) LORB_OFF_PATH("cv::copyMakeBorder form");
  const int h = src.rows + top + bottom, w = src.cols + left + right;
  const Mat s = src.clone();  // src may be the centre of dst (ORBextractor.cpp:1174)
  if (dst.rows != h || dst.cols != w || dst.type() != CV_8U) dst.create(h, w, CV_8U);
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
  };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) dst.ptr(y)[x] = s.ptr(refl(y - top, s.rows))[refl(x - left, s.cols)];
}
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  if (src.type() != CV_8U || ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 ||
      borderType != BORDER_REFLECT_101)
    LORB_OFF_PATH("cv::GaussianBlur form");
  const Mat s = src.clone();  // in-place call (ORBextractor.cpp:1132)
  if (dst.rows != s.rows || dst.cols != s.cols || dst.type() != CV_8U) dst.create(s.rows, s.cols, CV_8U);
  orc::gaussian7_u8(s.ptr(0), s.cols, s.rows, (int)s.step, dst.ptr(0), (int)dst.step);
}
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  if (image.type() != CV_8U || !nonmaxSuppression) LORB_OFF_PATH("cv::FAST form");
  keypoints.clear();
  if (image.rows < 7 || image.cols < 7) return;
  const int cap = image.rows * image.cols;
  std::vector<int> x(cap), y(cap), sc(cap);
  const int n = orc::fast9_nms(image.ptr(0), image.cols, image.rows, (int)image.step, threshold, x.data(), y.data(),
                               sc.data(), cap);
  for (int i = 0; i < n; i++) {  // KeyPoint((float)j, (float)(i-1), 7.f, -1, (float)score)
    KeyPoint kp;
    kp.pt = Point2f((float)x[i], (float)y[i]);
    kp.size = 7.f;
    kp.angle = -1;
    kp.response = (float)sc[i];
    keypoints.push_back(kp);
  }
}
void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) { LORB_OFF_PATH("cv::KeyPointsFilter::retainBest"); }
}  // namespace cv
namespace Simple_ORB_SLAM {
// (ORBextractor itself is compiled from the reference in ref_orb_harness.cpp)
bool Camera::Project(const Point3f&, Point2f&) { LORB_OFF_PATH("Camera::Project"); }
}  // namespace Simple_ORB_SLAM
