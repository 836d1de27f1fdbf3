/*
 * oracle/match_ref.c — CPU restatement of the reference's descriptor matching.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lorb_slam_b200/,
 * include/) may call, link or import this file.  It is the checker that the
 * CUDA path is compared against (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).
 *
 * Parity status
 *   - brute-force cross-check + knn2: PINNED against cv2 4.13 BFMatcher (the
 *     same OpenCV routine the reference calls, src/matcher.cpp:36-39) through
 *     the golden vectors in tests/golden/ (tests/golden/make_golden.py).
 *   - fp32 projection arithmetic: PINNED against cv2.gemm through golden vectors.
 *   - projection-guided searches as a whole: restated line by line from the
 *     reference (cited below); the reference has no tests or fixtures for them
 *      and cannot be built here -> "parity unpinned" beyond the
 *     citations.
 *
 * All citations are file:line in /root/reference.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * The reference itself builds -O0 on baseline x86-64 (CMakeLists.txt:5-6), so
 * no FMA contraction may happen in the float code below.
 */
#include <limits.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GRID_COLS 64 /* include/frame.h:14 */
#define GRID_ROWS 48 /* include/frame.h:13 */
#define TH_HIGH 100  /* src/matcher.cpp:6 */
#define TH_LOW 50    /* src/matcher.cpp:7 */
#define HISTO_LENGTH 30 /* src/matcher.cpp:8 */

/* ---- a2: Matcher::DescriptorDistance, src/matcher.cpp:369-385 (SWAR popcount
 * over eight 32-bit words). */
int orc_hamming256(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t wa, wb;
    memcpy(&wa, a + 4 * i, 4);
    memcpy(&wb, b + 4 * i, 4);
    uint32_t v = wa ^ wb;
    v = v - ((v >> 1) & 0x55555555u);
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
    dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
  }
  return dist;
}

/* Faster equivalent used for the big distance matrices (same value; the test
 * suite checks it against orc_hamming256). */
static inline int ham_fast(const uint8_t* a, const uint8_t* b) {
  uint64_t x[4], y[4];
  memcpy(x, a, 32);
  memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
         __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

/* Nearest neighbour both ways: fwd[i] = argmin_t d(i,t), bwd[t] = argmin_i d(i,t),
 * lowest index on ties (cv::batchDistance keeps the first strict minimum). */
static void nn_both(const uint8_t* q, int nq, const uint8_t* t, int nt, int* fwd, int* fwd_d,
                    int* bwd, int* bwd_d) {
  for (int j = 0; j < nt; j++) {
    bwd[j] = -1;
    bwd_d[j] = 1 << 30;
  }
  for (int i = 0; i < nq; i++) {
    int best = -1, bd = 1 << 30;
    const uint8_t* qi = q + 32 * (size_t)i;
    for (int j = 0; j < nt; j++) {
      int d = ham_fast(qi, t + 32 * (size_t)j);
      if (d < bd) {
        bd = d;
        best = j;
      }
      if (d < bwd_d[j]) {
        bwd_d[j] = d;
        bwd[j] = i;
      }
    }
    fwd[i] = best;
    fwd_d[i] = bd;
  }
}

/*
 * a3/a4: cv::BFMatcher(NORM_HAMMING, crossCheck=true).match(query, train)
 * as called at src/matcher.cpp:36-39 and :342-345.
 * mode 0 (mutual): OpenCV >= 3.4 semantics, pinned against cv2 4.13.
 * mode 1 (legacy): OpenCV 3.1 batchDistance cross-check as recalled (for every
 *   train t, idx = its nearest query; dist[idx] takes the strictly smaller d,
 *   scanning t ascending).  Unpinned: OpenCV 3.1 is not executable here.
 * Emits matches in ascending query index (the order knnMatch's result is
 * flattened in).  Returns the number of matches.
 */
int orc_bf_crosscheck(const uint8_t* q, int nq, const uint8_t* t, int nt, int mode, int* out_q,
                      int* out_t, int* out_d) {
  if (nq <= 0 || nt <= 0) return 0;
  int* fwd = (int*)malloc(sizeof(int) * (size_t)nq * 2);
  int* fwd_d = fwd + nq;
  int* bwd = (int*)malloc(sizeof(int) * (size_t)nt * 2);
  int* bwd_d = bwd + nt;
  nn_both(q, nq, t, nt, fwd, fwd_d, bwd, bwd_d);
  int n = 0;
  if (mode == 0) {
    for (int i = 0; i < nq; i++) {
      int j = fwd[i];
      if (j >= 0 && bwd[j] == i) {
        out_q[n] = i;
        out_t[n] = j;
        out_d[n] = fwd_d[i];
        n++;
      }
    }
  } else {
    int* nidx = (int*)malloc(sizeof(int) * (size_t)nq * 2);
    int* dist = nidx + nq;
    for (int i = 0; i < nq; i++) {
      nidx[i] = -1;
      dist[i] = 0x7fffffff;
    }
    for (int j = 0; j < nt; j++) {
      int idx = bwd[j];
      int d = bwd_d[j];
      if (d < dist[idx]) {
        dist[idx] = d;
        nidx[idx] = j;
      }
    }
    for (int i = 0; i < nq; i++)
      if (nidx[i] >= 0) {
        out_q[n] = i;
        out_t[n] = nidx[i];
        out_d[n] = dist[i];
        n++;
      }
    free(nidx);
  }
  free(fwd);
  free(bwd);
  return n;
}

/* src/matcher.cpp:42-56: minDist scan, then keep iff !(distance > max(2*minDist, 30.0))
 * evaluated in double.  Returns the kept count (the reference's return value). */
int orc_bf_filter(const int* d, int n, int* min_dist_out, uint8_t* keep) {
  double minDist = 1.7976931348623157e308; /* DBL_MAX */
  for (int i = 0; i < n; i++)
    if ((double)(float)d[i] < minDist) minDist = (double)(float)d[i];
  int kept = 0;
  for (int i = 0; i < n; i++) {
    double thr = 2 * minDist > 30.0 ? 2 * minDist : 30.0;
    int k = !((double)(float)d[i] > thr);
    keep[i] = (uint8_t)k;
    kept += k;
  }
  *min_dist_out = n > 0 ? (int)minDist : -1;
  return kept;
}

/* kNN-2 (cv::BFMatcher::knnMatch k=2, crossCheck=false): neighbours sorted by
 * (distance, train index).  idx = -1 / dist = 256 where there is no neighbour. */
void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) {
  for (int i = 0; i < nq; i++) {
    int b0 = -1, d0 = 256 + 1, b1 = -1, d1 = 256 + 1;
    const uint8_t* qi = q + 32 * (size_t)i;
    for (int j = 0; j < nt; j++) {
      int d = ham_fast(qi, t + 32 * (size_t)j);
      if (d < d0) {
        d1 = d0;
        b1 = b0;
        d0 = d;
        b0 = j;
      } else if (d < d1) {
        d1 = d;
        b1 = j;
      }
    }
    idx[2 * i] = b0;
    dist[2 * i] = b0 >= 0 ? d0 : 256;
    idx[2 * i + 1] = b1;
    dist[2 * i + 1] = b1 >= 0 ? d1 : 256;
  }
}

/* Sweep unit (BASELINE config 5): one keyframe pair -> (kept, matches, minDist). */
void orc_sweep_pair(const uint8_t* a, const uint8_t* b, int n_desc, int* kept, int* matches,
                    int* min_dist) {
  int* oq = (int*)malloc(sizeof(int) * (size_t)n_desc * 3);
  int* ot = oq + n_desc;
  int* od = ot + n_desc;
  uint8_t* keep = (uint8_t*)malloc((size_t)n_desc + 1);
  int n = orc_bf_crosscheck(a, n_desc, b, n_desc, 0, oq, ot, od);
  *kept = orc_bf_filter(od, n, min_dist, keep);
  *matches = n;
  free(oq);
  free(keep);
}

/* ------------------------------------------------------------------ a7: grid */

typedef struct {
  int n_kp;
  const float *x, *y;
  const int* octave;
  float min_x, min_y, inv_w, inv_h;
  int* cell_start; /* GRID_COLS*GRID_ROWS + 1, cell id = ix*GRID_ROWS + iy */
  int* cell_items; /* keypoint indices, ascending inside a cell */
} grid_t;

/* ComputeImageBounds src/frame.cpp:83-84, AssignFeaturesToGrid :87-103,
 * PosInGrid :105-115 (round(), not floor; out-of-range cells drop the keypoint). */
static void grid_build(grid_t* g, int n_kp, const float* x, const float* y, const int* octave,
                       float min_x, float max_x, float min_y, float max_y) {
  g->n_kp = n_kp;
  g->x = x;
  g->y = y;
  g->octave = octave;
  g->min_x = min_x;
  g->min_y = min_y;
  g->inv_w = (float)GRID_COLS / (max_x - min_x);
  g->inv_h = (float)GRID_ROWS / (max_y - min_y);
  int ncell = GRID_COLS * GRID_ROWS;
  g->cell_start = (int*)calloc((size_t)ncell + 1, sizeof(int));
  g->cell_items = (int*)malloc(sizeof(int) * (size_t)(n_kp > 0 ? n_kp : 1));
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)(n_kp > 0 ? n_kp : 1));
  for (int i = 0; i < n_kp; i++) {
    int px = (int)roundf((x[i] - min_x) * g->inv_w);
    int py = (int)roundf((y[i] - min_y) * g->inv_h);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) {
      cell_of[i] = -1;
      continue;
    }
    cell_of[i] = px * GRID_ROWS + py;
    g->cell_start[cell_of[i] + 1]++;
  }
  for (int c = 0; c < ncell; c++) g->cell_start[c + 1] += g->cell_start[c];
  int* fill = (int*)malloc(sizeof(int) * (size_t)ncell);
  memcpy(fill, g->cell_start, sizeof(int) * (size_t)ncell);
  for (int i = 0; i < n_kp; i++)
    if (cell_of[i] >= 0) g->cell_items[fill[cell_of[i]]++] = i;
  free(fill);
  free(cell_of);
}

static void grid_free(grid_t* g) {
  free(g->cell_start);
  free(g->cell_items);
}

/* Frame::GetFeaturesInArea src/frame.cpp:370-423.  Writes indices in the
 * reference's iteration order (ix outer, iy inner, cell order); returns count. */
static int grid_query(const grid_t* g, float x, float y, float r, int minLevel, int maxLevel,
                      int* out) {
  int n = 0;
  int nMinCellX = (int)floorf((x - g->min_x - r) * g->inv_w);
  if (nMinCellX < 0) nMinCellX = 0;
  if (nMinCellX >= GRID_COLS) return 0;
  int nMaxCellX = (int)ceilf((x - g->min_x + r) * g->inv_w);
  if (nMaxCellX > GRID_COLS - 1) nMaxCellX = GRID_COLS - 1;
  if (nMaxCellX < 0) return 0;
  int nMinCellY = (int)floorf((y - g->min_y - r) * g->inv_h);
  if (nMinCellY < 0) nMinCellY = 0;
  if (nMinCellY >= GRID_ROWS) return 0;
  int nMaxCellY = (int)ceilf((y - g->min_y + r) * g->inv_h);
  if (nMaxCellY > GRID_ROWS - 1) nMaxCellY = GRID_ROWS - 1;
  if (nMaxCellY < 0) return 0;
  const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      int c = ix * GRID_ROWS + iy;
      for (int j = g->cell_start[c]; j < g->cell_start[c + 1]; j++) {
        int k = g->cell_items[j];
        if (bCheckLevels) {
          if (g->octave[k] < minLevel) continue;
          if (maxLevel >= 0)
            if (g->octave[k] > maxLevel) continue;
        }
        const float distx = g->x[k] - x;
        const float disty = g->y[k] - y;
        if (fabsf(distx) < r && fabsf(disty) < r) out[n++] = k;
      }
    }
  return n;
}

/* Exposed for tests: candidate list of one window query. */
int orc_features_in_area(int n_kp, const float* kx, const float* ky, const int* octave, float min_x,
                         float max_x, float min_y, float max_y, float x, float y, float r,
                         int minLevel, int maxLevel, int* out) {
  grid_t g;
  grid_build(&g, n_kp, kx, ky, octave, min_x, max_x, min_y, max_y);
  int n = grid_query(&g, x, y, r, minLevel, maxLevel, out);
  grid_free(&g);
  return n;
}

/* ---- a6: Matcher::SearchByProjection(Frame*, const set<MapPoint*>&, th),
 * src/matcher.cpp:220-316.  Points arrive in set-iteration order. */
int orc_search_proj_points(int n_kp, const float* kx, const float* ky, const int* koct,
                           const float* kuright, const uint8_t* kdesc, const int* kclaim_obs,
                           float min_x, float max_x, float min_y, float max_y,
                           const float* scale_factors, int n_pts, const float* proj_x,
                           const float* proj_y, const float* proj_xr, const int* level,
                           const float* view_cos, const uint8_t* active, const uint8_t* mp_desc,
                           const int* mp_nobs, float th, int* out_kp_for_point,
                           int* out_point_for_kp, long long* n_candidates) {
  grid_t g;
  grid_build(&g, n_kp, kx, ky, koct, min_x, max_x, min_y, max_y);
  int* holder_obs = (int*)malloc(sizeof(int) * (size_t)(n_kp > 0 ? n_kp : 1));
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_kp > 0 ? n_kp : 1));
  for (int i = 0; i < n_kp; i++) {
    holder_obs[i] = kclaim_obs[i];
    out_point_for_kp[i] = -1;
  }
  long long ncand = 0;
  int nmatches = 0;
  const int bFactor = th != 1.0; /* :224 (float promoted to double) */
  for (int k = 0; k < n_pts; k++) {
    out_kp_for_point[k] = -1;
    if (!active[k]) continue; /* :231-235 */
    const int lvl = level[k];
    float r = ((double)view_cos[k] > 0.998) ? 2.5f : 4.0f; /* :430-436 */
    if (bFactor) r *= th;                                   /* :244 */
    const float rs = r * scale_factors[lvl];
    int nc = grid_query(&g, proj_x[k], proj_y[k], rs, lvl - 1, lvl, cand); /* :252-253 */
    ncand += nc;
    if (nc == 0) continue;
    const uint8_t* d_mp = mp_desc + 32 * (size_t)k;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int c = 0; c < nc; c++) {
      const int idx = cand[c];
      if (holder_obs[idx] > 0) continue; /* :273-275 */
      if (kuright[idx] > 0) {            /* :277-282 */
        const float er = fabsf(proj_xr[k] - kuright[idx]);
        if (er > r * scale_factors[lvl]) continue;
      }
      const int dist = orc_hamming256(d_mp, kdesc + 32 * (size_t)idx);
      if (dist < bestDist) { /* :289-301 */
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = koct[idx];
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = koct[idx];
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) { /* :305-312 */
      if (bestLevel == bestLevel2 && (double)bestDist > 0.8 * (double)bestDist2) continue;
      holder_obs[bestIdx] = mp_nobs[k];
      out_point_for_kp[bestIdx] = k;
      out_kp_for_point[k] = bestIdx;
      nmatches++;
    }
  }
  if (n_candidates) *n_candidates = ncand;
  free(holder_obs);
  free(cand);
  grid_free(&g);
  return nmatches;
}

/* cv::Mat 3x3 * 3x1 (+ c) float product as OpenCV's small-matrix gemm path
 * evaluates it: sequential fp32 multiply/add, k ascending, then + c
 * (pinned against cv2.gemm in tests/golden). */
static inline float dot3_seq(const float* a, int sa, const float* b, int sb) {
  float t = a[0] * b[0];
  t = t + a[sa] * b[sb];
  t = t + a[2 * sa] * b[2 * sb];
  return t;
}

/* Exposed for the golden test of the fp32 projection: xc = R*x + t. */
void orc_project_rt(const float* tcw, const float* xw, float* xc) {
  for (int r = 0; r < 3; r++) xc[r] = dot3_seq(tcw + 4 * r, 1, xw, 1) + tcw[4 * r + 3];
}

/* tlc of src/matcher.cpp:74-83: twc = -Rcw^T tcw ; tlc = Rlw*twc + tlw. */
void orc_tlc(const float* tcw_cur, const float* tcw_last, float* tlc) {
  float twc[3];
  float tc[3] = {tcw_cur[3], tcw_cur[7], tcw_cur[11]};
  for (int i = 0; i < 3; i++) twc[i] = -dot3_seq(tcw_cur + i, 4, tc, 1);
  for (int r = 0; r < 3; r++) tlc[r] = dot3_seq(tcw_last + 4 * r, 1, twc, 1) + tcw_last[4 * r + 3];
}

/* a8: Matcher::ComputeThreeMaxima src/matcher.cpp:387-428 on bin counts. */
void orc_three_maxima(const int* histo, int L, int* ind1, int* ind2, int* ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  int i1 = -1, i2 = -1, i3 = -1;
  for (int i = 0; i < L; i++) {
    const int s = histo[i];
    if (s > max1) {
      max3 = max2;
      max2 = max1;
      max1 = s;
      i3 = i2;
      i2 = i1;
      i1 = i;
    } else if (s > max2) {
      max3 = max2;
      max2 = s;
      i3 = i2;
      i2 = i;
    } else if (s > max3) {
      max3 = s;
      i3 = i;
    }
  }
  if ((float)max2 < 0.1f * (float)max1) {
    i2 = -1;
    i3 = -1;
  } else if ((float)max3 < 0.1f * (float)max1) {
    i3 = -1;
  }
  *ind1 = i1;
  *ind2 = i2;
  *ind3 = i3;
}

/* ---- a5: Matcher::SearchByProjection(Frame* Cur, Frame* Last, th),
 * src/matcher.cpp:64-218. */
int orc_search_proj_frame(int n_kp, const float* kx, const float* ky, const int* koct,
                          const float* kangle, const float* kuright, const uint8_t* kdesc,
                          const int* kclaim_obs, float min_x, float max_x, float min_y,
                          float max_y, const float* scale_factors, const float* tcw_cur,
                          const float* tcw_last, float fx, float fy, float cx, float cy, float mbf,
                          float mb, int n_last, const uint8_t* last_valid, const float* last_xw,
                          const int* last_octave, const float* last_angle, const uint8_t* mp_desc,
                          const int* mp_nobs, float th, int* out_kp_for_item,
                          int* out_state_for_kp, long long* n_candidates) {
  grid_t g;
  grid_build(&g, n_kp, kx, ky, koct, min_x, max_x, min_y, max_y);
  int* holder_obs = (int*)malloc(sizeof(int) * (size_t)(n_kp > 0 ? n_kp : 1));
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_kp > 0 ? n_kp : 1));
  for (int i = 0; i < n_kp; i++) {
    holder_obs[i] = kclaim_obs[i];
    out_state_for_kp[i] = -1;
  }
  /* rotation histogram :69-72 — entries are keypoint indices */
  int* hist_items = (int*)malloc(sizeof(int) * (size_t)(n_last > 0 ? n_last : 1));
  int* hist_bin = (int*)malloc(sizeof(int) * (size_t)(n_last > 0 ? n_last : 1));
  int hist_count[HISTO_LENGTH];
  memset(hist_count, 0, sizeof(hist_count));
  int n_hist = 0;
  const float factor = HISTO_LENGTH / 360.0f;

  float tlc[3];
  orc_tlc(tcw_cur, tcw_last, tlc);
  const int bForward = tlc[2] > mb;   /* :86 */
  const int bBackward = -tlc[2] > mb; /* :87 */

  long long ncand = 0;
  int nmatches = 0;
  for (int i = 0; i < n_last; i++) {
    out_kp_for_item[i] = -1;
    if (!last_valid[i]) continue; /* :91-95 */
    float x3Dc[3];
    orc_project_rt(tcw_cur, last_xw + 3 * (size_t)i, x3Dc); /* :99-101 */
    const float xc = x3Dc[0];
    const float yc = x3Dc[1];
    const float invzc = (float)(1.0 / (double)x3Dc[2]); /* :105 */
    if (invzc < 0) continue;                            /* :107 */
    float u = fx * xc * invzc + cx;                     /* :110 */
    float v = fy * yc * invzc + cy;                     /* :111 */
    if (u < min_x || u > max_x) continue;               /* :113 */
    if (v < min_y || v > max_y) continue;               /* :115 */
    const int nLastOctave = last_octave[i];
    const float radius = th * scale_factors[nLastOctave]; /* :121 */
    int nc;
    if (bForward) /* :129-134; default maxLevel = -1 (include/frame.h:73) */
      nc = grid_query(&g, u, v, radius, nLastOctave, -1, cand);
    else if (bBackward)
      nc = grid_query(&g, u, v, radius, 0, nLastOctave, cand);
    else
      nc = grid_query(&g, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
    ncand += nc;
    if (nc == 0) continue;
    const uint8_t* dMP = mp_desc + 32 * (size_t)i;
    int bestDist = 256, bestIdx2 = -1;
    for (int c = 0; c < nc; c++) {
      const int i2 = cand[c];
      if (holder_obs[i2] > 0) continue; /* :149-151 */
      if (kuright[i2] > 0) {            /* :153-160 */
        const float ur = u - mbf * invzc;
        const float er = fabsf(ur - kuright[i2]);
        if (er > radius) continue;
      }
      const int dist = orc_hamming256(dMP, kdesc + 32 * (size_t)i2);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx2 = i2;
      }
    }
    if (bestDist <= TH_HIGH) { /* :174-191 */
      holder_obs[bestIdx2] = mp_nobs[i];
      out_state_for_kp[bestIdx2] = i;
      out_kp_for_item[i] = bestIdx2;
      nmatches++;
      float rot = last_angle[i] - kangle[bestIdx2];
      if (rot < 0.0) rot += 360.0f;
      int bin = (int)roundf(rot * factor);
      if (bin == HISTO_LENGTH) bin = 0;
      hist_items[n_hist] = bestIdx2;
      hist_bin[n_hist] = bin;
      n_hist++;
      hist_count[bin]++;
    }
  }
  int ind1, ind2, ind3;
  orc_three_maxima(hist_count, HISTO_LENGTH, &ind1, &ind2, &ind3); /* :196-202 */
  for (int b = 0; b < HISTO_LENGTH; b++) {
    if (b != ind1 && b != ind2 && b != ind3) {
      for (int j = 0; j < n_hist; j++)
        if (hist_bin[j] == b) { /* :204-214 */
          out_state_for_kp[hist_items[j]] = -2;
          nmatches--;
        }
    }
  }
  if (n_candidates) *n_candidates = ncand;
  free(hist_items);
  free(hist_bin);
  free(holder_obs);
  free(cand);
  grid_free(&g);
  return nmatches;
}

/* ---- Frame::IsInFrustum src/frame.cpp:425-494 with
 * MapPoint::PredictScale src/map_point.cpp:267-284, GetMin/MaxDistanceInvariance :209-217.
 * ow = camera centre (translation of mTcw.inv(), given by the caller). */
void orc_frustum_project(const float* tcw, const float* ow, float fx, float fy, float cx, float cy,
                         float mbf, float min_x, float max_x, float min_y, float max_y, int n,
                         const float* xw, const float* normal, const float* min_dist,
                         const float* max_dist, float cos_limit, float log_sf, int n_levels,
                         uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int* level,
                         float* view_cos) {
  for (int i = 0; i < n; i++) {
    in_view[i] = 0;
    const float* P = xw + 3 * (size_t)i;
    float Pc[3];
    for (int r = 0; r < 3; r++) { /* 4x4 * 4x1 float gemm, sequential, no FMA */
      float t = tcw[4 * r] * P[0];
      t = t + tcw[4 * r + 1] * P[1];
      t = t + tcw[4 * r + 2] * P[2];
      t = t + tcw[4 * r + 3] * 1.0f;
      Pc[r] = t;
    }
    if (Pc[2] < 0.0f) continue;
    const float invz = 1.0f / Pc[2];
    const float u = fx * Pc[0] * invz + cx;
    const float v = fy * Pc[1] * invz + cy;
    if (u < min_x || u > max_x) continue;
    if (v < min_y || v > max_y) continue;
    const float maxDistance = 1.2f * max_dist[i];
    const float minDistance = 0.8f * min_dist[i];
    const float PO[3] = {P[0] - ow[0], P[1] - ow[1], P[2] - ow[2]};
    const float dist =
        (float)sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist < minDistance || dist > maxDistance) continue;
    const float* Pn = normal + 3 * (size_t)i;
    const float viewCos = (PO[0] * Pn[0] + PO[1] * Pn[1] + PO[2] * Pn[2]) / dist;
    if (viewCos < cos_limit) continue;
    const float ratio = max_dist[i] / dist;
    int nScale = (int)ceilf(logf(ratio) / log_sf);
    if (nScale < 0)
      nScale = 0;
    else if (nScale >= n_levels)
      nScale = n_levels - 1;
    in_view[i] = 1;
    proj_x[i] = u;
    proj_xr[i] = u - mbf * invz;
    proj_y[i] = v;
    level[i] = nScale;
    view_cos[i] = viewCos;
  }
}

/* ---- MapPoint::ComputeDescriptor src/map_point.cpp:69-129 for one
 * map point with m observation descriptors.  Returns the chosen index, writes its median. */
static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }
int orc_compute_descriptor(const uint8_t* desc, int m, int* median_out) {
  if (m <= 0) {
    *median_out = -1;
    return -1;
  }
  int* distances = (int*)malloc(sizeof(int) * (size_t)m * m);
  int* v = (int*)malloc(sizeof(int) * (size_t)m);
  for (int i = 0; i < m; i++) {
    distances[i * m + i] = 0;
    for (int j = i + 1; j < m; j++) {
      const int d = orc_hamming256(desc + 32 * (size_t)i, desc + 32 * (size_t)j); /* :88-100 */
      distances[i * m + j] = d;
      distances[j * m + i] = d;
    }
  }
  int bestIdx = 0, bestMedian = 0x7fffffff;
  for (int i = 0; i < m; i++) {
    memcpy(v, distances + (size_t)i * m, sizeof(int) * (size_t)m);
    qsort(v, (size_t)m, sizeof(int), cmp_int);
    const int median = v[(size_t)(0.5 * (m - 1))]; /* :115 */
    if (median < bestMedian) {
      bestMedian = median;
      bestIdx = i;
    }
  }
  free(distances);
  free(v);
  *median_out = bestMedian;
  return bestIdx;
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline must use every host
 * thread it can, so the caller sets the count explicitly. */
void orc_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
int orc_get_max_threads(void) { return omp_get_max_threads(); }

/* ---- multi-threaded driver for the CPU baseline of the sweep (bench.py
 * --impl reference): pairs are split over OpenMP threads. */
void orc_sweep(const uint8_t* bank, int n_desc, const int* pair_a, const int* pair_b, int n_pairs,
               int* kept, int* matches, int* min_dist) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int p = 0; p < n_pairs; p++)
    orc_sweep_pair(bank + (size_t)pair_a[p] * n_desc * 32, bank + (size_t)pair_b[p] * n_desc * 32,
                   n_desc, &kept[p], &matches[p], &min_dist[p]);
}

/* ---- Frame::ComputeStereoMatches src/frame.cpp:125-333, line by line on
 * flat arrays.  pyr_*[l] -> pixel (0,0) of level l (8-bit, `step` bytes per row), as
 * ORBextractor::mvImagePyramid holds them.  Returns the number of matches left after the
 * outlier step.  A correlation patch that would leave its level image makes the reference's
 * cv::Mat::rowRange/colRange throw; here that keypoint is skipped (same rule as the CUDA path). */
typedef struct { int first, second; } orc_pair_t;
static int cmp_pair(const void* a, const void* b) {
  const orc_pair_t* x = (const orc_pair_t*)a; const orc_pair_t* y = (const orc_pair_t*)b;
  if (x->first != y->first) return x->first < y->first ? -1 : 1;
  return (x->second > y->second) - (x->second < y->second);
}
int orc_stereo_matches(int n_levels, const int* lw, const int* lh, const int* lstep,
                       const uint8_t* const* pyr_left, const int* rw, const int* rh, const int* rstep,
                       const uint8_t* const* pyr_right, const float* mvScaleFactors,
                       const float* mvInvScaleFactors, float mbf, float mb, int N, const float* lx,
                       const float* ly, const int* loct, const uint8_t* ldesc, int Nr, const float* rx,
                       const float* ry, const int* roct, const uint8_t* rdesc, float* mvuRight,
                       float* mvDepth) {
  (void)n_levels; (void)rh;
  for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; } /* :127-128 */
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                              /* :130 */
  const int nRows = lh[0];                                                   /* :132 */
  /* :141-160 row table, right keypoints appended in ascending iR */
  int* rowCount = (int*)calloc((size_t)nRows + 1, sizeof(int));
  for (int iR = 0; iR < Nr; iR++) {
    const float kpY = ry[iR];
    const float r = 2.0f * mvScaleFactors[roct[iR]];
    const int maxr = (int)ceil(kpY + r), minr = (int)floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rowCount[yi + 1]++;
  }
  for (int y = 0; y < nRows; y++) rowCount[y + 1] += rowCount[y];
  int* rowItems = (int*)malloc(sizeof(int) * (size_t)(rowCount[nRows] > 0 ? rowCount[nRows] : 1));
  int* fill = (int*)malloc(sizeof(int) * (size_t)(nRows > 0 ? nRows : 1));
  for (int y = 0; y < nRows; y++) fill[y] = rowCount[y];
  for (int iR = 0; iR < Nr; iR++) {
    const float kpY = ry[iR];
    const float r = 2.0f * mvScaleFactors[roct[iR]];
    const int maxr = (int)ceil(kpY + r), minr = (int)floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rowItems[fill[yi]++] = iR;
  }
  const float minZ = mb, minD = 0;       /* :164-165 */
  const float maxD = mbf / minZ;         /* :166 */
  orc_pair_t* vDistIdx = (orc_pair_t*)malloc(sizeof(orc_pair_t) * (size_t)(N > 0 ? N : 1));
  int nDist = 0;
  for (int iL = 0; iL < N; iL++) {
    const int levelL = loct[iL];
    const float vL = ly[iL], uL = lx[iL];
    if (vL < 0.0f || (int)vL >= nRows) continue; /* vRowIndices[vL] would be out of range */
    const int row = (int)vL;                     /* :185 */
    if (rowCount[row + 1] == rowCount[row]) continue; /* :187 */
    const float minU = uL - maxD, maxU = uL - minD;   /* :190-191 */
    if (maxU < 0) continue;                           /* :193 */
    int bestDist = TH_HIGH;
    int bestIdxR = 0;
    const uint8_t* dL = ldesc + 32 * (size_t)iL;
    for (int c = rowCount[row]; c < rowCount[row + 1]; c++) {
      const int iR = rowItems[c];
      if (roct[iR] < levelL - 1 || roct[iR] > levelL + 1) continue; /* :209 */
      const float uR = rx[iR];
      if (uR >= minU && uR <= maxU) {
        const int dist = orc_hamming256(dL, rdesc + 32 * (size_t)iR);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDist < thOrbDist) { /* :231 */
      const float uR0 = rx[bestIdxR];
      const float scaleFactor = mvInvScaleFactors[levelL];
      const float scaleduL = roundf(uL * scaleFactor);
      const float scaledvL = roundf(vL * scaleFactor);
      const float scaleduR0 = roundf(uR0 * scaleFactor);
      const int w = 5, L = 5;
      const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
      const uint8_t* IL = pyr_left[levelL];
      const uint8_t* IR = pyr_right[levelL];
      /* where cv::Mat::rowRange/colRange of the left patch would assert (:240) */
      if (cv - w < 0 || cv + w >= lh[levelL] || cu - w < 0 || cu + w >= lw[levelL]) continue;
      float ILn[11][11];
      const float cL = (float)IL[(size_t)cv * lstep[levelL] + cu];
      for (int a = 0; a < 11; a++)
        for (int b = 0; b < 11; b++)
          ILn[a][b] = (float)IL[(size_t)(cv - w + a) * lstep[levelL] + cu - w + b] - cL; /* :241-242 */
      int bestDistS = INT_MAX, bestincR = 0;
      float vDists[11];
      const float iniu = scaleduR0 + L - w;     /* :252 */
      const float endu = scaleduR0 + L + w + 1; /* :253 */
      if (iniu < 0 || endu >= rw[levelL]) continue; /* :254 */
      if (cv + w >= rh[levelL] || cr - L - w < 0 || cr + L + w >= rw[levelL]) continue; /* asserts of :260 */
      for (int incR = -L; incR <= +L; incR++) {
        const float cR = (float)IR[(size_t)cv * rstep[levelL] + cr + incR];
        double s = 0;
        for (int a = 0; a < 11; a++)
          for (int b = 0; b < 11; b++) {
            const float v = (float)IR[(size_t)(cv - w + a) * rstep[levelL] + cr + incR - w + b] - cR;
            s += fabs((double)ILn[a][b] - (double)v); /* cv::norm(IL, IR, NORM_L1) :264 */
          }
        const float dist = (float)s;
        if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; } /* :265-269 */
        vDists[L + incR] = dist;
      }
      if (bestincR == -L || bestincR == L) continue; /* :276 */
      const float dist1 = vDists[L + bestincR - 1];
      const float dist2 = vDists[L + bestincR];
      const float dist3 = vDists[L + bestincR + 1];
      const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2)); /* :287 */
      if (deltaR < -1 || deltaR > 1) continue;                                        /* :290 */
      float bestuR = mvScaleFactors[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR); /* :297 */
      float disparity = (uL - bestuR);                                                /* :300 */
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
          disparity = 0.01;
          bestuR = uL - 0.01;
        }
        mvDepth[iL] = mbf / disparity;
        mvuRight[iL] = bestuR;
        vDistIdx[nDist].first = bestDistS;
        vDistIdx[nDist].second = iL;
        nDist++;
      }
    }
  }
  int kept = 0;
  if (nDist > 0) { /* :320-337 (the reference reads vDistIdx[0] of an empty vector otherwise) */
    qsort(vDistIdx, (size_t)nDist, sizeof(orc_pair_t), cmp_pair);
    const float median = vDistIdx[nDist / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    kept = nDist;
    for (int i = nDist - 1; i >= 0; i--) {
      if (vDistIdx[i].first < thDist) break;
      mvuRight[vDistIdx[i].second] = -1;
      mvDepth[vDistIdx[i].second] = -1;
      kept--;
    }
  }
  free(vDistIdx);
  free(fill);
  free(rowItems);
  free(rowCount);
  return kept;
}

/* ---- SURVEY 8(f) rank 5 (descriptor stage): IC_Angle + computeOrbDescriptor of
 * src/ORBextractor.cpp:79-149 for keypoints given in the coordinates of their pyramid level.
 * `pattern` (512 x,y pairs = bit_pattern_31_) and `umax` (16 entries) are the tables the
 * reference's ORBextractor constructor builds (:464-482); they are inputs here. */
static float orc_fast_atan2(float y, float x) { /* cv::fastAtan2, no FMA (pinned against cv2) */
  const float PI_F = (float)(180 / 3.1415926535897932384626433832795);
  const float p1 = 0.9997878412794807f * PI_F, p3 = -0.3258083974640975f * PI_F,
              p5 = 0.1555786518463281f * PI_F, p7 = -0.04432655554792128f * PI_F;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}
float orc_fast_atan2_export(float y, float x) { return orc_fast_atan2(y, x); }

void orc_orb_describe(int n_levels, const int* w, const int* h, const int* step_raw,
                      const uint8_t* const* raw, const int* step_blur, const uint8_t* const* blurred,
                      int n_kp, const float* kx, const float* ky, const int* klevel, const int* pattern,
                      const int* umax, float* out_angle, uint8_t* out_desc) {
  (void)n_levels; (void)w; (void)h;
  const int HALF_PATCH_SIZE = 15;
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f); /* :109 */
  for (int i = 0; i < n_kp; i++) {
    const int l = klevel[i];
    /* IC_Angle :79-106 */
    {
      int m_01 = 0, m_10 = 0;
      const int step = step_raw[l];
      const uint8_t* center = raw[l] + (size_t)lrintf(ky[i]) * step + lrintf(kx[i]);
      for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
      for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
          const int val_plus = center[u + v * step], val_minus = center[u - v * step];
          v_sum += (val_plus - val_minus);
          m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
      }
      out_angle[i] = orc_fast_atan2((float)m_01, (float)m_10);
    }
    /* computeOrbDescriptor :110-149 */
    {
      const float angle = out_angle[i] * factorPI;
      const float a = cosf(angle), b = sinf(angle); /* (float)cos(float): std::cos(float) overload */
      const int step = step_blur[l];
      const uint8_t* center = blurred[l] + (size_t)lrintf(ky[i]) * step + lrintf(kx[i]);
      const int* pat = pattern;
      for (int k = 0; k < 32; ++k, pat += 32) {
        int val = 0;
        for (int bit = 0; bit < 8; bit++) {
          const int x0 = pat[4 * bit], y0 = pat[4 * bit + 1], x1 = pat[4 * bit + 2], y1 = pat[4 * bit + 3];
          const int t0 = center[lrintf(x0 * b + y0 * a) * step + lrintf(x0 * a - y0 * b)];
          const int t1 = center[lrintf(x1 * b + y1 * a) * step + lrintf(x1 * a - y1 * b)];
          val |= (t0 < t1) << bit;
        }
        out_desc[32 * (size_t)i + k] = (uint8_t)val;
      }
    }
  }
}
