// orb_quadtree_gpu.cuh — ORBextractor::DistributeOctTree (reference src/ORBextractor.cpp:496-797) as ONE
// CTA per pyramid level, so that a frame never leaves the device between the FAST kernel and the
// descriptor kernel.  Same result as the sequential CPU restatement (oracle/orb_quadtree_ref.h, itself
// pinned to the compiled reference), bit for bit, including the list order of the surviving nodes.
//
// The reference walks a std::list and splits nodes one at a time, but what a whole pass does is a
// function of its input that can be evaluated in parallel:
//   * keys never need to move: every split is a stable partition, so the keys of a node are always in
//     their original (candidate) order; a key only carries the id of its current node;
//   * a round takes the current list L and a sequence Sel of nodes to split (a regular pass: every node
//     with more than one key, in list order; a "final" round: the nodes created in the previous round,
//     sorted by (size, creation id) descending, cut where the node budget N is reached) and produces
//         L' = reverse(children of Sel, in (Sel order, n1..n4) order, empty ones dropped) ++ (L minus Sel)
//     because the reference pushes children to the FRONT while it erases the parent;
//   * creation ids grow in that same (Sel order, n1..n4) order, which is the tie-break among nodes of
//     equal size (the reference compares node addresses, :695; under an allocator that never reuses
//     memory that is creation order, last created first -- how oracle/_ref runs the reference);
//   * a leaf's keypoint is its first key of maximal response: a 64-bit atomicMax on (response, ~index).
// Each of these is a histogram, a prefix sum or a rank-by-counting over at most a few thousand items:
// block-wide primitives, all state in global scratch (L1-resident), shared memory only for the scans.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace lorb {

constexpr int QT_THREADS = 1024;

struct QtLevel {
  const uint32_t* keys;  // packed score<<24 | y<<12 | x, in candidate order
  int n_keys;
  int width, height;     // of the bordered region (max_x - min_x, max_y - min_y)
  int n_features;        // N
  // scratch (device), sized by the host: see qt_scratch_ints()
  int* scratch;
  int node_cap;          // capacity of the node table
  // optional: the candidates still sit in the FAST kernel's per-cell slots (cells of this level in
  // (row, column) order, row-major inside a cell); they are flattened into keys_flat first and
  // `keys` / `n_keys` are ignored
  const uint32_t* slots;
  const int* cell_cnt;
  int n_cells, slot_cap;
  uint32_t* keys_flat;
  int key_cap;           // capacity the scratch was sized for (>= number of keys)
  // outputs
  int* out_index;        // [qt_out_cap] chosen key indices in list order
  int* out_count;        // device
  int* out_count_host;   // mapped pinned copy (may be NULL)
};

struct QtArgs {
  QtLevel lv[16];
};

// Node ids are never reused.  The list holds at most min(4 * (N + n_ini), K) nodes (a regular pass may
// overshoot N by splitting every node once more; every node holds a key), a round creates at most that
// many, and there are at most ~log2(side) + a few rounds.  Running out is reported (out_count = -1),
// never silent.
inline int qt_node_cap(int n_keys, int n_features, int n_ini) {
  const long long live = std::min<long long>(4ll * ((long long)n_features + n_ini), n_keys) + 8;
  return (int)(20 * live + n_ini + 64);
}

// Largest possible result: N + 3 (the careful phase stops at the first split that reaches N), except
// that the very first pass is unconditional and may turn the n_ini initial nodes into 4 * n_ini.
inline int qt_out_cap(int n_features, int n_ini) { return std::max(n_features + 3, 4 * n_ini) + 4; }

__host__ __device__ inline size_t qt_scratch_ints(int n_keys, int node_cap) {  // n_keys = key_cap
  // per key: node id, quadrant; per node: 10 fields, best (2 + alignment), two lists, three 4-wide slot arrays
  return (size_t)2 * n_keys + (size_t)27 * node_cap + 64;
}

// In-place exclusive prefix sum of data[0..n) by the whole CTA; returns the total (to every thread).
__device__ inline int qt_block_scan(int* data, int n, int* s_warp, int* s_total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + QT_THREADS - 1) / QT_THREADS;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; i++) sum += data[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < QT_THREADS / 32 ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    s_warp[lane] = winc - w;  // exclusive offset of each warp
    if (lane == 31) *s_total = winc;
  }
  __syncthreads();
  int run = s_warp[warp] + inc - sum;
  for (int i = lo; i < hi; i++) {
    const int v = data[i];
    data[i] = run;
    run += v;
  }
  __syncthreads();
  return *s_total;
}

__global__ void __launch_bounds__(QT_THREADS) orb_quadtree_kernel(QtArgs A) {
  const QtLevel L = A.lv[blockIdx.x];
  __shared__ int s_warp[32];
  __shared__ int s_total, s_misc[8];
  const int tid = threadIdx.x;
  const int N = L.n_features, cap = L.node_cap;
  const int n_ini = (int)roundf(__fdiv_rn((float)L.width, (float)L.height));
  // ---- carve the scratch
  int* q = L.scratch;
  int* key_node = q;            q += L.key_cap;
  int* key_quad = q;            q += L.key_cap;
  int* nx0 = q;                 q += cap;
  int* nx1 = q;                 q += cap;
  int* ny0 = q;                 q += cap;
  int* ny1 = q;                 q += cap;
  int* ncnt = q;                q += cap;
  int* nsel = q;                q += cap;  // rank in Sel, -1 if not selected
  int* nflag = q;               q += cap;  // 1 = no_more
  int* cand = q;                q += cap;  // candidate node ids of a final round
  int* cand_sorted = q;         q += cap;
  int* tmp = q;                 q += cap;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(q + ((reinterpret_cast<uintptr_t>(q) & 4) ? 1 : 0));
  q += 2 * cap + 2;
  int* list_a = q;              q += cap;
  int* list_b = q;              q += cap;
  int* ccnt = q;                q += 4 * cap;  // children of Sel: counts ...
  int* crank = q;               q += 4 * cap;  // ... creation ranks of the non-empty ones
  int* scan = q;                q += 4 * cap;  // general scan buffer

  // ---- candidates: given, or flattened from the per-cell slots
  const uint32_t* keys = L.keys;
  int K = L.n_keys;
  if (L.slots) {
    for (int i = tid; i < L.n_cells; i += QT_THREADS) scan[i] = min(L.cell_cnt[i], L.slot_cap);
    __syncthreads();
    K = qt_block_scan(scan, L.n_cells, s_warp, &s_total);
    if (K > L.key_cap) K = 0;  // cannot happen (key_cap = cells x slot_cap); reported through the count
    for (int c = tid >> 5; c < L.n_cells; c += QT_THREADS / 32) {
      const int n = min(L.cell_cnt[c], L.slot_cap), off = scan[c];
      for (int e = tid & 31; e < n && off + e < L.key_cap; e += 32)
        L.keys_flat[off + e] = L.slots[(size_t)c * L.slot_cap + e];
    }
    __syncthreads();
    keys = L.keys_flat;
  }
  if (K == 0 || n_ini < 1) {  // (the reference divides by zero for n_ini == 0)
    if (tid == 0) {
      *L.out_count = 0;
      if (L.out_count_host) *L.out_count_host = 0;
    }
    return;
  }
  const float h_x = __fdiv_rn((float)L.width, (float)n_ini);
  // ---- initial nodes (:571-594): key -> column (int)(x / hX)
  for (int i = tid; i < n_ini; i += QT_THREADS) {
    nx0[i] = (int)__fmul_rn(h_x, (float)i);
    nx1[i] = (int)__fmul_rn(h_x, (float)(i + 1));
    ny0[i] = 0;
    ny1[i] = L.height;
    ncnt[i] = 0;
  }
  __syncthreads();
  for (int k = tid; k < K; k += QT_THREADS) {
    const float x = (float)(keys[k] & 0xfff);
    const int col = min((int)__fdiv_rn(x, h_x), n_ini - 1);
    key_node[k] = col;
    atomicAdd(&ncnt[col], 1);
  }
  __syncthreads();
  // list = non-empty columns in order (:598-609); one key -> no_more
  for (int i = tid; i < n_ini; i += QT_THREADS) scan[i] = ncnt[i] > 0;
  __syncthreads();
  int size = qt_block_scan(scan, n_ini, s_warp, &s_total);
  for (int i = tid; i < n_ini; i += QT_THREADS) {
    nflag[i] = ncnt[i] == 1;
    nsel[i] = -1;
    if (ncnt[i] > 0) list_a[scan[i]] = i;
  }
  __syncthreads();
  int next_id = n_ini;
  int* list = list_a;
  int* list_new = list_b;
  int n_cand = 0;        // nodes created by the last round that can still be split
  bool final_phase = false, finish = false;

  while (!finish) {
    const int prev_size = size;
    int S;  // number of nodes to split this round
    if (!final_phase) {
      // regular pass: every node with more than one key, in list order
      for (int p = tid; p < size; p += QT_THREADS) scan[p] = nflag[list[p]] ? 0 : 1;
      __syncthreads();
      S = qt_block_scan(scan, size, s_warp, &s_total);
      for (int p = tid; p < size; p += QT_THREADS) {
        const int nd = list[p];
        if (!nflag[nd]) {
          nsel[nd] = scan[p];
          cand_sorted[scan[p]] = nd;
        }
      }
      __syncthreads();
    } else {
      // final round: candidates sorted by (size, creation id) descending (rank by counting)
      for (int i = tid; i < n_cand; i += QT_THREADS) {
        const int nd = cand[i], sz = ncnt[nd];
        int r = 0;
        for (int j = 0; j < n_cand; j++) {
          const int nj = cand[j], sj = ncnt[nj];
          r += (sj > sz) || (sj == sz && nj > nd);
        }
        cand_sorted[r] = nd;
        nsel[nd] = r;
      }
      S = n_cand;
      __syncthreads();
    }
    if (S == 0) break;  // nothing left to split: size == prev_size in the reference
    if (next_id + 4 * S > cap) {  // cannot happen for node_cap sized by the host; fail loudly
      if (tid == 0) {
        *L.out_count = -1;
        if (L.out_count_host) *L.out_count_host = -1;
      }
      return;
    }
    // ---- quadrant of every key of a selected node, child counts (DivideNode :496-551)
    for (int c = tid; c < 4 * S; c += QT_THREADS) ccnt[c] = 0;
    __syncthreads();
    for (int k = tid; k < K; k += QT_THREADS) {
      const int nd = key_node[k], s = nsel[nd];
      if (s < 0) continue;
      const uint32_t e = keys[k];
      const float x = (float)(e & 0xfff), y = (float)((e >> 12) & 0xfff);
      const int hx = (int)ceilf(__fdiv_rn((float)(nx1[nd] - nx0[nd]), 2.0f));
      const int hy = (int)ceilf(__fdiv_rn((float)(ny1[nd] - ny0[nd]), 2.0f));
      const int qd = (int)(x >= (float)(nx0[nd] + hx)) + 2 * (int)(y >= (float)(ny0[nd] + hy));
      key_quad[k] = qd;
      atomicAdd(&ccnt[4 * s + qd], 1);
    }
    __syncthreads();
    if (final_phase) {
      // split in sorted order until the list holds N nodes (:731-752): a split replaces one node
      // by its non-empty children
      for (int s = tid; s < S; s += QT_THREADS)
        scan[s] = (ccnt[4 * s] > 0) + (ccnt[4 * s + 1] > 0) + (ccnt[4 * s + 2] > 0) + (ccnt[4 * s + 3] > 0) - 1;
      __syncthreads();
      qt_block_scan(scan, S, s_warp, &s_total);  // scan[s] = growth before split s
      if (tid == 0) s_misc[0] = S;
      __syncthreads();
      for (int s = tid; s < S; s += QT_THREADS) {
        const int grow = (ccnt[4 * s] > 0) + (ccnt[4 * s + 1] > 0) + (ccnt[4 * s + 2] > 0) + (ccnt[4 * s + 3] > 0) - 1;
        if (size + scan[s] + grow >= N) atomicMin(&s_misc[0], s + 1);  // first split that reaches N
      }
      __syncthreads();
      const int S_cut = s_misc[0];
      __syncthreads();
      for (int s = S_cut + tid; s < S; s += QT_THREADS) {
        nsel[cand_sorted[s]] = -1;  // not split after all
        ccnt[4 * s] = ccnt[4 * s + 1] = ccnt[4 * s + 2] = ccnt[4 * s + 3] = 0;
      }
      S = S_cut;
      __syncthreads();
    }
    // ---- creation ranks of the non-empty children, in (Sel order, n1..n4) order
    for (int c = tid; c < 4 * S; c += QT_THREADS) crank[c] = ccnt[c] > 0;
    __syncthreads();
    const int M = qt_block_scan(crank, 4 * S, s_warp, &s_total);
    // ---- old list minus Sel keeps its order behind the new children
    for (int p = tid; p < size; p += QT_THREADS) scan[p] = nsel[list[p]] >= 0;
    __syncthreads();
    qt_block_scan(scan, size, s_warp, &s_total);
    for (int p = tid; p < size; p += QT_THREADS) {
      const int nd = list[p];
      if (nsel[nd] < 0) list_new[M + p - scan[p]] = nd;
    }
    // ---- the children: ids by creation rank, pushed to the front = reversed
    if (tid == 0) s_misc[1] = 0;
    __syncthreads();
    for (int c = tid; c < 4 * S; c += QT_THREADS) {
      const int n = ccnt[c];
      if (n == 0) continue;
      const int s = c >> 2, qd = c & 3, nd = cand_sorted[s], id = next_id + crank[c];
      const int hx = (int)ceilf(__fdiv_rn((float)(nx1[nd] - nx0[nd]), 2.0f));
      const int hy = (int)ceilf(__fdiv_rn((float)(ny1[nd] - ny0[nd]), 2.0f));
      const int xm = nx0[nd] + hx, ym = ny0[nd] + hy;
      nx0[id] = (qd & 1) ? xm : nx0[nd];
      nx1[id] = (qd & 1) ? nx1[nd] : xm;
      ny0[id] = (qd & 2) ? ym : ny0[nd];
      ny1[id] = (qd & 2) ? ny1[nd] : ym;
      ncnt[id] = n;
      nflag[id] = n == 1;
      nsel[id] = -1;
      list_new[M - 1 - crank[c]] = id;
      if (n > 1) atomicAdd(&s_misc[1], 1);
    }
    __syncthreads();
    for (int k = tid; k < K; k += QT_THREADS) {
      const int s = nsel[key_node[k]];
      if (s >= 0) key_node[k] = next_id + crank[4 * s + key_quad[k]];
    }
    __syncthreads();
    // candidates of a following final round: the children with more than one key, in creation order
    for (int r = tid; r < M; r += QT_THREADS) tmp[r] = ncnt[next_id + r] > 1;
    __syncthreads();
    const int to_expand = s_misc[1];
    qt_block_scan(tmp, M, s_warp, &s_total);
    for (int r = tid; r < M; r += QT_THREADS)
      if (ncnt[next_id + r] > 1) cand[tmp[r]] = next_id + r;
    n_cand = to_expand;
    size = M + size - S;
    next_id += M;
    int* t = list;
    list = list_new;
    list_new = t;
    __syncthreads();
    // ---- :675-753
    if (size >= N || size == prev_size) {
      finish = true;
    } else if (!final_phase && size + to_expand * 3 > N) {
      final_phase = true;
    }
  }
  // ---- each surviving node yields its first key of maximal response (:776-794)
  for (int i = tid; i < next_id; i += QT_THREADS) best[i] = 0ull;
  __syncthreads();
  for (int k = tid; k < K; k += QT_THREADS) {
    const unsigned long long v = ((unsigned long long)(keys[k] >> 24) << 32) | (0xffffffffu - (unsigned)k);
    atomicMax(&best[key_node[k]], v);
  }
  __syncthreads();
  for (int p = tid; p < size; p += QT_THREADS)
    L.out_index[p] = (int)(0xffffffffu - (unsigned)(best[list[p]] & 0xffffffffull));
  if (tid == 0) {
    *L.out_count = size;
    if (L.out_count_host) *L.out_count_host = size;
  }
}

}  // namespace lorb
