// orb_quadtree_ref.h — CPU restatement of ORBextractor::DistributeOctTree (reference
// src/ORBextractor.cpp:496-797), sequential like the reference.  TEST INFRASTRUCTURE ONLY (part of
// oracle/orb_ref.cpp): the oracle of the device quadtree (lorb_slam_b200/csrc/orb_quadtree_gpu.cuh);
// itself pinned to the compiled reference's DistributeOctTree (tests/golden/orb_golden.npz qt/*,
// tests/test_orb_cpu.py).
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace orc {

struct QKey {
  float x, y, response;
};

// ORBextractor::DistributeOctTree (:554-797) on an index-linked node list.
//   * the node list keeps the reference's order: children are pushed to the FRONT in the order
//     n1..n4 as their parent is erased, a pass walks from the (old) front to the back;
//   * a pass expands every node with more than one key; when the next pass could overshoot N
//     (size + 3*expandable > N) the nodes are expanded largest first instead (:687-753).  The
//     reference orders equal sizes by node ADDRESS (std::sort of (size, pointer) pairs); under an
//     allocator that never reuses memory that is creation order, which is the rule here (and how
//     oracle/_ref runs the reference): among equal sizes the node created LAST goes first;
//   * each surviving node yields its first key of maximal response (:776-794).
// Keys of a node are a contiguous run of `perm`, children are a stable 4-way partition of it.
struct QNode {
  int x0, x1, y0, y1;
  int k0, k1;      // keys perm[k0 .. k1)
  int prev, next;  // list links (-1 = none)
  bool no_more;
};

struct QItem {
  float x, y, response;
  int idx;  // position in the input
};

static void distribute_quadtree(const std::vector<QKey>& keys, int min_x, int max_x, int min_y, int max_y, int N,
                                std::vector<int>* result) {
  result->clear();
  const int n_keys = (int)keys.size();
  const int n_ini = (int)roundf((float)(max_x - min_x) / (max_y - min_y));
  if (n_ini < 1 || n_keys == 0) return;  // (the reference divides by zero for n_ini == 0)
  const float h_x = (float)(max_x - min_x) / n_ini;
  // the keys travel with their node: items[k0..k1) of a node are contiguous and keep input order
  // scratch is kept per thread: a frame calls this once per level, every frame
  static thread_local std::vector<QItem> items, moved;
  static thread_local std::vector<uint8_t> quad;
  static thread_local std::vector<QNode> nodes;
  if ((int)items.size() < n_keys) {
    items.resize(n_keys);
    moved.resize(n_keys);
    quad.resize(n_keys);
  }
  nodes.clear();
  int head = -1, tail = -1, size = 0;
  auto push_back = [&](int id) {
    nodes[id].prev = tail;
    nodes[id].next = -1;
    if (tail >= 0) nodes[tail].next = id; else head = id;
    tail = id;
    ++size;
  };
  auto push_front = [&](int id) {
    nodes[id].prev = -1;
    nodes[id].next = head;
    if (head >= 0) nodes[head].prev = id; else tail = id;
    head = id;
    ++size;
  };
  auto erase = [&](int id) {  // returns the next node
    const int p = nodes[id].prev, nx = nodes[id].next;
    if (p >= 0) nodes[p].next = nx; else head = nx;
    if (nx >= 0) nodes[nx].prev = p; else tail = p;
    --size;
    return nx;
  };
  // initial nodes (:571-594): keys go to column (int)(x / hX), in input order
  {
    std::vector<int> cnt(n_ini + 1, 0), col(n_keys);
    for (int i = 0; i < n_keys; i++) {
      col[i] = std::min((int)(keys[i].x / h_x), n_ini - 1);
      cnt[col[i] + 1]++;
    }
    for (int i = 0; i < n_ini; i++) cnt[i + 1] += cnt[i];
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < n_keys; i++) items[fill[col[i]]++] = QItem{keys[i].x, keys[i].y, keys[i].response, i};
    for (int i = 0; i < n_ini; i++) {
      QNode nd;
      nd.x0 = (int)(h_x * (float)i);
      nd.x1 = (int)(h_x * (float)(i + 1));
      nd.y0 = 0;
      nd.y1 = max_y - min_y;
      nd.k0 = cnt[i];
      nd.k1 = cnt[i + 1];
      nd.no_more = false;
      nodes.push_back(nd);
      push_back((int)nodes.size() - 1);
    }
  }
  for (int id = head; id >= 0;) {  // :598-609
    const int nk = nodes[id].k1 - nodes[id].k0;
    if (nk == 1) {
      nodes[id].no_more = true;
      id = nodes[id].next;
    } else if (nk == 0) {
      id = erase(id);
    } else {
      id = nodes[id].next;
    }
  }
  // DivideNode (:496-551) + the four push_front blocks; appends children with > 1 keys to `grown`.
  // Quadrants 0..3 = n1 (left/top), n2 (right/top), n3 (left/bottom), n4 (right/bottom).
  std::vector<std::pair<int, int>> grown, prev_grown;  // (size, node id = creation order)
  auto divide = [&](int id) {
    const QNode nd = nodes[id];
    const int half_x = (int)ceilf((float)(nd.x1 - nd.x0) / 2), half_y = (int)ceilf((float)(nd.y1 - nd.y0) / 2);
    const int xm = nd.x0 + half_x, ym = nd.y0 + half_y;
    const float fxm = (float)xm, fym = (float)ym;
    // stable 4-way split: count (branch-free), then scatter
    int c1 = 0, c2 = 0, c3 = 0;
    for (int k = nd.k0; k < nd.k1; k++) {
      const int right = items[k].x >= fxm, low = items[k].y >= fym;
      quad[k] = (uint8_t)(right + 2 * low);
      c1 += right & (1 - low);
      c2 += (1 - right) & low;
      c3 += right & low;
    }
    const int cnt[4] = {nd.k1 - nd.k0 - c1 - c2 - c3, c1, c2, c3};
    int start[4], fill[4];
    start[0] = nd.k0;
    for (int q = 1; q < 4; q++) start[q] = start[q - 1] + cnt[q - 1];
    for (int q = 0; q < 4; q++) fill[q] = start[q];
    for (int k = nd.k0; k < nd.k1; k++) moved[fill[quad[k]]++] = items[k];
    std::copy(moved.begin() + nd.k0, moved.begin() + nd.k1, items.begin() + nd.k0);
    const int bx0[4] = {nd.x0, xm, nd.x0, xm}, bx1[4] = {xm, nd.x1, xm, nd.x1};
    const int by0[4] = {nd.y0, nd.y0, ym, ym}, by1[4] = {ym, ym, nd.y1, nd.y1};
    int expandable = 0;
    for (int q = 0; q < 4; q++) {
      if (cnt[q] == 0) continue;
      QNode ch;
      ch.x0 = bx0[q];
      ch.x1 = bx1[q];
      ch.y0 = by0[q];
      ch.y1 = by1[q];
      ch.k0 = start[q];
      ch.k1 = start[q] + cnt[q];
      ch.no_more = cnt[q] == 1;
      nodes.push_back(ch);
      const int cid = (int)nodes.size() - 1;
      push_front(cid);
      if (cnt[q] > 1) {
        ++expandable;
        grown.emplace_back(cnt[q], cid);
      }
    }
    return expandable;
  };
  bool finish = false;
  while (!finish) {
    const int prev_size = size;
    int to_expand = 0;
    grown.clear();
    for (int id = head; id >= 0;) {
      if (nodes[id].no_more) {
        id = nodes[id].next;
        continue;
      }
      to_expand += divide(id);
      id = erase(id);
    }
    if (size >= N || size == prev_size) {
      finish = true;
    } else if (size + to_expand * 3 > N) {
      while (!finish) {
        const int prev_size2 = size;
        prev_grown = grown;
        grown.clear();
        std::sort(prev_grown.begin(), prev_grown.end());  // (size, creation order) ascending
        for (int j = (int)prev_grown.size() - 1; j >= 0; j--) {
          divide(prev_grown[j].second);
          erase(prev_grown[j].second);
          if (size >= N) break;
        }
        if (size >= N || size == prev_size2) finish = true;
      }
    }
  }
  result->reserve(size);
  for (int id = head; id >= 0; id = nodes[id].next) {
    int best = nodes[id].k0;
    for (int k = nodes[id].k0 + 1; k < nodes[id].k1; k++)
      if (items[k].response > items[best].response) best = k;
    result->push_back(items[best].idx);
  }
}

}  // namespace orc
