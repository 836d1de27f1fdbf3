// match_tc.cu — the keyframe-pair matching sweep on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by 1-D TMA bulk copies).
//
// What it computes is exactly what sweep_kernel (match_bf.cu) computes: for every keyframe
// pair the cv::BFMatcher(NORM_HAMMING, crossCheck=true) matches (reference
// src/matcher.cpp:36-39), minDist (:42-47) and the number that survive
// `distance > max(2*minDist, 30.0)` (:49-56) -- bit for bit, including the lowest-index
// tie-break on both sides.
//
// How.  A 256-bit descriptor becomes a row of 256 signed bytes, +32 for a set bit and -32 for
// a clear one, so that  sum_k a_k b_k = 1024 * (256 - 2 * hamming(a, b)).  Three more 16-byte
// columns carry (X | Y | X) with X = (-16, -1, 0...) and Y = (idx >> 4, idx & 15, 0...): the
// streamed (M) side reads (X, Y), the resident (N) side reads (Y, X), which adds  -i - j  to
// the int32 accumulator.  Hence
//     acc[i][j] = 1024 * (256 - 2 d(i, j)) - i - j        (exact: int8 x int8 -> int32)
// and  max_j acc[i][j]  is "smallest distance, then lowest j" for row i while  max_i acc[i][j]
// is "smallest distance, then lowest i" for column j: both argmins of the cross-check are
// plain integer maxima of the accumulator tile, one VIMNMX / VIMNMX3 per element and side.
//
// Work decomposition.  A keyframe's operand image is [n_pad / 8 row groups][19 core matrices]
// [8 rows][16 bytes] (K-major, no swizzle: SBO = 2432 B, LBO = 128 B), so any run of row
// groups is a valid UMMA operand and one contiguous TMA bulk copy.  A unit of work is
// (keyframe pair (a, b), slab s): 256 descriptors of `a` stay resident in shared memory as the
// N operand (77 824 B) while the 128-row tiles of `b` stream through a 3-stage ring as the M
// operand (38 912 B each); one tile is 9 MMAs of 128 x 256 x 32.  Per CTA: warp 0 = TMA
// producer, warp 1 = MMA issuer (one thread), warps 2..3 = flusher (finished units -> global
// memory, off the epilogue's path), warps 4..11 = epilogue (TMEM -> registers, row
// maxima to shared memory, column maxima kept in 128 registers per thread across the tiles of
// a unit), TMEM double-buffered (2 x 256 columns).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lorb {
namespace tc {

constexpr int KC = 19;                       // core matrices (16 B of K each) per row group
constexpr int RG_BYTES = KC * 128;           // 8 descriptors
constexpr int MT_ROWS = 128;                 // streamed tile (UMMA M)
constexpr int SLAB_COLS = 256;               // resident slab (UMMA N)
constexpr int MTILE_BYTES = MT_ROWS / 8 * RG_BYTES;   // 38 912
constexpr int SLAB_BYTES = SLAB_COLS / 8 * RG_BYTES;  // 77 824
constexpr int STAGES = 3;
constexpr int MAX_PAD = 2048;
constexpr int KEY_OFFSET = 1024 * 256;       // acc + i + j = KEY_OFFSET - 2048 d ... see decode
constexpr int TC_THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int INT_LOWEST = -2147483647 - 1;

// smem carve-up (bytes)
constexpr int OFF_SLAB = 0;
constexpr int OFF_RING = OFF_SLAB + SLAB_BYTES;
constexpr int OFF_ROWACC = OFF_RING + STAGES * MTILE_BYTES;  // 2 x MAX_PAD ints
constexpr int OFF_COLACC = OFF_ROWACC + 2 * MAX_PAD * 4;     // 2 x SLAB_COLS ints
constexpr int OFF_BARS = OFF_COLACC + 2 * SLAB_COLS * 4;     // 16 mbarriers
constexpr int OFF_TMEM = OFF_BARS + 16 * 8;
constexpr int TC_SMEM_BYTES = OFF_TMEM + 16;

// ---------------------------------------------------------------- operand images
// One thread writes one 16-byte piece: (keyframe, row group g, core matrix kc, row r).
__global__ void __launch_bounds__(256)
    tc_expand_kernel(const uint8_t* __restrict__ bank, int n_kf, int n_desc, int n_pad,
                     uint4* __restrict__ img) {
  const long long pieces_per_kf = (long long)(n_pad / 8) * KC * 8;
  const long long total = pieces_per_kf * n_kf;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total;
       id += (long long)gridDim.x * blockDim.x) {
    const int kf = (int)(id / pieces_per_kf);
    const int rem = (int)(id % pieces_per_kf);
    const int r = rem & 7, kc = (rem >> 3) % KC, g = (rem >> 3) / KC;
    const int x = g * 8 + r;  // index of this row inside the keyframe
    uint32_t w[4];
    if (kc < 16) {
      const int src = min(x, n_desc - 1);  // padding rows replicate the last descriptor
      const uint8_t* d = bank + ((size_t)kf * n_desc + src) * 32 + kc * 2;
      const uint32_t bits = (uint32_t)d[0] | ((uint32_t)d[1] << 8);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const uint32_t bit = (bits >> (4 * q + e)) & 1u;
          v |= (bit ? 0x20u : 0xE0u) << (8 * e);  // +32 / -32
        }
        w[q] = v;
      }
    } else if (kc == 17) {
      w[0] = (uint32_t)(x >> 4) | ((uint32_t)(x & 15) << 8);
      w[1] = w[2] = w[3] = 0;
    } else {
      w[0] = 0xF0u | (0xFFu << 8);  // (-16, -1)
      w[1] = w[2] = w[3] = 0;
    }
    img[id] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void tc_fill_kernel(int* __restrict__ p, long long n, int v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// ---------------------------------------------------------------- the sweep
__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_sweep_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ img_b, long long kf_bytes,
                    const int4* __restrict__ units,
                    int n_units, int n_mtiles, int n_pad, int* __restrict__ rowkey,
                    int* __restrict__ colkey) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint8_t* slab = smem + OFF_SLAB;
  uint8_t* ring = smem + OFF_RING;
  int* rowacc = reinterpret_cast<int*>(smem + OFF_ROWACC);
  int* colacc = reinterpret_cast<int*>(smem + OFF_COLACC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full = bars;            // [STAGES] TMA -> MMA
  uint64_t* empty = bars + 3;       // [STAGES] MMA -> TMA
  uint64_t* res_full = bars + 6;    // slab landed
  uint64_t* res_empty = bars + 7;   // every MMA reading the slab has completed
  uint64_t* tmem_full = bars + 8;   // [2] MMA -> epilogue
  uint64_t* tmem_empty = bars + 10; // [2] epilogue -> MMA
  uint64_t* acc_full = bars + 12;   // [2] epilogue -> flusher: the unit's row / column maxima are complete
  uint64_t* acc_empty = bars + 14;  // [2] flusher -> epilogue: the buffer has been flushed and reset
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = (int)(((long long)n_units * blockIdx.x) / gridDim.x);
  const int u1 = (int)(((long long)n_units * (blockIdx.x + 1)) / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(res_full, 1);
    mbar_init(res_empty, 1);
    for (int b = 0; b < 2; b++) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], EPI_THREADS / 32);
      mbar_init(&acc_full[b], EPI_THREADS / 32);
      mbar_init(&acc_empty[b], 2);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * MAX_PAD; i += TC_THREADS) rowacc[i] = INT_LOWEST;
  for (int i = tid; i < 2 * SLAB_COLS; i += TC_THREADS) colacc[i] = INT_LOWEST;
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer
      uint32_t g = 0, t = 0;
      int4 nxt = u0 < u1 ? units[u0] : make_int4(0, 0, 0, 0);
      for (int u = u0; u < u1; u++) {
        const int4 un = nxt;  // (a, b, slab | first-of-group << 8, pair)
        if (u + 1 < u1) nxt = units[u + 1];
        if (u == u0 || (un.z >> 8)) {
          mbar_wait_parity(res_empty, (g & 1u) ^ 1u);
          mbar_arrive_expect_tx(res_full, SLAB_BYTES);
          tma_load_1d(slab, img + (size_t)un.x * kf_bytes + (size_t)(un.z & 255) * SLAB_BYTES, SLAB_BYTES,
                      res_full);
          g++;
        }
        const uint8_t* src = img_b + (size_t)un.y * kf_bytes;
        for (int m = 0; m < n_mtiles; m++, t++) {
          const uint32_t stage = t % STAGES, k = t / STAGES;
          mbar_wait_parity(&empty[stage], (k & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full[stage], MTILE_BYTES);
          tma_load_1d(ring + stage * MTILE_BYTES, src + (size_t)m * MTILE_BYTES, MTILE_BYTES, &full[stage]);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer
      constexpr uint32_t IDESC = idesc_s8(MT_ROWS, SLAB_COLS);
      const uint32_t slab_addr = smem_u32(slab), ring_addr = smem_u32(ring);
      uint32_t g = 0, t = 0;
      int4 nxt = u0 < u1 ? units[u0] : make_int4(0, 0, 0, 0);
      for (int u = u0; u < u1; u++) {
        const int4 un = nxt;
        if (u + 1 < u1) nxt = units[u + 1];
        if (u == u0 || (un.z >> 8)) {
          mbar_wait_parity(res_full, g & 1u);
          g++;
        }
        for (int m = 0; m < n_mtiles; m++, t++) {
          const uint32_t acc = t & 1u, ka = t >> 1;
          const uint32_t stage = t % STAGES, ks = t / STAGES;
          mbar_wait_parity(&tmem_empty[acc], (ka & 1u) ^ 1u);
          mbar_wait_parity(&full[stage], ks & 1u);
          fence_after_sync();
          const uint32_t a_addr = ring_addr + stage * MTILE_BYTES;
          const uint32_t d_addr = tmem_base + acc * SLAB_COLS;
#pragma unroll
          for (int k = 0; k < 8; k++)
            mma_s8(d_addr, smem_desc(a_addr + k * 256, 128, RG_BYTES),
                   smem_desc(slab_addr + k * 256, 128, RG_BYTES), IDESC, k > 0 ? 1u : 0u);
          // index columns: M side reads (X, Y) at core matrices 16, 17; N side (Y, X) at 17, 18
          mma_s8(d_addr, smem_desc(a_addr + 16 * 128, 128, RG_BYTES),
                 smem_desc(slab_addr + 17 * 128, 128, RG_BYTES), IDESC, 1u);
          mma_commit(&empty[stage]);
          mma_commit(&tmem_full[acc]);
        }
        if (u + 1 == u1 || (nxt.z >> 8)) mma_commit(res_empty);  // last unit on this slab
      }
    } else if (warp >= 2) {
      // ===================== flusher (warps 2, 3): hands a finished unit's maxima to global memory
      // while the epilogue warps are already on the next unit (the buffers alternate): this slab's
      // 256 columns are final; rows are merged over the slabs by RED.MAX
      const int ft = tid - 64;  // 0..63
      for (int u = u0; u < u1; u++) {
        const int4 un = units[u];
        const int buf = (u - u0) & 1;
        mbar_wait_parity(&acc_full[buf], (uint32_t)(((u - u0) >> 1) & 1));
        int* racc = rowacc + buf * MAX_PAD;
        int* cacc = colacc + buf * SLAB_COLS;
        const size_t po = (size_t)un.w * n_pad;
        for (int j = ft; j < SLAB_COLS; j += 64) {
          colkey[po + (size_t)(un.z & 255) * SLAB_COLS + j] = cacc[j];
          cacc[j] = INT_LOWEST;
        }
        for (int i = ft; i < n_mtiles * MT_ROWS; i += 64) {
          atomicMax(&rowkey[po + i], racc[i]);
          racc[i] = INT_LOWEST;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    }
  } else {
    // ===================== epilogue: 8 warps; quarter q owns TMEM lanes [32q, 32q+32),
    // half h the accumulator columns [128h, 128h+128)
    reg_inc<232>();
    const int ew = warp - 4, q = ew & 3, h = ew >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 128);
    uint32_t t = 0;
    for (int u = u0; u < u1; u++) {
      const int buf = (u - u0) & 1;
      int* racc = rowacc + buf * MAX_PAD;
      int* cacc = colacc + buf * SLAB_COLS;
      // the flusher has emptied this buffer (it was last used two units ago)
      mbar_wait_parity(&acc_empty[buf], (uint32_t)((((u - u0) >> 1) & 1) ^ 1));
      int colmax[128];
#pragma unroll
      for (int k = 0; k < 128; k++) colmax[k] = INT_LOWEST;
      for (int m = 0; m < n_mtiles; m++, t++) {
        const uint32_t acc = t & 1u, ka = t >> 1;
        mbar_wait_parity(&tmem_full[acc], ka & 1u);
        fence_after_sync();
        const uint32_t taddr = lane_base + acc * SLAB_COLS;
        int rmax = INT_LOWEST;
        int v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tmem_ld_wait(v0);
        tmem_ld32(taddr + 32, v1);
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          rmax = max3(rmax, v0[k], v0[k + 1]);
          colmax[k] = max(colmax[k], v0[k]);
          colmax[k + 1] = max(colmax[k + 1], v0[k + 1]);
        }
        tmem_ld_wait(v1);
        tmem_ld32(taddr + 64, v0);
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          rmax = max3(rmax, v1[k], v1[k + 1]);
          colmax[32 + k] = max(colmax[32 + k], v1[k]);
          colmax[32 + k + 1] = max(colmax[32 + k + 1], v1[k + 1]);
        }
        tmem_ld_wait(v0);
        tmem_ld32(taddr + 96, v1);
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          rmax = max3(rmax, v0[k], v0[k + 1]);
          colmax[64 + k] = max(colmax[64 + k], v0[k]);
          colmax[64 + k + 1] = max(colmax[64 + k + 1], v0[k + 1]);
        }
        tmem_ld_wait(v1);
        // every load of this warp from the accumulator buffer has completed: hand it back
        fence_before_sync();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          rmax = max3(rmax, v1[k], v1[k + 1]);
          colmax[96 + k] = max(colmax[96 + k], v1[k]);
          colmax[96 + k + 1] = max(colmax[96 + k + 1], v1[k + 1]);
        }
        atomicMax(&racc[m * MT_ROWS + q * 32 + lane], rmax);  // the other half adds its columns
      }
      // column maxima over the warp's 32 rows, then over the four quarters through shared memory
      int keep[4] = {INT_LOWEST, INT_LOWEST, INT_LOWEST, INT_LOWEST};
#pragma unroll
      for (int k = 0; k < 128; k++) {
        const int r = __reduce_max_sync(0xffffffffu, colmax[k]);
        if (lane == (k & 31)) keep[k >> 5] = r;
      }
#pragma unroll
      for (int c = 0; c < 4; c++) atomicMax(&cacc[h * 128 + c * 32 + lane], keep[c]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_full[buf]);  // release: this warp's shared-memory atomics are done
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Cross-check, minDist and the max(2*minDist, 30) filter of one keyframe pair from its row /
// column keys (reference src/matcher.cpp:42-56).  Rows are b's descriptors, columns a's; the
// three numbers are symmetric in the roles.  Resets the row keys for the next launch.
__global__ void __launch_bounds__(256)
    tc_finalize_kernel(int* __restrict__ rowkey, const int* __restrict__ colkey, int n_desc, int n_pad,
                       int* __restrict__ out) {
  __shared__ int s_min[8], s_cnt[8], s_res[2];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* rk = rowkey + (size_t)p * n_pad;
  const int* ck = colkey + (size_t)p * n_pad;
  // the kernel is two dependent memory round trips (row key -> column -> column key): all eight
  // row keys of a thread are fetched before the first is used, then all eight column keys
  // (one iteration at a time this took 0.78 ms per 8128 pairs, 14 % of a sweep step)
  int key[8], ckey[8], dd[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int i = tid + k * 256;
    key[k] = i < n_pad ? __ldcg(rk + i) : INT_LOWEST;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int i = tid + k * 256;
    const int j = (KEY_OFFSET - (key[k] + i)) & 2047;  // 2048 d + j
    ckey[k] = i < n_desc ? __ldcg(ck + j) : 0;
  }
  int cnt = 0, lmin = 1 << 30;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int i = tid + k * 256;
    dd[k] = -1;
    if (i < n_pad) rk[i] = INT_LOWEST;
    if (i < n_desc) {
      const int ur = KEY_OFFSET - (key[k] + i);
      const int j = ur & 2047;
      const int uc = KEY_OFFSET - (ckey[k] + j);  // 2048 d' + i'
      if ((uc & 2047) == i) {
        dd[k] = ur >> 11;
        cnt++;
        lmin = min(lmin, dd[k]);
      }
    }
  }
  lmin = __reduce_min_sync(0xffffffffu, lmin);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) {
    s_min[warp] = lmin;
    s_cnt[warp] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    int v = 1 << 30, c = 0;
    for (int w = 0; w < 8; w++) {
      v = min(v, s_min[w]);
      c += s_cnt[w];
    }
    s_res[0] = v;
    s_res[1] = c;
  }
  __syncthreads();
  const int min_dist = s_res[0], n_match = s_res[1];
  const int thr = max(2 * min_dist, 30);
  int kept = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) kept += (dd[k] >= 0 && !(dd[k] > thr)) ? 1 : 0;
  kept = __reduce_add_sync(0xffffffffu, kept);
  __syncthreads();
  if (lane == 0) s_cnt[warp] = kept;
  __syncthreads();
  if (tid == 0) {
    int ksum = 0;
    for (int w = 0; w < 8; w++) ksum += s_cnt[w];
    out[3 * (size_t)p + 0] = ksum;
    out[3 * (size_t)p + 1] = n_match;
    out[3 * (size_t)p + 2] = n_match > 0 ? min_dist : -1;
  }
}

// ---------------------------------------------------------------- tensor-pipe micro-benchmark
// The sweep's own MMA stream (9 x tcgen05.mma kind::i8 128x256x32 per tile, two accumulator
// buffers) on zeroed operands with no TMA and no epilogue: the rate at which this instruction
// mix can possibly run on this part.
__global__ void __launch_bounds__(128, 1) tc_mma_bench_kernel(int tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SLAB_BYTES + MTILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (SLAB_BYTES + MTILE_BYTES) / 16; i += 128)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();  // generic-proxy zero fill -> async-proxy (tensor core) reads
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    constexpr uint32_t IDESC = idesc_s8(MT_ROWS, SLAB_COLS);
    const uint32_t b_addr = smem_u32(smem), a_addr = b_addr + SLAB_BYTES;
    for (int t = 0; t < tiles; t++) {
      const uint32_t d_addr = tmem_base + (t & 1) * SLAB_COLS;
#pragma unroll
      for (int k = 0; k < 9; k++)
        mma_s8(d_addr, smem_desc(a_addr + k * 256, 128, RG_BYTES), smem_desc(b_addr + k * 256, 128, RG_BYTES),
               IDESC, k > 0 ? 1u : 0u);
    }
    mma_commit(bar);
    mbar_wait_parity(bar, 0);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
int pad_of(int n_desc) { return (n_desc + SLAB_COLS - 1) / SLAB_COLS * SLAB_COLS; }
size_t image_bytes(int n_desc) { return (size_t)(pad_of(n_desc) / 8) * RG_BYTES; }

// Operand images of the resident bank (call after the bank has been uploaded).
int bank_expand(lorb_ctx* c) {
  const int n_kf = c->bank_n_kf, n_desc = c->bank_n_desc;
  const int n_pad = pad_of(n_desc);
  const size_t bytes = image_bytes(n_desc) * (size_t)n_kf;
  LORB_TRY(c->tc_img.reserve(bytes));
  const long long pieces = (long long)(bytes / 16);
  const int grid = (int)std::min<long long>((pieces + 255) / 256, (long long)c->sm_count * 32);
  LORB_LAUNCH(c, tc_expand_kernel, grid, 256, 0, c->bank.as<uint8_t>(), n_kf, n_desc, n_pad,
              c->tc_img.as<uint4>());
  c->tc_img_n_kf = n_kf;
  c->tc_img_n_desc = n_desc;
  return LORB_OK;
}

// Unit list of a pair plan: for every run of equal `a`, slab by slab, the pairs of the run.
// A unit is (a, b, slab | first-of-its-(a, slab)-group << 8, pair index).
// A launch handles at most TC_CHUNK_PAIRS keyframe pairs (its key scratch is 16 KB per pair); a longer
// plan runs as several launches over consecutive chunks of the pair list.
constexpr int TC_CHUNK_PAIRS = 16384;

int plan_build(lorb_ctx* c, const int* pa, const int* pb, int n_pairs) {
  const int n_pad = pad_of(c->bank_n_desc), n_slabs = n_pad / SLAB_COLS;
  std::vector<int4> units;
  units.reserve((size_t)n_pairs * n_slabs);
  c->tc_chunk_units.clear();
  for (int c0 = 0; c0 < n_pairs; c0 += TC_CHUNK_PAIRS) {
    const int c1 = std::min(n_pairs, c0 + TC_CHUNK_PAIRS);
    c->tc_chunk_units.push_back((long long)units.size());
    for (int p0 = c0; p0 < c1;) {
      int p1 = p0 + 1;
      while (p1 < c1 && pa[p1] == pa[p0]) p1++;
      for (int s = 0; s < n_slabs; s++)
        for (int p = p0; p < p1; p++) units.push_back(make_int4(pa[p], pb[p], s | (p == p0 ? 256 : 0), p - c0));
      p0 = p1;
    }
  }
  c->tc_chunk_units.push_back((long long)units.size());
  c->tc_n_units = (long long)units.size();
  if (units.empty()) return LORB_OK;
  LORB_TRY(c->tc_units.reserve(units.size() * sizeof(int4)));
  LORB_CUDA_TRY(cudaMemcpyAsync(c->tc_units.p, units.data(), units.size() * sizeof(int4),
                                cudaMemcpyHostToDevice, c->stream));
  // row / column keys of one chunk: [pairs][n_pad] each; the row keys start (and are left by the
  // finalize kernel) at INT_MIN
  const size_t keys = (size_t)std::min(n_pairs, TC_CHUNK_PAIRS) * n_pad;
  LORB_TRY(c->tc_keys.reserve(2 * keys * 4));
  LORB_LAUNCH(c, tc_fill_kernel, c->sm_count * 8, 256, 0, c->tc_keys.as<int>(), (long long)keys, INT_LOWEST);
  c->tc_keys_rows = keys;
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));  // `units` goes out of scope
  return LORB_OK;
}

int launch_sweep(lorb_ctx* c, int kf_base_a, int kf_base_b, int n_pairs, int* out) {
  if (n_pairs == 0 || c->tc_n_units == 0) return LORB_OK;
  const int n_desc = c->bank_n_desc, n_pad = pad_of(n_desc);
  const size_t kf_bytes = image_bytes(n_desc);
  LORB_CUDA_TRY(cudaFuncSetAttribute(tc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_BYTES));
  int* rowkey = c->tc_keys.as<int>();
  int* colkey = rowkey + c->tc_keys_rows;
  const int n_chunks = (int)c->tc_chunk_units.size() - 1;
  for (int ch = 0; ch < n_chunks; ch++) {
    const long long u0 = c->tc_chunk_units[ch], nu = c->tc_chunk_units[ch + 1] - u0;
    const int base = ch * TC_CHUNK_PAIRS, np = std::min(n_pairs - base, TC_CHUNK_PAIRS);
    const int grid = (int)std::min<long long>(nu, c->sm_count);
    prof_begin(c, 3);
    LORB_LAUNCH(c, tc_sweep_kernel, grid, TC_THREADS, TC_SMEM_BYTES,
                c->tc_img.as<uint8_t>() + (size_t)kf_base_a * kf_bytes,
                c->tc_img.as<uint8_t>() + (size_t)kf_base_b * kf_bytes, (long long)kf_bytes,
                c->tc_units.as<int4>() + u0, (int)nu, n_pad / MT_ROWS, n_pad, rowkey, colkey);
    prof_end(c, 3);
    LORB_LAUNCH(c, tc_finalize_kernel, np, 256, 0, rowkey, colkey, n_desc, n_pad, out + 3 * (size_t)base);
  }
  return LORB_OK;
}

}  // namespace tc
}  // namespace lorb

extern "C" int lorb_microbench_tensor_i8(lorb_ctx* c, int tiles, double* ops_per_s) {
  using namespace lorb;
  using namespace lorb::tc;
  LORB_REQUIRE(c && ops_per_s, "ctx/out");
  LORB_REQUIRE(tiles > 0, "tiles");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  const int smem = SLAB_BYTES + MTILE_BYTES + 64;
  LORB_CUDA_TRY(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  LORB_CUDA_TRY(cudaEventCreate(&e0));
  LORB_CUDA_TRY(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    LORB_CUDA_TRY(cudaEventRecord(e0, c->stream));
    LORB_LAUNCH(c, tc_mma_bench_kernel, c->sm_count, 128, smem, tiles);
    LORB_CUDA_TRY(cudaEventRecord(e1, c->stream));
    LORB_CUDA_TRY(cudaEventSynchronize(e1));
  }
  float ms = 0;
  LORB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // 9 MMAs of 128 x 256 x 32 multiply-adds per tile and SM, 2 operations each
  *ops_per_s = (double)c->sm_count * tiles * 9.0 * MT_ROWS * SLAB_COLS * 32.0 * 2.0 / (ms * 1e-3);
  return LORB_OK;
}
