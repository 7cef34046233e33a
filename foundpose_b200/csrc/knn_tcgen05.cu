// Brute-force k-NN (squared L2 / inner product) on tcgen05 with a running top-k epilogue.
//
// Replaces faiss.IndexFlatL2.search / IndexFlatIP.search as called from utils/knn_util.py:83,95
// (faiss 1.8.0, CPU).  faiss evaluates ||q||^2 + ||x||^2 - 2<q,x> with an sgemm; here the <q,x>
// tile is a tcgen05.mma (fp16 operands, fp32 accumulate in TMEM) and the epilogue keeps the k best
// candidates of each query in registers while the bank streams through shared memory - the
// distance matrix never exists in HBM.
//
// Work is described by "items": one item = up to 128 consecutive query rows searched against one
// contiguous segment of bank rows.  That covers every k-NN on the path with one kernel:
//   K1  all queries  vs the visual-word centroids           (utils/template_util.py:13-29)
//   K2  a crop's queries vs one template's rows             (utils/corresp_util.py:46)
//   K3  a template's rows vs a crop's queries               (utils/corresp_util.py:47)
//   K4  queries vs the full bank                            (BASELINE.json configs 3, 5)
// Items may be produced on the device (knn_build_items_*), so the retrieval pipeline needs no
// host synchronisation.
//
// CTA layout (384 threads, persistent over items):
//   warp 0   TMA producer   (query k-block + bank-tile k-block per stage, 128B swizzle)
//   warp 1   MMA issuer     (128 x 256 x 16 tcgen05.mma, accumulators double-buffered in TMEM)
//   warp 2   TMEM allocator
//   warps 4-11 epilogue     (thread = query row x column half: tcgen05.ld dots, d = ||x||^2 - 2 dot,
//                            tree argmin for k = 1 / prefiltered sorted insert for k > 1)
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

int g_knn_flags = 0;      // experiment switches (fp_knn_set_flags), 0 in production

constexpr int BQ = 128;   // query rows per item
constexpr int BX = 256;   // bank rows per tile (128 x 256 x 16 UMMA: 96 B/clk of smem operand reads)
constexpr int BKK = 64;   // K elements per stage
constexpr int kStages = 4;
constexpr int kMaxK = 16;
// Candidate lists of the k > 1 scan: kCandCap distances per epilogue thread, laid out [slot][256 threads]
// (conflict-free) in the 16 KB that the half-merge uses at the end of an item.  A list is filled from 16 columns at
// a time, so it cannot overflow.
constexpr int kCandCap = 16;
static_assert(kCandCap * 256 * 4 <= BQ * kMaxK * 8, "candidate lists must fit the half-merge buffer");
constexpr int kXnTiles = 4;                        // bank tiles per staged group of ||x||^2 (one barrier per group)
constexpr int kXnSlot = kXnTiles * BX + kXnTiles * (BX / 32);   // floats per slot: the norms + the min of every 32-column chunk
constexpr uint32_t kXnBytes = 2 * kXnSlot * 4;         // ||x||^2 of the current and the next group of tiles
constexpr uint32_t kMergeBytes = BQ * kMaxK * 8 + kXnBytes;   // (dist, idx) lists of the second column half + xn_s
constexpr uint32_t kQStage = BQ * BKK * 2;
constexpr uint32_t kXStage = BX * BKK * 2;
constexpr uint32_t kStageBytes = kQStage + kXStage;
constexpr uint32_t kKnnSmem = kStages * kStageBytes + kMergeBytes + 256 + 1024;
constexpr int kKnnThreads = 384;   // 4 control warps + 8 epilogue warps

template <int K>
struct TopK {
  float d[K];
  int i[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; ++j) { d[j] = INFINITY; i[j] = -1; }
  }
  // Sorted ascending; a later candidate with an equal distance never displaces an earlier one,
  // so ties resolve to the lower index (candidates arrive in ascending index order).
  // Every slot is a select of old values and the candidate (p[] is monotone because the list is sorted), so the
  // K slots update independently: depth 3 instead of a K-deep chain of conditional swaps, and a candidate that
  // does not belong in the list (+inf for an idle lane) is a no-op without a guard branch.
  __device__ __forceinline__ void insert(float cd, int ci) {
    bool p[K];
#pragma unroll
    for (int j = 0; j < K; ++j) p[j] = cd < d[j];
#pragma unroll
    for (int j = K - 1; j >= 1; --j) {
      d[j] = p[j - 1] ? d[j - 1] : (p[j] ? cd : d[j]);
      i[j] = p[j - 1] ? i[j - 1] : (p[j] ? ci : i[j]);
    }
    d[0] = p[0] ? cd : d[0];
    i[0] = p[0] ? ci : i[0];
  }
  __device__ __forceinline__ void push(float cd, int ci) {
    if (cd < d[K - 1]) insert(cd, ci);
  }
  // Lexicographic (distance, index) insert: used when merging lists whose index ranges interleave.
  __device__ __forceinline__ void push_lex(float cd, int ci) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (cd < d[j] || (cd == d[j] && ci < i[j])) {
        const float td = d[j]; d[j] = cd; cd = td;
        const int ti = i[j]; i[j] = ci; ci = ti;
      }
    }
  }
};


// Shared-memory loads by 32-bit shared address (pointers that went through uintptr_t arithmetic compile to generic loads).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ float min16(const float* d) {
  const float a = fminf(fminf(d[0], d[1]), d[2]);
  const float b = fminf(fminf(d[3], d[4]), d[5]);
  const float c = fminf(fminf(d[6], d[7]), d[8]);
  const float e = fminf(fminf(d[9], d[10]), d[11]);
  const float f = fminf(fminf(d[12], d[13]), d[14]);
  return fminf(fminf(fminf(fminf(a, b), c), fminf(e, f)), d[15]);
}

// Inserts the candidates (distance < current k-th best) among 16 consecutive columns of this thread's row into its
// sorted list.  Inserting straight from the registers costs one insert per column POSITION at which any of the
// warp's 32 rows has a candidate - while the lists warm up (the whole sweep for the short sweeps of the HBM-bound
// pass) that is nearly all of them.  Instead every row appends ITS candidates (usually 0-2) to a small list in
// shared memory - four predicated instructions per column, no branches; which columns they were is a 16-bit mask -
// and the warp then runs the insert max-over-rows-of-count times, idle rows inserting +inf.  Same insert order as
// a scan in column order (ascending index), so ties resolve identically.  The warp must be converged.
template <int K>
__device__ __forceinline__ void insert_candidates(const float* dist, int idx0, TopK<K>& best, uint32_t cand) {
  const float thr = best.d[K - 1];
  uint32_t ptr = cand, mask = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.lt.f32 p, %2, %3;\n"
        "@p st.shared.f32 [%0], %2;\n"
        "@p or.b32 %1, %1, %4;\n"
        "@p add.u32 %0, %0, 1024;\n"
        "}"
        : "+r"(ptr), "+r"(mask)
        : "f"(dist[i]), "f"(thr), "r"(1u << i)
        : "memory");
  }
  const int rounds = __reduce_max_sync(0xffffffffu, __popc(mask));
  for (int j = 0; j < rounds; ++j) {
    float cd;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cd) : "r"(cand + j * 1024) : "memory");
    const int bit = __ffs(mask) - 1;
    cd = mask ? cd : INFINITY;
    mask &= mask - 1;
    best.insert(cd, idx0 + bit);
  }
}

// One epilogue thread's share of a 128 x 256 distance tile: its query row (TMEM lane) x one half (128) of the
// tile's bank columns, starting at TMEM address `taddr`.  `ncols` = valid columns in this half (may be <= 0),
// `xn` = ||x||^2 of the half's 128 columns STAGED IN SHARED MEMORY by the epilogue warps one tile ahead (a global
// load per 32-column chunk on the critical path of every tile cost more than the chunk's arithmetic),
// `col_base` = index of the half's first column relative to the item's segment.
// L2: ||x||^2 - 2<q,x> (||q||^2 is added once per item); IP: -<q,x>.
template <int K>
__device__ __forceinline__ void scan_tile_half(uint32_t taddr, int ncols, uint32_t xn /* shared address */,
                                               uint32_t xn_min /* shared address: min ||x||^2 per 32 columns */,
                                               int col_base, int metric_ip, TopK<K>& best,
                                               uint32_t cand /* shared address: this thread's candidate list */) {
  // dist = scale * <q,x> + ||x||^2 with scale = -2 (L2) or -1 (IP, where the staged norms are zero): one FFMA per
  // candidate, no per-element select on the metric.
  const float scale = metric_ip ? -1.0f : -2.0f;
  __syncwarp();   // the warp-wide votes and the .sync.aligned TMEM loads below need the 32 rows converged
#pragma unroll 1
  for (int c = 0; c < BX / 64; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + c * 32, v);
    tmem_ld_wait();
    if (c * 32 < ncols) {
      float dist[32];
      if (c * 32 + 32 <= ncols) {
        // Safe prefilter on the raw inner products: every distance of the chunk is >= scale * max<q,x> + min||x||^2
        // (scale < 0; fmaf is monotone in each argument, so the bound also holds for the rounded values).  When
        // that bound cannot enter the list - almost always once the list has warmed up - the chunk costs one
        // 3-input max tree instead of 32 FFMA + the norm loads + the scan.
        float m[11];
#pragma unroll
        for (int i = 0; i < 10; ++i)
          m[i] = fmaxf(fmaxf(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1])), __uint_as_float(v[3 * i + 2]));
        m[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
        m[0] = fmaxf(fmaxf(m[0], m[1]), m[2]);
        m[3] = fmaxf(fmaxf(m[3], m[4]), m[5]);
        m[6] = fmaxf(fmaxf(m[6], m[7]), m[8]);
        m[9] = fmaxf(m[9], m[10]);
        const float smax = fmaxf(fmaxf(fmaxf(m[0], m[3]), m[6]), m[9]);
        // (warp-uniform: the rows of a warp stay converged for the collectives of the insert below)
        if (!__any_sync(0xffffffffu, fmaf(scale, smax, lds_f1(xn_min + c * 4)) < best.d[K - 1])) continue;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 n4 = lds_f4(xn + (c * 32 + g * 4) * 4);
          dist[g * 4 + 0] = fmaf(scale, __uint_as_float(v[g * 4 + 0]), n4.x);
          dist[g * 4 + 1] = fmaf(scale, __uint_as_float(v[g * 4 + 1]), n4.y);
          dist[g * 4 + 2] = fmaf(scale, __uint_as_float(v[g * 4 + 2]), n4.z);
          dist[g * 4 + 3] = fmaf(scale, __uint_as_float(v[g * 4 + 3]), n4.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          dist[i] = col < ncols ? fmaf(scale, __uint_as_float(v[i]), lds_f1(xn + col * 4)) : INFINITY;
        }
      }
      if constexpr (K == 1) {
        // argmin of the 32 candidates by a tree (depth 5) instead of a 32-deep serial chain;
        // the left operand (lower index) wins ties.
        int idx[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool lt = dist[2 * i + 1] < dist[2 * i];
          dist[i] = lt ? dist[2 * i + 1] : dist[2 * i];
          idx[i] = lt ? 2 * i + 1 : 2 * i;
        }
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const bool lt = dist[2 * i + 1] < dist[2 * i];
            dist[i] = lt ? dist[2 * i + 1] : dist[2 * i];
            idx[i] = lt ? idx[2 * i + 1] : idx[2 * i];
          }
        }
        if (dist[0] < best.d[0]) {
          best.d[0] = dist[0];
          best.i[0] = col_base + c * 32 + idx[0];
        }
      } else {
        // The k-th best bounds what can enter the list.  Test the two 16-column halves of the chunk (min trees of
        // 3-input minima) and only where some row of the warp has a candidate run the insert for that half.
        const int idx0 = col_base + c * 32;
        const float h0 = min16(dist);
        const float h1 = min16(dist + 16);
        if (__any_sync(0xffffffffu, h0 < best.d[K - 1])) insert_candidates<K>(dist, idx0, best, cand);
        if (__any_sync(0xffffffffu, h1 < best.d[K - 1])) insert_candidates<K>(dist + 16, idx0 + 16, best, cand);
      }
    }
  }
}

// ||x||^2 of a group of kXnTiles bank tiles (kXnTiles * 256 columns): epilogue thread et = 0..255 stages columns
// et, et + 256, ... of the group; zero past the segment (and for the inner-product metric).
struct XnGroup { float v[kXnTiles]; };
__device__ __forceinline__ XnGroup load_group_xnorm(const float* __restrict__ xnorm, const KnnItem& item, int group,
                                                    int et, int metric_ip) {
  XnGroup g;
#pragma unroll
  for (int i = 0; i < kXnTiles; ++i) {
    const int col = (group * kXnTiles + i) * BX + et;
    g.v[i] = (metric_ip || col >= item.b_rows) ? 0.f : __ldg(xnorm + item.b_row0 + col);
  }
  return g;
}
// Also stores the minimum of every 32-column chunk (a warp's 32 consecutive columns are exactly one chunk) behind
// the norms: slot[kXnTiles * BX + tile * 8 + chunk].
__device__ __forceinline__ void store_group_xnorm(float* slot, const XnGroup& g, int et) {
#pragma unroll
  for (int i = 0; i < kXnTiles; ++i) {
    slot[i * BX + et] = g.v[i];
    float mn = g.v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((et & 31) == 0) slot[kXnTiles * BX + i * (BX / 32) + (et >> 5)] = mn;
  }
}

// Merges the two column halves of an item through shared memory (their index ranges interleave tile by tile ->
// lexicographic order) and writes the k_out best of query row `r` (valid iff r < q_rows).  All 256 epilogue
// threads of the CTA call it.
template <int K>
__device__ __forceinline__ void merge_halves_and_store(TopK<K>& best, int half, int r, int q_rows, long out_row,
                                                       float qn, int metric_ip, int k_out, uint8_t* merge_buf,
                                                       float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  float* merge_d = reinterpret_cast<float*>(merge_buf);
  int* merge_i = reinterpret_cast<int*>(merge_buf + BQ * kMaxK * 4);
  if constexpr (K > 1) asm volatile("bar.sync 1, 256;" ::: "memory");   // the buffer held the scans' candidate lists
  if (half == 1) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      merge_d[j * BQ + r] = best.d[j];
      merge_i[j * BQ + r] = best.i[j];
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (half == 0) {
#pragma unroll
    for (int j = 0; j < K; ++j) best.push_lex(merge_d[j * BQ + r], merge_i[j * BQ + r]);
    if (r < q_rows) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j < k_out) {
          float dv;
          if (metric_ip) dv = -best.d[j];                 // similarity, descending
          else dv = fmaxf(best.d[j] + qn, 0.f);           // faiss clamps negative distances to 0
          if (best.i[j] < 0) dv = metric_ip ? -INFINITY : INFINITY;  // fewer than k bank rows
          out_d[out_row * k_out + j] = dv;
          out_i[out_row * k_out + j] = best.i[j];
        }
      }
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");   // merge buffer reusable
}

template <int K>
__global__ void __launch_bounds__(kKnnThreads, 1)
knn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
           const KnnItem* __restrict__ items, int num_items, int dim,
           const float* __restrict__ qnorm, const float* __restrict__ xnorm, int metric_ip,
           int k_out, float* __restrict__ out_d, int64_t* __restrict__ out_i, int flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* merge_buf = smem + kStages * kStageBytes;
  float* xn_s = reinterpret_cast<float*>(merge_buf + BQ * kMaxK * 8);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(merge_buf + kMergeBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BX);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = dim / BKK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const KnnItem item = items[it];
        if (item.q_rows <= 0 || item.b_rows <= 0) continue;
        const int num_tiles = (item.b_rows + BX - 1) / BX;
        for (int t = 0; t < num_tiles; ++t) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sq = smem + stage * kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_2d(sq, &tmQ, &full_bar[stage], kb * BKK, item.q_row0);
            tma_load_2d(sq + kQStage, &tmX, &full_bar[stage], kb * BKK, item.b_row0 + t * BX);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BQ, BX);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const KnnItem item = items[it];
        if (item.q_rows <= 0 || item.b_rows <= 0) continue;
        const int num_tiles = (item.b_rows + BX - 1) / BX;
        for (int t = 0; t < num_tiles; ++t) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + acc * BX;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            const uint32_t sq = smem_u32(smem + stage * kStageBytes);
            const uint64_t adesc = make_smem_desc_sw128(sq);
            const uint64_t bdesc = make_smem_desc_sw128(sq + kQStage);
#pragma unroll
            for (int k = 0; k < BKK / 16; ++k)
              umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // 8 epilogue warps: warp % 4 selects the TMEM sub-partition (32 query rows), (warp - 4) / 4 the
    // half of the tile's 256 bank columns.  Each thread keeps the k best of ITS columns; the two
    // halves are merged through shared memory once per item.
    const int sub = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = sub * 32 + lane;
    const int et = threadIdx.x - 128;          // 0..255: the tile column whose ||x||^2 this thread stages
    const uint32_t cand = smem_u32(merge_buf) + et * 4;   // candidate list of this thread (k > 1 scans)
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const KnnItem item = items[it];
      if (item.q_rows <= 0 || item.b_rows <= 0) continue;
      const int num_tiles = (item.b_rows + BX - 1) / BX;
      TopK<K> best;
      best.init();
      store_group_xnorm(xn_s, load_group_xnorm(xnorm, item, 0, et, metric_ip), et);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      XnGroup xn_next;
      for (int t = 0; t < num_tiles; ++t) {
        const int grp = t / kXnTiles, tg = t % kXnTiles;
        if (tg == 0) xn_next = load_group_xnorm(xnorm, item, grp + 1, et, metric_ip);   // lands during the scans
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after_sync();
        const int col_base = t * BX + half * (BX / 2);
        if (!(flags & 1))   // experiment switch: epilogue skips the scan (fp_knn_set_flags bit 0)
        scan_tile_half<K>(tmem_base + lane_addr + acc * BX + half * (BX / 2), item.b_rows - col_base,
                          smem_u32(xn_s + (grp & 1) * kXnSlot + tg * BX + half * (BX / 2)),
                          smem_u32(xn_s + (grp & 1) * kXnSlot + kXnTiles * BX + tg * (BX / 32) + half * (BX / 64)),
                          col_base, metric_ip, best, cand);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        if (tg == kXnTiles - 1) {
          store_group_xnorm(xn_s + ((grp + 1) & 1) * kXnSlot, xn_next, et);
          asm volatile("bar.sync 1, 256;" ::: "memory");   // next group's slot complete; this group's no longer read
        }
      }
      merge_halves_and_store<K>(best, half, r, item.q_rows, static_cast<long>(item.out_row0) + r,
                                (metric_ip || r >= item.q_rows) ? 0.f : qnorm[static_cast<long>(item.q_row0) + r],
                                metric_ip, k_out, merge_buf, out_d, out_i);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * BX);
  }
}

template <int K>
int launch_knn(const CUtensorMap& tmQ, const CUtensorMap& tmX, const KnnItem* items, int num_items,
               int dim, const float* qnorm, const float* xnorm, int metric_ip, int k_out,
               float* out_d, int64_t* out_i, cudaStream_t stream) {
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(knn_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kKnnSmem));
  }
  const int grid = num_items < num_sms() ? num_items : num_sms();
  ProfScope prof(PROF_KNN, stream, 0.0);
  knn_kernel<K><<<grid, kKnnThreads, kKnnSmem, stream>>>(tmQ, tmX, items, num_items, dim, qnorm,
                                                        xnorm, metric_ip, k_out, out_d, out_i, g_knn_flags);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}


// ----------------------------------------------------------------------------------------
// Pair kernel (tcgen05 cta_group::2) for the tensor-bound regime: many queries against a large bank
// (K4 of BASELINE configs 3 and 5: 900 queries per crop x the whole bank).
//
// The 1-CTA kernel above re-streams the query block with every bank tile and reads both operands of a
// 128 x 256 x 16 MMA from its own shared memory: 96 B/clk of operand reads + 96 B/clk of TMA writes
// against a 128 B/clk port - it tops out near 1000 TFLOP/s.  Here a cluster of two CTAs computes a
// 256-query x 256-bank-row tile per MMA (M256 N256 K16): each CTA keeps ITS 128 query rows RESIDENT
// in shared memory for the whole bank sweep (d <= 576; wider descriptors stream Q like the GEMM
// does) and stages only HALF of each bank tile, so per SM the port sees 64 B/clk of operand reads +
// 32 B/clk of TMA writes.  Work item = up to 256 query rows x one bank segment; epilogue as above
// (each CTA scans its own 128 x 256 accumulator out of its own TMEM).
// ----------------------------------------------------------------------------------------
constexpr int kPairMaxStages = 8;
constexpr uint32_t kPairHalfX = 128 * BKK * 2;     // this CTA's half of a bank tile k-block: 16 KB
constexpr uint32_t kPairQBlock = BQ * BKK * 2;     // one k-block of this CTA's 128 query rows: 16 KB
constexpr uint32_t kPairSmemBudget = 232448 - 1024 /*align*/ - kMergeBytes - 512 /*barriers*/;

struct PairLayout {
  int q_resident;        // 1: queries stay in smem for the whole item; 0: streamed with the bank
  int stages;
  uint32_t q_bytes;      // resident query region
  uint32_t stage_bytes;  // 16 KB (resident) or 32 KB (streaming: [X half][Q block])
  uint32_t smem_bytes;   // dynamic shared memory to request
  int flags;             // experiment switches (g_knn_flags)
};

inline PairLayout pair_layout(int dim) {
  PairLayout L;
  const uint32_t qb = static_cast<uint32_t>(dim / BKK) * kPairQBlock;
  const int st_res = qb < kPairSmemBudget ? static_cast<int>((kPairSmemBudget - qb) / kPairHalfX) : 0;
  if (st_res >= 4) {
    L.q_resident = 1;
    L.stages = st_res > kPairMaxStages ? kPairMaxStages : st_res;
    L.q_bytes = qb;
    L.stage_bytes = kPairHalfX;
  } else {
    L.q_resident = 0;
    L.stages = static_cast<int>(kPairSmemBudget / (kPairHalfX + kPairQBlock));
    if (L.stages > kPairMaxStages) L.stages = kPairMaxStages;
    L.q_bytes = 0;
    L.stage_bytes = kPairHalfX + kPairQBlock;
  }
  L.smem_bytes = L.q_bytes + L.stages * L.stage_bytes + kMergeBytes + 512 + 1024;
  L.flags = 0;
  return L;
}

template <int K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kKnnThreads, 1)
knn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                const KnnItem* __restrict__ items, int num_items, int dim,
                const float* __restrict__ qnorm, const float* __restrict__ xnorm, int metric_ip,
                int k_out, float* __restrict__ out_d, int64_t* __restrict__ out_i, const PairLayout L,
                unsigned long long* __restrict__ sync_counter, int sync_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* qbuf = smem;
  uint8_t* stages = smem + L.q_bytes;
  uint8_t* merge_buf = stages + L.stages * L.stage_bytes;
  float* xn_s = reinterpret_cast<float*>(merge_buf + BQ * kMaxK * 8);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(merge_buf + kMergeBytes);
  uint64_t* empty_bar = full_bar + kPairMaxStages;
  uint64_t* tfull_bar = empty_bar + kPairMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* qfull_bar = tempty_bar + 2;
  uint64_t* qempty_bar = qfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader (issues the MMAs)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < L.stages; ++s) {
      mbar_init(&full_bar[s], 1);      // leader: one arrive.expect_tx for both CTAs' bytes
      mbar_init(&empty_bar[s], 1);     // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);     // multicast tcgen05.commit
      mbar_init(&tempty_bar[a], 16);   // leader: epilogue warps of both CTAs
    }
    mbar_init(qfull_bar, 1);
    mbar_init(qempty_bar, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 2 * BX);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = dim / BKK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int row_off = static_cast<int>(rank) * BQ;   // this CTA's share of the item's query rows / tile's bank rows

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, qphase = 0;
      unsigned long long sync_expected = 0;
      bool sync_on = sync_tiles > 0;
      for (int it = cluster_id; it < num_items; it += num_clusters) {
        const KnnItem item = items[it];
        if (item.q_rows <= 0 || item.b_rows <= 0) continue;
        if (L.q_resident) {
          mbar_wait(qempty_bar, qphase ^ 1);   // the previous item's MMAs have all retired
          if (rank == 0) mbar_arrive_expect_tx(qfull_bar, 2u * L.q_bytes);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_load_2d_2sm(qbuf + kb * kPairQBlock, &tmQ, qfull_bar, kb * BKK, item.q_row0 + row_off);
          qphase ^= 1;
        }
        const int num_tiles = (item.b_rows + BX - 1) / BX;
        // Items of one wave sweep the same bank rows: every `sync_tiles` tiles the leaders' producers of the
        // `item.pad` participating clusters meet at a grid barrier (a monotonic counter in global memory), so that
        // the 74 sweeps stay within a few MB of each other and the bank is read from HBM once per wave instead of
        // once per cluster (ncu: 790 GB of DRAM reads per 57 600-query launch without it, L2 hit rate 57%).
        const unsigned long long participants = static_cast<unsigned long long>(item.pad);
        for (int t = 0; t < num_tiles; ++t) {
          if (sync_on && participants > 0 && rank == 0 && (t % sync_tiles) == 0) {
            // sync_counter[0] = arrivals (monotonic), sync_counter[1] = "give up" flag.  The barrier is an
            // optimisation, never a correctness requirement: if the other clusters do not show up within a few ms
            // (the grid is not co-resident because another kernel holds SMs, or a cluster was delayed), this
            // cluster raises the flag and every cluster lets its sweep run free for the rest of the launch.
            sync_expected += participants;
            asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(sync_counter) : "memory");
            unsigned long long seen = 0, quit = 0;
            uint32_t polls = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(sync_counter) : "memory");
              if (seen >= sync_expected) break;
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(quit) : "l"(sync_counter + 1) : "memory");
              if (quit != 0 || ++polls > 8192) {
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(sync_counter + 1), "l"(1ull) : "memory");
                sync_on = false;
                break;
              }
              __nanosleep(100);
            } while (true);
          }
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sx = stages + stage * L.stage_bytes;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * L.stage_bytes);
            tma_load_2d_2sm(sx, &tmX, &full_bar[stage], kb * BKK, item.b_row0 + t * BX + row_off);
            if (!L.q_resident)
              tma_load_2d_2sm(sx + kPairHalfX, &tmQ, &full_bar[stage], kb * BKK, item.q_row0 + row_off);
            if (++stage == L.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * BQ, BX);
      int stage = 0;
      uint32_t phase = 0, qphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = cluster_id; it < num_items; it += num_clusters) {
        const KnnItem item = items[it];
        if (item.q_rows <= 0 || item.b_rows <= 0) continue;
        if (L.q_resident) {
          mbar_wait(qfull_bar, qphase);
          qphase ^= 1;
        }
        const int num_tiles = (item.b_rows + BX - 1) / BX;
        for (int t = 0; t < num_tiles; ++t) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + acc * BX;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            const uint32_t sx = smem_u32(stages + stage * L.stage_bytes);
            const uint64_t bdesc = make_smem_desc_sw128(sx);
            const uint64_t adesc = make_smem_desc_sw128(L.q_resident ? smem_u32(qbuf) + kb * kPairQBlock
                                                                     : sx + kPairHalfX);
#pragma unroll
            for (int k = 0; k < BKK / 16; ++k)
              umma_f16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm(&empty_bar[stage]);
            if (++stage == L.stages) { stage = 0; phase ^= 1; }
          }
          umma_commit_2sm(&tfull_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        if (L.q_resident) umma_commit_2sm(qempty_bar);   // both CTAs' query regions are free again
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int sub = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = sub * 32 + lane;
    const int et = threadIdx.x - 128;          // 0..255: the tile column whose ||x||^2 this thread stages
    const uint32_t cand = smem_u32(merge_buf) + et * 4;   // candidate list of this thread (k > 1 scans)
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = cluster_id; it < num_items; it += num_clusters) {
      const KnnItem item = items[it];
      if (item.q_rows <= 0 || item.b_rows <= 0) continue;
      const int num_tiles = (item.b_rows + BX - 1) / BX;
      const int my_rows = item.q_rows - row_off;   // valid query rows of this CTA (may be <= 0)
      TopK<K> best;
      best.init();
      store_group_xnorm(xn_s, load_group_xnorm(xnorm, item, 0, et, metric_ip), et);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      XnGroup xn_next;
      for (int t = 0; t < num_tiles; ++t) {
        const int grp = t / kXnTiles, tg = t % kXnTiles;
        if (tg == 0) xn_next = load_group_xnorm(xnorm, item, grp + 1, et, metric_ip);   // lands during the scans
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after_sync();
        const int col_base = t * BX + half * (BX / 2);
#ifdef FP_KNN_EXPERIMENTS   // make EXTRA=-DFP_KNN_EXPERIMENTS: TMEM reads only (fp_knn_set_flags bit 2); costs registers
        if (L.flags & 4) {
#pragma unroll 1
          for (int c = 0; c < BX / 64; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + acc * BX + half * (BX / 2) + c * 32, v);
            tmem_ld_wait();
            if (__uint_as_float(v[0]) + __uint_as_float(v[31]) == 1.2345e30f) best.d[0] = 0.f;   // keep the load alive
          }
        } else
#endif
        if (!(L.flags & 1))
          scan_tile_half<K>(tmem_base + lane_addr + acc * BX + half * (BX / 2), item.b_rows - col_base,
                            smem_u32(xn_s + (grp & 1) * kXnSlot + tg * BX + half * (BX / 2)),
                            smem_u32(xn_s + (grp & 1) * kXnSlot + kXnTiles * BX + tg * (BX / 32) + half * (BX / 64)),
                            col_base, metric_ip, best, cand);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        if (tg == kXnTiles - 1) {
          store_group_xnorm(xn_s + ((grp + 1) & 1) * kXnSlot, xn_next, et);
          asm volatile("bar.sync 1, 256;" ::: "memory");   // next group's slot complete; this group's no longer read
        }
      }
      merge_halves_and_store<K>(best, half, r, my_rows, static_cast<long>(item.out_row0) + row_off + r,
                                (metric_ip || r >= my_rows) ? 0.f
                                                            : qnorm[static_cast<long>(item.q_row0) + row_off + r],
                                metric_ip, k_out, merge_buf, out_d, out_i);
    }
  }

  // Neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers.
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, 2 * BX);
  }
}

// Experiment switches of the pair kernel (fp_knn_set_flags; 0 in production):
//   bit 0  epilogue warps skip the scan (main-loop speed alone; results are garbage)
//   bit 1  stream the queries with the bank even when they would fit in shared memory
//   bit 2  epilogue warps only read the accumulators out of TMEM (no arithmetic)
//   bit 3  no sweep barrier: the clusters' bank sweeps run free

template <int K>
int launch_knn_pair(const CUtensorMap& tmQ, const CUtensorMap& tmX, const KnnItem* items, int num_items,
                    int dim, const float* qnorm, const float* xnorm, int metric_ip, int k_out,
                    float* out_d, int64_t* out_i, unsigned long long* sync_counter, int sync_tiles,
                    cudaStream_t stream) {
  PairLayout L = pair_layout(dim);
  if ((g_knn_flags & 2) && L.q_resident) {
    L.q_resident = 0;
    L.stages = 6;
    L.q_bytes = 0;
    L.stage_bytes = kPairHalfX + kPairQBlock;
    L.smem_bytes = L.stages * L.stage_bytes + kMergeBytes + 512 + 1024;
  }
  L.flags = g_knn_flags;
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(knn_pair_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       232448));
  }
  const int max_clusters = num_sms() / 2;
  const int clusters = num_items < max_clusters ? num_items : max_clusters;
  if (sync_counter == nullptr || (g_knn_flags & 8)) sync_tiles = 0;   // bit 3: experiment, sweeps run free
  if (sync_tiles > 0) FP_CUDA_CHECK(cudaMemsetAsync(sync_counter, 0, 2 * sizeof(unsigned long long), stream));
  ProfScope prof(PROF_KNN_PAIR, stream, 0.0);
  knn_pair_kernel<K><<<2 * clusters, kKnnThreads, L.smem_bytes, stream>>>(
      tmQ, tmX, items, num_items, dim, qnorm, xnorm, metric_ip, k_out, out_d, out_i, L, sync_counter, sync_tiles);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ||row||^2 in fp32 of fp16 rows: one warp per row.
__global__ void row_sqnorm_kernel(const __half* __restrict__ x, float* __restrict__ out, long rows,
                                  int dim) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long row = blockIdx.x * static_cast<long>(wpb) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long>(gridDim.x) * wpb) {
    const __half2* p = reinterpret_cast<const __half2*>(x + row * dim);
    float s = 0.f;
    for (int i = lane; i < dim / 2; i += 32) {
      const float2 f = __half22float2(p[i]);
      s = fmaf(f.x, f.x, s);
      s = fmaf(f.y, f.y, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
  }
}

// fp32 -> fp16 row conversion with optional L2 normalisation (cosine metric,
// utils/knn_util.py:57, 93) - one warp per row.
__global__ void convert_rows_kernel(const float* __restrict__ x, __half* __restrict__ y, long rows,
                                    int dim, int l2_normalize) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long row = blockIdx.x * static_cast<long>(wpb) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long>(gridDim.x) * wpb) {
    const float* xr = x + row * dim;
    float inv = 1.f;
    if (l2_normalize) {
      float s = 0.f;
      for (int i = lane; i < dim; i += 32) s = fmaf(xr[i], xr[i], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      inv = 1.0f / sqrtf(s);
    }
    for (int i = lane; i < dim; i += 32) y[row * dim + i] = __float2half_rn(xr[i] * inv);
  }
}

// fp16 rows -> L2-normalised fp16 rows + their ||.||^2 (fp32, of the ROUNDED unit rows) in one pass: the query side
// of the cosine visual-word metric (utils/knn_util.py:93: faiss.normalize_L2 on the search vectors), fused so
// that the projected descriptors are read once.  One warp per row.
__global__ void unit_rows_kernel(const __half* __restrict__ x, __half* __restrict__ y, float* __restrict__ sqn,
                                 long rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long row = blockIdx.x * static_cast<long>(wpb) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long>(gridDim.x) * wpb) {
    const __half2* p = reinterpret_cast<const __half2*>(x + row * dim);
    float s = 0.f;
    for (int i = lane; i < dim / 2; i += 32) {
      const float2 f = __half22float2(p[i]);
      s = fmaf(f.x, f.x, fmaf(f.y, f.y, s));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / sqrtf(s);
    __half2* q = reinterpret_cast<__half2*>(y + row * dim);
    float t = 0.f;
    for (int i = lane; i < dim / 2; i += 32) {
      const float2 f = __half22float2(p[i]);
      const __half2 h = __floats2half2_rn(f.x * inv, f.y * inv);
      q[i] = h;
      const float2 g = __half22float2(h);
      t = fmaf(g.x, g.x, fmaf(g.y, g.y, t));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sqn[row] = t;
  }
}

// fp32 rows -> THREE fp16 column blocks whose pairwise products reproduce the fp32 inner product on the tensor
// cores: x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significant bits), and
//     <a, b> ~= <a_hi, b_hi> + <a_lo, b_hi> + <a_hi, b_lo>       (the lo.lo term is below fp32 resolution)
// pattern 0 writes [hi | lo | hi] (bank side), pattern 1 writes [hi | hi | lo] (query side), so ONE contraction
// over 3W columns yields the three terms.  Rows are optionally L2-normalised first (cosine similarity,
// utils/template_util.py:160-164: zero rows stay zero, like x / max(||x||, eps)) and multiplied by `scale` (a
// power of two that lifts the lo parts out of the fp16 subnormal range).  One warp per row.
__global__ void split_rows_kernel(const float* __restrict__ x, __half* __restrict__ y, long rows, int dim,
                                  int pattern, int l2_normalize, float scale) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long row = blockIdx.x * static_cast<long>(wpb) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long>(gridDim.x) * wpb) {
    const float* xr = x + row * dim;
    float mul = scale;
    if (l2_normalize) {
      float s = 0.f;
      for (int i = lane; i < dim; i += 32) s = fmaf(xr[i], xr[i], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      mul = s > 0.f ? scale / sqrtf(s) : 0.f;
    }
    __half* yr = y + row * 3L * dim;
    for (int i = lane; i < dim; i += 32) {
      const float v = xr[i] * mul;
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      yr[i] = hi;
      yr[dim + i] = pattern == 0 ? lo : hi;
      yr[2 * dim + i] = pattern == 0 ? hi : lo;
    }
  }
}

// Items for one dense problem: all query rows against one bank segment.
__global__ void build_items_dense_kernel(KnnItem* items, int num_items, int q_total, int b_row0,
                                         int b_rows) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_items; i += gridDim.x * blockDim.x) {
    KnnItem it;
    it.q_row0 = i * BQ;
    it.q_rows = min(BQ, q_total - i * BQ);
    it.b_row0 = b_row0;
    it.b_rows = b_rows;
    it.out_row0 = static_cast<long long>(i) * BQ;
    it.pad = 0;
    items[i] = it;
  }
}

// Items for a bank split into chunks (few queries, large bank: every SM streams its own slice of
// the bank so the search is HBM-bound instead of running on one CTA per query block).
__global__ void build_items_split_kernel(KnnItem* items, int q_blocks, int num_chunks, int q_total,
                                         int b_row0, int b_rows, int chunk_rows) {
  const int n = q_blocks * num_chunks;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / q_blocks, m = i - c * q_blocks;
    KnnItem it;
    it.q_row0 = m * BQ;
    it.q_rows = min(BQ, q_total - m * BQ);
    it.b_row0 = b_row0 + c * chunk_rows;
    it.b_rows = max(0, min(chunk_rows, b_rows - c * chunk_rows));
    it.out_row0 = static_cast<long long>(c) * q_blocks * BQ + m * BQ;
    it.pad = 0;
    items[i] = it;
  }
}

// Merge per-chunk top-k lists: part_* [num_chunks, q_pad, k] (indices relative to the chunk) ->
// out_* [nq, k], ordered by (distance, global index).  One WARP per query: every lane folds the lists of chunks
// lane, lane + 32, ... into its own sorted top-K, then k rounds of a warp-wide lexicographic argmin pop the global
// winners (a single thread per query walking num_chunks * k candidates through a 16-deep insert took 175 us for the
// 64 x 40 x 5 lists of the template scoring, ncu profiles/r02_ncu_retrieval.md).
template <int K>
__global__ void __launch_bounds__(256)
knn_merge_kernel(const float* __restrict__ part_d, const int64_t* __restrict__ part_i,
                 int num_chunks, int q_pad, int nq, int k, int chunk_rows, int b_rows,
                 int descending, float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  TopK<K> best;
#pragma unroll
  for (int j = 0; j < K; ++j) { best.d[j] = INFINITY; best.i[j] = 0x7fffffff; }
  for (int c = lane; c < num_chunks; c += 32) {
    if (static_cast<long>(c) * chunk_rows >= b_rows) break;
    const long base = (static_cast<long>(c) * q_pad + q) * k;
    for (int j = 0; j < k; ++j) {
      const long li = part_i[base + j];
      if (li < 0) continue;
      const float d = part_d[base + j];
      best.push_lex(descending ? -d : d, static_cast<int>(c * chunk_rows + li));
    }
  }
  for (int r = 0; r < k; ++r) {
    float wd = best.d[0];
    int wi = best.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, wd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (od < wd || (od == wd && oi < wi)) { wd = od; wi = oi; }
    }
    if (best.i[0] == wi && wi != 0x7fffffff) {   // this lane held the winner: pop it
#pragma unroll
      for (int j = 0; j + 1 < K; ++j) { best.d[j] = best.d[j + 1]; best.i[j] = best.i[j + 1]; }
      best.d[K - 1] = INFINITY;
      best.i[K - 1] = 0x7fffffff;
    }
    if (lane == 0) {
      const bool ok = wi != 0x7fffffff;
      out_d[static_cast<long>(q) * k + r] = ok ? (descending ? -wd : wd) : (descending ? -INFINITY : INFINITY);
      out_i[static_cast<long>(q) * k + r] = ok ? wi : -1;
    }
  }
}

}  // namespace

void knn_set_flags(int flags) { g_knn_flags = flags; }

int knn_items_per_rows(int rows) { return (rows + BQ - 1) / BQ; }

int knn_build_items_split(KnnItem* items, int q_total, int b_row0, int b_rows, int num_chunks,
                          int chunk_rows, cudaStream_t stream) {
  const int q_blocks = knn_items_per_rows(q_total);
  const int n = q_blocks * num_chunks;
  if (n == 0) return 0;
  FP_REQUIRE(chunk_rows % BX == 0, "knn: chunk_rows=%d must be a multiple of %d", chunk_rows, BX);
  ProfScope prof(PROF_RETRIEVAL, stream, n * 32.0);
  build_items_split_kernel<<<(n + 255) / 256, 256, 0, stream>>>(items, q_blocks, num_chunks, q_total, b_row0,
                                                               b_rows, chunk_rows);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int knn_merge(const float* part_d, const int64_t* part_i, int num_chunks, int q_pad, int nq, int k,
              int chunk_rows, int b_rows, int descending, float* out_d, int64_t* out_i, cudaStream_t stream) {
  FP_REQUIRE(k >= 1 && k <= 16, "knn_merge: k=%d is outside [1,16]", k);
  if (nq <= 0) return 0;
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(num_chunks) * nq * k * 12);
  const int blocks = (nq + 7) / 8;   // 8 warps = 8 queries per block
#define FP_MERGE_CASE(KK)                                                                                   \
  knn_merge_kernel<KK><<<blocks, 256, 0, stream>>>(part_d, part_i, num_chunks, q_pad, nq, k, chunk_rows, b_rows, \
                                                   descending, out_d, out_i)
  if (k == 1) FP_MERGE_CASE(1);
  else if (k <= 3) FP_MERGE_CASE(3);
  else if (k <= 5) FP_MERGE_CASE(5);
  else if (k <= 8) FP_MERGE_CASE(8);
  else FP_MERGE_CASE(16);
#undef FP_MERGE_CASE
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int knn_build_items_dense(KnnItem* items, int q_total, int b_row0, int b_rows, cudaStream_t stream) {
  const int n = knn_items_per_rows(q_total);
  if (n == 0) return 0;
  ProfScope prof(PROF_RETRIEVAL, stream, n * 32.0);
  build_items_dense_kernel<<<(n + 255) / 256, 256, 0, stream>>>(items, n, q_total, b_row0, b_rows);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int row_sqnorm_f16(const __half* x, float* out, long rows, int dim, cudaStream_t stream) {
  if (rows == 0) return 0;
  FP_REQUIRE(dim % 2 == 0, "row_sqnorm: dim must be even");
  long blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(rows) * dim * 2);
  row_sqnorm_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, out, rows, dim);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int convert_rows_f16(const float* x, __half* y, long rows, int dim, int l2_normalize,
                     cudaStream_t stream) {
  if (rows == 0) return 0;
  long blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(rows) * dim * 6);
  convert_rows_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, y, rows, dim, l2_normalize);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int unit_rows_f16(const __half* x, __half* y, float* sqnorm, long rows, int dim, cudaStream_t stream) {
  if (rows == 0) return 0;
  FP_REQUIRE(dim % 2 == 0, "unit_rows: dim must be even");
  long blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(rows) * dim * 4);
  unit_rows_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, y, sqnorm, rows, dim);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int split_rows_f16(const float* x, __half* y, long rows, int dim, int pattern, int l2_normalize, float scale,
                   cudaStream_t stream) {
  if (rows == 0) return 0;
  FP_REQUIRE(pattern == 0 || pattern == 1, "split_rows: pattern must be 0 (bank side) or 1 (query side)");
  long blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(rows) * dim * 10);
  split_rows_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, y, rows, dim, pattern, l2_normalize, scale);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int knn_search_items(const __half* q, long q_rows_total, const __half* x, long x_rows_total, int dim,
                     const KnnItem* items, int num_items, const float* qnorm, const float* xnorm,
                     int metric_ip, int k, float* out_d, int64_t* out_i, cudaStream_t stream) {
  FP_REQUIRE(k >= 1 && k <= 16, "knn: k=%d is outside the supported range [1,16]", k);
  FP_REQUIRE(dim % BKK == 0 && dim >= BKK, "knn: dim=%d must be a positive multiple of %d", dim, BKK);
  if (num_items <= 0 || q_rows_total <= 0) return 0;
  FP_REQUIRE(x_rows_total > 0, "knn: the index is empty");
  CUtensorMap tmQ, tmX;
  if (make_tma_2d_f16(&tmQ, q, q_rows_total, dim, dim, BQ) != 0) return 3;
  if (make_tma_2d_f16(&tmX, x, x_rows_total, dim, dim, BX) != 0) return 3;
#define FP_KNN_CASE(KK)                                                                          \
  return launch_knn<KK>(tmQ, tmX, items, num_items, dim, qnorm, xnorm, metric_ip, k, out_d, out_i, \
                        stream)
  if (k == 1) FP_KNN_CASE(1);
  if (k == 2) FP_KNN_CASE(2);
  if (k == 3) FP_KNN_CASE(3);
  if (k <= 5) FP_KNN_CASE(5);
  if (k <= 8) FP_KNN_CASE(8);
  FP_KNN_CASE(16);
#undef FP_KNN_CASE
}


int knn_search_pair_items(const __half* q, long q_rows_total, const __half* x, long x_rows_total, int dim,
                          const KnnItem* items, int num_items, const float* qnorm, const float* xnorm,
                          int metric_ip, int k, float* out_d, int64_t* out_i, unsigned long long* sync_counter,
                          int sync_tiles, cudaStream_t stream) {
  FP_REQUIRE(k >= 1 && k <= 16, "knn: k=%d is outside the supported range [1,16]", k);
  FP_REQUIRE(dim % BKK == 0 && dim >= BKK, "knn: dim=%d must be a positive multiple of %d", dim, BKK);
  if (num_items <= 0 || q_rows_total <= 0) return 0;
  FP_REQUIRE(x_rows_total > 0, "knn: the index is empty");
  CUtensorMap tmQ, tmX;
  if (make_tma_2d_f16(&tmQ, q, q_rows_total, dim, dim, BQ) != 0) return 3;
  if (make_tma_2d_f16(&tmX, x, x_rows_total, dim, dim, 128) != 0) return 3;
#define FP_KNN_CASE(KK)                                                                               \
  return launch_knn_pair<KK>(tmQ, tmX, items, num_items, dim, qnorm, xnorm, metric_ip, k, out_d, out_i, \
                             sync_counter, sync_tiles, stream)
  if (k == 1) FP_KNN_CASE(1);
  if (k <= 3) FP_KNN_CASE(3);
  if (k <= 5) FP_KNN_CASE(5);
  if (k <= 8) FP_KNN_CASE(8);
  FP_KNN_CASE(16);
#undef FP_KNN_CASE
}

}  // namespace fp
