// Template retrieval (tf-idf bag of visual words) and cyclic-buddies correspondence kernels.
//
//   tfidf_histogram   : calc_tfidf (utils/template_util.py:31-71), batched over crops
//   bow_scores        : cosine_similarity(template_descs, tile(query_tfidf)) (template_util.py:167-169)
//   topk_rows         : torch.topk(..., sorted=True) per crop (template_util.py:172-174)
//   build_pair_items  : per (crop, retrieved template) k-NN work items for the two 1-NN searches
//                       of cyclic_buddies_matching (utils/corresp_util.py:46-47)
//   cyclic_buddies    : cycle distance, top-k best buddies, scores, 2D/3D gathers
//                       (corresp_util.py:50-68, 135-155)
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

// ---------------------------------------------------------------------------------------
// tf-idf histogram: one CTA per crop, histogram in shared memory.
// Hard assignment (the shipped default): every (query, word) pair contributes the same
// tf = (1/sqrt(k)) / Nq, so the histogram is an INTEGER count per word (warp-aggregated shared
// atomics via __match_any_sync, deterministic) scaled once by tf * idf[word].
// Soft assignment: float shared atomics of exp(-d^2 / 2 sigma^2) weights.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tfidf_kernel(const int64_t* __restrict__ word_ids, const float* __restrict__ word_dists, int k,
             const int* __restrict__ row_start, const int* __restrict__ row_count,
             const float* __restrict__ idf, int W, int soft, float sigma2, int sqrt_input,
             float* __restrict__ out) {
  extern __shared__ float hist[];  // W floats (soft) or W ints (hard)
  int* ihist = reinterpret_cast<int*>(hist);
  const int b = blockIdx.x;
  const int start = row_start[b];
  const int n = row_count[b];
  for (int w = threadIdx.x; w < W; w += blockDim.x) hist[w] = 0.f;  // 0.0f and 0 share a bit pattern
  __syncthreads();
  const long total = static_cast<long>(n) * k;
  if (!soft) {
    const long padded = (total + 31) / 32 * 32;
    for (long e = threadIdx.x; e < padded; e += blockDim.x) {
      const bool active = e < total;
      const int word = active ? static_cast<int>(word_ids[static_cast<long>(start) * k + e]) : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, word);
      const int leader = __ffs(peers) - 1;
      if (active && (threadIdx.x & 31) == leader) atomicAdd(&ihist[word], __popc(peers));
    }
    __syncthreads();
    // F.normalize(ones(k)) = 1/sqrt(k); tf = w / Nq.
    const float tf = (1.0f / sqrtf(static_cast<float>(k))) / static_cast<float>(n);
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
      const int c = ihist[w];
      // Words that never occur stay exactly 0 (scatter_add_ into zeros), even if idf is inf.
      out[static_cast<long>(b) * W + w] = c ? (static_cast<float>(c) * tf) * idf[w] : 0.f;
    }
  } else {
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const long base = (static_cast<long>(start) + q) * k;
      float norm2 = 0.f;
      for (int j = 0; j < k; ++j) {
        // calc_tfidf always squares its input; on the query side that input is the sqrt of the
        // faiss (squared) distance (template_util.py:27), at bank-build time it is the squared
        // distance itself.
        float d = word_dists[base + j];
        if (sqrt_input) d = sqrtf(d);
        const float w = expf(-(d * d) / (2.0f * sigma2));
        norm2 += w * w;
      }
      const float inv = 1.0f / fmaxf(sqrtf(norm2), 1e-12f);  // F.normalize eps
      for (int j = 0; j < k; ++j) {
        float d = word_dists[base + j];
        if (sqrt_input) d = sqrtf(d);
        const float w = expf(-(d * d) / (2.0f * sigma2)) * inv;
        const int word = static_cast<int>(word_ids[base + j]);
        atomicAdd(&hist[word], (w / static_cast<float>(n)) * idf[word]);
      }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) out[static_cast<long>(b) * W + w] = hist[w];
  }
}

// ---------------------------------------------------------------------------------------
// Cosine scores: out[b, t] = <desc_t, q_b> / (max(|desc_t|, eps) * max(|q_b|, eps)).
// Tile: 32 templates x 64 crops per CTA (256 threads, 2x4 outputs each), W in chunks of 32.
// template_descs (T x W fp32) is read exactly once per 64 crops -> HBM-bound for B <= 64.
// ---------------------------------------------------------------------------------------
constexpr int BT = 32, BC = 64, BW = 32;

__global__ void __launch_bounds__(256)
bow_scores_kernel(const float* __restrict__ descs, const float* __restrict__ desc_norm,
                  const float* __restrict__ q, int T, int B, int W, float eps,
                  float* __restrict__ out) {
  __shared__ float sd[BT][BW + 1];
  __shared__ float sq[BC][BW + 1];
  __shared__ float sqn[BC];
  const int t0 = blockIdx.x * BT, c0 = blockIdx.y * BC;
  const int tx = threadIdx.x & 15;   // crop group: crops tx*4 .. tx*4+3
  const int ty = threadIdx.x >> 4;   // template group: templates ty*2, ty*2+1
  float acc[2][4] = {};
  float qn_part = 0.f;               // thread (ty==0 rows) accumulates |q|^2 for its crops
  for (int w0 = 0; w0 < W; w0 += BW) {
    for (int e = threadIdx.x; e < BT * BW; e += 256) {
      const int r = e / BW, c = e - r * BW;
      const int t = t0 + r, w = w0 + c;
      sd[r][c] = (t < T && w < W) ? descs[static_cast<long>(t) * W + w] : 0.f;
    }
    for (int e = threadIdx.x; e < BC * BW; e += 256) {
      const int r = e / BW, c = e - r * BW;
      const int bb = c0 + r, w = w0 + c;
      sq[r][c] = (bb < B && w < W) ? q[static_cast<long>(bb) * W + w] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int w = 0; w < BW; ++w) {
      const float d0 = sd[ty * 2][w], d1 = sd[ty * 2 + 1][w];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float qv = sq[tx * 4 + j][w];
        acc[0][j] = fmaf(d0, qv, acc[0][j]);
        acc[1][j] = fmaf(d1, qv, acc[1][j]);
      }
    }
    if (threadIdx.x < BC) {
      for (int w = 0; w < BW; ++w) qn_part = fmaf(sq[threadIdx.x][w], sq[threadIdx.x][w], qn_part);
    }
    __syncthreads();
  }
  if (threadIdx.x < BC) sqn[threadIdx.x] = sqrtf(qn_part);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int t = t0 + ty * 2 + i;
    if (t >= T) continue;
    const float dn = fmaxf(desc_norm[t], eps);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int bb = c0 + tx * 4 + j;
      if (bb < B) out[static_cast<long>(bb) * T + t] = acc[i][j] / (dn * fmaxf(sqn[tx * 4 + j], eps));
    }
  }
}

__global__ void row_norm_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int rows,
                                    int dim) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) {
      const float v = x[static_cast<long>(row) * dim + i];
      s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = sqrtf(s);
  }
}

// ---------------------------------------------------------------------------------------
// Per-row top-k (largest first, ties -> lower index), k <= 16. One CTA per row.
// ---------------------------------------------------------------------------------------
constexpr int kMaxTopK = 16;

struct LexTop {
  float v[kMaxTopK];
  int i[kMaxTopK];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j) { v[j] = -INFINITY; i[j] = 0x7fffffff; }
  }
  // Descending by value, ascending by index among equal values.
  __device__ __forceinline__ void push(float cv, int ci, int k) {
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j) {
      if (j < k && (cv > v[j] || (cv == v[j] && ci < i[j]))) {
        const float tv = v[j]; v[j] = cv; cv = tv;
        const int ti = i[j]; i[j] = ci; ci = ti;
      }
    }
  }
};

__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ x, int cols, int k, float* __restrict__ out_v,
                 int64_t* __restrict__ out_i) {
  __shared__ float sv[8][kMaxTopK];
  __shared__ int si[8][kMaxTopK];
  const int row = blockIdx.x;
  const float* xr = x + static_cast<long>(row) * cols;
  LexTop best;
  best.init();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    best.push(xr[c], c, k);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    float ov[kMaxTopK];
    int oi[kMaxTopK];
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j) {
      ov[j] = __shfl_xor_sync(0xffffffffu, best.v[j], o);
      oi[j] = __shfl_xor_sync(0xffffffffu, best.i[j], o);
    }
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j)
      if (j < k) best.push(ov[j], oi[j], k);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j) { sv[warp][j] = best.v[j]; si[warp][j] = best.i[j]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      for (int j = 0; j < k; ++j) best.push(sv[w][j], si[w][j], k);
    for (int j = 0; j < k; ++j) {
      out_v[static_cast<long>(row) * k + j] = best.v[j];
      out_i[static_cast<long>(row) * k + j] = best.i[j] == 0x7fffffff ? -1 : best.i[j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// k-NN work items for the (crop, retrieved template) pairs.
// ---------------------------------------------------------------------------------------
__global__ void build_pair_items_kernel(const int64_t* __restrict__ top_ids, int num_pairs, int topn,
                                        const int* __restrict__ tpl_off,
                                        const int* __restrict__ q_start,
                                        const int* __restrict__ q_count, int max_q, int max_p,
                                        KnnItem* __restrict__ items_q2o,
                                        KnnItem* __restrict__ items_o2q) {
  const int per_q = (max_q + 127) / 128, per_p = (max_p + 127) / 128;
  const int total = num_pairs * (per_q + per_p);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pair = e / (per_q + per_p);
    const int m = e - pair * (per_q + per_p);
    const int b = pair / topn;
    const long t = top_ids[pair];
    const bool ok = t >= 0;
    const int ts = ok ? tpl_off[t] : 0;
    const int tn = ok ? tpl_off[t + 1] - ts : 0;
    const int qs = q_start[b], qn = q_count[b];
    KnnItem it;
    it.pad = 0;
    if (m < per_q) {          // query -> object: crop's queries vs the template's bank rows
      it.q_row0 = qs + m * 128;
      it.q_rows = max(0, min(128, qn - m * 128));
      it.b_row0 = ts;
      it.b_rows = tn;
      it.out_row0 = static_cast<long long>(pair) * max_q + m * 128;
      items_q2o[pair * per_q + m] = it;
    } else {                  // object -> query: the template's rows vs the crop's queries
      const int mm = m - per_q;
      it.q_row0 = ts + mm * 128;
      it.q_rows = max(0, min(128, tn - mm * 128));
      it.b_row0 = qs;
      it.b_rows = qn;
      it.out_row0 = static_cast<long long>(pair) * max_p + mm * 128;
      items_o2q[pair * per_p + mm] = it;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Cyclic buddies: one CTA per (crop, template) pair; bitonic sort of (dist, query id) keys in smem.
// ---------------------------------------------------------------------------------------
constexpr int kCycMax = 4096;

// 64-bit sort key of query i: (cycle distance bits << 32) | i.  The distance is non-negative so its
// bit pattern is monotonic; ties order by query id (canonical rule).
__device__ __forceinline__ unsigned long long cycle_key(const float* __restrict__ points, int qs,
                                                        const int64_t* __restrict__ q2o_p,
                                                        const int64_t* __restrict__ o2q_p, int i) {
  const long o = q2o_p[i];
  const long c = o2q_p[o];
  const float dx = points[(qs + i) * 2] - points[(qs + c) * 2];
  const float dy = points[(qs + i) * 2 + 1] - points[(qs + c) * 2 + 1];
  // torch.linalg.norm over 2 components: sqrt(dx^2 + dy^2).
  const float d = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  return (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | static_cast<unsigned>(i);
}

// Pre-selection for more than kCycMax keys per pair: each CTA sorts one chunk of kCycMax keys
// (level 0: computed from the cycle distances; later levels: read from `in`) and keeps its `keep`
// smallest.  out: [pairs, chunks * keep], padded with ~0.
__global__ void __launch_bounds__(512)
cyclic_preselect_kernel(const float* __restrict__ points, const int* __restrict__ q_start,
                        const int* __restrict__ q_count, const int64_t* __restrict__ q2o,
                        const int64_t* __restrict__ o2q, const int64_t* __restrict__ top_ids, int topn,
                        int max_q, int max_p, const unsigned long long* __restrict__ in, int n_in, int keep,
                        unsigned long long* __restrict__ out) {
  __shared__ unsigned long long keys[kCycMax];
  const int chunk = blockIdx.x, pair = blockIdx.y, chunks = gridDim.x;
  const int b = pair / topn;
  const int n = q_count[b];
  const int qs = q_start[b];
  const bool ok = top_ids[pair] >= 0;
  const int64_t* q2o_p = q2o + static_cast<long>(pair) * max_q;
  const int64_t* o2q_p = o2q + static_cast<long>(pair) * max_p;
  for (int i = threadIdx.x; i < kCycMax; i += blockDim.x) {
    const int g = chunk * kCycMax + i;
    unsigned long long key = ~0ull;
    if (in) {
      if (g < n_in) key = in[static_cast<long>(pair) * n_in + g];
    } else if (ok && g < n) {
      key = cycle_key(points, qs, q2o_p, o2q_p, g);
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= kCycMax; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < kCycMax / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < keep; j += blockDim.x)
    out[(static_cast<long>(pair) * chunks + chunk) * keep + j] = keys[j];
}

__global__ void __launch_bounds__(512)
cyclic_buddies_kernel(const float* __restrict__ points, const int* __restrict__ q_start,
                      const int* __restrict__ q_count, const int64_t* __restrict__ q2o,
                      const int64_t* __restrict__ o2q, const int64_t* __restrict__ top_ids, int topn,
                      const int* __restrict__ tpl_off, const int64_t* __restrict__ feat_perm,
                      const float* __restrict__ vertices, int max_q, int max_p, int top_k,
                      int64_t* __restrict__ out_qids, int64_t* __restrict__ out_vids,
                      float* __restrict__ out_dists, float* __restrict__ out_scores,
                      float* __restrict__ out_c2d, float* __restrict__ out_c3d,
                      int* __restrict__ out_count, const unsigned long long* __restrict__ cand, int n_cand) {
  extern __shared__ unsigned long long keys[];  // next pow2 >= number of keys sorted here
  const int pair = blockIdx.x;
  const int b = pair / topn;
  const int n = q_count[b];
  const long t = top_ids[pair];
  const int kk = min(top_k, n);
  // A retrieved template without bank rows (all-zero descriptor, cosine 0: it can enter the top-N when the other
  // scores are 0 too) has no 1-NN results - the k-NN kernel skips its items - so it yields no correspondences.
  const bool usable = t >= 0 && tpl_off[t + 1] > tpl_off[t];
  if (threadIdx.x == 0) out_count[pair] = usable ? kk : 0;
  if (n <= 0 || !usable) return;
  const int qs = q_start[b];
  const int ts = tpl_off[t];
  const int64_t* q2o_p = q2o + static_cast<long>(pair) * max_q;
  const int64_t* o2q_p = o2q + static_cast<long>(pair) * max_p;
  // Keys come either straight from the cycle distances (n <= kCycMax) or from the candidate list
  // produced by the pre-selection passes (large query sets, e.g. grid_cell_size = 1).
  const int n_keys = cand ? n_cand : n;
  int npow = 1;
  while (npow < n_keys) npow <<= 1;
  for (int i = threadIdx.x; i < npow; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (cand) {
      if (i < n_cand) key = cand[static_cast<long>(pair) * n_cand + i];
    } else if (i < n) {
      key = cycle_key(points, qs, q2o_p, o2q_p, i);
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= npow; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < npow / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  const float dmax = __uint_as_float(static_cast<unsigned>(keys[kk - 1] >> 32));
  for (int j = threadIdx.x; j < kk; j += blockDim.x) {
    const unsigned long long key = keys[j];
    const int qi = static_cast<int>(key & 0xffffffffu);
    const float d = __uint_as_float(static_cast<unsigned>(key >> 32));
    const long o = q2o_p[qi];
    const long feat = feat_perm ? feat_perm[ts + o] : static_cast<long>(ts) + o;
    const long orow = static_cast<long>(pair) * top_k + j;
    out_qids[orow] = qi;
    out_vids[orow] = feat;
    out_dists[orow] = d;
    out_scores[orow] = 1.0f - d / dmax;   // NaN when dmax == 0, as in the reference
    out_c2d[orow * 2] = points[(qs + qi) * 2];
    out_c2d[orow * 2 + 1] = points[(qs + qi) * 2 + 1];
    out_c3d[orow * 3] = vertices[feat * 3];
    out_c3d[orow * 3 + 1] = vertices[feat * 3 + 1];
    out_c3d[orow * 3 + 2] = vertices[feat * 3 + 2];
  }
}

}  // namespace

int tfidf_histogram(const int64_t* word_ids, const float* word_dists, int k, const int* row_start,
                    const int* row_count, int B, const float* idf, int W, int soft, float sigma2,
                    int sqrt_input, float* out, cudaStream_t stream) {
  FP_REQUIRE(W * 4 <= 200 * 1024, "tfidf: %d visual words do not fit in shared memory", W);
  if (B <= 0) return 0;
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(tfidf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
  }
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(B) * W * 4);
  tfidf_kernel<<<B, 256, W * 4, stream>>>(word_ids, word_dists, k, row_start, row_count, idf, W, soft,
                                          sigma2, sqrt_input, out);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int row_norm_f32(const float* x, float* out, int rows, int dim, cudaStream_t stream) {
  if (rows <= 0) return 0;
  int blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(rows) * dim * 4);
  row_norm_f32_kernel<<<blocks, 256, 0, stream>>>(x, out, rows, dim);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int bow_scores(const float* descs, const float* desc_norm, const float* q, int T, int B, int W,
               float* out, cudaStream_t stream) {
  if (T <= 0 || B <= 0) return 0;
  dim3 grid((T + BT - 1) / BT, (B + BC - 1) / BC);
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(T) * W * 4 * ((B + BC - 1) / BC) + static_cast<double>(B) * W * 4);
  bow_scores_kernel<<<grid, 256, 0, stream>>>(descs, desc_norm, q, T, B, W, 1e-8f, out);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int topk_rows(const float* x, int rows, int cols, int k, float* out_v, int64_t* out_i,
              cudaStream_t stream) {
  FP_REQUIRE(k >= 1 && k <= kMaxTopK, "topk: k=%d is outside [1,%d]", k, kMaxTopK);
  FP_REQUIRE(k <= cols, "topk: selected index k out of range (k=%d, size=%d)", k, cols);
  if (rows <= 0) return 0;
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(rows) * cols * 4);
  topk_rows_kernel<<<rows, 256, 0, stream>>>(x, cols, k, out_v, out_i);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int build_pair_items(const int64_t* top_ids, int num_pairs, int topn, const int* tpl_off,
                     const int* q_start, const int* q_count, int max_q, int max_p,
                     KnnItem* items_q2o, KnnItem* items_o2q, cudaStream_t stream) {
  if (num_pairs <= 0) return 0;
  const int total = num_pairs * ((max_q + 127) / 128 + (max_p + 127) / 128);
  ProfScope prof(PROF_RETRIEVAL, stream, total * 32.0);
  build_pair_items_kernel<<<(total + 255) / 256, 256, 0, stream>>>(
      top_ids, num_pairs, topn, tpl_off, q_start, q_count, max_q, max_p, items_q2o, items_o2q);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

size_t cyclic_buddies_workspace_bytes(int num_pairs, int max_q, int top_k) {
  if (max_q <= kCycMax) return 0;
  const int keep = top_k < kCycMax ? top_k : kCycMax;
  const size_t chunks0 = (max_q + kCycMax - 1) / kCycMax;
  // two ping-pong candidate buffers, the first level is the largest
  return 2 * static_cast<size_t>(num_pairs) * chunks0 * keep * sizeof(unsigned long long);
}

int cyclic_buddies(const float* points, const int* q_start, const int* q_count, const int64_t* q2o,
                   const int64_t* o2q, const int64_t* top_ids, int num_pairs, int topn,
                   const int* tpl_off, const int64_t* feat_perm, const float* vertices, int max_q,
                   int max_p, int top_k, int64_t* out_qids, int64_t* out_vids, float* out_dists,
                   float* out_scores, float* out_c2d, float* out_c3d, int* out_count, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
  if (num_pairs <= 0) return 0;
  FP_REQUIRE(top_k >= 1, "cyclic_buddies: top_k must be positive");
  const unsigned long long* cand = nullptr;
  int n_cand = 0;
  if (max_q > kCycMax) {
    // Large query sets (e.g. the reference's default grid_cell_size = 1 -> 176 400 points): reduce to
    // <= kCycMax candidates with chunk-wise pre-selection passes, then run the final sort.
    // every pass must shrink the candidate set: keep at most half a chunk
    FP_REQUIRE(top_k <= kCycMax / 2, "cyclic_buddies: top_k=%d exceeds %d for more than %d query points",
               top_k, kCycMax / 2, kCycMax);
    const size_t need = cyclic_buddies_workspace_bytes(num_pairs, max_q, top_k);
    FP_REQUIRE(workspace != nullptr && workspace_bytes >= need,
               "cyclic_buddies: %d query points per crop need a workspace of %zu bytes", max_q, need);
    unsigned long long* buf0 = static_cast<unsigned long long*>(workspace);
    unsigned long long* buf1 = buf0 + need / (2 * sizeof(unsigned long long));
    const unsigned long long* in = nullptr;
    int n_in = max_q;
    unsigned long long* out = buf0;
    while (n_in > kCycMax) {
      const int chunks = (n_in + kCycMax - 1) / kCycMax;
      dim3 grid(chunks, num_pairs);
      ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(num_pairs) * n_in * 8);
      cyclic_preselect_kernel<<<grid, 512, 0, stream>>>(points, q_start, q_count, q2o, o2q, top_ids, topn, max_q,
                                                        max_p, in, n_in, top_k, out);
      FP_CUDA_CHECK(cudaGetLastError());
      in = out;
      n_in = chunks * top_k;
      out = (out == buf0) ? buf1 : buf0;
    }
    cand = in;
    n_cand = n_in;
  }
  int npow = 1;
  const int n_sort = cand ? n_cand : max_q;
  while (npow < n_sort) npow <<= 1;
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(num_pairs) * (max_q + max_p) * 8);
  cyclic_buddies_kernel<<<num_pairs, 512, npow * 8, stream>>>(
      points, q_start, q_count, q2o, o2q, top_ids, topn, tpl_off, feat_perm, vertices, max_q, max_p,
      top_k, out_qids, out_vids, out_dists, out_scores, out_c2d, out_c3d, out_count, cand, n_cand);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
