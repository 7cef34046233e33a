// extern "C" entry points of libfoundpose_b200.so (declared in include/foundpose_b200.h).
// Plain pointers and sizes only: no torch types cross this boundary.
#include "../../include/foundpose_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace fp {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- launch accounting / profiling --------------------------------------------------------
struct ProfRecord {
  int category;
  double work;
  cudaEvent_t start, stop;
};
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<unsigned long long> g_launches_by_cat[PROF_NUM_CATEGORIES];
static bool g_profiling = false;
static std::mutex g_prof_mutex;
static std::vector<ProfRecord*> g_records;

ProfScope::ProfScope(int category, cudaStream_t stream, double work)
    : category_(category), stream_(stream), rec_(nullptr) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_launches_by_cat[category].fetch_add(1, std::memory_order_relaxed);
  if (g_profiling) {
    ProfRecord* r = new ProfRecord();
    r->category = category;
    r->work = work;
    cudaEventCreate(&r->start);
    cudaEventCreate(&r->stop);
    cudaEventRecord(r->start, stream);
    rec_ = r;
  }
}

ProfScope::~ProfScope() {
  if (rec_ != nullptr) {
    ProfRecord* r = static_cast<ProfRecord*>(rec_);
    cudaEventRecord(r->stop, stream_);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_records.push_back(r);
  }
}

}  // namespace fp

extern "C" {

unsigned long long fp_launch_count(void) { return fp::g_launches.load(); }

int fp_num_sms(void) { return fp::num_sms(); }

unsigned long long fp_launch_count_category(int category) {
  if (category < 0 || category >= fp::PROF_NUM_CATEGORIES) return 0;
  return fp::g_launches_by_cat[category].load();
}

int fp_profile_enable(int on) {
  fp::g_profiling = on != 0;
  return 0;
}

int fp_profile_read(int category, double* total_ms, double* total_work, int* launches, int reset) {
  if (category < 0 || category >= fp::PROF_NUM_CATEGORIES) {
    fp::set_last_error("fp_profile_read: bad category %d", category);
    return 1;
  }
  std::lock_guard<std::mutex> lock(fp::g_prof_mutex);
  double ms = 0.0, work = 0.0;
  int n = 0;
  for (fp::ProfRecord* r : fp::g_records) {
    if (r->category != category) continue;
    if (cudaEventSynchronize(r->stop) != cudaSuccess) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r->start, r->stop) == cudaSuccess) {
      ms += t;
      work += r->work;
      ++n;
    }
  }
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  if (launches) *launches = n;
  if (reset) {
    std::vector<fp::ProfRecord*> keep;
    for (fp::ProfRecord* r : fp::g_records) {
      if (r->category == category) {
        cudaEventDestroy(r->start);
        cudaEventDestroy(r->stop);
        delete r;
      } else {
        keep.push_back(r);
      }
    }
    fp::g_records.swap(keep);
  }
  return 0;
}


const char* fp_last_error(void) { return fp::g_last_error; }

int fp_version(void) { return FP_B200_VERSION; }

int fp_gemm_tn_f16(int epilogue, const void* A, int lda, const void* B, int ldb, int M, int N,
                   int K, const float* bias, const float* gamma, void* out_f16, int ld_f16,
                   float* out_f32, int ld_f32, void* stream) {
  fp::GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.gamma = gamma;
  p.out_f16 = static_cast<__half*>(out_f16); p.ld_f16 = ld_f16;
  p.out_f32 = out_f32; p.ld_f32 = ld_f32;
  if (epilogue == fp::EPI_PATCH_F32) {
    fp::set_last_error("fp_gemm_tn_f16: the patch-embed epilogue is internal to fp_vit_forward");
    return 1;
  }
  return fp::gemm_tn(epilogue, static_cast<const __half*>(A), lda, static_cast<const __half*>(B),
                     ldb, p, static_cast<cudaStream_t>(stream));
}

int fp_attention_debug_buffer(void* device_buffer) {
  fp::attention_set_debug_buffer(device_buffer);
  return 0;
}

int fp_gemm_force_1sm(int on) {
  fp::gemm_force_1sm(on);
  return 0;
}

int fp_umma_probe(const void* A, const void* B, float* out, int b_mn_major, void* stream) {
  return fp::umma_probe(static_cast<const __half*>(A), static_cast<const __half*>(B), out,
                        b_mn_major, static_cast<cudaStream_t>(stream));
}

int fp_layernorm_f16(const float* x, void* y_f16, const float* weight, const float* bias, int M,
                     int D, float eps, void* stream) {
  return fp::layernorm_f16(x, static_cast<__half*>(y_f16), weight, bias, M, D, eps,
                           static_cast<cudaStream_t>(stream));
}

int fp_attention_f16(const void* qkv_f16, void* out_f16, int B, int N, int heads, void* stream) {
  return fp::attention_f16(static_cast<const __half*>(qkv_f16), static_cast<__half*>(out_f16), B, N,
                           heads, static_cast<cudaStream_t>(stream));
}

static_assert(sizeof(fp_knn_item) == sizeof(fp::KnnItem), "fp_knn_item layout mismatch");

int fp_convert_rows_f16(const float* x, void* y_f16, int64_t rows, int dim, int l2_normalize,
                        void* stream) {
  return fp::convert_rows_f16(x, static_cast<__half*>(y_f16), rows, dim, l2_normalize,
                              static_cast<cudaStream_t>(stream));
}

int fp_row_sqnorm_f16(const void* x_f16, float* out, int64_t rows, int dim, void* stream) {
  return fp::row_sqnorm_f16(static_cast<const __half*>(x_f16), out, rows, dim,
                            static_cast<cudaStream_t>(stream));
}

int fp_unit_rows_f16(const void* x_f16, void* y_f16, float* sqnorm, int64_t rows, int dim, void* stream) {
  return fp::unit_rows_f16(static_cast<const __half*>(x_f16), static_cast<__half*>(y_f16), sqnorm, rows, dim,
                           static_cast<cudaStream_t>(stream));
}

int fp_split_rows_f16(const float* x, void* y_f16, int64_t rows, int dim, int pattern, int l2_normalize,
                      float scale, void* stream) {
  return fp::split_rows_f16(x, static_cast<__half*>(y_f16), rows, dim, pattern, l2_normalize, scale,
                            static_cast<cudaStream_t>(stream));
}

int fp_pca_project(const void* x_f16, const void* components_f16, const float* bias, int M, int D,
                   int d, float* out_f32, void* out_f16, void* stream) {
  if (M <= 0) return 0;
  if (out_f32 == nullptr && out_f16 == nullptr) {
    fp::set_last_error("fp_pca_project: out_f32 and out_f16 are both NULL");
    return 1;
  }
  fp::GemmParams p;
  p.M = M; p.N = d; p.K = D;
  p.bias = bias;
  p.out_f32 = out_f32; p.ld_f32 = d;
  p.out_f16 = static_cast<__half*>(out_f16); p.ld_f16 = d;
  return fp::gemm_tn(fp::EPI_BIAS_F32, static_cast<const __half*>(x_f16), D,
                     static_cast<const __half*>(components_f16), D, p,
                     static_cast<cudaStream_t>(stream));
}

int fp_knn_num_items(int q_rows) { return fp::knn_items_per_rows(q_rows); }

int fp_knn_items_dense(fp_knn_item* items, int q_total, int b_row0, int b_rows, void* stream) {
  return fp::knn_build_items_dense(reinterpret_cast<fp::KnnItem*>(items), q_total, b_row0, b_rows,
                                   static_cast<cudaStream_t>(stream));
}

int fp_knn_items_split(fp_knn_item* items, int q_total, int b_row0, int b_rows, int num_chunks,
                       int chunk_rows, void* stream) {
  return fp::knn_build_items_split(reinterpret_cast<fp::KnnItem*>(items), q_total, b_row0, b_rows, num_chunks,
                                   chunk_rows, static_cast<cudaStream_t>(stream));
}

int fp_knn_merge(const float* part_d, const int64_t* part_i, int num_chunks, int q_pad, int nq, int k,
                 int chunk_rows, int b_rows, int descending, float* out_d, int64_t* out_i, void* stream) {
  return fp::knn_merge(part_d, part_i, num_chunks, q_pad, nq, k, chunk_rows, b_rows, descending, out_d, out_i,
                       static_cast<cudaStream_t>(stream));
}

int fp_knn_search_items(const void* q_f16, int64_t q_rows_total, const float* q_sqnorm,
                        const void* bank_f16, int64_t bank_rows_total, const float* bank_sqnorm,
                        int dim, const fp_knn_item* items, int num_items, int metric, int k,
                        float* out_d, int64_t* out_i, void* stream) {
  if (metric != 0 && metric != 1) {
    fp::set_last_error("Metric %d is not supported.", metric);
    return 1;
  }
  return fp::knn_search_items(static_cast<const __half*>(q_f16), q_rows_total,
                              static_cast<const __half*>(bank_f16), bank_rows_total, dim,
                              reinterpret_cast<const fp::KnnItem*>(items), num_items, q_sqnorm,
                              bank_sqnorm, metric, k, out_d, out_i,
                              static_cast<cudaStream_t>(stream));
}

int fp_knn_search_pair_items(const void* q_f16, int64_t q_rows_total, const float* q_sqnorm,
                             const void* bank_f16, int64_t bank_rows_total, const float* bank_sqnorm,
                             int dim, const fp_knn_item* items, int num_items, int metric, int k,
                             float* out_d, int64_t* out_i, uint64_t* sync_counter, int sync_tiles, void* stream) {
  if (metric != 0 && metric != 1) {
    fp::set_last_error("Metric %d is not supported.", metric);
    return 1;
  }
  return fp::knn_search_pair_items(static_cast<const __half*>(q_f16), q_rows_total,
                                   static_cast<const __half*>(bank_f16), bank_rows_total, dim,
                                   reinterpret_cast<const fp::KnnItem*>(items), num_items, q_sqnorm,
                                   bank_sqnorm, metric, k, out_d, out_i,
                                   reinterpret_cast<unsigned long long*>(sync_counter), sync_tiles,
                                   static_cast<cudaStream_t>(stream));
}

int fp_knn_set_flags(int flags) {
  fp::knn_set_flags(flags);
  return 0;
}

int fp_crop_warp(const void* images, int src_is_f32, int num_images, int src_h, int src_w,
                 int channels, const uint8_t* masks, const double* params, int B, int crop_w,
                 int crop_h, float* out_images, uint8_t* out_masks, float* out_boxes,
                 int32_t* box_workspace, void* stream) {
  return fp::crop_warp(images, src_is_f32, num_images, src_h, src_w, channels, masks, params, B,
                       crop_w, crop_h, out_images, out_masks, out_boxes, box_workspace,
                       static_cast<cudaStream_t>(stream));
}

int fp_filter_points_by_mask(const float* points, int num_points, const uint8_t* masks, int B,
                             int H, int W, float* out_points, int32_t* out_ids,
                             int32_t* out_counts, int out_stride, void* stream) {
  return fp::filter_points_by_mask(points, num_points, masks, B, H, W, out_points, out_ids,
                                   out_counts, out_stride, static_cast<cudaStream_t>(stream));
}

int fp_sample_features(const float* tokens, int B, int Hp, int Wp, int C, const float* points,
                       const int32_t* counts, int stride, float img_w, float img_h, float* out_f32,
                       void* out_f16, void* stream) {
  return fp::sample_features(tokens, B, Hp, Wp, C, points, counts, stride, img_w, img_h, out_f32,
                             static_cast<__half*>(out_f16), static_cast<cudaStream_t>(stream));
}

int fp_calc_tfidf(const int64_t* word_ids, const float* word_dists, int k,
                  const int32_t* row_start, const int32_t* row_count, int B, const float* idf,
                  int W, int soft_assignment, float soft_sigma_squared, int sqrt_input, float* out,
                  void* stream) {
  return fp::tfidf_histogram(word_ids, word_dists, k, row_start, row_count, B, idf, W,
                             soft_assignment, soft_sigma_squared, sqrt_input, out,
                             static_cast<cudaStream_t>(stream));
}

int fp_row_norm_f32(const float* x, float* out, int rows, int dim, void* stream) {
  return fp::row_norm_f32(x, out, rows, dim, static_cast<cudaStream_t>(stream));
}

int fp_bow_scores(const float* descs, const float* desc_norm, const float* q, int T, int B, int W,
                  float* out, void* stream) {
  return fp::bow_scores(descs, desc_norm, q, T, B, W, out, static_cast<cudaStream_t>(stream));
}

int fp_kmeans_update(const float* samples, const int64_t* assign, int64_t n, int d, int k, uint64_t* sums,
                     int32_t* counts, float* centroids, void* stream) {
  return fp::kmeans_update(samples, assign, n, d, k, reinterpret_cast<unsigned long long*>(sums), counts, centroids,
                           static_cast<cudaStream_t>(stream));
}

int fp_pnp_ransac(const float* coord_2d, const float* coord_3d, const int32_t* counts,
                  const double* intrinsics, int P, int M, int iters, double thresh, double confidence,
                  uint64_t seed, int problem_offset, int32_t* success, double* out_R, double* out_t,
                  uint8_t* inlier_mask, int32_t* num_inliers, int32_t* iters_run, int32_t* best_hyp,
                  void* stream) {
  return fp::pnp_ransac(coord_2d, coord_3d, counts, intrinsics, P, M, iters, thresh, confidence, seed,
                        problem_offset, success, out_R, out_t, inlier_mask, num_inliers, iters_run, best_hyp,
                        static_cast<cudaStream_t>(stream));
}

int fp_topk_rows(const float* x, int rows, int cols, int k, float* out_v, int64_t* out_i,
                 void* stream) {
  return fp::topk_rows(x, rows, cols, k, out_v, out_i, static_cast<cudaStream_t>(stream));
}

int fp_build_pair_items(const int64_t* top_ids, int num_pairs, int topn, const int32_t* tpl_off,
                        const int32_t* q_start, const int32_t* q_count, int max_q, int max_p,
                        fp_knn_item* items_q2o, fp_knn_item* items_o2q, void* stream) {
  return fp::build_pair_items(top_ids, num_pairs, topn, tpl_off, q_start, q_count, max_q, max_p,
                              reinterpret_cast<fp::KnnItem*>(items_q2o),
                              reinterpret_cast<fp::KnnItem*>(items_o2q),
                              static_cast<cudaStream_t>(stream));
}

int fp_cyclic_buddies(const float* points, const int32_t* q_start, const int32_t* q_count,
                      const int64_t* q2o, const int64_t* o2q, const int64_t* top_ids,
                      int num_pairs, int topn, const int32_t* tpl_off, const int64_t* feat_perm,
                      const float* vertices, int max_q, int max_p, int top_k, int64_t* out_query_ids,
                      int64_t* out_vertex_ids, float* out_dists, float* out_scores,
                      float* out_coord_2d, float* out_coord_3d, int32_t* out_count, void* workspace,
                      uint64_t workspace_bytes, void* stream) {
  return fp::cyclic_buddies(points, q_start, q_count, q2o, o2q, top_ids, num_pairs, topn, tpl_off,
                            feat_perm, vertices, max_q, max_p, top_k, out_query_ids, out_vertex_ids,
                            out_dists, out_scores, out_coord_2d, out_coord_3d, out_count, workspace,
                            static_cast<size_t>(workspace_bytes), static_cast<cudaStream_t>(stream));
}

uint64_t fp_cyclic_buddies_workspace_bytes(int num_pairs, int max_q, int top_k) {
  return fp::cyclic_buddies_workspace_bytes(num_pairs, max_q, top_k);
}

}  // extern "C"
