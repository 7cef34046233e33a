// Internal C++ declarations shared by the translation units of libfoundpose_b200.so.
// The public, torch-free C ABI is include/foundpose_b200.h (implemented in api.cu).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fp {

// ---- gemm_tcgen05.cu --------------------------------------------------------------------
enum GemmEpilogue : int {
  EPI_BIAS_F16 = 0,       // out_f16 = acc + bias                         (qkv)
  EPI_BIAS_GELU_F16 = 1,  // out_f16 = gelu_erf(acc + bias)               (mlp.fc1 + act)
  EPI_RESID_F32 = 2,      // out_f32 += gamma * (acc + bias)              (proj/fc2 + LayerScale + residual)
  EPI_PATCH_F32 = 3,      // token stream row remap + bias + pos-embed    (patch embedding)
  EPI_BIAS_F32 = 4,       // out_f32 = acc + bias (+ fp16 copy)           (PCA projection)
  // LayerNorm fused into the GEMMs around it (layers/block.py:63,75 + the Linear that follows):
  //   LN(x) W^T + b = rstd * (x W'^T - mu * colsum(W')) + b'   with W' = W diag(gamma), b' = b + W beta,
  // so the CONSUMER runs on the raw fp16 copy of the residual rows and applies (mu, rstd) per row in its epilogue,
  // and the PRODUCER of the residual rows (patch embed / proj / fc2 epilogue) emits that fp16 copy and the
  // per-row partial sums (sum x, sum x^2 per 128 columns) the statistics are built from.
  EPI_LN_BIAS_F16 = 5,       // out_f16 = rstd_m * (acc - mu_m * colsum_n) + bias_n               (norm1 + qkv)
  EPI_LN_BIAS_GELU_F16 = 6,  // out_f16 = gelu_erf(rstd_m * (acc - mu_m * colsum_n) + bias_n)     (norm2 + fc1 + act)
  EPI_RESID_LN_F32 = 7,      // EPI_RESID_F32 (x read-modify-write through shared memory) + x16 + row partial sums
  EPI_PATCH_LN_F32 = 8,      // EPI_PATCH_F32 + x16 + row partial sums
};

struct GemmParams {
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;   // [N]
  const float* gamma = nullptr;  // [N]  (EPI_RESID_F32)
  __half* out_f16 = nullptr;
  int ld_f16 = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  // EPI_PATCH_F32: GEMM row m = b * patches_per_img + p  ->  token row b * tokens_per_img + tok_off + p
  int patches_per_img = 1;
  int tokens_per_img = 1;
  int tok_off = 0;
  const float* pos = nullptr;    // [patches_per_img, N] fp32
  int flags = 0;                 // A/B tuning switches (see gemm_force_1sm)
  // EPI_LN_*: statistics of the A rows.  ln_stats [M, ln_slots, 2] partial (sum, sum of squares) per 128 columns
  // of the ln_dim-wide rows, ln_colsum [N] = sum_k W'[n, k] of the fp16 weights actually multiplied.
  const float* ln_stats = nullptr;
  int ln_slots = 0;
  const float* ln_colsum = nullptr;
  int ln_dim = 0;
  float ln_eps = 1e-6f;
  // EPI_RESID_LN_F32 / EPI_PATCH_LN_F32: fp16 copy of the rows written to out_f32 and their partial sums
  // stats_out [rows, N / 128, 2] (row index = row of out_f32).
  __half* x16 = nullptr;
  int ld_x16 = 0;
  float* stats_out = nullptr;
};

int gemm_pick_bn(int M, int N);
void gemm_force_1sm(int on);   // A/B switch: disable the cta_group::2 kernel
int gemm_tn(int epi, const __half* A, int lda, const __half* B, int ldb, const GemmParams& p,
            cudaStream_t stream);
int umma_probe(const __half* A, const __half* B, float* out, int b_mn_major, cudaStream_t stream);


// ---- vit_ops.cu -------------------------------------------------------------------------
int patchify_normalize(const float* img, __half* out, int B, int H, int W, int ps, int Kpad,
                       cudaStream_t stream);
int init_special_tokens(float* x, const float* cls_pos, const float* reg, int B, int ntok, int R,
                        int D, cudaStream_t stream, __half* x16 = nullptr, float* stats = nullptr,
                        int stat_slots = 0);
int layernorm_f16(const float* x, __half* y, const float* w, const float* b, int M, int D, float eps,
                  cudaStream_t stream);
int final_norm_tokens(const float* x, const float* w, const float* b, float* out_tok,
                      __half* out_tok16, float* out_cls, int B, int ntok, int R, int P, int D,
                      int apply_norm, float eps, cudaStream_t stream);

// ---- attention_tcgen05.cu ---------------------------------------------------------------
int attention_f16(const __half* qkv, __half* out, int B, int N, int heads, cudaStream_t stream);
void attention_set_flags(int flags);   // experiment switches, 0 in production
void attention_set_debug_buffer(void* p);   // device buffer of 5 x 2048 u64 timeline events, or NULL


// ---- knn_tcgen05.cu ---------------------------------------------------------------------
// One k-NN work item: query rows [q_row0, q_row0+q_rows) (q_rows <= 128) against bank rows
// [b_row0, b_row0+b_rows); results go to output rows [out_row0, out_row0+q_rows); returned
// indices are relative to b_row0. Mirrors fp_knn_item in the public header.
struct KnnItem {
  int q_row0;
  int q_rows;
  int b_row0;
  int b_rows;
  long long out_row0;
  long long pad;   // pair kernel: number of clusters whose items sweep the same bank rows in lockstep (0 = none)
};
int knn_items_per_rows(int rows);
int knn_build_items_dense(KnnItem* items, int q_total, int b_row0, int b_rows, cudaStream_t stream);
int knn_build_items_split(KnnItem* items, int q_total, int b_row0, int b_rows, int num_chunks,
                          int chunk_rows, cudaStream_t stream);
int knn_merge(const float* part_d, const int64_t* part_i, int num_chunks, int q_pad, int nq, int k,
              int chunk_rows, int b_rows, int descending, float* out_d, int64_t* out_i, cudaStream_t stream);
int row_sqnorm_f16(const __half* x, float* out, long rows, int dim, cudaStream_t stream);
int convert_rows_f16(const float* x, __half* y, long rows, int dim, int l2_normalize,
                     cudaStream_t stream);
int unit_rows_f16(const __half* x, __half* y, float* sqnorm, long rows, int dim, cudaStream_t stream);
int split_rows_f16(const float* x, __half* y, long rows, int dim, int pattern, int l2_normalize, float scale,
                   cudaStream_t stream);
int knn_search_items(const __half* q, long q_rows_total, const __half* x, long x_rows_total, int dim,
                     const KnnItem* items, int num_items, const float* qnorm, const float* xnorm,
                     int metric_ip, int k, float* out_d, int64_t* out_i, cudaStream_t stream);


// Pair kernel (cta_group::2, queries resident in shared memory): items hold up to 256 query rows.
int knn_search_pair_items(const __half* q, long q_rows_total, const __half* x, long x_rows_total, int dim,
                          const KnnItem* items, int num_items, const float* qnorm, const float* xnorm,
                          int metric_ip, int k, float* out_d, int64_t* out_i, unsigned long long* sync_counter,
                          int sync_tiles, cudaStream_t stream);


void knn_set_flags(int flags);   // experiment switches of the pair kernel, 0 in production

// ---- crop_warp.cu -----------------------------------------------------------------------
int crop_warp(const void* images, int src_is_f32, int num_images, int src_h, int src_w, int channels,
              const uint8_t* masks, const double* params, int B, int crop_w, int crop_h,
              float* out_images, uint8_t* out_masks, float* out_boxes, int* box_workspace,
              cudaStream_t stream);

// ---- feature_ops.cu ---------------------------------------------------------------------
int filter_points_by_mask(const float* points, int num_points, const uint8_t* masks, int B, int H,
                          int W, float* out_points, int* out_ids, int* out_counts, int out_stride,
                          cudaStream_t stream);
int sample_features(const float* tokens, int B, int Hp, int Wp, int C, const float* points,
                    const int* counts, int stride, float img_w, float img_h, float* out_f32,
                    __half* out_f16, cudaStream_t stream);

// ---- retrieval.cu -----------------------------------------------------------------------
int tfidf_histogram(const int64_t* word_ids, const float* word_dists, int k, const int* row_start,
                    const int* row_count, int B, const float* idf, int W, int soft, float sigma2,
                    int sqrt_input, float* out, cudaStream_t stream);
int row_norm_f32(const float* x, float* out, int rows, int dim, cudaStream_t stream);
int bow_scores(const float* descs, const float* desc_norm, const float* q, int T, int B, int W,
               float* out, cudaStream_t stream);
int topk_rows(const float* x, int rows, int cols, int k, float* out_v, int64_t* out_i,
              cudaStream_t stream);
int build_pair_items(const int64_t* top_ids, int num_pairs, int topn, const int* tpl_off,
                     const int* q_start, const int* q_count, int max_q, int max_p,
                     KnnItem* items_q2o, KnnItem* items_o2q, cudaStream_t stream);
int cyclic_buddies(const float* points, const int* q_start, const int* q_count, const int64_t* q2o,
                   const int64_t* o2q, const int64_t* top_ids, int num_pairs, int topn,
                   const int* tpl_off, const int64_t* feat_perm, const float* vertices, int max_q,
                   int max_p, int top_k, int64_t* out_qids, int64_t* out_vids, float* out_dists,
                   float* out_scores, float* out_c2d, float* out_c3d, int* out_count, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);
size_t cyclic_buddies_workspace_bytes(int num_pairs, int max_q, int top_k);

// ---- kmeans.cu --------------------------------------------------------------------------
int kmeans_update(const float* x, const int64_t* assign, long long n, int d, int k, unsigned long long* sums,
                  int* counts, float* centroids, cudaStream_t stream);

// ---- pnp_ransac.cu ----------------------------------------------------------------------
int pnp_ransac(const float* coord_2d, const float* coord_3d, const int* counts, const double* intrinsics, int P,
               int M, int iters, double thresh, double confidence, unsigned long long seed, int problem_offset,
               int* success, double* out_R, double* out_t, unsigned char* inlier_mask, int* num_inliers,
               int* iters_run, int* best_hyp, cudaStream_t stream);

}  // namespace fp
