/* foundpose_b200 — C ABI of the B200-native FoundPose per-crop hot path.
 *
 * The reference (facebookresearch/foundpose) has no FFI layer: its "plugin interface" for this
 * path is the Python API of utils/{dinov2_utils,feature_util,projector_util,knn_util,
 * template_util,corresp_util}.py, which bottoms out in torch ATen, faiss and scikit-learn calls.
 * Each entry point below replaces one of those third-party calls and cites the reference call
 * site (file:line relative to the reference root) whose arithmetic it reproduces.  The Python
 * mirror of the reference modules (the foundpose_b200/utils package) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success; non-zero = failure, message via fp_last_error()
 *     (1 = bad argument, 2 = CUDA runtime error, 3 = driver/TMA descriptor error)
 *   - all data pointers are DEVICE pointers (tensor.data_ptr()), row-major, caller-owned and
 *     caller-allocated; nothing is allocated inside except per-handle workspaces
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream)
 *   - "f16" buffers are IEEE binary16; index outputs are int64 to match faiss / torch.topk
 */
#ifndef FOUNDPOSE_B200_H_
#define FOUNDPOSE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP_B200_VERSION 200

/* Last error message of the calling thread (never NULL). */
const char* fp_last_error(void);
/* Library version (FP_B200_VERSION). */
int fp_version(void);

/* ---- dense contraction on the 5th-gen tensor cores -------------------------------------
 * C[M,N] = A[M,K] . B[N,K]^T, fp16 in, fp32 accumulate (tcgen05.mma + TMA + TMEM).
 * Replaces torch.nn.functional.linear at external/dinov2/dinov2/layers/attention.py:58,67,
 * layers/mlp.py:35,38 and numpy sgemm inside sklearn PCA.transform (utils/projector_util.py:67).
 * epilogue: 0 out_f16 = acc+bias | 1 out_f16 = gelu_erf(acc+bias) |
 *           2 out_f32 += gamma*(acc+bias) | 4 out_f32 = acc+bias (and out_f16 copy if non-NULL)
 * N % 128 == 0, K % 64 == 0, lda/ldb % 8 == 0. */
int fp_gemm_tn_f16(int epilogue, const void* A, int lda, const void* B, int ldb, int M, int N,
                   int K, const float* bias, const float* gamma, void* out_f16, int ld_f16,
                   float* out_f32, int ld_f32, void* stream);

/* A/B switch for benchmarking: 1 = always use the 1-CTA kernel instead of the cta_group::2 pair
 * kernel (default 0). */
int fp_gemm_force_1sm(int on);

/* Debug: device buffer of 5 x 2048 uint64 that CTA 0 of the attention kernel fills with
 * (tag << 48 | globaltimer ns) events per warp role (tools/attn_timeline.py); NULL disables. */
int fp_attention_debug_buffer(void* device_buffer);

/* Single-tile UMMA descriptor probe used by the GPU tests (not on the hot path). */
int fp_umma_probe(const void* A, const void* B, float* out, int b_mn_major, void* stream);

/* ---- ViT building blocks (exported for unit tests; fp_vit_forward sequences them) -------- */
/* LayerNorm(eps) over the last dimension: x fp32 [M,D] -> y f16 [M,D].
 * Replaces torch layer_norm at external/dinov2/dinov2/layers/block.py:63,75. D % 128 == 0. */
int fp_layernorm_f16(const float* x, void* y_f16, const float* weight, const float* bias, int M,
                     int D, float eps, void* stream);

/* Fused softmax(q k^T / sqrt(64)) v for all heads (head_dim 64): qkv f16 [B*N, 3*heads*64]
 * (columns [q|k|v], each head-major) -> out f16 [B*N, heads*64].
 * Replaces the bmm/softmax/bmm of external/dinov2/dinov2/layers/attention.py:60-66. */
int fp_attention_f16(const void* qkv_f16, void* out_f16, int B, int N, int heads, void* stream);

/* ---- DINOv2 feature extractor -------------------------------------------------------------
 * Replaces DinoFeatureExtractor.forward (utils/dinov2_utils.py:115-158) = ImageNet normalise +
 * DinoVisionTransformer blocks 0..layer (external/dinov2/dinov2/models/vision_transformer.py:
 * 213-232, 254-270; layers/block.py:89-114) + final LayerNorm over [cls | patch] tokens. */
typedef struct fp_vit fp_vit; /* opaque */

typedef struct {
  int embed_dim;            /* D: 384 / 768 / 1024; must equal num_heads * 64 */
  int num_heads;
  int num_blocks;           /* number of blocks whose weights are supplied (>= layer + 1) */
  int num_register_tokens;  /* 0 or 4 */
  int patch_size;           /* 14 */
  int img_h, img_w;         /* multiples of patch_size */
  int fuse_layernorm;       /* 1: norm1 / norm2 (layers/block.py:63,75) are folded into the GEMMs around them - the
                               caller passes gamma-folded qkv_w / fc1_w, beta-folded qkv_b / fc1_b and the column
                               sums below; needs embed_dim % 256 == 0.  0: separate LayerNorm kernels. */
} fp_vit_config;

typedef struct {            /* device pointers; the caller keeps them alive while the handle lives */
  const void* patch_w;      /* f16 [D, Kpad]: PatchEmbed.proj.weight flattened (c,dy,dx), zero padded
                               to Kpad = fp_vit_patch_k() columns */
  const float* patch_b;     /* [D] */
  const float* cls_pos;     /* [D]  cls_token + pos_embed[0] */
  const float* reg_tokens;  /* [R, D] or NULL */
  const float* pos_patch;   /* [P, D] positional embedding resampled to the patch grid
                               (interpolate_pos_encoding, vision_transformer.py:179-211; done once
                               per image size by the host) */
  const float* norm_w;      /* [D] final LayerNorm */
  const float* norm_b;
} fp_vit_weights;

typedef struct {            /* one transformer block; weights f16 [out, in], vectors fp32 */
  const float* norm1_w; const float* norm1_b;
  const void* qkv_w;  const float* qkv_b;    /* [3D, D], [3D] */
  const void* proj_w; const float* proj_b;   /* [D, D],  [D]  */
  const float* ls1;                          /* [D] LayerScale gamma */
  const float* norm2_w; const float* norm2_b;
  const void* fc1_w; const float* fc1_b;     /* [4D, D], [4D] */
  const void* fc2_w; const float* fc2_b;     /* [D, 4D], [D]  */
  const float* ls2;
  /* fuse_layernorm = 1 only.  With W' = f16(W diag(norm_w)) passed as qkv_w / fc1_w and b' = b + W norm_b passed
   * as qkv_b / fc1_b:  LN(x) W^T + b = rstd (x W'^T - mu colsum) + b',  colsum[n] = sum_k W'[n,k] (fp32). */
  const float* qkv_colsum;                   /* [3D] */
  const float* fc1_colsum;                   /* [4D] */
} fp_vit_block_weights;

/* Allocates the activation workspace for up to max_batch images. */
int fp_vit_create(const fp_vit_config* cfg, const fp_vit_weights* weights,
                  const fp_vit_block_weights* blocks, int max_batch, fp_vit** out);
void fp_vit_destroy(fp_vit* handle);
/* Padded length of one flattened patch (columns of patch_w). */
int fp_vit_patch_k(const fp_vit* handle);
/* images fp32 [B,3,H,W] in [0,1] -> out_tokens fp32 [B, P, D] (patch tokens of block `layer`,
 * row-major over the patch grid; feature_maps = out_tokens viewed as B x Hp x Wp x D),
 * optional f16 copy, optional out_cls fp32 [B, D].
 * facet: 0 token, 1 query, 2 key, 3 value (utils/dinov2_utils.py:160-196). */
int fp_vit_forward(fp_vit* handle, const float* images, int batch, int layer, int facet,
                   int apply_norm, float* out_tokens, void* out_tokens_f16, float* out_cls,
                   void* stream);

/* ---- descriptor conversion ---------------------------------------------------------------- */
/* fp32 rows -> f16 rows, optionally L2-normalised first (cosine metric: utils/knn_util.py:57,93). */
int fp_convert_rows_f16(const float* x, void* y_f16, int64_t rows, int dim, int l2_normalize,
                        void* stream);
/* out[r] = ||x[r]||^2 (fp32) of f16 rows; the ||x||^2 term of faiss's L2 expansion. */
int fp_row_sqnorm_f16(const void* x_f16, float* out, int64_t rows, int dim, void* stream);

/* f16 rows -> L2-normalised f16 rows and their ||.||^2 (fp32) in one pass: the query side of the cosine metric
 * (faiss.normalize_L2 at utils/knn_util.py:93) for descriptors that are already packed as f16. */
int fp_unit_rows_f16(const void* x_f16, void* y_f16, float* sqnorm, int64_t rows, int dim, void* stream);

/* fp32 rows [rows, dim] -> f16 rows [rows, 3*dim] holding the split x = hi + lo as [hi | lo | hi] (pattern 0, index
 * side) or [hi | hi | lo] (pattern 1, query side): one inner-product search over the 3*dim columns then evaluates
 * <a,b> to fp32 accuracy on the tensor cores.  Rows are optionally L2-normalised first (zero rows stay zero) and
 * multiplied by `scale` (use a power of two, e.g. 1024, to keep the lo parts out of the f16 subnormals; the inner
 * products come back multiplied by scale_a * scale_b).  Used for the bag-of-words cosine scores
 * torch.nn.functional.cosine_similarity(template_descs, query_tfidf) of utils/template_util.py:160-164. */
int fp_split_rows_f16(const float* x, void* y_f16, int64_t rows, int dim, int pattern, int l2_normalize,
                      float scale, void* stream);

/* ---- PCA projection ------------------------------------------------------------------------
 * out = x . components^T + bias, bias = -(mean . components^T)  (sklearn PCA.transform with
 * whiten=False as called at utils/projector_util.py:66-69).  x f16 [M,D], components f16 [d,D],
 * d % 128 == 0, D % 64 == 0.  Writes fp32 [M,d] (if out_f32 != NULL) and an f16 copy (if out_f16 != NULL, the
 * k-NN operand); at least one of the two. */
int fp_pca_project(const void* x_f16, const void* components_f16, const float* bias, int M, int D,
                   int d, float* out_f32, void* out_f16, void* stream);

/* ---- brute-force k-NN ----------------------------------------------------------------------
 * Replaces faiss.IndexFlatL2.search / IndexFlatIP.search (utils/knn_util.py:83, 95).
 * One item = <=128 consecutive query rows against one contiguous run of bank rows; indices are
 * returned relative to b_row0 (a per-template index in the reference, scripts/infer.py:224-239). */
typedef struct {
  int32_t q_row0, q_rows;   /* query rows [q_row0, q_row0 + q_rows), q_rows <= 128 (0 = skip) */
  int32_t b_row0, b_rows;   /* bank rows  [b_row0, b_row0 + b_rows) */
  int64_t out_row0;         /* results go to output rows [out_row0, out_row0 + q_rows) */
  int64_t reserved;         /* fp_knn_search_pair_items: sweep-barrier participants of this item (see there); else 0 */
} fp_knn_item;

/* Number of items that cover q_rows query rows of one dense problem. */
int fp_knn_num_items(int q_rows);
/* Fills items[0 .. fp_knn_num_items(q_total)) for "all query rows vs bank rows [b_row0, +b_rows)". */
int fp_knn_items_dense(fp_knn_item* items, int q_total, int b_row0, int b_rows, void* stream);
/* Few queries vs a large bank: split the bank into num_chunks slices of chunk_rows rows
 * (chunk_rows % 256 == 0) so that every SM streams its own slice (HBM-bound search).  Fills
 * fp_knn_num_items(q_total) * num_chunks items; chunk c writes its partial top-k lists to output rows
 * [c * q_pad, c * q_pad + q_total), q_pad = fp_knn_num_items(q_total) * 128.  fp_knn_merge reduces
 * the partial lists [num_chunks, q_pad, k] to the final [nq, k] ordered by (distance, index). */
int fp_knn_items_split(fp_knn_item* items, int q_total, int b_row0, int b_rows, int num_chunks,
                       int chunk_rows, void* stream);
int fp_knn_merge(const float* part_d, const int64_t* part_i, int num_chunks, int q_pad, int nq, int k,
                 int chunk_rows, int b_rows, int descending, float* out_d, int64_t* out_i, void* stream);
/* metric 0: squared L2, ascending (out_d = max(||q||^2 + ||x||^2 - 2<q,x>, 0));
 * metric 1: inner product, descending (out_d = <q,x>).  k <= 16, dim % 64 == 0.
 * q_sqnorm / bank_sqnorm: fp_row_sqnorm_f16 of the respective rows (unused for metric 1).
 * out_d fp32 [rows,k], out_i int64 [rows,k]; ties resolve to the lower index. */
int fp_knn_search_items(const void* q_f16, int64_t q_rows_total, const float* q_sqnorm,
                        const void* bank_f16, int64_t bank_rows_total, const float* bank_sqnorm,
                        int dim, const fp_knn_item* items, int num_items, int metric, int k,
                        float* out_d, int64_t* out_i, void* stream);
/* Same search for the tensor-bound regime (many queries x a large bank: the benchmark search K4 of
 * BASELINE.json configs 3 and 5, queries of all crops vs the WHOLE bank; utils/knn_util.py:65-106 with
 * the full feat_vectors as the index).  Items hold up to 256 query rows (q_rows <= 256): a cluster of
 * two CTAs keeps them resident in shared memory and sweeps the bank segment with cta_group::2 MMAs.
 * Same outputs, tie rule and indices relative to b_row0 as fp_knn_search_items.
 * Sweep barrier (optional): items i, i + C, i + 2C, ... run on cluster i % C (C = fp_num_sms() / 2).  When the items of
 * one wave [wC, wC + C) cover the same bank rows, set their `reserved` field to the number of items in that wave and
 * pass a device `sync_counter` (TWO uint64 words, zeroed by the call) + `sync_tiles` > 0: every sync_tiles bank tiles (256 rows each) the clusters of a
 * wave meet at a grid barrier, so the bank streams from HBM once per wave instead of once per cluster.  Every item
 * of a wave must then have the same b_rows; items with reserved = 0 do not take part.  The barrier is an optimisation
 * only: a cluster that waits a few ms in vain (grid not co-resident) switches it off for the rest of the launch.
 * sync_counter may be NULL. */
int fp_knn_search_pair_items(const void* q_f16, int64_t q_rows_total, const float* q_sqnorm,
                             const void* bank_f16, int64_t bank_rows_total, const float* bank_sqnorm,
                             int dim, const fp_knn_item* items, int num_items, int metric, int k,
                             float* out_d, int64_t* out_i, uint64_t* sync_counter, int sync_tiles, void* stream);

/* Experiment switches of the k-NN kernels (tools/k4_probe.py, tools/knn_hbm_noscan.py): bit 0 = skip the epilogue scan,
 * bit 1 = stream the queries instead of keeping them resident, bit 3 = no sweep barrier; bit 2 (TMEM reads only) exists
 * only in builds with -DFP_KNN_EXPERIMENTS.  0 (default) in production. */
int fp_knn_set_flags(int flags);

/* ---- crop stage in front of the extractor ------------------------- */
/* warp_image x2 + array_to_tensor + calc_2d_box for B instances (scripts/infer.py:427-456,
 * utils/misc.py:458-519, cv2.remap semantics).  images: [num_images, src_h, src_w, channels]
 * interleaved, uint8 (scaled by 1/255 like infer.py:396) or fp32 when src_is_f32; masks: uint8
 * [B, src_h, src_w] modal masks or NULL.  params: double [B, 40] per instance:
 *   [0:2] f and [2:4] c of the virtual crop camera, [4:13] R and [13:16] t of its T_world_from_eye
 *   (row-major), [16:25] R and [25:28] t of the source camera's T_world_from_eye, [28:30] f and
 *   [30:32] c of the source camera, [32] index of the source image, [33:40] reserved.
 * Writes out_images fp32 [B, channels, crop_h, crop_w] (INTER_LINEAR == INTER_AREA inside remap,
 * constant border 0), out_masks uint8 [B, crop_h, crop_w] (INTER_NEAREST) and out_boxes fp32 [B,4]
 * = (x1,y1,x2,y2) of the warped mask, zeros when it is empty.  box_workspace: int32 [4*B]. */
int fp_crop_warp(const void* images, int src_is_f32, int num_images, int src_h, int src_w,
                 int channels, const uint8_t* masks, const double* params, int B, int crop_w,
                 int crop_h, float* out_images, uint8_t* out_masks, float* out_boxes,
                 int32_t* box_workspace, void* stream);

/* ---- query points and feature sampling ---------------------------------------------------- */
/* filter_points_by_mask (utils/feature_util.py:75-97) for B crops sharing one point grid:
 * keeps points whose rounded pixel lies strictly inside the canvas and inside masks[b] (uint8,
 * [B,H,W]); order preserved.  Outputs use a fixed stride of out_stride (>= num_points) rows per
 * crop: out_points fp32 [B,out_stride,2], out_ids int32 [B,out_stride], out_counts int32 [B]. */
int fp_filter_points_by_mask(const float* points, int num_points, const uint8_t* masks, int B,
                             int H, int W, float* out_points, int32_t* out_ids,
                             int32_t* out_counts, int out_stride, void* stream);
/* sample_feature_map_at_points (utils/feature_util.py:100-131): bilinear grid_sample,
 * align_corners=False, zero padding.  tokens fp32 [B, Hp*Wp, C] (token-major feature map);
 * points fp32 [B,stride,2] image coordinates; counts int32 [B] or NULL (= stride).  Rows past
 * counts[b] are zero-filled.  Writes fp32 and/or f16 [B*stride, C]. */
int fp_sample_features(const float* tokens, int B, int Hp, int Wp, int C, const float* points,
                       const int32_t* counts, int stride, float img_w, float img_h, float* out_f32,
                       void* out_f16, void* stream);

/* ---- tf-idf bag-of-words template retrieval ----------------------------------------------- */
/* calc_tfidf (utils/template_util.py:31-71) for B crops: rows [row_start[b], +row_count[b]) of
 * word_ids int64 [*,k] / word_dists fp32 [*,k] -> out fp32 [B,W].  sqrt_input=1 applies the sqrt
 * of template_util.py:27 to word_dists first (query side). */
int fp_calc_tfidf(const int64_t* word_ids, const float* word_dists, int k,
                  const int32_t* row_start, const int32_t* row_count, int B, const float* idf,
                  int W, int soft_assignment, float soft_sigma_squared, int sqrt_input, float* out,
                  void* stream);
/* out[r] = ||x[r]|| (fp32 rows). */
int fp_row_norm_f32(const float* x, float* out, int rows, int dim, void* stream);
/* torch.nn.functional.cosine_similarity(template_descs, tile(query_tfidf)) for B crops at once
 * (utils/template_util.py:167-169): descs fp32 [T,W], desc_norm [T], q fp32 [B,W] -> out [B,T]. */
int fp_bow_scores(const float* descs, const float* desc_norm, const float* q, int T, int B, int W,
                  float* out, void* stream);
/* torch.topk(x, k, sorted=True) per row (utils/template_util.py:172-174), ties -> lower index. */
int fp_topk_rows(const float* x, int rows, int cols, int k, float* out_v, int64_t* out_i,
                 void* stream);

/* ---- cyclic-buddies correspondences -------------------------------------------------------- */
/* k-NN work items of the two 1-NN searches per (crop, retrieved template) pair
 * (utils/corresp_util.py:46-47).  pair p = b * topn + j uses template top_ids[p];
 * items_q2o: ceil(max_q/128) items per pair (crop queries vs template rows, results at rows
 * p*max_q + i); items_o2q: ceil(max_p/128) items per pair (template rows vs crop queries,
 * results at rows p*max_p + o).  tpl_off int32 [T+1] = CSR offsets of the templates' bank rows. */
int fp_build_pair_items(const int64_t* top_ids, int num_pairs, int topn, const int32_t* tpl_off,
                        const int32_t* q_start, const int32_t* q_count, int max_q, int max_p,
                        fp_knn_item* items_q2o, fp_knn_item* items_o2q, void* stream);
/* cyclic_buddies_matching + the gathers of establish_correspondences
 * (utils/corresp_util.py:50-68, 135-155) for every pair; outputs have top_k rows per pair, the
 * first out_count[p] = min(top_k, q_count[b]) are valid, ordered by (cycle distance, query id).
 * feat_perm (int64, may be NULL) maps sorted bank rows back to original feature ids.
 * More than 4096 query points per crop (e.g. the reference default grid_cell_size = 1) need a scratch
 * buffer of fp_cyclic_buddies_workspace_bytes() bytes; otherwise workspace may be NULL. */
int fp_cyclic_buddies(const float* points, const int32_t* q_start, const int32_t* q_count,
                      const int64_t* q2o, const int64_t* o2q, const int64_t* top_ids,
                      int num_pairs, int topn, const int32_t* tpl_off, const int64_t* feat_perm,
                      const float* vertices, int max_q, int max_p, int top_k, int64_t* out_query_ids,
                      int64_t* out_vertex_ids, float* out_dists, float* out_scores,
                      float* out_coord_2d, float* out_coord_3d, int32_t* out_count, void* workspace,
                      uint64_t workspace_bytes, void* stream);
uint64_t fp_cyclic_buddies_workspace_bytes(int num_pairs, int max_q, int top_k);

/* ---- offline bank build: k-means centroid update -------------------- */
/* One update step of faiss.Kmeans.train as used by utils/cluster_util.py:37-52 (faiss Clustering.cpp
 * compute_centroids): centroids[c] = mean of the samples with assign == c, zero for empty clusters.
 * samples fp32 [n,d], assign int64 [n] (rows outside [0,k) are ignored), sums uint64 [k,d] and
 * counts int32 [k] are outputs / scratch (64-bit fixed-point sums, value * 2^24: order-independent,
 * hence bit-reproducible), centroids fp32 [k,d].  The assignment step is fp_knn_search_items. */
int fp_kmeans_update(const float* samples, const int64_t* assign, int64_t n, int d, int k, uint64_t* sums,
                     int32_t* counts, float* centroids, void* stream);

/* ---- coarse pose from the correspondences --------------------------- */
/* Replaces the per-template host loop scripts/infer.py:551-577 -> utils/pnp_util.py:42-72
 * (cv2.solvePnPRansac(flags=SOLVEPNP_ITERATIVE) + cv2.solvePnPRefineLM) for P problems at once,
 * one problem = one (crop, template) pair = the output of fp_cyclic_buddies for that pair.
 *   coord_2d fp32 [P,M,2], coord_3d fp32 [P,M,3], counts int32 [P] (valid correspondences, <= M),
 *   intrinsics fp64 [P,4] = fx, fy, cx, cy of the crop camera (utils/misc.py:325-341).
 * RANSAC: `iters` hypotheses per problem (hypothesis h of problem problem_offset + p samples 4
 * distinct correspondences from a splitmix64 stream keyed by (seed, problem, h), P3P on three,
 * fourth disambiguates), inlier <=> squared reprojection error <= thresh^2, sequential semantics
 * incl. OpenCV's RANSACUpdateNumIters(confidence) evaluated by one scan over the per-hypothesis
 * inlier counts; then Levenberg-Marquardt on the inliers (float64).  See oracle/pnp.py.
 * Outputs: success int32 [P] (>= 6 inliers), R fp64 [P,9] row-major and t fp64 [P,3] (model ->
 * camera), inlier_mask uint8 [P,M], num_inliers / iters_run / best_hyp int32 [P]. */
int fp_pnp_ransac(const float* coord_2d, const float* coord_3d, const int32_t* counts,
                  const double* intrinsics, int P, int M, int iters, double thresh, double confidence,
                  uint64_t seed, int problem_offset, int32_t* success, double* out_R, double* out_t,
                  uint8_t* inlier_mask, int32_t* num_inliers, int32_t* iters_run, int32_t* best_hyp,
                  void* stream);

/* ---- launch accounting and per-kernel timing (used by bench.py) ---------------------------- */
/* Number of kernels this library has launched since it was loaded (all threads). */
unsigned long long fp_launch_count(void);
/* Streaming multiprocessors of the current device (persistent grids are sized from it). */
int fp_num_sms(void);
/* Same, for one kernel family: 0 gemm, 1 attention, 2 layernorm, 3 vit misc, 4 knn,
 * 5 feature ops, 6 retrieval, 7 full-bank k-NN (pair kernel). */
unsigned long long fp_launch_count_category(int category);
/* When on, every launch is bracketed by CUDA events on its stream. */
int fp_profile_enable(int on);
/* Sums device time (ms), algorithmic work (flops or bytes, see csrc/common.cuh) and launches
 * recorded for a family since the last reset; synchronises on the recorded events. */
int fp_profile_read(int category, double* total_ms, double* total_work, int* launches, int reset);

#ifdef __cplusplus
}
#endif

#endif /* FOUNDPOSE_B200_H_ */
