// ViT feature-extractor handle: owns the activation workspace and sequences the kernels of one
// batched forward pass up to block `layer` (early exit: the reference runs the remaining blocks
// and discards them; skipping them is output-identical).
//
// Reference call chain reproduced: utils/dinov2_utils.py:115-158 (forward) ->
// external/dinov2/dinov2/models/vision_transformer.py:213-232 (prepare tokens) ->
// layers/block.py:89-114 (blocks) -> utils/dinov2_utils.py:138-153 (final LayerNorm, reshape).
#include "../../include/foundpose_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <new>

namespace fp {

struct VitHandle {
  fp_vit_config cfg;
  fp_vit_weights w;
  fp_vit_block_weights* blocks = nullptr;  // host copy of the per-block pointer table
  int max_batch = 0;
  int P = 0;      // patches per image
  int ntok = 0;   // 1 + registers + P
  int Kpad = 0;   // padded patch vector length (multiple of 64)
  // workspace (device)
  __half* patches = nullptr;  // [max_batch*P, Kpad]
  float* x = nullptr;         // [max_batch*ntok, D]     fp32 residual stream
  __half* xn = nullptr;       // [max_batch*ntok, D]     LayerNorm output (GEMM operand)
  __half* qkv = nullptr;      // [max_batch*ntok, 3D]
  __half* attn = nullptr;     // [max_batch*ntok, D]
  __half* hidden = nullptr;   // [max_batch*ntok, 4D]
  float* facet_x = nullptr;   // [max_batch*ntok, D]     only for key/query/value facets
  // LayerNorm-fused mode (cfg.fuse_layernorm): fp16 copy of the residual stream (the GEMM operand instead of xn)
  // and per-row partial (sum, sum of squares) per 128 columns, both written by the producing epilogues.
  bool fused_ln = false;
  float* stats = nullptr;     // [max_batch*ntok, D/128, 2]
};

namespace {

// out[row, d*H + h] = qkv[row, which*D + h*64 + d]  (utils/dinov2_utils.py:296-309: the
// per-head tensor B x h x t x d is permuted to B x t x d x h and flattened).
__global__ void facet_gather_kernel(const __half* __restrict__ qkv, float* __restrict__ out, long M,
                                    int D, int heads, int which) {
  const long total = M * D;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = idx / D;
    const int c = static_cast<int>(idx - row * D);
    const int d = c / heads, h = c - d * heads;
    out[idx] = __half2float(qkv[row * 3 * D + static_cast<long>(which) * D + h * 64 + d]);
  }
}

}  // namespace

}  // namespace fp

using fp::VitHandle;

extern "C" {

int fp_vit_create(const fp_vit_config* cfg, const fp_vit_weights* weights,
                  const fp_vit_block_weights* blocks, int max_batch, fp_vit** out) {
  FP_REQUIRE(cfg != nullptr && weights != nullptr && blocks != nullptr && out != nullptr,
             "fp_vit_create: null argument");
  FP_REQUIRE(cfg->embed_dim % 128 == 0 && cfg->embed_dim == cfg->num_heads * 64,
             "fp_vit_create: embed_dim=%d must be num_heads*64 and a multiple of 128", cfg->embed_dim);
  FP_REQUIRE(cfg->img_h % cfg->patch_size == 0 && cfg->img_w % cfg->patch_size == 0,
             "Input image size %dx%d is not a multiple of patch size %d", cfg->img_h, cfg->img_w,
             cfg->patch_size);
  FP_REQUIRE(cfg->num_blocks >= 1 && max_batch >= 1, "fp_vit_create: bad num_blocks/max_batch");
  FP_REQUIRE(!cfg->fuse_layernorm || cfg->embed_dim % 256 == 0,
             "fp_vit_create: fuse_layernorm needs embed_dim %% 256 == 0 (got %d)", cfg->embed_dim);
  if (cfg->fuse_layernorm) {
    for (int i = 0; i < cfg->num_blocks; ++i)
      FP_REQUIRE(blocks[i].qkv_colsum != nullptr && blocks[i].fc1_colsum != nullptr,
                 "fp_vit_create: fuse_layernorm needs qkv_colsum / fc1_colsum of block %d", i);
  }
  VitHandle* h = new (std::nothrow) VitHandle();
  FP_REQUIRE(h != nullptr, "fp_vit_create: out of host memory");
  h->cfg = *cfg;
  h->w = *weights;
  h->blocks = new fp_vit_block_weights[cfg->num_blocks];
  for (int i = 0; i < cfg->num_blocks; ++i) h->blocks[i] = blocks[i];
  h->max_batch = max_batch;
  h->fused_ln = cfg->fuse_layernorm != 0;
  h->P = (cfg->img_h / cfg->patch_size) * (cfg->img_w / cfg->patch_size);
  h->ntok = 1 + cfg->num_register_tokens + h->P;
  const int k = 3 * cfg->patch_size * cfg->patch_size;
  h->Kpad = (k + 63) / 64 * 64;
  const size_t D = cfg->embed_dim;
  const size_t rows = static_cast<size_t>(max_batch) * h->ntok;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  alloc(reinterpret_cast<void**>(&h->patches), static_cast<size_t>(max_batch) * h->P * h->Kpad * 2);
  alloc(reinterpret_cast<void**>(&h->x), rows * D * 4);
  alloc(reinterpret_cast<void**>(&h->xn), rows * D * 2);
  alloc(reinterpret_cast<void**>(&h->qkv), rows * 3 * D * 2);
  alloc(reinterpret_cast<void**>(&h->attn), rows * D * 2);
  alloc(reinterpret_cast<void**>(&h->hidden), rows * 4 * D * 2);
  if (h->fused_ln) alloc(reinterpret_cast<void**>(&h->stats), rows * (D / 128) * 2 * 4);
  if (e != cudaSuccess) {
    fp::set_last_error("fp_vit_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    fp_vit_destroy(reinterpret_cast<fp_vit*>(h));
    return 2;
  }
  *out = reinterpret_cast<fp_vit*>(h);
  return 0;
}

void fp_vit_destroy(fp_vit* handle) {
  VitHandle* h = reinterpret_cast<VitHandle*>(handle);
  if (h == nullptr) return;
  cudaFree(h->patches);
  cudaFree(h->x);
  cudaFree(h->xn);
  cudaFree(h->qkv);
  cudaFree(h->attn);
  cudaFree(h->hidden);
  cudaFree(h->facet_x);
  cudaFree(h->stats);
  delete[] h->blocks;
  delete h;
}

int fp_vit_patch_k(const fp_vit* handle) {
  return handle ? reinterpret_cast<const VitHandle*>(handle)->Kpad : -1;
}

int fp_vit_forward(fp_vit* handle, const float* images, int batch, int layer, int facet,
                   int apply_norm, float* out_tokens, void* out_tokens_f16, float* out_cls,
                   void* stream_) {
  VitHandle* h = reinterpret_cast<VitHandle*>(handle);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FP_REQUIRE(h != nullptr && images != nullptr && out_tokens != nullptr, "fp_vit_forward: null argument");
  FP_REQUIRE(batch >= 1 && batch <= h->max_batch, "fp_vit_forward: batch=%d exceeds max_batch=%d",
             batch, h->max_batch);
  FP_REQUIRE(layer >= 0 && layer < h->cfg.num_blocks, "fp_vit_forward: layer=%d out of range [0,%d)",
             layer, h->cfg.num_blocks);
  FP_REQUIRE(facet >= 0 && facet <= 3, "fp_vit_forward: facet=%d is not a supported facet", facet);
  const int D = h->cfg.embed_dim;
  const int heads = h->cfg.num_heads;
  const int R = h->cfg.num_register_tokens;
  const int M = batch * h->ntok;
  const float eps = 1e-6f;
  int rc;
  bool fused = false;   // LayerNorm folded into the GEMMs (h->xn then holds the RAW fp16 residual rows)

  // Token preparation: patch embedding (+bias +pos) straight into the token stream, cls/registers.
  if ((rc = fp::patchify_normalize(images, h->patches, batch, h->cfg.img_h, h->cfg.img_w,
                                   h->cfg.patch_size, h->Kpad, stream)) != 0) return rc;
  {
    fp::GemmParams p;
    p.M = batch * h->P; p.N = D; p.K = h->Kpad;
    p.bias = h->w.patch_b;
    p.out_f32 = h->x; p.ld_f32 = D;
    p.patches_per_img = h->P; p.tokens_per_img = h->ntok; p.tok_off = 1 + R;
    p.pos = h->w.pos_patch;
    // The pair kernel (the only one with the LayerNorm-fused epilogues) needs more than 256 rows.
    fused = h->fused_ln;
    FP_REQUIRE(!fused || batch * h->P > 256,
               "fp_vit_forward: a handle created with fuse_layernorm needs more than 256 patch rows per call");
    if (fused) {
      p.x16 = h->xn; p.ld_x16 = D; p.stats_out = h->stats;
    }
    if ((rc = fp::gemm_tn(fused ? fp::EPI_PATCH_LN_F32 : fp::EPI_PATCH_F32, h->patches, h->Kpad,
                          static_cast<const __half*>(h->w.patch_w), h->Kpad, p, stream)) != 0) return rc;
  }
  if ((rc = fp::init_special_tokens(h->x, h->w.cls_pos, h->w.reg_tokens, batch, h->ntok, R, D, stream,
                                    fused ? h->xn : nullptr, h->stats, D / 128)) != 0) return rc;

  const float* final_src = h->x;
  for (int i = 0; i <= layer; ++i) {
    const fp_vit_block_weights& bw = h->blocks[i];
    // x = x + ls1 * proj(attn(norm1(x)))
    if (!fused) {
      if ((rc = fp::layernorm_f16(h->x, h->xn, bw.norm1_w, bw.norm1_b, M, D, eps, stream)) != 0) return rc;
    }
    {
      fp::GemmParams p;
      p.M = M; p.N = 3 * D; p.K = D;
      p.bias = bw.qkv_b; p.out_f16 = h->qkv; p.ld_f16 = 3 * D;
      if (fused) {
        p.ln_stats = h->stats; p.ln_slots = D / 128; p.ln_colsum = bw.qkv_colsum; p.ln_dim = D; p.ln_eps = eps;
      }
      if ((rc = fp::gemm_tn(fused ? fp::EPI_LN_BIAS_F16 : fp::EPI_BIAS_F16, h->xn, D,
                            static_cast<const __half*>(bw.qkv_w), D, p, stream)) != 0) return rc;
    }
    if (i == layer && facet != 0) {
      // key/query/value facet: the hook recomputes qkv from the attention input of this block.
      if (h->facet_x == nullptr) {
        FP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->facet_x),
                                 static_cast<size_t>(h->max_batch) * h->ntok * D * 4));
      }
      const int which = facet - 1;  // 1 = query, 2 = key, 3 = value
      fp::ProfScope prof(fp::PROF_VIT_MISC, stream, static_cast<double>(M) * D * 6);
      fp::facet_gather_kernel<<<fp::num_sms() * 4, 256, 0, stream>>>(h->qkv, h->facet_x, M, D, heads, which);
      FP_CUDA_CHECK(cudaGetLastError());
      final_src = h->facet_x;
      break;
    }
    if ((rc = fp::attention_f16(h->qkv, h->attn, batch, h->ntok, heads, stream)) != 0) return rc;
    {
      fp::GemmParams p;
      p.M = M; p.N = D; p.K = D;
      p.bias = bw.proj_b; p.gamma = bw.ls1; p.out_f32 = h->x; p.ld_f32 = D;
      if (fused) {
        p.x16 = h->xn; p.ld_x16 = D; p.stats_out = h->stats;
      }
      if ((rc = fp::gemm_tn(fused ? fp::EPI_RESID_LN_F32 : fp::EPI_RESID_F32, h->attn, D,
                            static_cast<const __half*>(bw.proj_w), D, p, stream)) != 0) return rc;
    }
    // x = x + ls2 * fc2(gelu(fc1(norm2(x))))
    if (!fused) {
      if ((rc = fp::layernorm_f16(h->x, h->xn, bw.norm2_w, bw.norm2_b, M, D, eps, stream)) != 0) return rc;
    }
    {
      fp::GemmParams p;
      p.M = M; p.N = 4 * D; p.K = D;
      p.bias = bw.fc1_b; p.out_f16 = h->hidden; p.ld_f16 = 4 * D;
      if (fused) {
        p.ln_stats = h->stats; p.ln_slots = D / 128; p.ln_colsum = bw.fc1_colsum; p.ln_dim = D; p.ln_eps = eps;
      }
      if ((rc = fp::gemm_tn(fused ? fp::EPI_LN_BIAS_GELU_F16 : fp::EPI_BIAS_GELU_F16, h->xn, D,
                            static_cast<const __half*>(bw.fc1_w), D, p, stream)) != 0) return rc;
    }
    {
      fp::GemmParams p;
      p.M = M; p.N = D; p.K = 4 * D;
      p.bias = bw.fc2_b; p.gamma = bw.ls2; p.out_f32 = h->x; p.ld_f32 = D;
      const bool last = i == layer;   // nothing reads the fp16 copy / sums of the last block's output
      if (fused && !last) {
        p.x16 = h->xn; p.ld_x16 = D; p.stats_out = h->stats;
      }
      if ((rc = fp::gemm_tn((fused && !last) ? fp::EPI_RESID_LN_F32 : fp::EPI_RESID_F32, h->hidden, 4 * D,
                            static_cast<const __half*>(bw.fc2_w), 4 * D, p, stream)) != 0) return rc;
    }
  }
  return fp::final_norm_tokens(final_src, h->w.norm_w, h->w.norm_b, out_tokens,
                               static_cast<__half*>(out_tokens_f16), out_cls, batch, h->ntok, R, h->P,
                               D, apply_norm, eps, stream);
}

}  // extern "C"
