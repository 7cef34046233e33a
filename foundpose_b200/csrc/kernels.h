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
};

int gemm_pick_bn(int M, int N);
int gemm_tn(int epi, const __half* A, int lda, const __half* B, int ldb, const GemmParams& p,
            cudaStream_t stream);
int umma_probe(const __half* A, const __half* B, float* out, int b_mn_major, cudaStream_t stream);

}  // namespace fp
