// Memory-bound helper kernels of the ViT stage (everything that is not a contraction).
//
//   patchify_normalize : images fp32 [B,3,H,W] in [0,1]  -> fp16 im2col rows [B*P, Kpad]
//                        with the ImageNet normalisation of utils/dinov2_utils.py:111-113,123 fused
//                        (PatchEmbed.proj is a 14x14/14 conv = GEMM over these rows,
//                        external/dinov2/dinov2/layers/patch_embed.py:75)
//   init_special_tokens: cls (+pos[0]) and register rows of the token stream
//                        (models/vision_transformer.py:219-230)
//   layernorm_f16      : LayerNorm(eps=1e-6) fp32 residual stream -> fp16 GEMM operand
//                        (layers/block.py:63,75)
//   final_norm_tokens  : drop cls/register tokens, final LayerNorm over [cls | patches]
//                        (utils/dinov2_utils.py:126-153), fp32 + optional fp16 copy
//
// All of them are HBM-bound streaming kernels: one warp per row, 128-bit accesses, grid sized
// from the row count.
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One thread per PAIR of horizontally adjacent pixels: consecutive lanes read consecutive float2 of an image row
// (fully coalesced) and write one half2 of the im2col row of the patch the pair falls into (14 is even, so a pair
// never straddles two patches; 28-byte runs per patch row on the write side).  Threads (b, 0, py * ps, *) also zero
// the padding columns [3 ps^2, Kpad) of their patch.  The first two versions (one thread per 14-pixel run, patch row
// or patch column fastest) issued 32-sector loads and ran at 1.0 / 0.55 TB/s under ncu.
__global__ void patchify_normalize_kernel(const float* __restrict__ img, __half* __restrict__ out,
                                          int B, int H, int W, int ps, int Kpad) {
  const int Hp = H / ps, Wp = W / ps;
  const int W2 = W / 2;
  const long total = static_cast<long>(B) * 3 * H * W2;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W2) * 2;
    const long row = idx / W2;                       // (b, c, y)
    const int y = static_cast<int>(row % H);
    const int c = static_cast<int>((row / H) % 3);
    const int b = static_cast<int>(row / (3L * H));
    const int py = y / ps, dy = y - py * ps;
    const int px = x / ps, dx = x - px * ps;
    const float2 v = *reinterpret_cast<const float2*>(img + row * W + x);
    const long patch = (static_cast<long>(b) * Hp + py) * Wp + px;
    // (x - mean) / std exactly as torchvision's normalize (sub then div).
    const __half2 h = __floats2half2_rn((v.x - mean[c]) / stdv[c], (v.y - mean[c]) / stdv[c]);
    *reinterpret_cast<__half2*>(out + patch * Kpad + c * ps * ps + dy * ps + dx) = h;
    if (c == 0 && dy == 0) {
      const int pad0 = 3 * ps * ps;
      for (int k = pad0 + dx; k < Kpad; k += ps) {          // the 7 pairs of the row share the padding columns
        out[patch * Kpad + k] = __float2half_rn(0.f);
        if (k + 1 < Kpad) out[patch * Kpad + k + 1] = __float2half_rn(0.f);
      }
    }
  }
}

// LayerNorm-fused mode: the cls / register rows also need their fp16 copy and (sum x, sum x^2): one warp per row,
// the full-row sums go to slot 0, the other slots are zero.
__global__ void special_tokens_ln_kernel(const float* __restrict__ x, __half* __restrict__ x16,
                                         float* __restrict__ stats, int B, int ntok, int R, int D, int slots) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int total = B * (1 + R);
  for (int o = blockIdx.x * wpb + (threadIdx.x >> 5); o < total; o += gridDim.x * wpb) {
    const long row = static_cast<long>(o / (1 + R)) * ntok + (o % (1 + R));
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float v = x[row * D + i];
      x16[row * D + i] = __float2half_rn(v);
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    for (int t = lane; t < slots; t += 32) {
      stats[(row * slots + t) * 2] = t == 0 ? s1 : 0.f;
      stats[(row * slots + t) * 2 + 1] = t == 0 ? s2 : 0.f;
    }
  }
}

// x[b, 0, :] = cls_pos ; x[b, 1..R, :] = reg[r]
__global__ void init_special_tokens_kernel(float* __restrict__ x, const float* __restrict__ cls_pos,
                                           const float* __restrict__ reg, int B, int ntok, int R,
                                           int D) {
  const long total = static_cast<long>(B) * (1 + R) * D;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(idx % D);
    const int t = static_cast<int>((idx / D) % (1 + R));
    const int b = static_cast<int>(idx / (static_cast<long>(D) * (1 + R)));
    const float v = (t == 0) ? cls_pos[d] : reg[static_cast<long>(t - 1) * D + d];
    x[(static_cast<long>(b) * ntok + t) * D + d] = v;
  }
}

// LayerNorm over D (multiple of 128, <= 2048): one warp per row, row cached in registers.
// MAX_VEC float4 per lane; with kExact the row has exactly MAX_VEC * 128 elements (no predicates,
// registers sized to the row), otherwise D / 128 <= MAX_VEC vectors are used.
template <int MAX_VEC, bool kExact = false>
__device__ __forceinline__ void ln_row(const float* __restrict__ xr, int D, const float* __restrict__ w,
                                       const float* __restrict__ bvec, float eps, int lane,
                                       float4 (&v)[MAX_VEC], int& nvec) {
  nvec = kExact ? MAX_VEC : D / 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    if (i < nvec) {
      v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    if (i < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    if (i < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(w + (i * 32 + lane) * 4));
      const float4 be = __ldg(reinterpret_cast<const float4*>(bvec + (i * 32 + lane) * 4));
      v[i].x = (v[i].x - mean) * rstd * g.x + be.x;
      v[i].y = (v[i].y - mean) * rstd * g.y + be.y;
      v[i].z = (v[i].z - mean) * rstd * g.z + be.z;
      v[i].w = (v[i].w - mean) * rstd * g.w + be.w;
    }
  }
}

template <int NVEC, bool kExact>
__global__ void __launch_bounds__(256)
layernorm_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, const float* __restrict__ w,
                     const float* __restrict__ b, int M, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long row = blockIdx.x * static_cast<long>(warps_per_block) + (threadIdx.x >> 5); row < M;
       row += static_cast<long>(gridDim.x) * warps_per_block) {
    float4 v[NVEC];
    int nvec;
    ln_row<NVEC, kExact>(x + row * D, D, w, b, eps, lane, v, nvec);
    __half* yr = y + row * D;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      if (i < nvec) {
        __half2 h0 = __floats2half2_rn(v[i].x, v[i].y);
        __half2 h1 = __floats2half2_rn(v[i].z, v[i].w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(yr + (i * 32 + lane) * 4) = pk;
      }
    }
  }
}

// Output row j of image b: j == 0 -> cls token (row b*ntok), j >= 1 -> patch token
// (row b*ntok + 1 + R + j - 1). Optional LayerNorm. Writes cls to out_cls[b], patches to
// out_tok[b*P + j-1] (fp32) and out_tok16 (fp16, optional).
__global__ void __launch_bounds__(256)
final_norm_tokens_kernel(const float* __restrict__ x, const float* __restrict__ w,
                         const float* __restrict__ bvec, float* __restrict__ out_tok,
                         __half* __restrict__ out_tok16, float* __restrict__ out_cls, int B,
                         int ntok, int R, int P, int D, int apply_norm, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long total = static_cast<long>(B) * (P + 1);
  for (long o = blockIdx.x * static_cast<long>(warps_per_block) + (threadIdx.x >> 5); o < total;
       o += static_cast<long>(gridDim.x) * warps_per_block) {
    const int b = static_cast<int>(o / (P + 1));
    const int j = static_cast<int>(o - static_cast<long>(b) * (P + 1));
    const long src_row = static_cast<long>(b) * ntok + (j == 0 ? 0 : R + j);
    float4 v[16];
    int nvec = D / 128;
    if (apply_norm) {
      ln_row<16>(x + src_row * D, D, w, bvec, eps, lane, v, nvec);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nvec) v[i] = *reinterpret_cast<const float4*>(x + src_row * D + (i * 32 + lane) * 4);
    }
    float* dst = (j == 0) ? (out_cls != nullptr ? out_cls + static_cast<long>(b) * D : nullptr)
                          : out_tok + (static_cast<long>(b) * P + (j - 1)) * D;
    if (dst != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nvec) *reinterpret_cast<float4*>(dst + (i * 32 + lane) * 4) = v[i];
    }
    if (j > 0 && out_tok16 != nullptr) {
      __half* d16 = out_tok16 + (static_cast<long>(b) * P + (j - 1)) * D;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < nvec) {
          __half2 h0 = __floats2half2_rn(v[i].x, v[i].y);
          __half2 h1 = __floats2half2_rn(v[i].z, v[i].w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(d16 + (i * 32 + lane) * 4) = pk;
        }
      }
    }
  }
}

inline int grid_for(long work_items, int per_block, int max_blocks = num_sms() * 8) {
  long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  return static_cast<int>(g > max_blocks ? max_blocks : g);
}

}  // namespace

int patchify_normalize(const float* img, __half* out, int B, int H, int W, int ps, int Kpad,
                       cudaStream_t stream) {
  FP_REQUIRE(H % ps == 0, "Input image height %d is not a multiple of patch height %d", H, ps);
  FP_REQUIRE(W % ps == 0, "Input image width %d is not a multiple of patch width: %d", W, ps);
  FP_REQUIRE(Kpad >= 3 * ps * ps, "patchify: Kpad too small");
  FP_REQUIRE(ps % 2 == 0 && W % 2 == 0 && (reinterpret_cast<uintptr_t>(img) & 7) == 0,
             "patchify: needs an even patch size / image width and an 8-byte aligned image");
  const long total = static_cast<long>(B) * 3 * H * (W / 2);
  ProfScope prof(PROF_VIT_MISC, stream, static_cast<double>(B) * 3 * H * W * 4 + static_cast<double>(B) * (H / ps) * (W / ps) * Kpad * 2);
  patchify_normalize_kernel<<<grid_for(total, 256), 256, 0, stream>>>(img, out, B, H, W, ps, Kpad);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int init_special_tokens(float* x, const float* cls_pos, const float* reg, int B, int ntok, int R,
                        int D, cudaStream_t stream, __half* x16, float* stats, int stat_slots) {
  const long total = static_cast<long>(B) * (1 + R) * D;
  {
    ProfScope prof(PROF_VIT_MISC, stream, static_cast<double>(total) * 4);
    init_special_tokens_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, cls_pos, reg, B, ntok, R, D);
    FP_CUDA_CHECK(cudaGetLastError());
  }
  if (x16 != nullptr) {
    ProfScope prof(PROF_VIT_MISC, stream, static_cast<double>(total) * 6);
    special_tokens_ln_kernel<<<grid_for(static_cast<long>(B) * (1 + R), 8), 256, 0, stream>>>(
        x, x16, stats, B, ntok, R, D, stat_slots);
    FP_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

int layernorm_f16(const float* x, __half* y, const float* w, const float* b, int M, int D, float eps,
                  cudaStream_t stream) {
  FP_REQUIRE(D % 128 == 0 && D <= 2048, "layernorm: D=%d must be a multiple of 128 and <= 2048", D);
  ProfScope prof(PROF_LAYERNORM, stream, static_cast<double>(M) * D * 6);
  // Persistent grid (a one-block-per-8-rows launch was measured 20% slower on B200).
  const int grid = grid_for(M, 8);
  switch (D / 128) {   // registers sized to the row for the ViT widths
    case 3: layernorm_f16_kernel<3, true><<<grid, 256, 0, stream>>>(x, y, w, b, M, D, eps); break;
    case 6: layernorm_f16_kernel<6, true><<<grid, 256, 0, stream>>>(x, y, w, b, M, D, eps); break;
    case 8: layernorm_f16_kernel<8, true><<<grid, 256, 0, stream>>>(x, y, w, b, M, D, eps); break;
    default: layernorm_f16_kernel<16, false><<<grid, 256, 0, stream>>>(x, y, w, b, M, D, eps); break;
  }
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int final_norm_tokens(const float* x, const float* w, const float* b, float* out_tok,
                      __half* out_tok16, float* out_cls, int B, int ntok, int R, int P, int D,
                      int apply_norm, float eps, cudaStream_t stream) {
  FP_REQUIRE(D % 128 == 0 && D <= 2048, "final_norm: D=%d must be a multiple of 128 and <= 2048", D);
  const long total = static_cast<long>(B) * (P + 1);
  ProfScope prof(PROF_LAYERNORM, stream, static_cast<double>(total) * D * (out_tok16 ? 10 : 8));
  final_norm_tokens_kernel<<<grid_for(total, 8), 256, 0, stream>>>(x, w, b, out_tok, out_tok16,
                                                                  out_cls, B, ntok, R, P, D,
                                                                  apply_norm, eps);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
