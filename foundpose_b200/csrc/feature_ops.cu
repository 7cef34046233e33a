// Query-point selection and feature sampling (reference utils/feature_util.py:55-131).
//
//   filter_points_by_mask : ordered compaction of the grid points that fall inside the object mask
//                           (feature_util.py:75-97), batched: one CTA per crop, fixed output stride
//   sample_features       : bilinear `grid_sample` (align_corners=False, zero padding) of a
//                           patch-token map at 2D image points (feature_util.py:100-131), reading the
//                           token-major [Hp*Wp, C] layout the ViT kernel writes (no CHW transpose)
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

__global__ void __launch_bounds__(256)
filter_points_kernel(const float* __restrict__ points, int num_points,
                     const uint8_t* __restrict__ masks, int H, int W, float* __restrict__ out_points,
                     int* __restrict__ out_ids, int* __restrict__ out_counts, int out_stride) {
  __shared__ int warp_counts[8];
  __shared__ int base;
  const int b = blockIdx.x;
  const uint8_t* mask = masks + static_cast<long>(b) * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < num_points; start += blockDim.x) {
    const int i = start + threadIdx.x;
    bool keep = false;
    float x = 0.f, y = 0.f;
    if (i < num_points) {
      x = points[2 * i];
      y = points[2 * i + 1];
      // (points + 0.5).int(): float add then truncation toward zero.
      const int xi = static_cast<int>(x + 0.5f);
      const int yi = static_cast<int>(y + 0.5f);
      // filter_points_by_box with strict inequalities on the integer coordinates.
      if (xi > 0 && xi < W && yi > 0 && yi < H) keep = mask[static_cast<long>(yi) * W + xi] != 0;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_counts[warp] = __popc(ballot);
    __syncthreads();
    int offset = base;
    for (int w = 0; w < warp; ++w) offset += warp_counts[w];
    if (keep) {
      const int pos = offset + __popc(ballot & ((1u << lane) - 1));
      out_points[(static_cast<long>(b) * out_stride + pos) * 2] = x;
      out_points[(static_cast<long>(b) * out_stride + pos) * 2 + 1] = y;
      out_ids[static_cast<long>(b) * out_stride + pos] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_counts[w];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[b] = base;
}

// One warp per (crop, point). tokens: [B, Hp*Wp, C] fp32. Rows >= counts[b] are zero-filled so
// that fixed-stride buffers are fully defined.
__global__ void __launch_bounds__(256)
sample_features_kernel(const float* __restrict__ tokens, int Hp, int Wp, int C,
                       const float* __restrict__ points, const int* __restrict__ counts,
                       int stride, int B, float img_w, float img_h, float* __restrict__ out_f32,
                       __half* __restrict__ out_f16) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long total = static_cast<long>(B) * stride;
  for (long o = blockIdx.x * static_cast<long>(wpb) + (threadIdx.x >> 5); o < total;
       o += static_cast<long>(gridDim.x) * wpb) {
    const int b = static_cast<int>(o / stride);
    const int i = static_cast<int>(o - static_cast<long>(b) * stride);
    const int n = counts ? counts[b] : stride;
    float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
    int x0 = 0, y0 = 0;
    const bool valid = i < n;
    if (valid) {
      const float px = points[o * 2], py = points[o * 2 + 1];
      // uv = 2/size * p - 1  (feature_util.py:116); unnormalise with align_corners=False.
      const float u = (2.0f / img_w) * px - 1.0f;
      const float v = (2.0f / img_h) * py - 1.0f;
      const float ix = ((u + 1.0f) * Wp - 1.0f) / 2.0f;
      const float iy = ((v + 1.0f) * Hp - 1.0f) / 2.0f;
      const float fx = floorf(ix), fy = floorf(iy);
      x0 = static_cast<int>(fx);
      y0 = static_cast<int>(fy);
      // Same weight expressions as ATen's grid_sampler_2d: (ix_se - ix) * (iy_se - iy), ...
      const float ex = fx + 1.0f - ix, ey = fy + 1.0f - iy;   // distance to the se corner
      const float tx = ix - fx, ty = iy - fy;                 // distance to the nw corner
      w00 = ex * ey;  // nw
      w01 = tx * ey;  // ne
      w10 = ex * ty;  // sw
      w11 = tx * ty;  // se
    }
    const bool in_x0 = x0 >= 0 && x0 < Wp, in_x1 = x0 + 1 >= 0 && x0 + 1 < Wp;
    const bool in_y0 = y0 >= 0 && y0 < Hp, in_y1 = y0 + 1 >= 0 && y0 + 1 < Hp;
    const float* base = tokens + static_cast<long>(b) * Hp * Wp * C;
    const float* p00 = base + (static_cast<long>(y0) * Wp + x0) * C;
    const float* p01 = p00 + C;
    const float* p10 = p00 + static_cast<long>(Wp) * C;
    const float* p11 = p10 + C;
    for (int c = lane * 4; c < C; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        if (in_x0 && in_y0) {
          const float4 t = *reinterpret_cast<const float4*>(p00 + c);
          acc.x += t.x * w00; acc.y += t.y * w00; acc.z += t.z * w00; acc.w += t.w * w00;
        }
        if (in_x1 && in_y0) {
          const float4 t = *reinterpret_cast<const float4*>(p01 + c);
          acc.x += t.x * w01; acc.y += t.y * w01; acc.z += t.z * w01; acc.w += t.w * w01;
        }
        if (in_x0 && in_y1) {
          const float4 t = *reinterpret_cast<const float4*>(p10 + c);
          acc.x += t.x * w10; acc.y += t.y * w10; acc.z += t.z * w10; acc.w += t.w * w10;
        }
        if (in_x1 && in_y1) {
          const float4 t = *reinterpret_cast<const float4*>(p11 + c);
          acc.x += t.x * w11; acc.y += t.y * w11; acc.z += t.z * w11; acc.w += t.w * w11;
        }
      }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o * C + c) = acc;
      if (out_f16) {
        __half2 h0 = __floats2half2_rn(acc.x, acc.y);
        __half2 h1 = __floats2half2_rn(acc.z, acc.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(out_f16 + o * C + c) = pk;
      }
    }
  }
}

}  // namespace

int filter_points_by_mask(const float* points, int num_points, const uint8_t* masks, int B, int H,
                          int W, float* out_points, int* out_ids, int* out_counts, int out_stride,
                          cudaStream_t stream) {
  FP_REQUIRE(out_stride >= num_points, "filter_points_by_mask: out_stride < num_points");
  if (B <= 0) return 0;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(B) * num_points * 13);
  filter_points_kernel<<<B, 256, 0, stream>>>(points, num_points, masks, H, W, out_points, out_ids,
                                              out_counts, out_stride);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sample_features(const float* tokens, int B, int Hp, int Wp, int C, const float* points,
                    const int* counts, int stride, float img_w, float img_h, float* out_f32,
                    __half* out_f16, cudaStream_t stream) {
  FP_REQUIRE(C % 4 == 0, "sample_features: channel count %d must be a multiple of 4", C);
  const long total = static_cast<long>(B) * stride;
  if (total <= 0) return 0;
  long blocks = (total + 7) / 8;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(total) * C * (4 + (out_f32 ? 4 : 0) + (out_f16 ? 2 : 0)));
  sample_features_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      tokens, Hp, Wp, C, points, counts, stride, B, img_w, img_h, out_f32, out_f16);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
