// Crop stage in front of the feature extractor (reference scripts/infer.py:427-456, utils/misc.py:458-519).
//
// One launch warps every instance of a batch from its source image into its virtual crop camera:
//   * the destination pixel is unprojected, rotated into the source camera and projected, in fp64 with
//     the reference's operation order (structs.py:477-500), then cast to fp32 (misc.py:513);
//   * the image is sampled the way cv2.remap samples fp32 pixels with INTER_LINEAR / INTER_AREA:
//     source coordinates quantised to 1/32 pixel with round-half-even, four fp32 taps weighted by the
//     exact bilinear table and summed left to right without fused multiply-adds, constant border 0;
//     uint8 sources are scaled by 1/255 first (infer.py:396) and the result is written planar (CHW)
//     because that is what the extractor reads (infer.py:464);
//   * the modal mask is sampled like cv2.remap INTER_NEAREST;
//   * the box of the warped mask (infer.py:449-456) is reduced with atomics and finalised by a second
//     tiny kernel.
// HBM-bound streaming work: 13 B written per destination pixel, reads are gathers served by L2.
#include <climits>

#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

constexpr int kCropParamStride = 40;   // doubles per crop, see include/foundpose_b200.h

__device__ __forceinline__ int round_to_fixed(float v) {
  // cvRound: round half to even; NaN converts to INT_MIN like cvtss2si.
  return (v != v) ? INT_MIN : __float2int_rn(v);
}

__device__ __forceinline__ int saturate_short(int v) { return max(-32768, min(32767, v)); }

template <typename SrcT>
__device__ __forceinline__ float load_pixel(const SrcT* p);
template <>
__device__ __forceinline__ float load_pixel<uint8_t>(const uint8_t* p) {
  return __fdiv_rn(static_cast<float>(*p), 255.0f);
}
template <>
__device__ __forceinline__ float load_pixel<float>(const float* p) { return *p; }

template <typename SrcT>
__global__ void __launch_bounds__(256)
crop_warp_kernel(const SrcT* __restrict__ images, int num_images, int src_h, int src_w, int channels,
                 const uint8_t* __restrict__ masks, const double* __restrict__ params, int crop_w,
                 int crop_h, float* __restrict__ out_images, uint8_t* __restrict__ out_masks,
                 int* __restrict__ box_acc) {
  __shared__ double prm[kCropParamStride];
  __shared__ int s_box[4];
  const int b = blockIdx.y;
  if (threadIdx.x < kCropParamStride) prm[threadIdx.x] = params[static_cast<long>(b) * kCropParamStride + threadIdx.x];
  if (threadIdx.x < 4) s_box[threadIdx.x] = INT_MIN;
  __syncthreads();
  const int npix = crop_w * crop_h;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  int image_index = static_cast<int>(prm[32]);
  image_index = max(0, min(num_images - 1, image_index));
  const SrcT* image = images ? images + static_cast<long>(image_index) * src_h * src_w * channels : nullptr;
  const uint8_t* mask = masks ? masks + static_cast<long>(b) * src_h * src_w : nullptr;

  int mx_lo = INT_MIN, my_lo = INT_MIN, mx_hi = INT_MIN, my_hi = INT_MIN;
  if (pix < npix) {
    const int px = pix % crop_w, py = pix / crop_w;
    // window -> unit eye ray of the virtual camera
    const double qx = __ddiv_rn(__dsub_rn(static_cast<double>(px), prm[2]), prm[0]);
    const double qy = __ddiv_rn(__dsub_rn(static_cast<double>(py), prm[3]), prm[1]);
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), 1.0));
    const double v0 = __ddiv_rn(qx, nrm), v1 = __ddiv_rn(qy, nrm), v2 = __ddiv_rn(1.0, nrm);
    // eye -> world (R_dst v + t_dst), world -> source eye (R_src^T (p - t_src))
    double d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(v0, prm[4 + 3 * i]), __dmul_rn(v1, prm[5 + 3 * i])),
                                           __dmul_rn(v2, prm[6 + 3 * i])), prm[13 + i]);
      d[i] = __dsub_rn(w, prm[25 + i]);
    }
    double e[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      e[j] = __dadd_rn(__dadd_rn(__dmul_rn(d[0], prm[16 + j]), __dmul_rn(d[1], prm[19 + j])), __dmul_rn(d[2], prm[22 + j]));
    double sx = __dadd_rn(__dmul_rn(__ddiv_rn(e[0], e[2]), prm[28]), prm[30]);
    double sy = __dadd_rn(__dmul_rn(__ddiv_rn(e[1], e[2]), prm[29]), prm[31]);
    if (e[2] < 0.0) { sx = -1.0; sy = -1.0; }          // depth check (misc.py:507-510)
    const float mapx = static_cast<float>(sx), mapy = static_cast<float>(sy);

    if (image) {
      const int fxp = round_to_fixed(__fmul_rn(mapx, 32.0f)), fyp = round_to_fixed(__fmul_rn(mapy, 32.0f));
      const int ix = saturate_short(fxp >> 5), iy = saturate_short(fyp >> 5);
      const float ax = static_cast<float>(fxp & 31) * 0.03125f, ay = static_cast<float>(fyp & 31) * 0.03125f;
      const float w0 = __fmul_rn(1.0f - ay, 1.0f - ax), w1 = __fmul_rn(1.0f - ay, ax);
      const float w2 = __fmul_rn(ay, 1.0f - ax), w3 = __fmul_rn(ay, ax);
      const bool x0 = ix >= 0 && ix < src_w, x1 = ix + 1 >= 0 && ix + 1 < src_w;
      const bool y0 = iy >= 0 && iy < src_h, y1 = iy + 1 >= 0 && iy + 1 < src_h;
      const long r0 = static_cast<long>(iy) * src_w, r1 = static_cast<long>(iy + 1) * src_w;
      float* dst = out_images + (static_cast<long>(b) * channels) * npix + pix;
      for (int c = 0; c < channels; ++c) {
        const float t0 = (x0 && y0) ? load_pixel<SrcT>(image + (r0 + ix) * channels + c) : 0.0f;
        const float t1 = (x1 && y0) ? load_pixel<SrcT>(image + (r0 + ix + 1) * channels + c) : 0.0f;
        const float t2 = (x0 && y1) ? load_pixel<SrcT>(image + (r1 + ix) * channels + c) : 0.0f;
        const float t3 = (x1 && y1) ? load_pixel<SrcT>(image + (r1 + ix + 1) * channels + c) : 0.0f;
        float acc = __fmul_rn(t0, w0);
        acc = __fadd_rn(acc, __fmul_rn(t1, w1));
        acc = __fadd_rn(acc, __fmul_rn(t2, w2));
        acc = __fadd_rn(acc, __fmul_rn(t3, w3));
        dst[static_cast<long>(c) * npix] = acc;
      }
    }
    if (mask) {
      const int ix = saturate_short(round_to_fixed(mapx)), iy = saturate_short(round_to_fixed(mapy));
      uint8_t m = 0;
      if (ix >= 0 && ix < src_w && iy >= 0 && iy < src_h) m = mask[static_cast<long>(iy) * src_w + ix];
      out_masks[static_cast<long>(b) * npix + pix] = m;
      if (m) { mx_lo = -px; my_lo = -py; mx_hi = px; my_hi = py; }
    }
  }
  if (mask) {
    // box of the warped mask: max(-x), max(-y), max(x), max(y) -> warp, CTA, then one global atomic each
    mx_lo = __reduce_max_sync(0xffffffffu, mx_lo);
    my_lo = __reduce_max_sync(0xffffffffu, my_lo);
    mx_hi = __reduce_max_sync(0xffffffffu, mx_hi);
    my_hi = __reduce_max_sync(0xffffffffu, my_hi);
    if ((threadIdx.x & 31) == 0 && mx_hi != INT_MIN) {
      atomicMax(&s_box[0], mx_lo);
      atomicMax(&s_box[1], my_lo);
      atomicMax(&s_box[2], mx_hi);
      atomicMax(&s_box[3], my_hi);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_box[threadIdx.x] != INT_MIN) atomicMax(&box_acc[4 * b + threadIdx.x], s_box[threadIdx.x]);
  }
}

__global__ void crop_box_init_kernel(int* __restrict__ box_acc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) box_acc[i] = INT_MIN;
}

__global__ void crop_box_finalize_kernel(const int* __restrict__ box_acc, float* __restrict__ out_boxes, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int* a = box_acc + 4 * b;
  const bool empty = a[2] == INT_MIN;               // calc_2d_box of no points: zeros (misc.py:296-297)
  out_boxes[4 * b + 0] = empty ? 0.f : static_cast<float>(-a[0]);
  out_boxes[4 * b + 1] = empty ? 0.f : static_cast<float>(-a[1]);
  out_boxes[4 * b + 2] = empty ? 0.f : static_cast<float>(a[2]);
  out_boxes[4 * b + 3] = empty ? 0.f : static_cast<float>(a[3]);
}

}  // namespace

int crop_warp(const void* images, int src_is_f32, int num_images, int src_h, int src_w, int channels,
              const uint8_t* masks, const double* params, int B, int crop_w, int crop_h,
              float* out_images, uint8_t* out_masks, float* out_boxes, int* box_workspace,
              cudaStream_t stream) {
  FP_REQUIRE(src_h > 0 && src_w > 0 && src_h <= 32767 && src_w <= 32767, "crop_warp: source size %dx%d out of range", src_w, src_h);
  FP_REQUIRE(crop_w > 0 && crop_h > 0, "crop_warp: empty viewport");
  FP_REQUIRE(images == nullptr || (out_images != nullptr && num_images > 0 && channels > 0), "crop_warp: image output missing");
  FP_REQUIRE(masks == nullptr || (out_masks != nullptr && out_boxes != nullptr && box_workspace != nullptr),
             "crop_warp: mask outputs / workspace missing");
  if (B <= 0) return 0;
  const int npix = crop_w * crop_h;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(B) * npix * ((images ? 4.0 * channels : 0.0) + (masks ? 1.0 : 0.0)));
  if (masks) {
    crop_box_init_kernel<<<(4 * B + 255) / 256, 256, 0, stream>>>(box_workspace, 4 * B);
    FP_CUDA_CHECK(cudaGetLastError());
  }
  const dim3 grid((npix + 255) / 256, B);
  if (src_is_f32)
    crop_warp_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(images), num_images, src_h, src_w,
                                                      channels, masks, params, crop_w, crop_h, out_images,
                                                      out_masks, box_workspace);
  else
    crop_warp_kernel<uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(images), num_images, src_h,
                                                        src_w, channels, masks, params, crop_w, crop_h,
                                                        out_images, out_masks, box_workspace);
  FP_CUDA_CHECK(cudaGetLastError());
  if (masks) {
    crop_box_finalize_kernel<<<(B + 127) / 128, 128, 0, stream>>>(box_workspace, out_boxes, B);
    FP_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

}  // namespace fp
