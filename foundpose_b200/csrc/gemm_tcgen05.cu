// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[M, N] = A[M, K] . B[N, K]^T          (A, B fp16 row-major "K-major"; fp32 accumulate)
//
// This one kernel carries every dense contraction on the FoundPose hot path:
//   * ViT patch embedding   (PatchEmbed.proj as a GEMM over 14x14x3 patches,
//                            reference external/dinov2/dinov2/layers/patch_embed.py:68-81)
//   * qkv / proj / fc1 / fc2 (external/dinov2/dinov2/layers/attention.py:58,67; mlp.py:35-38)
//     with bias, exact-erf GELU, and LayerScale*residual fused into the epilogue
//     (layers/block.py:111-113, layers/layer_scale.py:27)
//   * PCA projection         (utils/projector_util.py:66-69: X @ C^T - mean @ C^T)
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      : MMA issuer    (tcgen05.mma, 128 x BN x 16 per instruction, accum in TMEM)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue      (tcgen05.ld -> registers -> fused math -> global)
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
// overlaps the main loop of tile i+1.
#include "common.cuh"
#include "kernels.h"

#include <stdarg.h>

namespace fp {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 bytes = one swizzle-128B row
constexpr int kGemmThreads = 384;
constexpr int kEpilogueWarps = 8;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr uint32_t kABytes = BM * BK * 2;
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BN;  // 512 or 256: powers of two
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + 1024 /*align*/;
};

// Exact (erf) GELU as torch.nn.GELU() computes it (external/dinov2/dinov2/layers/mlp.py:36).
__device__ __forceinline__ float gelu_erf_libm(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

// Same function through erfc(z) = 2^P(z), z = |x| / sqrt(2), P a degree-5 polynomial without constant term
// fitted (minimax in erfc) on [0, 4.3]: max abs error 6e-7 in erf - the class of Abramowitz-Stegun 7.1.26 and
// far below the fp16 rounding of the stored activation - with ONE MUFU (ex2) and 10 FP32 ops per element
// (A-S 7.1.26 needs rcp + ex2 and 13; erff() ~50 with branches).  The leading coefficient is negative, so
// 2^P -> 0 for large |x| without a clamp.  It matters because the epilogue of a 128x256 tile shares its SM
// sub-partitions with nothing else but must finish within the 8192 cycles the tile's MMAs take: at 2 MUFU +
// 13 FMA per element the fc1 kernel ran at 72% tensor-pipe activity against 85% for the qkv kernel
// (profiles/r01_final_summary.md).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(-0.002944134408608079f, z, 0.029589958488941193f);
  p = fmaf(p, z, -0.14866553246974945f);
  p = fmaf(p, z, -0.9185094237327576f);
  p = fmaf(p, z, -1.6278890371322632f);
  p *= z;
  float e;   // erfc(z)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  const float half_x = 0.5f * x;
  return fmaf(fabsf(half_x), 1.0f - e, half_x);   // 0.5 x (1 + sign(x) erf(z))
}

template <int EPI>
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, int row, int n0,
                                                 const uint32_t (&r)[32], float& s1, float& s2) {
  // One thread owns 32 consecutive output columns [n0, n0+32) of one row.
  if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16) {
    __half* dst = p.out_f16 + static_cast<size_t>(row) * p.ld_f16 + n0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j * 8 + i]);
      if (p.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 8 + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if constexpr (EPI == EPI_BIAS_GELU_F16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
      }
      uint4 pk;
      __half2 h0 = __floats2half2_rn(v[0], v[1]);
      __half2 h1 = __floats2half2_rn(v[2], v[3]);
      __half2 h2 = __floats2half2_rn(v[4], v[5]);
      __half2 h3 = __floats2half2_rn(v[6], v[7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2);
      pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst + j * 8) = pk;
    }
  } else if constexpr (EPI == EPI_RESID_F32) {
    float* dst = p.out_f32 + static_cast<size_t>(row) * p.ld_f32 + n0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 x = *reinterpret_cast<float4*>(dst + j * 4);
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n0 + j * 4));
      x.x += g.x * (__uint_as_float(r[j * 4 + 0]) + b.x);
      x.y += g.y * (__uint_as_float(r[j * 4 + 1]) + b.y);
      x.z += g.z * (__uint_as_float(r[j * 4 + 2]) + b.z);
      x.w += g.w * (__uint_as_float(r[j * 4 + 3]) + b.w);
      *reinterpret_cast<float4*>(dst + j * 4) = x;
    }
  } else if constexpr (EPI == EPI_PATCH_F32 || EPI == EPI_PATCH_LN_F32) {
    const int b_img = row / p.patches_per_img;
    const int pidx = row - b_img * p.patches_per_img;
    const size_t drow = static_cast<size_t>(b_img) * p.tokens_per_img + p.tok_off + pidx;
    float* dst = p.out_f32 + drow * p.ld_f32 + n0;
    const float* pos = p.pos + static_cast<size_t>(pidx) * p.N + n0;
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
      const float4 e = __ldg(reinterpret_cast<const float4*>(pos + j * 4));
      float4 x;
      x.x = __uint_as_float(r[j * 4 + 0]) + b.x + e.x;
      x.y = __uint_as_float(r[j * 4 + 1]) + b.y + e.y;
      x.z = __uint_as_float(r[j * 4 + 2]) + b.z + e.z;
      x.w = __uint_as_float(r[j * 4 + 3]) + b.w + e.w;
      *reinterpret_cast<float4*>(dst + j * 4) = x;
      v[j * 4 + 0] = x.x; v[j * 4 + 1] = x.y; v[j * 4 + 2] = x.z; v[j * 4 + 3] = x.w;
    }
    if constexpr (EPI == EPI_PATCH_LN_F32) {
      // fp16 copy + partial sums for the first block's LayerNorm (once per forward: plain stores).
      __half* d16 = p.x16 + drow * p.ld_x16 + n0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __half2 h0 = __floats2half2_rn(v[j * 8 + 0], v[j * 8 + 1]);
        __half2 h1 = __floats2half2_rn(v[j * 8 + 2], v[j * 8 + 3]);
        __half2 h2 = __floats2half2_rn(v[j * 8 + 4], v[j * 8 + 5]);
        __half2 h3 = __floats2half2_rn(v[j * 8 + 6], v[j * 8 + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(d16 + j * 8) = pk;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s1 += v[i];
        s2 = fmaf(v[i], v[i], s2);
      }
    }
  } else {  // EPI_BIAS_F32 (+ optional fp16 copy)
    float* dst = p.out_f32 + static_cast<size_t>(row) * p.ld_f32 + n0;
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
      v[j * 4 + 0] = __uint_as_float(r[j * 4 + 0]) + b.x;
      v[j * 4 + 1] = __uint_as_float(r[j * 4 + 1]) + b.y;
      v[j * 4 + 2] = __uint_as_float(r[j * 4 + 2]) + b.z;
      v[j * 4 + 3] = __uint_as_float(r[j * 4 + 3]) + b.w;
      if (p.out_f32 != nullptr)   // the fp32 copy is optional (the k-NN stages read the fp16 one)
        *reinterpret_cast<float4*>(dst + j * 4) =
            make_float4(v[j * 4 + 0], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
    }
    if (p.out_f16 != nullptr) {
      __half* d16 = p.out_f16 + static_cast<size_t>(row) * p.ld_f16 + n0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __half2 h0 = __floats2half2_rn(v[j * 8 + 0], v[j * 8 + 1]);
        __half2 h1 = __floats2half2_rn(v[j * 8 + 2], v[j * 8 + 3]);
        __half2 h2 = __floats2half2_rn(v[j * 8 + 4], v[j * 8 + 5]);
        __half2 h3 = __floats2half2_rn(v[j * 8 + 6], v[j * 8 + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(d16 + j * 8) = pk;
      }
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpilogueWarps);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / num_n;
        const int n_blk = t - m_blk * num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes along K inside the 128B swizzle row = +2 in the (addr >> 4) field
            umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int sub = warp & 3;            // TMEM subpartition this warp may access
    const int half = (warp - 4) >> 2;    // which half of the BN columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / num_n;
      const int n_blk = t - m_blk * num_n;
      const int row = m_blk * BM + sub * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col0 = half * (BN / 2) + c * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + col0, r);
        tmem_ld_wait();
        float unused1 = 0.f, unused2 = 0.f;
        if (row < p.M) epilogue_store32<EPI>(p, row, n_blk * BN + col0, r, unused1, unused2);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------
// 2-CTA variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256
// output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256
// weight rows); the pair's tensor cores read both halves, so per-SM shared-memory operand traffic
// drops from 96 to 64 B/clk and the 128 B/clk port no longer throttles the MMA (see
// profiles/r01_baseline_summary.md).  Only the leader CTA issues MMAs; completion is multicast to
// both CTAs' barriers; both CTAs run their own TMA producer and epilogue warps.
// ----------------------------------------------------------------------------------------
template <int EPI>
struct Gemm2Cfg {
  // The residual epilogue stages its output through shared memory for TMA reduce-add stores
  // (8 epilogue warps x 2 buffers x 4 KB), paid for with one pipeline stage.
  static constexpr bool kTmaReduce = (EPI == EPI_RESID_F32);
  // Residual epilogue that also feeds the next LayerNorm: x is read-modify-written through shared memory (the SM
  // has to see the updated rows to emit their fp16 copy and partial sums), 3 staging buffers per epilogue warp.
  static constexpr bool kResidLn = (EPI == EPI_RESID_LN_F32);
  // LayerNorm statistics applied to the accumulator rows (consumer side of the fusion).
  static constexpr bool kLnFold = (EPI == EPI_LN_BIAS_F16 || EPI == EPI_LN_BIAS_GELU_F16);
  static constexpr bool kGelu = (EPI == EPI_BIAS_GELU_F16 || EPI == EPI_LN_BIAS_GELU_F16);
  // fp16 outputs (qkv, fc1) leave through smem + TMA stores as well: per-thread 16-byte global
  // stores of a row-per-thread fragment (half sectors, 32 rows per instruction) were measured to
  // cost 70-100 us per GEMM; the TMA writes whole 128-byte lines asynchronously.
  static constexpr bool kTmaStore16 = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || kLnFold);
  static constexpr bool kStaged = kTmaReduce || kTmaStore16 || kResidLn;
  static constexpr int kStages = kResidLn ? 4 : (kStaged ? 5 : 6);
  static constexpr int BN = 256;                       // output tile columns (pair)
  static constexpr uint32_t kABytes = BM * BK * 2;     // 16 KB: this CTA's 128 rows of A
  static constexpr uint32_t kBBytes = 128 * BK * 2;    // 16 KB: this CTA's half of the B tile
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 512;           // 2 accumulators x 256 columns
  static constexpr uint32_t kWarpStaging = kResidLn ? 3 * 4096 : 2 * 4096;   // bytes per epilogue warp
  static constexpr uint32_t kStagingBytes = kStaged ? kEpilogueWarps * kWarpStaging : 0;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + 256 + 1024;
};

// x[rows, cols] += tile: fp32 add performed by the TMA / L2 (cp.reduce.async.bulk.tensor).  The SM
// never loads the residual stream, so the epilogue has no global-load latency on its critical path,
// and rows past M are clipped by the tensor map.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD,
                const GemmParams p) {
  using Cfg = Gemm2Cfg<EPI>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int BN = Cfg::BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + STAGES * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* ld_bar = tempty_bar + 3;              // residual read-modify-write mode: 2 per epilogue warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader (issues the MMAs)
  const bool resid_rmw = Cfg::kTmaReduce && (p.flags & 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                 // leader: one arrive.expect_tx for both CTAs' bytes
      mbar_init(&empty_bar[s], 1);                // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);                // multicast tcgen05.commit
      mbar_init(&tempty_bar[a], 2 * kEpilogueWarps);   // leader: epilogue warps of both CTAs
    }
    if constexpr (Cfg::kTmaReduce || Cfg::kResidLn) {
      for (int i = 0; i < 2 * kEpilogueWarps; ++i) mbar_init(&ld_bar[i], 1);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / BK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int m_blk = t / num_n;
        const int n_blk = t - m_blk * num_n;
        const int row_a = m_blk * 2 * BM + static_cast<int>(rank) * BM;
        const int row_b = n_blk * BN + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, row_a);
          tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * BK, row_b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int sub = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t ld_phase[2] = {0, 0};
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m_blk = t / num_n;
      const int n_blk = t - m_blk * num_n;
      const int row0 = m_blk * 2 * BM + static_cast<int>(rank) * BM + sub * 32;
      const int row = row0 + lane;
      if constexpr (Cfg::kTmaReduce || Cfg::kResidLn) {
        // Read-modify-write mode: the x blocks of the first two chunks are requested while the tile's MMAs run.
        if ((resid_rmw || Cfg::kResidLn) && lane == 0) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c == 0) tma_store_wait_read<1>(); else tma_store_wait_read<0>();   // the store that last read it
            uint8_t* buf = staging + (warp - 4) * Cfg::kWarpStaging + c * 4096;
            uint64_t* bar = &ld_bar[(warp - 4) * 2 + c];
            mbar_arrive_expect_tx(bar, 4096);
            tma_load_2d(buf, &tmC, bar, n_blk * BN + half * (BN / 2) + c * 32, row0);
          }
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BN;
      if constexpr (Cfg::kResidLn) {
        // x_new = x + gamma * (acc + bias), read-modify-written through shared memory (x block fetched by the TMA
        // while the tile's MMAs run, plain TMA store back), PLUS what the next LayerNorm needs: the fp16 copy of
        // the new rows (64-column blocks through a third staging buffer) and sum x / sum x^2 of this thread's 128
        // columns, written to stats_out[row, n_blk * 2 + half] - one slot per 128 columns, summed in a fixed
        // order by the consumer (deterministic, no atomics).
        float s1 = 0.f, s2 = 0.f;
        uint8_t* wbuf = staging + (warp - 4) * Cfg::kWarpStaging;
        uint8_t* xbuf = wbuf + 8192;
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          const int col0 = half * (BN / 2) + c * 32;
          const int n0 = n_blk * BN + col0;
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + col0, r);
          tmem_ld_wait();
          uint8_t* buf = wbuf + (c & 1) * 4096;
          mbar_wait(&ld_bar[(warp - 4) * 2 + (c & 1)], ld_phase[c & 1]);
          ld_phase[c & 1] ^= 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n0 + j * 4));
            float4* cell = reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4));
            float4 y = *cell;
            y.x += g.x * (__uint_as_float(r[j * 4 + 0]) + b.x);
            y.y += g.y * (__uint_as_float(r[j * 4 + 1]) + b.y);
            y.z += g.z * (__uint_as_float(r[j * 4 + 2]) + b.z);
            y.w += g.w * (__uint_as_float(r[j * 4 + 3]) + b.w);
            *cell = y;
            s1 += (y.x + y.y) + (y.z + y.w);
            s2 = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, s2))));
            __half2 h0 = __floats2half2_rn(y.x, y.y);
            __half2 h1 = __floats2half2_rn(y.z, y.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            // 64-column fp16 block (128-byte rows, 128B swizzle): 16-byte chunk (c&1)*4 + j/2, half j&1
            *reinterpret_cast<uint2*>(xbuf + lane * 128 + ((((c & 1) * 4 + (j >> 1)) ^ (lane & 7)) << 4) +
                                      (j & 1) * 8) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, buf, n0, row0);
            if (c & 1) tma_store_2d(&tmD, xbuf, n_blk * BN + half * (BN / 2) + (c >> 1) * 64, row0);
            tma_store_commit();
            if (c + 2 < BN / 64) {   // x block of chunk c + 2 into the buffer just stored from
              tma_store_wait_read<0>();
              uint64_t* bar = &ld_bar[(warp - 4) * 2 + (c & 1)];
              mbar_arrive_expect_tx(bar, 4096);
              tma_load_2d(buf, &tmC, bar, n0 + 64, row0);
            }
          }
          __syncwarp();   // xbuf / buf are rewritten by the next chunk only after lane 0 has waited
        }
        if (row < p.M) {
          float2* st = reinterpret_cast<float2*>(p.stats_out) +
                       static_cast<size_t>(row) * (p.N / 128) + n_blk * 2 + half;
          *st = make_float2(s1, s2);
        }
      } else if constexpr (Cfg::kTmaStore16) {
        // Two 32 x 64 fp16 blocks per warp: TMEM -> bias (+GELU) -> 128B-swizzled smem -> TMA store.
        // LN fold: the accumulator row is LN-normalised here, out = rstd (acc - mu colsum) + bias.
        float ln_a = 1.f, ln_g = 0.f;
        if constexpr (Cfg::kLnFold) {
          ln_a = 0.f;
          if (row < p.M) {
            const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + static_cast<size_t>(row) * p.ln_slots;
            float t1 = 0.f, t2 = 0.f;
            for (int t = 0; t < p.ln_slots; ++t) {
              const float2 v2 = __ldg(st + t);
              t1 += v2.x;
              t2 += v2.y;
            }
            const float inv_d = 1.0f / static_cast<float>(p.ln_dim);
            const float mu = t1 * inv_d;
            const float var = fmaxf(fmaf(-mu, mu, t2 * inv_d), 0.f);
            ln_a = rsqrtf(var + p.ln_eps);
            ln_g = -ln_a * mu;
          }
        }
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          const int col0 = half * (BN / 2) + blk * 64;
          const int n0 = n_blk * BN + col0;
          uint32_t r[64];
          uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
          tmem_ld_32x32b_x32(taddr + col0, r0);
          tmem_ld_32x32b_x32(taddr + col0 + 32, r1);
          tmem_ld_wait();
          uint8_t* buf = staging + (warp - 4) * 8192 + blk * 4096;
          if (lane == 0) tma_store_wait_read<1>();   // the store that last read this buffer is done
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j * 8 + i]);
            if constexpr (Cfg::kLnFold) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 8 + 4));
              const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0 + j * 8));
              const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0 + j * 8 + 4));
              v[0] = fmaf(ln_a, v[0], fmaf(ln_g, c0.x, b0.x)); v[1] = fmaf(ln_a, v[1], fmaf(ln_g, c0.y, b0.y));
              v[2] = fmaf(ln_a, v[2], fmaf(ln_g, c0.z, b0.z)); v[3] = fmaf(ln_a, v[3], fmaf(ln_g, c0.w, b0.w));
              v[4] = fmaf(ln_a, v[4], fmaf(ln_g, c1.x, b1.x)); v[5] = fmaf(ln_a, v[5], fmaf(ln_g, c1.y, b1.y));
              v[6] = fmaf(ln_a, v[6], fmaf(ln_g, c1.z, b1.z)); v[7] = fmaf(ln_a, v[7], fmaf(ln_g, c1.w, b1.w));
            } else if (p.bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 8 + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if constexpr (Cfg::kGelu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
            }
            uint4 pk;
            __half2 h0 = __floats2half2_rn(v[0], v[1]);
            __half2 h1 = __floats2half2_rn(v[2], v[3]);
            __half2 h2 = __floats2half2_rn(v[4], v[5]);
            __half2 h3 = __floats2half2_rn(v[6], v[7]);
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2);
            pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, buf, n0, row0);
            tma_store_commit();
          }
        }
      } else {
      float ps1 = 0.f, ps2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col0 = half * (BN / 2) + c * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + col0, r);
        tmem_ld_wait();
        if constexpr (Cfg::kTmaReduce) {
          // y = gamma * (acc + bias) -> 128B-swizzled 32 x 32 fp32 block in smem -> TMA reduce-add.
          const int n0 = n_blk * BN + col0;
          uint8_t* buf = staging + (warp - 4) * 8192 + (c & 1) * 4096;
          if (resid_rmw) {
            // Alternative: x block fetched by the TMA (requested two chunks ahead), x += y in shared memory,
            // plain TMA store - same HBM traffic as the reduce-add, none of its L2 atomics.
            mbar_wait(&ld_bar[(warp - 4) * 2 + (c & 1)], ld_phase[c & 1]);
            ld_phase[c & 1] ^= 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n0 + j * 4));
              float4* cell = reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4));
              float4 y = *cell;
              y.x += g.x * (__uint_as_float(r[j * 4 + 0]) + b.x);
              y.y += g.y * (__uint_as_float(r[j * 4 + 1]) + b.y);
              y.z += g.z * (__uint_as_float(r[j * 4 + 2]) + b.z);
              y.w += g.w * (__uint_as_float(r[j * 4 + 3]) + b.w);
              *cell = y;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, buf, n0, row0);
              tma_store_commit();
              if (c + 2 < BN / 64) {   // request the x block of chunk c + 2 into the buffer just stored from
                tma_store_wait_read<0>();
                uint64_t* bar = &ld_bar[(warp - 4) * 2 + (c & 1)];
                mbar_arrive_expect_tx(bar, 4096);
                tma_load_2d(buf, &tmC, bar, n0 + 64, row0);
              }
            }
          } else {
          if (lane == 0) tma_store_wait_read<1>();   // the store that last read this buffer is done
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 4));
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + n0 + j * 4));
            float4 y;
            y.x = g.x * (__uint_as_float(r[j * 4 + 0]) + b.x);
            y.y = g.y * (__uint_as_float(r[j * 4 + 1]) + b.y);
            y.z = g.z * (__uint_as_float(r[j * 4 + 2]) + b.z);
            y.w = g.w * (__uint_as_float(r[j * 4 + 3]) + b.w);
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = y;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmC, buf, n0, row0);
            tma_store_commit();
          }
          }
        } else {
          if (row < p.M) epilogue_store32<EPI>(p, row, n_blk * BN + col0, r, ps1, ps2);
        }
      }
      if constexpr (EPI == EPI_PATCH_LN_F32) {
        if (row < p.M) {   // one slot per 128 columns of the DESTINATION token row
          const int b_img = row / p.patches_per_img;
          const size_t drow = static_cast<size_t>(b_img) * p.tokens_per_img + p.tok_off + (row - b_img * p.patches_per_img);
          reinterpret_cast<float2*>(p.stats_out)[drow * (p.N / 128) + n_blk * 2 + half] = make_float2(ps1, ps2);
        }
      }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  if constexpr (Cfg::kStaged) {
    if (warp >= 4 && lane == 0) tma_store_wait_read<0>();   // smem must outlive the bulk reads
  }
  // Neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers.
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int EPI>
int launch_2sm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<EPI>;
  CUtensorMap tmC = tmA;   // only read by the staged epilogues
  CUtensorMap tmD = tmA;   // fp16 copy of the residual rows (EPI_RESID_LN_F32)
  if (Cfg::kTmaReduce || Cfg::kResidLn) {
    if (make_tma_2d_f32_sw128(&tmC, p.out_f32, p.M, p.N, p.ld_f32, 32) != 0) return 3;
  }
  if (Cfg::kResidLn) {
    FP_REQUIRE(p.x16 != nullptr && p.stats_out != nullptr && p.ld_x16 % 8 == 0,
               "gemm_tn: EPI_RESID_LN_F32 needs x16 (ld %% 8 == 0) and stats_out");
    if (make_tma_2d_f16(&tmD, p.x16, p.M, p.N, p.ld_x16, 32, 64) != 0) return 3;
  }
  if (Cfg::kLnFold) {
    FP_REQUIRE(p.ln_stats != nullptr && p.ln_colsum != nullptr && p.bias != nullptr && p.ln_slots > 0 &&
               p.ln_dim > 0, "gemm_tn: EPI_LN_* needs ln_stats, ln_colsum, bias, ln_slots and ln_dim");
  }
  if (EPI == EPI_PATCH_LN_F32) {
    FP_REQUIRE(p.x16 != nullptr && p.stats_out != nullptr && p.ld_x16 % 8 == 0,
               "gemm_tn: EPI_PATCH_LN_F32 needs x16 (ld %% 8 == 0) and stats_out");
  }
  if (Cfg::kTmaStore16) {
    FP_REQUIRE(p.ld_f16 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_f16) & 15) == 0,
               "gemm_tn: fp16 output must be 16-byte aligned with ld %% 8 == 0");
    if (make_tma_2d_f16(&tmC, p.out_f16, p.M, p.N, p.ld_f16, 32, 64) != 0) return 3;
  }
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(gemm2_tn_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
  }
  const int num_tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * (p.N / Cfg::BN);
  int clusters = num_tiles < num_sms() / 2 ? num_tiles : num_sms() / 2;
  ProfScope prof(PROF_GEMM, stream, 2.0 * p.M * p.N * p.K);
  gemm2_tn_kernel<EPI><<<2 * clusters, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, tmD, p);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int BN, int EPI>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
               cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  ProfScope prof(PROF_GEMM, stream, 2.0 * p.M * p.N * p.K);
  gemm_tn_kernel<BN, EPI><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int EPI>
int launch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
              cudaStream_t stream) {
  if (bn == 256) return launch_one<256, EPI>(tmA, tmB, p, stream);
  return launch_one<128, EPI>(tmA, tmB, p, stream);
}

}  // namespace

static bool g_force_1sm = false;
static int g_tuning_flags = 0;   // spare A/B switches for experiments (GemmParams::flags)
void gemm_force_1sm(int on) {
  g_force_1sm = (on & 1) != 0;
  g_tuning_flags = (on >> 1) & 0xff;
  attention_set_flags(on >> 16);   // upper half: attention experiment switches
}

int gemm_pick_bn(int M, int N) {
  if (N % 256 != 0) return 128;
  // Prefer the 128x256 tile unless it quantises badly onto 148 SMs.
  const long tiles256 = static_cast<long>((M + BM - 1) / BM) * (N / 256);
  const long waves256 = (tiles256 + num_sms() - 1) / num_sms();
  const double eff256 = static_cast<double>(tiles256) / (waves256 * num_sms());
  const long tiles128 = tiles256 * 2;
  const long waves128 = (tiles128 + num_sms() - 1) / num_sms();
  const double eff128 = static_cast<double>(tiles128) / (waves128 * num_sms());
  return (eff128 > eff256 + 0.04) ? 128 : 256;
}

int gemm_tn(int epi, const __half* A, int lda, const __half* B, int ldb, const GemmParams& p_in,
            cudaStream_t stream) {
  GemmParams p = p_in;
  p.flags = g_tuning_flags;
  FP_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_tn: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  FP_REQUIRE(p.N % 128 == 0, "gemm_tn: N=%d must be a multiple of 128", p.N);
  FP_REQUIRE(p.K % BK == 0, "gemm_tn: K=%d must be a multiple of %d", p.K, BK);
  FP_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm_tn: leading dimensions must be multiples of 8");
  FP_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
             "gemm_tn: operands must be 16-byte aligned");
  const bool use_2sm = (p.N % 256 == 0) && (p.M > 256) && !g_force_1sm;
  const int bn = use_2sm ? 128 : gemm_pick_bn(p.M, p.N);   // 2-CTA: each CTA stages 128 B rows
  CUtensorMap tmA, tmB;
  if (make_tma_2d_f16(&tmA, A, p.M, p.K, lda, BM) != 0) return 3;
  if (make_tma_2d_f16(&tmB, B, p.N, p.K, ldb, bn) != 0) return 3;
  if (use_2sm) {
    switch (epi) {
      case EPI_BIAS_F16: return launch_2sm<EPI_BIAS_F16>(tmA, tmB, p, stream);
      case EPI_BIAS_GELU_F16: return launch_2sm<EPI_BIAS_GELU_F16>(tmA, tmB, p, stream);
      case EPI_RESID_F32: return launch_2sm<EPI_RESID_F32>(tmA, tmB, p, stream);
      case EPI_PATCH_F32: return launch_2sm<EPI_PATCH_F32>(tmA, tmB, p, stream);
      case EPI_BIAS_F32: return launch_2sm<EPI_BIAS_F32>(tmA, tmB, p, stream);
      case EPI_LN_BIAS_F16: return launch_2sm<EPI_LN_BIAS_F16>(tmA, tmB, p, stream);
      case EPI_LN_BIAS_GELU_F16: return launch_2sm<EPI_LN_BIAS_GELU_F16>(tmA, tmB, p, stream);
      case EPI_RESID_LN_F32: return launch_2sm<EPI_RESID_LN_F32>(tmA, tmB, p, stream);
      case EPI_PATCH_LN_F32: return launch_2sm<EPI_PATCH_LN_F32>(tmA, tmB, p, stream);
      default: set_last_error("gemm_tn: unknown epilogue %d", epi); return 1;
    }
  }
  FP_REQUIRE(epi <= EPI_BIAS_F32, "gemm_tn: the LayerNorm-fused epilogues need N %% 256 == 0 and M > 256 "
             "(pair kernel); got M=%d N=%d", p.M, p.N);
  switch (epi) {
    case EPI_BIAS_F16: return launch_bn<EPI_BIAS_F16>(bn, tmA, tmB, p, stream);
    case EPI_BIAS_GELU_F16: return launch_bn<EPI_BIAS_GELU_F16>(bn, tmA, tmB, p, stream);
    case EPI_RESID_F32: return launch_bn<EPI_RESID_F32>(bn, tmA, tmB, p, stream);
    case EPI_PATCH_F32: return launch_bn<EPI_PATCH_F32>(bn, tmA, tmB, p, stream);
    case EPI_BIAS_F32: return launch_bn<EPI_BIAS_F32>(bn, tmA, tmB, p, stream);
    default: set_last_error("gemm_tn: unknown epilogue %d", epi); return 1;
  }
}

// ----------------------------------------------------------------------------
// Host: TMA descriptor creation through the runtime-resolved driver entry point.
// ----------------------------------------------------------------------------
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tensorMapEncodeTiled get_encode_fn() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                     cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(ptr);
  }
  return fn;
}

// fp32 row-major [rows, cols], box [box_rows, 32 columns = 128 bytes], 128B swizzle.
int make_tma_2d_f32_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows) {
  PFN_tensorMapEncodeTiled fn = get_encode_fn();
  if (fn == nullptr) return 3;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(f32) failed (%d): base=%p rows=%llu cols=%llu ld=%llu",
                   static_cast<int>(r), base, (unsigned long long)rows, (unsigned long long)cols,
                   (unsigned long long)ld);
    return 3;
  }
  return 0;
}

int make_tma_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols) {
  PFN_tensorMapEncodeTiled fn = get_encode_fn();
  if (fn == nullptr) return 3;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * sizeof(__half)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u",
                   static_cast<int>(r), base, (unsigned long long)rows, (unsigned long long)cols,
                   (unsigned long long)ld, box_rows, box_cols);
    return 3;
  }
  return 0;
}

// ----------------------------------------------------------------------------
// UMMA descriptor probe: a single-CTA, single-tile contraction used by the GPU tests to pin
// the K-major and MN-major operand descriptors independently of the pipelined kernels.
//   mode 0: D[128,64] = A[128,64] . B[64(n),64(k)]^T      (B K-major)
//   mode 1: D[128,64] = A[128,64] . B[64(k),64(n)]        (B MN-major, as V in attention)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ out, int b_mn_major) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;            // 128 x 64 fp16 = 16 KB
  uint8_t* sb = smem + 16384;    // 64 x 64 fp16  =  8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 16384 + 8192);
    tma_load_2d(sa, &tmA, &bars[0], 0, 0);
    tma_load_2d(sb, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_f16(128, 64, 0, b_mn_major);
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(sa));
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sb));
    for (int k = 0; k < 4; ++k) {
      // K-major: +32 B per 16-element K step. MN-major: +16 rows x 128 B = 2048 B per K step.
      const uint64_t boff = b_mn_major ? static_cast<uint64_t>(k) * (2048 >> 4) : 2ull * k;
      umma_f16_ss(tmem_base, adesc + 2 * k, bdesc + boff, idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[row * 64 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 64);
  }
}

int umma_probe(const __half* A, const __half* B, float* out, int b_mn_major, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  if (make_tma_2d_f16(&tmA, A, 128, 64, 64, 128) != 0) return 3;
  if (make_tma_2d_f16(&tmB, B, 64, 64, 64, 64) != 0) return 3;
  const int smem = 16384 + 8192 + 64 + 1024;
  FP_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem));
  umma_probe_kernel<<<1, 128, smem, stream>>>(tmA, tmB, out, b_mn_major);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
