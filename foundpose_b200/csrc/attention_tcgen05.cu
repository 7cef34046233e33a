// Fused multi-head self-attention for the DINOv2 blocks on tcgen05 (flash-style, no N x N
// matrix in HBM).  Reference arithmetic: external/dinov2/dinov2/layers/attention.py:56-69
//     attn = softmax((q * hd^-0.5) @ k^T);  out = (attn @ v).transpose(1,2).reshape(B,N,C)
//
// Input : qkv  fp16 [B*N, 3*D]  (row = token; columns [q | k | v], each h*64 + d) - exactly what
//               the qkv GEMM epilogue writes (no head split / transpose pass)
// Output: out  fp16 [B*N, D]    (column h*64 + d) - the operand layout of the proj GEMM
//
// One CTA = (128-query tile, head, image).  256 threads:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (128 keys x 64) through rings
//   warp 1      MMA issuer:   S_j = Q K_j^T      (M128 N128 K64, both operands K-major)
//                             PV_j = P_j V_j     (M128 N64 K128, V used MN-major straight from
//                                                 its natural [key, d] layout)
//   warp 2      TMEM allocator (512 columns: S x2, PV x2)
//   warps 4-7   softmax: one thread per query row; S_j from TMEM -> online softmax in fp32 ->
//               P_j (fp16) into 128B-swizzled smem as the A operand of the PV MMA; the output
//               accumulator O lives in registers and is rescaled there (O = O*alpha + PV_j), so
//               no TMEM read-modify-write / correction pass is needed.
// S and PV are double-buffered in TMEM, P in smem: the tensor core computes S_{j+1} while the
// softmax warps work on S_j.
#include "common.cuh"
#include "kernels.h"

namespace fp {

namespace {

constexpr int kHD = 64;        // head dim (all DINOv2 variants)
constexpr int kBQ = 128;       // queries per CTA
constexpr int kBKV = 128;      // keys per block
constexpr int kKVStages = 3;
constexpr int kAttnThreads = 256;

constexpr uint32_t kQBytes = kBQ * kHD * 2;      // 16 KB
constexpr uint32_t kKBytes = kBKV * kHD * 2;     // 16 KB
constexpr uint32_t kVBytes = kBKV * kHD * 2;     // 16 KB
constexpr uint32_t kPBytes = kBQ * kBKV * 2;     // 32 KB (two 64-key swizzle atoms)

struct AttnSmem {
  static constexpr uint32_t q_off = 0;
  static constexpr uint32_t k_off = q_off + kQBytes;
  static constexpr uint32_t v_off = k_off + kKVStages * kKBytes;
  static constexpr uint32_t p_off = v_off + kKVStages * kVBytes;
  static constexpr uint32_t bar_off = p_off + 2 * kPBytes;
  static constexpr uint32_t total = bar_off + 256 + 1024;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int N, int D,
                 float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + AttnSmem::q_off;
  uint8_t* sK = smem + AttnSmem::k_off;
  uint8_t* sV = smem + AttnSmem::v_off;
  uint8_t* sP = smem + AttnSmem::p_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::bar_off);
  uint64_t* q_full = bars;                    // 1
  uint64_t* k_full = bars + 1;                // kKVStages
  uint64_t* k_empty = k_full + kKVStages;     // kKVStages
  uint64_t* v_full = k_empty + kKVStages;     // kKVStages
  uint64_t* v_empty = v_full + kKVStages;     // kKVStages
  uint64_t* s_full = v_empty + kKVStages;     // 2
  uint64_t* s_empty = s_full + 2;             // 2
  uint64_t* p_full = s_empty + 2;             // 2
  uint64_t* pv_done = p_full + 2;             // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int q0 = q_tile * kBQ;
  const int row_base = img * N;            // first token row of this image in the qkv matrix
  const int num_kv = (N + kBKV - 1) / kBKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 4);   // one arrive per softmax warp
      mbar_init(&p_full[b], 4);
      mbar_init(&pv_done[b], 1);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // 2 x 128 columns
  const uint32_t tmem_PV = tmem_base + 256;   // 2 x 64 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_2d(sQ, &tmQKV, q_full, head * kHD, row_base + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < num_kv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], kKBytes);
        tma_load_2d(sK + stage * kKBytes, &tmQKV, &k_full[stage], D + head * kHD,
                    row_base + j * kBKV);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], kVBytes);
        tma_load_2d(sV + stage * kVBytes, &tmQKV, &v_full[stage], 2 * D + head * kHD,
                    row_base + j * kBKV);
        if (++stage == kKVStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(kBQ, kHD, 0, 1);  // B (=V) is MN-major
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      tc_fence_after_sync();
      int kstage = 0, vstage = 0;
      uint32_t kphase = 0, vphase = 0;
      // Software pipeline: S_j is issued before PV_{j-1} so the softmax of block j can start
      // while the tensor core still works on PV_{j-1}.
      for (int j = 0; j <= num_kv; ++j) {
        if (j < num_kv) {
          const int b = j & 1;
          const uint32_t use = static_cast<uint32_t>(j >> 1);  // how often buffer b was used before
          mbar_wait(&k_full[kstage], kphase);
          mbar_wait(&s_empty[b], (use & 1) ^ 1);
          tc_fence_after_sync();
          const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + kstage * kKBytes));
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k)
            umma_f16_ss(tmem_S + b * kBKV, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&k_empty[kstage]);
          umma_commit(&s_full[b]);
          if (++kstage == kKVStages) { kstage = 0; kphase ^= 1; }
        }
        if (j >= 1) {
          const int jj = j - 1;
          const int b = jj & 1;
          const uint32_t use = static_cast<uint32_t>(jj >> 1);
          mbar_wait(&v_full[vstage], vphase);
          mbar_wait(&p_full[b], use & 1);
          tc_fence_after_sync();
          const uint32_t p_addr = smem_u32(sP + b * kPBytes);
          const uint32_t v_addr = smem_u32(sV + vstage * kVBytes);
#pragma unroll
          for (int k = 0; k < kBKV / 16; ++k) {
            // A = P: K-major, two 64-key atoms of 16 KB; +32 B per 16 keys inside an atom.
            const uint64_t pdesc =
                make_smem_desc_sw128(p_addr + (k >> 2) * (kBQ * 128) + (k & 3) * 32);
            // B = V: MN-major, 16 key rows of 128 B per K step.
            const uint64_t vdesc = make_smem_desc_sw128(v_addr + k * 2048);
            umma_f16_ss(tmem_PV + b * kHD, pdesc, vdesc, idesc_pv, k != 0);
          }
          umma_commit(&v_empty[vstage]);
          umma_commit(&pv_done[b]);
          if (++vstage == kKVStages) { vstage = 0; vphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int sub = warp & 3;
    const int r = sub * 32 + lane;                      // query row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    float o[kHD];
#pragma unroll
    for (int i = 0; i < kHD; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    for (int j = 0; j <= num_kv; ++j) {
      if (j < num_kv) {
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        mbar_wait(&s_full[b], use & 1);
        tc_fence_after_sync();
        uint32_t s[kBKV];
#pragma unroll
        for (int c = 0; c < kBKV / 32; ++c) {
          uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]);
          tmem_ld_32x32b_x32(tmem_S + lane_addr + b * kBKV + c * 32, chunk);
        }
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[b]);

        const int valid = N - j * kBKV;  // keys >= valid are out of range in this block
        float m_blk = -INFINITY;
#pragma unroll
        for (int i = 0; i < kBKV; ++i) {
          float v = __uint_as_float(s[i]);
          if (i >= valid) v = -INFINITY;
          s[i] = __float_as_uint(v);
          m_blk = fmaxf(m_blk, v);
        }
        const float m_new = fmaxf(m_run, m_blk);
        const float alpha = fast_exp2((m_run - m_new) * scale_log2e);  // 0 on the first block
        const float neg_m = -m_new * scale_log2e;
        float l_blk = 0.f;
        // P buffer b was last read by PV_{j-2}: that MMA has retired (we waited on pv_done for
        // it when consuming PV_{j-2} at iteration j-1).
        uint8_t* prow = sP + b * kPBytes + r * 128;
#pragma unroll
        for (int c = 0; c < kBKV / 8; ++c) {
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            p[i] = fast_exp2(fmaf(__uint_as_float(s[c * 8 + i]), scale_log2e, neg_m));
            l_blk += p[i];
          }
          __half2 h0 = __floats2half2_rn(p[0], p[1]);
          __half2 h1 = __floats2half2_rn(p[2], p[3]);
          __half2 h2 = __floats2half2_rn(p[4], p[5]);
          __half2 h3 = __floats2half2_rn(p[6], p[7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          // 128B swizzle: 16-byte chunk index XOR (row % 8); atom = c / 8.
          const int atom = c >> 3, cc = c & 7;
          *reinterpret_cast<uint4*>(prow + atom * (kBQ * 128) + ((cc ^ (r & 7)) << 4)) = pk;
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the UMMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);

        // Consume PV_{j-1} (relative to m_{j-1}); alpha_prev rescales O from m_{j-2} to m_{j-1}.
        if (j >= 1) {
          const int pb = (j - 1) & 1;
          const uint32_t puse = static_cast<uint32_t>((j - 1) >> 1);
          mbar_wait(&pv_done[pb], puse & 1);
          tc_fence_after_sync();
          uint32_t t[kHD];
          uint32_t(&t0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&t[0]);
          uint32_t(&t1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&t[32]);
          tmem_ld_32x32b_x32(tmem_PV + lane_addr + pb * kHD, t0);
          tmem_ld_32x32b_x32(tmem_PV + lane_addr + pb * kHD + 32, t1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < kHD; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(t[i]));
          tc_fence_before_sync();
        }
        alpha_prev = alpha;
      } else {
        // Drain: PV of the last block.
        const int pb = (j - 1) & 1;
        const uint32_t puse = static_cast<uint32_t>((j - 1) >> 1);
        mbar_wait(&pv_done[pb], puse & 1);
        tc_fence_after_sync();
        uint32_t t[kHD];
        uint32_t(&t0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&t[0]);
        uint32_t(&t1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&t[32]);
        tmem_ld_32x32b_x32(tmem_PV + lane_addr + pb * kHD, t0);
        tmem_ld_32x32b_x32(tmem_PV + lane_addr + pb * kHD + 32, t1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < kHD; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(t[i]));
        tc_fence_before_sync();
      }
    }

    const int q = q0 + r;
    if (q < N) {
      const float inv_l = 1.0f / l_run;
      __half* dst = out + static_cast<size_t>(row_base + q) * D + head * kHD;
#pragma unroll
      for (int c = 0; c < kHD / 8; ++c) {
        __half2 h0 = __floats2half2_rn(o[c * 8 + 0] * inv_l, o[c * 8 + 1] * inv_l);
        __half2 h1 = __floats2half2_rn(o[c * 8 + 2] * inv_l, o[c * 8 + 3] * inv_l);
        __half2 h2 = __floats2half2_rn(o[c * 8 + 4] * inv_l, o[c * 8 + 5] * inv_l);
        __half2 h3 = __floats2half2_rn(o[c * 8 + 6] * inv_l, o[c * 8 + 7] * inv_l);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + c * 8) = pk;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int attention_f16(const __half* qkv, __half* out, int B, int N, int heads, cudaStream_t stream) {
  const int D = heads * kHD;
  FP_REQUIRE(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  CUtensorMap tm;
  // One descriptor over the whole [B*N, 3D] matrix; box = 128 rows x 64 columns (one head slice).
  if (make_tma_2d_f16(&tm, qkv, static_cast<uint64_t>(B) * N, 3ull * D, 3ull * D, kBQ) != 0) return 3;
  static bool configured = false;
  if (!configured) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttnSmem::total));
    configured = true;
  }
  dim3 grid((N + kBQ - 1) / kBQ, heads, B);
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // hd^-0.5 * log2(e), hd = 64
  ProfScope prof(PROF_ATTENTION, stream, 4.0 * B * heads * static_cast<double>(N) * N * kHD);
  attention_kernel<<<grid, kAttnThreads, AttnSmem::total, stream>>>(tm, out, N, D, scale_log2e);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
