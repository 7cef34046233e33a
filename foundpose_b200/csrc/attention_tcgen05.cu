// Fused multi-head self-attention for the DINOv2 blocks on tcgen05 (flash-style, no N x N
// matrix in HBM).  Reference arithmetic: external/dinov2/dinov2/layers/attention.py:56-69
//     attn = softmax((q * hd^-0.5) @ k^T);  out = (attn @ v).transpose(1,2).reshape(B,N,C)
//
// Input : qkv  fp16 [B*N, 3*D]  (row = token; columns [q | k | v], each h*64 + d) - exactly what
//               the qkv GEMM epilogue writes (no head split / transpose pass)
// Output: out  fp16 [B*N, D]    (column h*64 + d) - the operand layout of the proj GEMM
//
// Persistent CTAs (one per SM, 384 threads) loop over work items = (pair of 128-query tiles, head,
// image).  Warp roles:
//   warp 0       TMA producer: the two Q tiles of the item, then K_j / V_j tiles (128 keys x 64)
//                through 3-stage rings shared by both query tiles
//   warp 1       MMA issuer 1: S_X(j) = Q_X K_j^T   (M128 N<=128 K64, operands K-major)
//   warp 3       MMA issuer 2: PV_X(j) = P_X(j) V_j (M128 N64 K<=128, V MN-major from its natural
//                                                    [key, d] layout)
//                              L_X(j)  = P_X(j) 1   (M128 N16: exact fp32 row sums of the fp16 P)
//   warp 2       TMEM allocator (512 columns: S_A, S_B, PV_A, PV_B, L_A, L_B)
//   warps 4-7    softmax warpgroup of query tile A, warps 8-11 of tile B: one thread per query row.
//                S(j) TMEM -> registers, running max in fp32, P = exp2(...) packed to f16x2 and written to 128B-swizzled smem as the A operand of the next
//                MMAs; O lives in registers and is rescaled there (O = O*alpha + PV(j)), so there
//                is no TMEM read-modify-write / correction pass.
// The two warpgroups ping-pong on the tensor core: while one computes its softmax the MMAs of the
// other are in flight.  The last key block is only round_up(valid, 16) keys wide.
#include "common.cuh"
#include "kernels.h"

namespace fp {

static unsigned long long* g_attn_dbg = nullptr;   // timeline buffer (tools/attn_timeline.py)
void attention_set_debug_buffer(void* p) { g_attn_dbg = static_cast<unsigned long long*>(p); }
static int g_attn_flags = 0;   // experiment switches (tools/attn_bench.py); 0 in production
void attention_set_flags(int f) { g_attn_flags = f; }

namespace {

// Timeline probe: CTA 0 only, role r in [0,5), up to 2048 events per role.
__device__ __forceinline__ void dbg_event(unsigned long long* dbg, int role, int& n, int tag) {
  if (dbg != nullptr && blockIdx.x == 0 && n < 2048) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[role * 2048 + n] = (static_cast<unsigned long long>(tag) << 48) | (t & 0xffffffffffffull);
    ++n;
  }
}

constexpr int kHD = 64;        // head dim (all DINOv2 variants)
constexpr int kBQ = 128;       // queries per tile (two tiles per work item)
constexpr int kBKV = 128;      // keys per block
constexpr int kKVStages = 3;
constexpr int kAttnThreads = 384;

constexpr uint32_t kQBytes = kBQ * kHD * 2;      // 16 KB per query tile
constexpr uint32_t kKBytes = kBKV * kHD * 2;     // 16 KB
constexpr uint32_t kVBytes = kBKV * kHD * 2;     // 16 KB
constexpr uint32_t kPBytes = kBQ * kBKV * 2;     // 32 KB (two 64-key swizzle atoms)
constexpr uint32_t kOnesBytes = 4096;            // 16 x 128 fp16 ones (B operand of the row-sum MMA)

struct AttnSmem {
  static constexpr uint32_t q_off = 0;
  static constexpr uint32_t k_off = q_off + 2 * kQBytes;
  static constexpr uint32_t v_off = k_off + kKVStages * kKBytes;
  static constexpr uint32_t p_off = v_off + kKVStages * kVBytes;
  static constexpr uint32_t ones_off = p_off + 2 * kPBytes;
  static constexpr uint32_t bar_off = ones_off + kOnesBytes;
  static constexpr uint32_t total = bar_off + 256 + 1024;
};

// TMEM column map (512 allocated)
constexpr uint32_t kColS = 0;      // S_A at 0, S_B at 128
constexpr uint32_t kColPV = 256;   // PV_A at 256, PV_B at 320
constexpr uint32_t kColL = 384;    // L_A at 384, L_B at 400

template <uint32_t kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <uint32_t kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of two scores in fp32 (MUFU.EX2), packed to f16x2 for the P operand.
__device__ __forceinline__ uint32_t exp2_f16x2(float lo, float hi) {
  uint32_t p;
  const float elo = fast_exp2(lo), ehi = fast_exp2(hi);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(ehi), "f"(elo));
  return p;
}

__device__ __forceinline__ void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// O (64 columns) and the row sum L accumulate in TMEM across key blocks.  When the running max of
// a row moves by more than 2^8 the accumulators are rescaled in place: O *= alpha, L *= alpha
// (warp-collective TMEM load / multiply / store; rare after the first blocks).
__device__ __forceinline__ void rescale_accumulators(uint32_t tPV, uint32_t tL, float alpha) {
#pragma unroll
  for (int c = 0; c < kHD / 16; ++c) {
    uint32_t t[16];
    tmem_ld_32x32b_x16(tPV + c * 16, t);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
    tmem_st_32x32b_x16(tPV + c * 16, t);
  }
  uint32_t l;
  tmem_ld_32x32b_x1(tL, l);
  tmem_ld_wait();
  tmem_st_32x32b_x1(tL, __float_as_uint(__uint_as_float(l) * alpha));
  tmem_st_wait();
}

struct AttnBars {
  uint64_t q_full, q_empty;
  uint64_t k_full[kKVStages], k_empty[kKVStages];
  uint64_t v_full[kKVStages], v_empty[kKVStages];
  uint64_t s_full[2], s_empty[2], p_full[2], pv_done[2];
  uint32_t tmem_slot;
};

// Online softmax of one key block for one query row (= one thread).  The reference max is updated
// lazily: the TMEM accumulators are only rescaled when a row's max moved by more than 2^8, so
// P <= 256 stays well inside fp16 and the rescale is rare after the first blocks.
constexpr float kRescaleThreshold = 8.0f;   // log2 domain

// Decide the reference max of this block.  Returns true when the accumulators must be rescaled
// (by exp2((m_old - m_used) * c)) once PV(j-1) has retired; warp-uniform.
__device__ __forceinline__ bool choose_reference_max(bool first, float scale_log2e, float m_blk,
                                                     float& m_used, float& m_old) {
  const float m_new = fmaxf(m_used, m_blk);
  m_old = m_used;
  if (first) {
    m_used = m_new;
    return false;
  }
  const bool need = (m_new - m_used) * scale_log2e > kRescaleThreshold;
  const bool any = __any_sync(0xffffffffu, need);
  if (any) m_used = m_new;
  return any;
}

// PV(j-1) done: the P buffer may be overwritten and O / L are stable (and can be rescaled).
__device__ __forceinline__ void wait_prev_pv(AttnBars* bars, int x, uint32_t tPV, uint32_t tL, bool first,
                                             uint32_t par, bool rescale, float scale_log2e, float m_old,
                                             float m_used) {
  if (first) return;
  mbar_wait(&bars->pv_done[x], par ^ 1);
  tc_fence_after_sync();
  if (rescale) {
    rescale_accumulators(tPV, tL, fast_exp2((m_old - m_used) * scale_log2e));
    tc_fence_before_sync();
  }
}

__device__ __forceinline__ uint4 exp_pack8(const uint32_t* s8, float scale_log2e, float neg_m) {
  uint4 pk;
  pk.x = exp2_f16x2(fmaf(__uint_as_float(s8[0]), scale_log2e, neg_m), fmaf(__uint_as_float(s8[1]), scale_log2e, neg_m));
  pk.y = exp2_f16x2(fmaf(__uint_as_float(s8[2]), scale_log2e, neg_m), fmaf(__uint_as_float(s8[3]), scale_log2e, neg_m));
  pk.z = exp2_f16x2(fmaf(__uint_as_float(s8[4]), scale_log2e, neg_m), fmaf(__uint_as_float(s8[5]), scale_log2e, neg_m));
  pk.w = exp2_f16x2(fmaf(__uint_as_float(s8[6]), scale_log2e, neg_m), fmaf(__uint_as_float(s8[7]), scale_log2e, neg_m));
  return pk;
}

// 128B swizzle of the P operand: 16-byte chunk index XOR (row % 8); atom = chunk / 8.
__device__ __forceinline__ void store_p_chunk(uint8_t* prow, int r, int c, const uint4& pk) {
  const int atom = c >> 3, cc = c & 7;
  *reinterpret_cast<uint4*>(prow + atom * (kBQ * 128) + ((cc ^ (r & 7)) << 4)) = pk;
}

// Full block: all 128 keys valid - fully static code, S held in registers (one TMEM pass).
// The exponentials are computed BEFORE waiting for PV(j-1), so that wait is off the critical path;
// only the shared-memory stores of P come after it.
__device__ __forceinline__ void softmax_block_full(AttnBars* bars, int x, int lane, uint32_t tS, uint32_t tPV,
                                                   uint32_t tL, uint8_t* prow, int r, bool first,
                                                   uint32_t par, float scale_log2e, float& m_used,
                                                   unsigned long long* dbg, int& dn, bool dbg_me, int j) {
  uint32_t s[kBKV];
#pragma unroll
  for (int c = 0; c < kBKV / 32; ++c) {
    uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]);
    tmem_ld_32x32b_x32(tS + c * 32, chunk);
  }
  tmem_ld_wait();
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars->s_empty[x]);   // S may be recomputed for the next block
  if (dbg_me) dbg_event(dbg, 3 + x, dn, 200 + j);
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < kBKV; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[i]));
  const float m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  float m_old;
  const bool rescale = choose_reference_max(first, scale_log2e, m_blk, m_used, m_old);
  const float neg_m = -m_used * scale_log2e;
  if (dbg_me) dbg_event(dbg, 3 + x, dn, 300 + j);
  uint4 pk[kBKV / 8];
#pragma unroll
  for (int c = 0; c < kBKV / 8; ++c) pk[c] = exp_pack8(&s[c * 8], scale_log2e, neg_m);
  if (dbg_me) dbg_event(dbg, 3 + x, dn, 400 + j);
  wait_prev_pv(bars, x, tPV, tL, first, par, rescale, scale_log2e, m_old, m_used);
#pragma unroll
  for (int c = 0; c < kBKV / 8; ++c) store_p_chunk(prow, r, c, pk[c]);
  if (dbg_me) dbg_event(dbg, 3 + x, dn, 500 + j);
}

// Last (partial) block: `len` (multiple of 16) columns were computed, keys >= valid are masked.
// Two TMEM passes over 32-column chunks keep the register footprint small; it runs once per item.
__device__ __forceinline__ void softmax_block_tail(AttnBars* bars, int x, int lane, uint32_t tS, uint32_t tPV,
                                                   uint32_t tL, uint8_t* prow, int r, int len, int valid,
                                                   bool first, uint32_t par, float scale_log2e,
                                                   float& m_used) {
  const int chunks = (len + 31) / 32;
  float m_blk = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tS + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c * 32 + i < valid) m_blk = fmaxf(m_blk, __uint_as_float(v[i]));
  }
  float m_old;
  const bool rescale = choose_reference_max(first, scale_log2e, m_blk, m_used, m_old);
  const float neg_m = -m_used * scale_log2e;
  wait_prev_pv(bars, x, tPV, tL, first, par, rescale, scale_log2e, m_old, m_used);
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tS + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c * 32 + i >= valid) v[i] = 0xff800000u;   // -inf -> P = 0
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (c * 32 + g * 8 < len) store_p_chunk(prow, r, c * 4 + g, exp_pack8(&v[g * 8], scale_log2e, neg_m));
  }
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars->s_empty[x]);
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int N, int D,
                 int heads, int num_items, float scale_log2e, int flags, unsigned long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + AttnSmem::q_off;
  uint8_t* sK = smem + AttnSmem::k_off;
  uint8_t* sV = smem + AttnSmem::v_off;
  uint8_t* sP = smem + AttnSmem::p_off;
  uint8_t* sOnes = smem + AttnSmem::ones_off;
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + AttnSmem::bar_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kv = (N + kBKV - 1) / kBKV;
  const int pairs = (N + 2 * kBQ - 1) / (2 * kBQ);
  // Width of the last key block: only as many 16-key MMA steps as there are valid keys.
  const int last_valid = N - (num_kv - 1) * kBKV;
  const int last_len = (last_valid + 15) / 16 * 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->q_empty, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&bars->k_full[s], 1);
      mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1);
      mbar_init(&bars->v_empty[s], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bars->s_full[x], 1);
      mbar_init(&bars->s_empty[x], 4);   // one arrive per softmax warp of that tile
      mbar_init(&bars->p_full[x], 4);
      mbar_init(&bars->pv_done[x], 1);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(&bars->tmem_slot, 512);
    tmem_relinquish();
  } else if (warp == 3) {
    // fp16 1.0 everywhere: any descriptor pointing into this region reads a ones matrix.
    uint32_t* o32 = reinterpret_cast<uint32_t*>(sOnes);
    for (int i = lane; i < static_cast<int>(kOnesBytes / 4); i += 32) o32[i] = 0x3C003C00u;
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      uint32_t item_phase = 0;
      int dn = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int pair = it % pairs;
        const int head = (it / pairs) % heads;
        const int img = it / (pairs * heads);
        const int row_base = img * N;
        mbar_wait(&bars->q_empty, item_phase ^ 1);
        dbg_event(dbg, 0, dn, 100);
        mbar_arrive_expect_tx(&bars->q_full, 2 * kQBytes);
        tma_load_2d(sQ, &tmQKV, &bars->q_full, head * kHD, row_base + pair * 2 * kBQ);
        tma_load_2d(sQ + kQBytes, &tmQKV, &bars->q_full, head * kHD, row_base + pair * 2 * kBQ + kBQ);
        item_phase ^= 1;
        for (int j = 0; j < num_kv; ++j) {
          mbar_wait(&bars->k_empty[stage], phase ^ 1);
          dbg_event(dbg, 0, dn, 10 + j);
          mbar_arrive_expect_tx(&bars->k_full[stage], kKBytes);
          tma_load_2d(sK + stage * kKBytes, &tmQKV, &bars->k_full[stage], D + head * kHD,
                      row_base + j * kBKV);
          mbar_wait(&bars->v_empty[stage], phase ^ 1);
          dbg_event(dbg, 0, dn, 30 + j);
          mbar_arrive_expect_tx(&bars->v_full[stage], kVBytes);
          tma_load_2d(sV + stage * kVBytes, &tmQKV, &bars->v_full[stage], 2 * D + head * kHD,
                      row_base + j * kBKV);
          if (++stage == kKVStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer 1: S_X(j) = Q_X K_j^T for both query tiles =====
      // The WHOLE warp runs this loop with warp-uniform operands and one elected lane issues the
      // MMAs: issuing tcgen05.mma from a divergent `lane == 0` branch makes the compiler wrap every
      // instruction in a waterfall loop (R2UR.BROADCAST / BRA.U.ANY), ~90 cycles per MMA, which
      // serialised the 24+ small MMAs of every key block (profiles/r01_attention_timeline.md).
      const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(sQ));
      const uint64_t qdesc1 = make_smem_desc_sw128(smem_u32(sQ + kQBytes));
      const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK));
      constexpr uint32_t idesc_full = make_idesc_f16(kBQ, kBKV, 0, 0);
      const uint32_t idesc_last = make_idesc_f16(kBQ, last_len, 0, 0);
      int kstage = 0;
      uint32_t kphase = 0, item_phase = 0;
      uint32_t blk = 0;   // running count of key blocks processed by this CTA (barrier parity)
      int dn = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        mbar_wait(&bars->q_full, item_phase);
        if (lane == 0) dbg_event(dbg, 1, dn, 100);
        item_phase ^= 1;
        for (int j = 0; j < num_kv; ++j) {
          const uint32_t idesc_s = (j == num_kv - 1) ? idesc_last : idesc_full;
          const uint32_t par = (blk + j) & 1;
          mbar_wait(&bars->k_full[kstage], kphase);
          const uint64_t kd = kdesc0 + static_cast<uint64_t>(kstage * (kKBytes / 16));
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            mbar_wait(&bars->s_empty[x], par ^ 1);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint64_t qd = x == 0 ? qdesc0 : qdesc1;
#pragma unroll
              for (int k = 0; k < kHD / 16; ++k)
                umma_f16_ss(tmem_base + kColS + x * kBKV, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
              umma_commit(&bars->s_full[x]);
            }
            __syncwarp();
            if (lane == 0) dbg_event(dbg, 1, dn, 40 + 20 * x + j);
          }
          if (elect_one()) {
            umma_commit(&bars->k_empty[kstage]);
            if (j == num_kv - 1) umma_commit(&bars->q_empty);   // Q tiles may be overwritten
          }
          __syncwarp();
          if (++kstage == kKVStages) { kstage = 0; kphase ^= 1; }
        }
        blk += num_kv;
      }
    } else if (warp == 3) {
      // ===== MMA issuer 2: PV_X(j) = P_X(j) V_j and the row sums L_X(j) = P_X(j) 1 =====
      constexpr uint32_t idesc_pv = make_idesc_f16(kBQ, kHD, 0, 1);   // B (=V) is MN-major
      constexpr uint32_t idesc_l = make_idesc_f16(kBQ, 16, 0, 0);
      // A = P: K-major, two 64-key atoms of 16 KB; +32 B (= +2) per 16 keys inside an atom.
      const uint64_t pdesc0 = make_smem_desc_sw128(smem_u32(sP));
      // B = V: MN-major, 16 key rows of 128 B (= +128 in the address field) per K step.
      const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV));
      const uint64_t odesc = make_smem_desc_sw128(smem_u32(sOnes));
      int vstage = 0;
      uint32_t vphase = 0;
      uint32_t blk = 0;
      int dn = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        for (int j = 0; j < num_kv; ++j) {
          const uint32_t par = (blk + j) & 1;
          const uint32_t acc0 = j != 0;          // the first block of an item overwrites O / L
          const int ksteps = (j == num_kv - 1) ? last_len / 16 : kBKV / 16;
          mbar_wait(&bars->v_full[vstage], vphase);
          if (lane == 0) dbg_event(dbg, 2, dn, 10 + j);
          const uint64_t vd = vdesc0 + static_cast<uint64_t>(vstage * (kVBytes / 16));
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            mbar_wait(&bars->p_full[x], par);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint64_t pbase = pdesc0 + static_cast<uint64_t>(x * (kPBytes / 16));
              if (ksteps == kBKV / 16) {
#pragma unroll
                for (int k = 0; k < kBKV / 16; ++k) {
                  const uint64_t pd = pbase + static_cast<uint64_t>((k >> 2) * (kBQ * 128 / 16) + (k & 3) * 2);
                  umma_f16_ss(tmem_base + kColPV + x * kHD, pd, vd + k * 128, idesc_pv, acc0 | (k != 0));
                  umma_f16_ss(tmem_base + kColL + x * 16, pd, odesc + (k & 3) * 2, idesc_l, acc0 | (k != 0));
                }
              } else {
#pragma unroll 1
                for (int k = 0; k < ksteps; ++k) {
                  const uint64_t pd = pbase + static_cast<uint64_t>((k >> 2) * (kBQ * 128 / 16) + (k & 3) * 2);
                  umma_f16_ss(tmem_base + kColPV + x * kHD, pd, vd + k * 128, idesc_pv, acc0 | (k != 0));
                  umma_f16_ss(tmem_base + kColL + x * 16, pd, odesc + (k & 3) * 2, idesc_l, acc0 | (k != 0));
                }
              }
              umma_commit(&bars->pv_done[x]);
            }
            __syncwarp();
            if (lane == 0) dbg_event(dbg, 2, dn, 40 + 20 * x + j);
          }
          if (elect_one()) umma_commit(&bars->v_empty[vstage]);
          __syncwarp();
          if (++vstage == kKVStages) { vstage = 0; vphase ^= 1; }
        }
        blk += num_kv;
      }
    }
  } else {
    // ===== softmax warpgroups =====
    const int x = (warp - 4) >> 2;                      // query tile of this warpgroup (0 = A, 1 = B)
    const int sub = warp & 3;
    const int r = sub * 32 + lane;                      // query row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + kColS + x * kBKV;
    const uint32_t tPV = tmem_base + lane_addr + kColPV + x * kHD;
    const uint32_t tL = tmem_base + lane_addr + kColL + x * 16;
    uint8_t* prow = sP + x * kPBytes + r * 128;
    uint32_t blk = 0;
    int dn = 0;
    const bool dbg_me = (sub == 0 && lane == 0);
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int pair = it % pairs;
      const int head = (it / pairs) % heads;
      const int img = it / (pairs * heads);
      const int row_base = img * N;
      float m_used = -INFINITY;
      for (int j = 0; j < num_kv; ++j) {
        const uint32_t par = (blk + j) & 1;
        const int valid = N - j * kBKV;            // keys >= valid are out of range
        const int len = (j == num_kv - 1) ? last_len : kBKV;
        mbar_wait(&bars->s_full[x], par);
        if (dbg_me) dbg_event(dbg, 3 + x, dn, 10 + j);
        tc_fence_after_sync();
        if (len == kBKV && valid >= kBKV) {
          softmax_block_full(bars, x, lane, tS, tPV, tL, prow, r, j == 0, par, scale_log2e, m_used, dbg, dn, dbg_me, j);
        } else {
          softmax_block_tail(bars, x, lane, tS, tPV, tL, prow, r, len, valid, j == 0, par, scale_log2e, m_used);
        }
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the UMMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[x]);
        if (dbg_me) dbg_event(dbg, 3 + x, dn, 30 + j);
      }
      // Drain: O and the row sums are complete once PV of the last block has retired.
      mbar_wait(&bars->pv_done[x], (blk + num_kv - 1) & 1);
      if (dbg_me) dbg_event(dbg, 3 + x, dn, 90);
      tc_fence_after_sync();
      float o[kHD];
      uint32_t lsum;
      tmem_ld_32x32b_x1(tL, lsum);
#pragma unroll
      for (int c = 0; c < kHD / 16; ++c) {
        uint32_t(&t)[16] = *reinterpret_cast<uint32_t(*)[16]>(&o[c * 16]);
        tmem_ld_32x32b_x16(tPV + c * 16, t);
      }
      tmem_ld_wait();
      tc_fence_before_sync();
      const float l_run = __uint_as_float(lsum);
      blk += num_kv;

      const int q = pair * 2 * kBQ + x * kBQ + r;
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
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttnSmem::total));
  }
  const int pairs = (N + 2 * kBQ - 1) / (2 * kBQ);
  const int num_items = pairs * heads * B;
  const int grid = num_items < kNumSMs ? num_items : kNumSMs;
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // hd^-0.5 * log2(e), hd = 64
  ProfScope prof(PROF_ATTENTION, stream, 4.0 * B * heads * static_cast<double>(N) * N * kHD);
  attention_kernel<<<grid, kAttnThreads, AttnSmem::total, stream>>>(tm, out, N, D, heads, num_items,
                                                                    scale_log2e, g_attn_flags, g_attn_dbg);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
