// Fused multi-head self-attention for the DINOv2 blocks on tcgen05 (flash-style, no N x N
// matrix in HBM).  Reference arithmetic: external/dinov2/dinov2/layers/attention.py:56-69
//     attn = softmax((q * hd^-0.5) @ k^T);  out = (attn @ v).transpose(1,2).reshape(B,N,C)
//
// Input : qkv  fp16 [B*N, 3*D]  (row = token; columns [q | k | v], each h*64 + d) - exactly what
//               the qkv GEMM epilogue writes (no head split / transpose pass)
// Output: out  fp16 [B*N, D]    (column h*64 + d) - the operand layout of the proj GEMM
//
// Persistent CTAs (one per SM, 640 threads) loop over work items = (pair of 128-query tiles, head,
// image).  Warp roles:
//   warp 0       TMA producer: the two Q tiles of the item, then K_j / V_j tiles (128 keys x 64)
//                through 3-stage rings shared by both query tiles
//   warp 1       MMA issuer 1: S_X(j) = Q_X K_j^T   (M128 N<=128 K64, both operands from smem)
//   warp 3       MMA issuer 2: O_X += P_X(j) V_j    (M128 N64 K<=128; A = P read from TENSOR MEMORY,
//                                                    B = V MN-major from its natural [key, d] layout)
//   warp 2       TMEM allocator (512 columns: S_A, S_B fp32 | O_A, O_B fp32 | P_A, P_B packed fp16)
//   warps 4-19   softmax: FOUR warps per SM sub-partition.  A query row is handled by a pair of
//                threads in two different warps (same TMEM lane quarter), 64 of the 128 keys of a
//                block each: S(j) TMEM -> registers, half-row max exchanged through shared memory
//                (named barrier of the two warps), lazily updated reference max, P = exp2(...)
//                packed to f16x2 and written back to tensor memory with tcgen05.st as the A operand
//                of the PV MMA.  Row sums accumulate in registers (one partial per thread, joined
//                at the end of the item); O accumulates in TMEM across key blocks.
// Why this shape (profiles/r01_attention_timeline.md): with one thread per row (two softmax warps
// per sub-partition) every phase of a block - TMEM load, max, exponentials, TMEM store, barrier
// hand-offs - is a serial latency of that warp, and the 4-lane/clk MUFU pipe idled 45% of the time.
// Four warps per sub-partition overlap the latencies of one with the exponentials of the others.
// P never touches shared memory (the S MMA alone reads 128 B/clk of smem operands).  A share of
// the exponentials can be evaluated by a polynomial on the FMA pipe.  The last key block is only
// round_up(valid, 16) keys wide.
#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>

namespace fp {

static unsigned long long* g_attn_dbg = nullptr;   // timeline buffer (tools/attn_timeline.py)
void attention_set_debug_buffer(void* p) { g_attn_dbg = static_cast<unsigned long long*>(p); }
static int g_attn_flags = 0;   // experiment switches (tools/attn_bench.py); 0 in production
void attention_set_flags(int f) { g_attn_flags = f; }

namespace {

// Timeline probe (tools/attn_timeline.py): CTA 0 only, role r in [0,5), up to 2048 events per role.
// Compiled in only with -DFP_ATTN_TIMELINE (make EXTRA=-DFP_ATTN_TIMELINE): the probe costs registers
// and ~60 cycles per event, so the production kernel carries none of it.
__device__ __forceinline__ void dbg_event(unsigned long long* dbg, int role, int& n, int tag) {
#ifdef FP_ATTN_TIMELINE
  if (dbg != nullptr && blockIdx.x == 0 && n < 2048) {
    const unsigned long long t = static_cast<unsigned long long>(clock64());   // SM cycles
    dbg[role * 2048 + n] = (static_cast<unsigned long long>(tag) << 48) | (t & 0xffffffffffffull);
    ++n;
  }
#endif
}

constexpr int kHD = 64;        // head dim (all DINOv2 variants)
constexpr int kBQ = 128;       // queries per tile (two tiles per work item)
constexpr int kBKV = 128;      // keys per block
constexpr int kKVStages = 3;
constexpr int kDefaultVariant = 0;   // index into the launcher's variant table

// kSplit = number of threads (in different warps of the same TMEM lane quarter) that share one
// query row: 1 -> 8 softmax warps (384 threads), 2 -> 16 softmax warps (640 threads).
__host__ __device__ constexpr int attn_threads(int split) { return 128 + 256 * split; }

constexpr uint32_t kQBytes = kBQ * kHD * 2;      // 16 KB per query tile
constexpr uint32_t kKBytes = kBKV * kHD * 2;     // 16 KB
constexpr uint32_t kVBytes = kBKV * kHD * 2;     // 16 KB

struct AttnBars {
  uint64_t q_full[2], q_empty[2];   // the Q pair of the NEXT item loads while the current one computes
  uint64_t k_full[kKVStages], k_empty[kKVStages];
  uint64_t v_full[kKVStages], v_empty[kKVStages];
  uint64_t s_full[2], s_empty[2], p_full[2], pv_done[2];
  uint32_t tmem_slot;
};

struct AttnSmem {
  static constexpr uint32_t q_off = 0;
  static constexpr uint32_t k_off = q_off + 2 * 2 * kQBytes;   // two stages of a Q pair
  static constexpr uint32_t v_off = k_off + kKVStages * kKBytes;
  static constexpr uint32_t max_off = v_off + kKVStages * kVBytes;   // float [2 par][2 tile][2 half][128]
  static constexpr uint32_t sum_off = max_off + 2 * 2 * 2 * kBQ * 4; // float [2 tile][2 half][128]
  static constexpr uint32_t bar_off = sum_off + 2 * 2 * kBQ * 4;
  static constexpr uint32_t total = bar_off + 256 + 1024;
};

// TMEM column map (512 allocated)
constexpr uint32_t kColS = 0;      // S_A at 0, S_B at 128 (fp32)
constexpr uint32_t kColO = 256;    // O_A at 256, O_B at 320 (fp32)
constexpr uint32_t kColP = 384;    // P_A at 384, P_B at 448 (two fp16 keys per 32-bit column)

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

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2: one issue slot for two fp32 lanes on sm_100) ----
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// exp2 of a pair on the FMA / ALU pipes instead of the 4-lane/clk/sub-partition MUFU: round-to-
// nearest range reduction with the 1.5*2^23 magic constant (the integer part lands in the low
// mantissa bits), degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 7.5e-5,
// below the fp16 half-ulp of P), and the integer part added straight into the exponent field.
// Inputs are clamped at -60: such P underflow to 0 in fp16 either way.
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& e0, float& e1) {
  const uint64_t x = pack2(fmaxf(x0, -60.0f), fmaxf(x1, -60.0f));
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);
  const uint64_t t = add2(x, magic);
  const uint64_t f = sub2(x, sub2(t, magic));
  uint64_t p = fma2(f, pack2(0.05517132207751274f, 0.05517132207751274f),
                    pack2(0.24261054396629333f, 0.24261054396629333f));
  p = fma2(p, f, pack2(0.6932609677314758f, 0.6932609677314758f));
  p = fma2(p, f, pack2(0.9999281167984009f, 0.9999281167984009f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
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
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Named barrier of the two warps that share a set of 32 query rows.
__device__ __forceinline__ void pair_sync(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// O (this thread's kCols of the 64 columns) accumulates in TMEM across key blocks.  When the running
// max of a row moves by more than 2^8 the accumulator is rescaled in place: O *= alpha
// (warp-collective TMEM load / multiply / store; rare after the first blocks).
template <int kCols>
__device__ __forceinline__ void rescale_accumulator(uint32_t tO_half, float alpha) {
#pragma unroll
  for (int c = 0; c < kCols / 16; ++c) {
    uint32_t t[16];
    tmem_ld_32x32b_x16(tO_half + c * 16, t);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
    tmem_st_32x32b_x16(tO_half + c * 16, t);
  }
  tmem_st_wait();
}

// Online softmax with a lazily updated reference max: the accumulator is only rescaled when a
// row's max moved by more than 2^8, so P <= 256 stays well inside fp16 and the rescale is rare
// after the first blocks.
constexpr float kRescaleThreshold = 8.0f;   // log2 domain

// Decide the reference max of this block.  Returns true when the accumulators must be rescaled
// (by exp2((m_old - m_used) * c)) once PV(j-1) has retired; warp-uniform, and identical in the
// two warps of a pair (they see the same 32 row maxima).
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

// Two scores -> (packed fp16 P pair, fp32 exponentials added to the row-sum accumulators).
// kPoly selects the FMA-pipe polynomial instead of MUFU.EX2.
template <bool kPoly>
__device__ __forceinline__ uint32_t exp_pair(uint32_t s0, uint32_t s1, uint64_t scale2, uint64_t negm2,
                                             float& sum0, float& sum1) {
  const uint64_t x = fma2(pack2u(s0, s1), scale2, negm2);
  float x0, x1, e0, e1;
  unpack2(x, x0, x1);
  if (kPoly) {
    exp2_poly_pair(x0, x1, e0, e1);
  } else {
    e0 = fast_exp2(x0);
    e1 = fast_exp2(x1);
  }
  sum0 += e0;
  sum1 += e1;
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}

// Per-thread context of a softmax thread (row r of tile x, key half h).
struct SoftmaxCtx {
  AttnBars* bars;
  float* max_buf;     // [2 par][this tile][2 half][128]: base of this tile, parity 0
  int x, half, r, lane, pair_bar;
  uint32_t tS, tO, tP;   // TMEM addresses of this thread's S half, O half and P half
  float scale_log2e;
};

// One key block.  kPartial: the last block of an item - `len` (multiple of 16) columns were
// computed by the S MMA and keys >= valid are masked; kPoly64 of every 64 pairs go through the
// FMA-pipe polynomial (full blocks only).  The exponentials are computed BEFORE waiting for
// PV(j-1), so that wait is off the critical path; only the TMEM stores of P come after it.
template <int kSplit, int kPoly64, bool kPartial>
__device__ __forceinline__ void softmax_block(const SoftmaxCtx& c, int len, int valid, bool first, uint32_t par,
                                              float& m_used, float& l_part, unsigned long long* dbg, int& dn,
                                              bool dbg_me, int j) {
  constexpr int kHalf = kBKV / kSplit;   // keys per softmax thread and block
  uint32_t s[kHalf];
#pragma unroll
  for (int q = 0; q < kHalf / 32; ++q) {
    uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[q * 32]);
    const int col0 = c.half * kHalf + q * 32;
    if (!kPartial || col0 < len) {
      tmem_ld_32x32b_x32(c.tS + q * 32, chunk);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) chunk[i] = 0xff800000u;
    }
  }
  tmem_ld_wait();
  tc_fence_before_sync();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(&c.bars->s_empty[c.x]);   // S may be recomputed for the next block
  if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 200 + j);
  if (kPartial) {
#pragma unroll
    for (int i = 0; i < kHalf; ++i)
      if (c.half * kHalf + i >= valid) s[i] = 0xff800000u;   // -inf -> P = 0
  }
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < kHalf; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[i]));
  const float m_loc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  float m_blk = m_loc;
  if (kSplit == 2) {
    // Row max = max over the two key halves: exchange with the partner thread (other warp).
    float* mb = c.max_buf + par * (2 * 2 * kBQ);
    mb[c.half * kBQ + c.r] = m_loc;
    pair_sync(c.pair_bar);
    m_blk = fmaxf(m_loc, mb[(c.half ^ 1) * kBQ + c.r]);
  }
  float m_old;
  const bool rescale = choose_reference_max(first, c.scale_log2e, m_blk, m_used, m_old);
  const float neg_m = -m_used * c.scale_log2e;
  if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 300 + j);
  const uint64_t scale2 = pack2(c.scale_log2e, c.scale_log2e), negm2 = pack2(neg_m, neg_m);
  float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  uint32_t pk[kHalf / 2];
#pragma unroll
  for (int i = 0; i < kHalf / 2; ++i) {
    const bool poly = !kPartial && ((i + 1) * kPoly64) / 64 != (i * kPoly64) / 64;
    pk[i] = poly ? exp_pair<true>(s[2 * i], s[2 * i + 1], scale2, negm2, sum[(2 * i) & 3], sum[(2 * i + 1) & 3])
                 : exp_pair<false>(s[2 * i], s[2 * i + 1], scale2, negm2, sum[(2 * i) & 3], sum[(2 * i + 1) & 3]);
  }
  if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 400 + j);
  const float alpha = rescale ? fast_exp2((m_old - m_used) * c.scale_log2e) : 1.0f;
  l_part = l_part * alpha + ((sum[0] + sum[1]) + (sum[2] + sum[3]));
  // PV(j-1) done: the P columns may be overwritten and O is stable (and can be rescaled).
  if (!first) {
    mbar_wait(&c.bars->pv_done[c.x], par ^ 1);
    tc_fence_after_sync();
    if (rescale) rescale_accumulator<kHD / kSplit>(c.tO, alpha);
  }
#pragma unroll
  for (int q = 0; q < kHalf / 32; ++q) {
    const uint32_t(&t)[16] = *reinterpret_cast<const uint32_t(*)[16]>(&pk[q * 16]);
    tmem_st_32x32b_x16(c.tP + q * 16, t);   // columns beyond `len` are written but never read
  }
  tmem_st_wait();
  if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 500 + j);
}

template <int kSplit, int kPolyA, int kPolyB>
__global__ void __launch_bounds__(attn_threads(kSplit), 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int N, int D,
                 int heads, int num_items, float scale_log2e, int flags, unsigned long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + AttnSmem::q_off;
  uint8_t* sK = smem + AttnSmem::k_off;
  uint8_t* sV = smem + AttnSmem::v_off;
  float* sMax = reinterpret_cast<float*>(smem + AttnSmem::max_off);
  float* sSum = reinterpret_cast<float*>(smem + AttnSmem::sum_off);
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + AttnSmem::bar_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kv = (N + kBKV - 1) / kBKV;
  const int pairs = (N + 2 * kBQ - 1) / (2 * kBQ);
  // Width of the last key block: only as many 16-key MMA steps as there are valid keys.
  const int last_valid = N - (num_kv - 1) * kBKV;
  const int last_len = (last_valid + 15) / 16 * 16;

  // Issue mode (flags bit 14 = 0, default): ONE ISSUER WARP PER QUERY TILE - warp 1 issues every MMA of tile A
  // (S_A(j+1), then PV_A(j)), warp 3 those of tile B - so that the two tiles' pipelines are independent and can run
  // out of phase (tile B starts `skew` cycles late, flags bits 8-13 x 128): while one softmax warp of a sub-partition
  // computes exponentials the other one sits in its serial TMEM-load / max / TMEM-store phases.  Bit 14 = 1 selects
  // the round-1 split (warp 1 = S of both tiles, warp 3 = PV of both tiles), whose in-order loops re-align the tiles
  // at every key block.  K / V / Q stages are released by both issuers in the per-tile mode (count 2).
  const bool per_tile_issue = (flags & 0x4000) == 0;
  const uint32_t release_count = per_tile_issue ? 2 : 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->q_full[s], 1);
      mbar_init(&bars->q_empty[s], release_count);
    }
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&bars->k_full[s], 1);
      mbar_init(&bars->k_empty[s], release_count);
      mbar_init(&bars->v_full[s], 1);
      mbar_init(&bars->v_empty[s], release_count);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bars->s_full[x], 1);
      mbar_init(&bars->s_empty[x], 4 * kSplit);   // one arrive per softmax warp of that tile
      mbar_init(&bars->p_full[x], 4 * kSplit);
      mbar_init(&bars->pv_done[x], 1);
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(&bars->tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp < 4) {
    // producer / issuer warps need few registers; the softmax rows take them.  The pool is the
    // launch allocation (threads x launch registers): 384 x 168 or 640 x 96.
    if (kSplit == 1) reg_dealloc<112>(); else reg_dealloc<64>();
    if (warp == 0 && lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      int dn = 0;
      // Q pair of item number `n` (n-th item of this CTA) goes to Q stage n & 1.
      auto load_q = [&](int it, int n) {
        const int pair = it % pairs;
        const int head = (it / pairs) % heads;
        const int img = it / (pairs * heads);
        const int qs = n & 1;
        mbar_wait(&bars->q_empty[qs], ((n >> 1) & 1) ^ 1);
        dbg_event(dbg, 0, dn, 100);
        mbar_arrive_expect_tx(&bars->q_full[qs], 2 * kQBytes);
        uint8_t* dst = sQ + qs * 2 * kQBytes;
        tma_load_2d(dst, &tmQKV, &bars->q_full[qs], head * kHD, img * N + pair * 2 * kBQ);
        tma_load_2d(dst + kQBytes, &tmQKV, &bars->q_full[qs], head * kHD, img * N + pair * 2 * kBQ + kBQ);
      };
      int n = 0;
      if (static_cast<int>(blockIdx.x) < num_items) load_q(blockIdx.x, 0);
      for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
        const int head = (it / pairs) % heads;
        const int img = it / (pairs * heads);
        const int row_base = img * N;
        // The next item's Q pair is requested after three K/V blocks of this item: the K/V ring has three
        // stages, so by then the last S MMA of the previous item - the one that frees that Q stage - has
        // retired and the request never stalls the K/V stream.
        const int q_prefetch_at = num_kv < 3 ? num_kv - 1 : 2;
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
          if (j == q_prefetch_at && it + static_cast<int>(gridDim.x) < num_items) load_q(it + gridDim.x, n + 1);
        }
      }
    } else if (per_tile_issue && (warp == 1 || warp == 3)) {
      // ===== MMA issuer of ONE query tile: S_x(0), then per key block { S_x(j+1) ; O_x += P_x(j) V_j } =====
      // (whole warp runs the loop with warp-uniform operands, one elected lane issues - see the note below)
      const int x = (warp == 1) ? 0 : 1;
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + x * kQBytes));
      const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK));
      const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV));
      constexpr uint32_t idesc_full = make_idesc_f16(kBQ, kBKV, 0, 0);
      const uint32_t idesc_last = make_idesc_f16(kBQ, last_len, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(kBQ, kHD, 0, 1);   // B (=V) is MN-major
      const uint32_t tS = tmem_base + kColS + x * kBKV;
      const uint32_t tO = tmem_base + kColO + x * kHD;
      const uint32_t tP = tmem_base + kColP + x * (kBKV / 2);   // 8 columns per 16 keys
      int kstage = 0, vstage = 0;
      uint32_t kphase = 0, vphase = 0;
      uint32_t blk = 0;   // running count of key blocks processed by this CTA (barrier parity)
      int n = 0;          // items done by this CTA: Q stage n & 1, phase (n >> 1) & 1
      if (x == 1) {       // start tile B out of phase with tile A
        // default (no skew bits set): 1024 cycles, about half a key block of one tile
        const long long skew = static_cast<long long>(((flags >> 8) & 0x3f) ? ((flags >> 8) & 0x3f) : 8) * 128;
        const long long t0 = clock64();
        while (clock64() - t0 < skew) {}
      }
      for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
        const int qs = n & 1;
        mbar_wait(&bars->q_full[qs], (n >> 1) & 1);
        const uint64_t qd = qdesc + static_cast<uint64_t>(qs * (2 * kQBytes / 16));
        auto issue_s = [&](int j) {
          const uint32_t idesc_s = (j == num_kv - 1) ? idesc_last : idesc_full;
          mbar_wait(&bars->k_full[kstage], kphase);
          mbar_wait(&bars->s_empty[x], ((blk + j) & 1) ^ 1);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint64_t kd = kdesc0 + static_cast<uint64_t>(kstage * (kKBytes / 16));
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) umma_f16_ss(tS, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
            umma_commit(&bars->s_full[x]);
            umma_commit(&bars->k_empty[kstage]);
            if (j == num_kv - 1) umma_commit(&bars->q_empty[qs]);   // this Q stage may be overwritten
          }
          __syncwarp();
          if (++kstage == kKVStages) { kstage = 0; kphase ^= 1; }
        };
        issue_s(0);
        for (int j = 0; j < num_kv; ++j) {
          if (j + 1 < num_kv) issue_s(j + 1);
          const uint32_t acc0 = j != 0;          // the first block of an item overwrites O
          const int ksteps = (j == num_kv - 1) ? last_len / 16 : kBKV / 16;
          mbar_wait(&bars->v_full[vstage], vphase);
          mbar_wait(&bars->p_full[x], (blk + j) & 1);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint64_t vd = vdesc0 + static_cast<uint64_t>(vstage * (kVBytes / 16));
            if (ksteps == kBKV / 16) {
#pragma unroll
              for (int k = 0; k < kBKV / 16; ++k)
                umma_f16_ts(tO, tP + k * 8, vd + k * 128, idesc_pv, acc0 | (k != 0));
            } else {
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k)
                umma_f16_ts(tO, tP + k * 8, vd + k * 128, idesc_pv, acc0 | (k != 0));
            }
            umma_commit(&bars->pv_done[x]);
            umma_commit(&bars->v_empty[vstage]);
          }
          __syncwarp();
          if (++vstage == kKVStages) { vstage = 0; vphase ^= 1; }
        }
        blk += num_kv;
      }
    } else if (warp == 1) {
      // ===== MMA issuer 1: S_X(j) = Q_X K_j^T for both query tiles =====
      // The WHOLE warp runs this loop with warp-uniform operands and one elected lane issues the
      // MMAs: issuing tcgen05.mma from a divergent `lane == 0` branch makes the compiler wrap every
      // instruction in a waterfall loop (R2UR.BROADCAST / BRA.U.ANY), ~90 cycles per MMA, which
      // serialised the small MMAs of every key block (profiles/r01_attention_timeline.md).
      const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(sQ));
      const uint64_t qdesc1 = make_smem_desc_sw128(smem_u32(sQ + kQBytes));
      const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK));
      constexpr uint32_t idesc_full = make_idesc_f16(kBQ, kBKV, 0, 0);
      const uint32_t idesc_last = make_idesc_f16(kBQ, last_len, 0, 0);
      int kstage = 0;
      uint32_t kphase = 0;
      uint32_t blk = 0;   // running count of key blocks processed by this CTA (barrier parity)
      int dn = 0;
      int n = 0;          // items done by this CTA: Q stage n & 1, phase (n >> 1) & 1
      for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
        const int qs = n & 1;
        mbar_wait(&bars->q_full[qs], (n >> 1) & 1);
        if (lane == 0) dbg_event(dbg, 1, dn, 100);
        const uint64_t qoff = static_cast<uint64_t>(qs * (2 * kQBytes / 16));
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
              const uint64_t qd = (x == 0 ? qdesc0 : qdesc1) + qoff;
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
            if (j == num_kv - 1) umma_commit(&bars->q_empty[qs]);   // this Q stage may be overwritten
          }
          __syncwarp();
          if (++kstage == kKVStages) { kstage = 0; kphase ^= 1; }
        }
        blk += num_kv;
      }
    } else if (warp == 3) {
      // ===== MMA issuer 2: O_X(j) += P_X(j) V_j with A = P read from tensor memory =====
      constexpr uint32_t idesc_pv = make_idesc_f16(kBQ, kHD, 0, 1);   // B (=V) is MN-major
      // B = V: MN-major, 16 key rows of 128 B (= +128 in the address field) per K step.
      const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV));
      int vstage = 0;
      uint32_t vphase = 0;
      uint32_t blk = 0;
      int dn = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        for (int j = 0; j < num_kv; ++j) {
          const uint32_t par = (blk + j) & 1;
          const uint32_t acc0 = j != 0;          // the first block of an item overwrites O
          const int ksteps = (j == num_kv - 1) ? last_len / 16 : kBKV / 16;
          mbar_wait(&bars->v_full[vstage], vphase);
          if (lane == 0) dbg_event(dbg, 2, dn, 10 + j);
          const uint64_t vd = vdesc0 + static_cast<uint64_t>(vstage * (kVBytes / 16));
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            mbar_wait(&bars->p_full[x], par);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t tO = tmem_base + kColO + x * kHD;
              const uint32_t tP = tmem_base + kColP + x * (kBKV / 2);   // 8 columns per 16 keys
              if (ksteps == kBKV / 16) {
#pragma unroll
                for (int k = 0; k < kBKV / 16; ++k)
                  umma_f16_ts(tO, tP + k * 8, vd + k * 128, idesc_pv, acc0 | (k != 0));
              } else {
#pragma unroll 1
                for (int k = 0; k < ksteps; ++k)
                  umma_f16_ts(tO, tP + k * 8, vd + k * 128, idesc_pv, acc0 | (k != 0));
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
    // ===== softmax warps =====
    if (kSplit == 1) reg_alloc<192>(); else reg_alloc<104>();   // 128x112 + 256x192 | 128x64 + 512x104
    constexpr int kHalf = kBKV / kSplit;
    const int sw = warp - 4;
    const int sub = warp & 3;                           // TMEM lane quarter of this warp
    SoftmaxCtx c;
    c.bars = bars;
    c.x = sw / (4 * kSplit);                            // query tile (0 = A, 1 = B)
    c.half = (sw >> 2) & (kSplit - 1);                  // key half inside every block
    c.r = sub * 32 + lane;                              // query row inside the tile
    c.lane = lane;
    c.pair_bar = 1 + c.x * 4 + sub;                     // named barriers 1..8
    c.max_buf = sMax + c.x * (2 * kBQ);
    c.scale_log2e = scale_log2e;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    c.tS = tmem_base + lane_addr + kColS + c.x * kBKV + c.half * kHalf;
    c.tO = tmem_base + lane_addr + kColO + c.x * kHD + c.half * (kHD / kSplit);
    c.tP = tmem_base + lane_addr + kColP + c.x * (kBKV / 2) + c.half * (kHalf / 2);
    float* sum_buf = sSum + c.x * (2 * kBQ);
    uint32_t blk = 0;
    int dn = 0;
    const bool dbg_me = (sub == 0 && lane == 0 && c.half == 0);
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int pair = it % pairs;
      const int head = (it / pairs) % heads;
      const int img = it / (pairs * heads);
      const int row_base = img * N;
      float m_used = -INFINITY;
      float l_part = 0.0f;
      for (int j = 0; j < num_kv; ++j) {
        const uint32_t par = (blk + j) & 1;
        const int valid = N - j * kBKV;            // keys >= valid are out of range
        const int len = (j == num_kv - 1) ? last_len : kBKV;
        mbar_wait(&bars->s_full[c.x], par);
        if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 10 + j);
        tc_fence_after_sync();
        if (len == kBKV && valid >= kBKV) {
          if (kPolyA == kPolyB || c.x == 0) softmax_block<kSplit, kPolyA, false>(c, len, valid, j == 0, par, m_used, l_part, dbg, dn, dbg_me, j);
          else softmax_block<kSplit, kPolyB, false>(c, len, valid, j == 0, par, m_used, l_part, dbg, dn, dbg_me, j);
        } else {
          softmax_block<kSplit, 0, true>(c, len, valid, j == 0, par, m_used, l_part, dbg, dn, dbg_me, j);
        }
        tc_fence_before_sync();     // P (tcgen05.st, waited) -> visible to the PV MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[c.x]);
        if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 30 + j);
      }
      float l_run = l_part;
      if (kSplit == 2) {
        // Join the two partial row sums of the row while the last PV retires.
        sum_buf[c.half * kBQ + c.r] = l_part;
        pair_sync(c.pair_bar);
        l_run = l_part + sum_buf[(c.half ^ 1) * kBQ + c.r];
      }
      // Drain: O is complete once PV of the last block has retired.
      mbar_wait(&bars->pv_done[c.x], (blk + num_kv - 1) & 1);
      if (dbg_me) dbg_event(dbg, 3 + c.x, dn, 90);
      tc_fence_after_sync();
      constexpr int kOC = kHD / kSplit;   // output columns of this thread
      float o[kOC];
#pragma unroll
      for (int q = 0; q < kOC / 16; ++q) {
        uint32_t(&t)[16] = *reinterpret_cast<uint32_t(*)[16]>(&o[q * 16]);
        tmem_ld_32x32b_x16(c.tO + q * 16, t);
      }
      tmem_ld_wait();
      tc_fence_before_sync();
      blk += num_kv;

      const int q = pair * 2 * kBQ + c.x * kBQ + c.r;
      if (q < N) {
        const float inv_l = 1.0f / l_run;
        __half* dst = out + static_cast<size_t>(row_base + q) * D + head * kHD + c.half * kOC;
#pragma unroll
        for (int g = 0; g < kOC / 8; ++g) {
          __half2 h0 = __floats2half2_rn(o[g * 8 + 0] * inv_l, o[g * 8 + 1] * inv_l);
          __half2 h1 = __floats2half2_rn(o[g * 8 + 2] * inv_l, o[g * 8 + 3] * inv_l);
          __half2 h2 = __floats2half2_rn(o[g * 8 + 4] * inv_l, o[g * 8 + 5] * inv_l);
          __half2 h3 = __floats2half2_rn(o[g * 8 + 6] * inv_l, o[g * 8 + 7] * inv_l);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst + g * 8) = pk;
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
  const int pairs = (N + 2 * kBQ - 1) / (2 * kBQ);
  const int num_items = pairs * heads * B;
  const int grid = num_items < num_sms() ? num_items : num_sms();
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // hd^-0.5 * log2(e), hd = 64
  ProfScope prof(PROF_ATTENTION, stream, 4.0 * B * heads * static_cast<double>(N) * N * kHD);
  // Kernel variants: (threads per row, share of the exponentials of tile A / tile B evaluated on
  // the FMA pipe, in pairs out of 64 per row and key block).  g_attn_flags (tools/attn_bench.py)
  // overrides the production choice: bits 0-7 = variant + 1.
  static const int env_flags = []() {   // FOUNDPOSE_ATTN_FLAGS: same bits as attention_set_flags, for A/B runs of bench.py
    const char* e = getenv("FOUNDPOSE_ATTN_FLAGS");
    return e ? static_cast<int>(strtol(e, nullptr, 0)) : 0;
  }();
  if (g_attn_flags == 0 && env_flags != 0) g_attn_flags = env_flags;
  const int variant = (g_attn_flags & 0xff) ? (g_attn_flags & 0xff) - 1 : kDefaultVariant;
#define FP_ATTN_LAUNCH(IDX, SPLIT, PA, PB)                                                              \
  case IDX: {                                                                                           \
    static bool configured[64] = {};                                                                    \
    if (per_device_once(configured)) {                                                                  \
      FP_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<SPLIT, PA, PB>,                               \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::total)); \
    }                                                                                                   \
    attention_kernel<SPLIT, PA, PB><<<grid, attn_threads(SPLIT), AttnSmem::total, stream>>>(            \
        tm, out, N, D, heads, num_items, scale_log2e, g_attn_flags, g_attn_dbg);                        \
    break;                                                                                              \
  }
  switch (variant) {
    FP_ATTN_LAUNCH(0, 1, 24, 24)   // production: measured best on B200 (profiles/r01_attention_variants.md)
    FP_ATTN_LAUNCH(1, 1, 0, 0)
    FP_ATTN_LAUNCH(2, 1, 16, 16)
    FP_ATTN_LAUNCH(3, 1, 32, 32)
    FP_ATTN_LAUNCH(4, 2, 0, 0)
    FP_ATTN_LAUNCH(5, 2, 16, 16)
    FP_ATTN_LAUNCH(6, 2, 8, 24)
    default:
      FP_REQUIRE(false, "attention: unknown kernel variant %d", variant);
  }
#undef FP_ATTN_LAUNCH
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
