// Shared device-side primitives for the sm_100a kernels of foundpose_b200:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM) wrappers, and the
// UMMA shared-memory / instruction descriptor builders.
//
// Everything here is inline PTX for sm_100a; there is no fallback path.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace fp {

// SM count of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a B200).
// Persistent grids and grid caps are sized from it.
int num_sms();

// ----------------------------------------------------------------------------
// Error plumbing shared by all translation units (defined in api.cu).
// ----------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define FP_CUDA_CHECK(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      fp::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                         cudaGetErrorString(_e));                                  \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

#define FP_REQUIRE(cond, ...)                                                      \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      fp::set_last_error(__VA_ARGS__);                                             \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

// ----------------------------------------------------------------------------
// Launch accounting / optional per-kernel timing (defined in api.cu).  Every kernel launcher of
// the library opens one scope: it always bumps the launch counter (bench.py reports it as
// `gpu_launches`) and, when profiling is enabled, brackets the launch with CUDA events on the
// launching stream so bench.py can attribute device time and algorithmic work per kernel family.
// ----------------------------------------------------------------------------
enum ProfCategory : int {
  PROF_GEMM = 0,        // gemm_tn_kernel          (work = flops)
  PROF_ATTENTION = 1,   // attention_kernel        (work = flops)
  PROF_LAYERNORM = 2,   // layernorm / final norm  (work = bytes)
  PROF_VIT_MISC = 3,    // patchify, special tokens, facet gather (work = bytes)
  PROF_KNN = 4,         // knn_kernel              (work = flops, upper bound for device-built items)
  PROF_FEATURE = 5,     // mask filter, sampling, conversions, norms (work = bytes)
  PROF_RETRIEVAL = 6,   // tf-idf, cosine scores, top-k, items, cyclic buddies (work = bytes)
  PROF_KNN_PAIR = 7,    // knn_pair_kernel: tensor-bound full-bank search (work = 0, the caller knows the flops)
  PROF_NUM_CATEGORIES = 8,
};

struct ProfScope {
  ProfScope(int category, cudaStream_t stream, double work);
  ~ProfScope();
  int category_;
  cudaStream_t stream_;
  void* rec_;
};

// ----------------------------------------------------------------------------
// Small helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Make generic-proxy smem writes visible to the async proxy (TMA / UMMA reads).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware
      : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must trap (and surface as a CUDA error) rather than hang
// the device. 2^28 polls is several seconds; any legitimate wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("foundpose_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x,
             threadIdx.x);
      __trap();
    }
  }
}

// Latency-critical wait (softmax <-> MMA hand-offs): plain try_wait polling, no suspend-time hint
// (the hinted form compiles to TRYWAIT + NANOSLEEP.SYNCS, whose wake-up is slower).
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 28)) {
      printf("foundpose_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  } while (!ok);
}

// ----------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// 2D tiled load: c0 = coordinate along the innermost (contiguous) dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], fp16/bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M x K, fp16 packed two per 32-bit column, lane =
// row) is read from tensor memory, e.g. softmax probabilities written with tcgen05.st.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread retire.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 32b, 32 consecutive columns: thread i of the warp receives lane
// (subpartition_base + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------
// cta_group::2 (a cluster of two CTAs on one TPC sharing one MMA): cluster helpers, TMEM allocation,
// TMA loads that complete on the leader's barrier, the paired MMA and its multicast commit.
// ----------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the leader's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// TMA load into this CTA's smem, completing on the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all prior MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// Arrive on the leader CTA's copy of a barrier (local arrive when executed by the leader).
// Default semantics (.release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id) issues it: what the
// consumer orders against is tcgen05 state (guarded by tcgen05.fence::before_thread_sync in front of the arrive),
// not generic-proxy memory.  The `.release.cluster` form compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in
// front of the arrive - 20% of the k-NN pair kernel's stall samples (profiles/r02_k4_pair_ncu.md).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// ----------------------------------------------------------------------------
// UMMA descriptors (layout per the PTX ISA "tcgen05 matrix descriptor" tables)
// ----------------------------------------------------------------------------
// Shared-memory operand descriptor for a 128B-swizzled tile whose rows are 128 bytes
// (64 x 16-bit) and whose 8-row groups are 1024 bytes apart.
//   K-major  operand: rows = M/N index, the 128 bytes run along K.
//   MN-major operand: rows = K index,   the 128 bytes run along M/N.
// Both use SBO = 1024 B (distance between 8-row groups); LBO is unused for a single
// 128-byte-wide swizzle atom and is set to 1 as CUTLASS does.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 16,
                                                         uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);            // [0,14) start address
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;       // [16,30) leading byte off
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;       // [32,46) stride byte off
  d |= static_cast<uint64_t>(1) << 46;                               // [46,48) version = 1
  d |= static_cast<uint64_t>(2) << 61;                               // [61,64) SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: fp16 A/B (format 0), fp32 accumulate.
// a_mn_major / b_mn_major: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, int a_mn_major = 0,
                                                      int b_mn_major = 0) {
  return (1u << 4)                                   // c_format = F32
         | (0u << 7)                                 // a_format = F16
         | (0u << 10)                                // b_format = F16
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(n >> 3) << 17)     // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);    // m_dim
}

// ----------------------------------------------------------------------------
// Host-side TMA descriptor creation (driver entry point resolved at run time so the
// library does not link against libcuda).
// ----------------------------------------------------------------------------
// 2D fp16 row-major tensor [rows, cols] with leading dimension ld (elements), box
// [box_rows, 64 cols], 128B swizzle. Returns 0 on success.
int make_tma_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols = 64);

int make_tma_2d_f32_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows);

// True the first time it is called for (flag array, current device): kernel attributes such as
// cudaFuncAttributeMaxDynamicSharedMemorySize are per device, so "configured once" must be too.
inline bool per_device_once(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

}  // namespace fp
