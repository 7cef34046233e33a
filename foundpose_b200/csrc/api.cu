// extern "C" entry points of libfoundpose_b200.so (declared in include/foundpose_b200.h).
// Plain pointers and sizes only: no torch types cross this boundary.
#include "../../include/foundpose_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace fp {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace fp

extern "C" {

const char* fp_last_error(void) { return fp::g_last_error; }

int fp_version(void) { return FP_B200_VERSION; }

int fp_gemm_tn_f16(int epilogue, const void* A, int lda, const void* B, int ldb, int M, int N,
                   int K, const float* bias, const float* gamma, void* out_f16, int ld_f16,
                   float* out_f32, int ld_f32, void* stream) {
  fp::GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.gamma = gamma;
  p.out_f16 = static_cast<__half*>(out_f16); p.ld_f16 = ld_f16;
  p.out_f32 = out_f32; p.ld_f32 = ld_f32;
  if (epilogue == fp::EPI_PATCH_F32) {
    fp::set_last_error("fp_gemm_tn_f16: the patch-embed epilogue is internal to fp_vit_forward");
    return 1;
  }
  return fp::gemm_tn(epilogue, static_cast<const __half*>(A), lda, static_cast<const __half*>(B),
                     ldb, p, static_cast<cudaStream_t>(stream));
}

int fp_umma_probe(const void* A, const void* B, float* out, int b_mn_major, void* stream) {
  return fp::umma_probe(static_cast<const __half*>(A), static_cast<const __half*>(B), out,
                        b_mn_major, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
