/* foundpose_b200 — C ABI of the B200-native FoundPose per-crop hot path.
 *
 * The reference (facebookresearch/foundpose) has no FFI layer: its "plugin interface" for this
 * path is the Python API of utils/{dinov2_utils,feature_util,projector_util,knn_util,
 * template_util,corresp_util}.py, which bottoms out in torch ATen, faiss and scikit-learn calls.
 * Each entry point below replaces one of those third-party calls and cites the reference call
 * site (file:line relative to the reference root) whose arithmetic it reproduces.  The Python
 * mirror of the reference modules (the foundpose_b200/utils package) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success; non-zero = failure, message via fp_last_error()
 *     (1 = bad argument, 2 = CUDA runtime error, 3 = driver/TMA descriptor error)
 *   - all data pointers are DEVICE pointers (tensor.data_ptr()), row-major, caller-owned and
 *     caller-allocated; nothing is allocated inside except per-handle workspaces
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream)
 *   - "f16" buffers are IEEE binary16; index outputs are int64 to match faiss / torch.topk
 */
#ifndef FOUNDPOSE_B200_H_
#define FOUNDPOSE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP_B200_VERSION 100

/* Last error message of the calling thread (never NULL). */
const char* fp_last_error(void);
/* Library version (FP_B200_VERSION). */
int fp_version(void);

/* ---- dense contraction on the 5th-gen tensor cores -------------------------------------
 * C[M,N] = A[M,K] . B[N,K]^T, fp16 in, fp32 accumulate (tcgen05.mma + TMA + TMEM).
 * Replaces torch.nn.functional.linear at external/dinov2/dinov2/layers/attention.py:58,67,
 * layers/mlp.py:35,38 and numpy sgemm inside sklearn PCA.transform (utils/projector_util.py:67).
 * epilogue: 0 out_f16 = acc+bias | 1 out_f16 = gelu_erf(acc+bias) |
 *           2 out_f32 += gamma*(acc+bias) | 4 out_f32 = acc+bias (and out_f16 copy if non-NULL)
 * N % 128 == 0, K % 64 == 0, lda/ldb % 8 == 0. */
int fp_gemm_tn_f16(int epilogue, const void* A, int lda, const void* B, int ldb, int M, int N,
                   int K, const float* bias, const float* gamma, void* out_f16, int ld_f16,
                   float* out_f32, int ld_f32, void* stream);

/* Single-tile UMMA descriptor probe used by the GPU tests (not on the hot path). */
int fp_umma_probe(const void* A, const void* B, float* out, int b_mn_major, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* FOUNDPOSE_B200_H_ */
