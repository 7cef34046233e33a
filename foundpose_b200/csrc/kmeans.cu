// Centroid update of k-means for the offline bank build: per-cluster sums of the
// assigned samples and the mean.  Reference: utils/cluster_util.py:13-68 -> faiss.Kmeans.train
// (faiss 1.8.0 Clustering.cpp: compute_centroids), called from scripts/gen_repre.py:289-300.
// The assignment step of every iteration is the visual-word search K1 (knn_tcgen05.cu).
//
// Sums are accumulated in 64-bit fixed point (value * 2^24, round-half-even): integer addition is
// associative, so the result does not depend on the order in which the atomics land - the update
// is bit-reproducible from run to run and equal to oracle/cluster.py, which no floating-point
// atomic scheme can offer.  |x| < 2^15 and < 2^24 samples per cluster keep the sums inside int64.
//
// HBM-bound: n*d*4 bytes of samples read once + n*8 bytes of assignments; k*d 64-bit atomics per
// ~n/k samples resolve in L2.
#include "common.cuh"
#include "kernels.h"

namespace fp {
namespace {

constexpr double kFixedScale = 16777216.0;   // 2^24

__global__ void __launch_bounds__(256)
kmeans_accumulate_kernel(const float* __restrict__ x, const int64_t* __restrict__ assign, long long n, int d,
                         int k, unsigned long long* __restrict__ sums, int* __restrict__ counts) {
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const long long c = assign[r];
    if (c < 0 || c >= k) continue;   // unassigned rows do not contribute
    if (threadIdx.x == 0) atomicAdd(&counts[c], 1);
    const float* row = x + r * d;
    unsigned long long* dst = sums + c * d;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      const long long q = __double2ll_rn(static_cast<double>(row[j]) * kFixedScale);
      atomicAdd(&dst[j], static_cast<unsigned long long>(q));
    }
  }
}

__global__ void kmeans_finalize_kernel(const unsigned long long* __restrict__ sums, const int* __restrict__ counts,
                                       int k, int d, float* __restrict__ centroids) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(k) * d) return;
  const int c = static_cast<int>(i / d);
  const int cnt = counts[c];
  // empty clusters stay at zero (faiss compute_centroids); split_clusters re-seeds them on the host
  centroids[i] = cnt == 0 ? 0.0f
                          : static_cast<float>(static_cast<double>(static_cast<long long>(sums[i])) /
                                               (kFixedScale * static_cast<double>(cnt)));
}

}  // namespace

int kmeans_update(const float* x, const int64_t* assign, long long n, int d, int k, unsigned long long* sums,
                  int* counts, float* centroids, cudaStream_t stream) {
  FP_REQUIRE(n >= 0 && d > 0 && k > 0, "kmeans_update: bad sizes n=%lld d=%d k=%d", n, d, k);
  FP_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * static_cast<size_t>(k) * d, stream));
  FP_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int) * static_cast<size_t>(k), stream));
  if (n > 0) {
    const long long want = n < static_cast<long long>(num_sms()) * 16 ? n : static_cast<long long>(num_sms()) * 16;
    ProfScope prof(PROF_FEATURE, stream, static_cast<double>(n) * d * 4.0);
    kmeans_accumulate_kernel<<<static_cast<unsigned>(want), 256, 0, stream>>>(x, assign, n, d, k, sums, counts);
    FP_CUDA_CHECK(cudaGetLastError());
  }
  const long long total = static_cast<long long>(k) * d;
  ProfScope prof(PROF_FEATURE, stream, static_cast<double>(total) * 12.0);
  kmeans_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(sums, counts, k, d, centroids);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
