// Batched coarse pose from 2D-3D correspondences: RANSAC over P3P hypotheses + Levenberg-Marquardt
// on the inliers, one CTA per (crop, template) problem.  SURVEY.md S8f N2.
//
// Replaces the per-template host loop of the reference: scripts/infer.py:551-577 ->
// utils/pnp_util.py:42-72 -> cv2.solvePnPRansac(flags=SOLVEPNP_ITERATIVE) + cv2.solvePnPRefineLM.
// The arithmetic of that step lives in OpenCV; oracle/pnp.py states which parts of its published
// algorithm are kept (sequential RANSAC semantics incl. RANSACUpdateNumIters, inlier test
// err^2 <= thresh^2, strict improvement, non-linear least squares on the inliers) and which are
// made explicit because cv::RNG cannot be reproduced outside OpenCV (counter-based splitmix64
// sampling of 4 distinct correspondences, P3P + 4th-point disambiguation as the minimal solver).
// This file follows oracle/pnp.py function by function, in float64.
//
// Parallel shape: the hypotheses of a problem are independent, so thread t scores hypotheses
// t, t+256, ...; their inlier counts land in shared memory and ONE sequential scan over the
// counts reproduces the order-dependent part of RANSAC (best-so-far + shrinking iteration budget)
// exactly.  The LM normal equations (21 + 6 + 1 sums over <= M points) are reduced with warp
// shuffles and a fixed-order sum over the 8 warps, so every thread holds identical values and the
// 6x6 solve is done redundantly without a broadcast.  The work is O(iters * M) flops per problem
// on <= 12 KB of correspondences: latency-bound, no HBM or tensor roofline applies.
#include "common.cuh"
#include "kernels.h"

namespace fp {
namespace {

constexpr int kPnpThreads = 256;
constexpr int kPnpWarps = kPnpThreads / 32;
constexpr int kModelPoints = 4;
constexpr int kMinInliersForPose = 6;
constexpr int kLmMaxIters = 20;
constexpr int kNormalTerms = 28;   // 21 (upper H) + 6 (g) + 1 (cost)

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ void sample_indices(unsigned long long seed, int problem, int hyp, int n, int (&ids)[kModelPoints]) {
  const unsigned long long base = seed ^ (static_cast<unsigned long long>(problem) * 0xD1B54A32D192ED03ull) ^
                                  (static_cast<unsigned long long>(hyp) * 0x8CB92BA72F3D8DD7ull);
  int have = 0;
  unsigned long long ctr = 0;
  while (have < kModelPoints) {
    const int idx = static_cast<int>(splitmix64(base + ctr * 0x2545F4914F6CDD1Dull) % static_cast<unsigned long long>(n));
    ++ctr;
    bool dup = false;
    for (int k = 0; k < have; ++k) dup |= (ids[k] == idx);
    if (!dup) ids[have++] = idx;
  }
}

// Real roots of c4 x^4 + c3 x^3 + c2 x^2 + c1 x + c0 (Ferrari via the resolvent cubic + 2 Newton steps).
__device__ int solve_quartic_real(double c4, double c3, double c2, double c1, double c0, double (&out)[4]) {
  if (c4 == 0.0 || !isfinite(c4)) return 0;
  const double b = c3 / c4, c = c2 / c4, d = c1 / c4, e = c0 / c4;
  const double p = c - 3.0 * b * b / 8.0;
  const double q = d - b * c / 2.0 + b * b * b / 8.0;
  const double r = e - b * d / 4.0 + b * b * c / 16.0 - 3.0 * b * b * b * b / 256.0;
  const double a2 = 2.0 * p, a1 = p * p - 4.0 * r, a0 = -q * q;
  const double Q = (3.0 * a1 - a2 * a2) / 9.0;
  const double R = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
  const double D = Q * Q * Q + R * R;
  double z;
  if (D >= 0.0) {
    const double sd = sqrt(D);
    z = cbrt(R + sd) + cbrt(R - sd) - a2 / 3.0;
  } else {
    const double th = acos(fmax(-1.0, fmin(1.0, R / sqrt(-Q * Q * Q))));
    z = 2.0 * sqrt(-Q) * cos(th / 3.0) - a2 / 3.0;
  }
  int n = 0;
  if (z > 1e-300) {
    const double s = sqrt(z);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double sign = k == 0 ? 1.0 : -1.0;
      const double bb = sign * s;
      const double cc = 0.5 * (p + z - sign * q / s);
      const double disc = bb * bb - 4.0 * cc;
      if (disc >= 0.0) {
        const double sq = sqrt(disc);
        out[n++] = 0.5 * (-bb + sq) - b / 4.0;
        out[n++] = 0.5 * (-bb - sq) - b / 4.0;
      }
    }
  } else {
    const double disc = p * p - 4.0 * r;
    if (disc >= 0.0) {
      const double sq = sqrt(disc);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double y2 = k == 0 ? 0.5 * (-p + sq) : 0.5 * (-p - sq);
        if (y2 >= 0.0 && n <= 2) {
          const double y = sqrt(y2);
          out[n++] = y - b / 4.0;
          out[n++] = -y - b / 4.0;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double x = out[i];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const double f = (((c4 * x + c3) * x + c2) * x + c1) * x + c0;
      const double df = ((4.0 * c4 * x + 3.0 * c3) * x + 2.0 * c2) * x + c1;
      if (df != 0.0) x = x - f / df;
    }
    out[i] = x;
  }
  return n;
}

struct Vec3 { double x, y, z; };
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Pose { double R[9]; double t[3]; };   // row-major R

// Orthonormal frame of a triangle, columns (e1, e2, e3) stored as F[row*3 + col].
__device__ bool tri_frame(Vec3 p1, Vec3 p2, Vec3 p3, double (&F)[9]) {
  Vec3 e1 = p2 - p1;
  const double n1 = sqrt(dot(e1, e1));
  if (n1 == 0.0) return false;
  e1 = (1.0 / n1) * e1;
  Vec3 e3 = cross(e1, p3 - p1);
  const double n3 = sqrt(dot(e3, e3));
  if (n3 == 0.0) return false;
  e3 = (1.0 / n3) * e3;
  const Vec3 e2 = cross(e3, e1);
  F[0] = e1.x; F[1] = e2.x; F[2] = e3.x;
  F[3] = e1.y; F[4] = e2.y; F[5] = e3.y;
  F[6] = e1.z; F[7] = e2.z; F[8] = e3.z;
  return true;
}

__device__ __forceinline__ double reproj_err2(const Pose& P, const double* K4, Vec3 X, double u0, double v0) {
  const double xc = P.R[0] * X.x + P.R[1] * X.y + P.R[2] * X.z + P.t[0];
  const double yc = P.R[3] * X.x + P.R[4] * X.y + P.R[5] * X.z + P.t[1];
  const double zc = P.R[6] * X.x + P.R[7] * X.y + P.R[8] * X.z + P.t[2];
  const double u = K4[0] * xc / zc + K4[2];
  const double v = K4[1] * yc / zc + K4[3];
  const double e = (u - u0) * (u - u0) + (v - v0) * (v - v0);
  return isfinite(e) ? e : INFINITY;
}

// Pose of hypothesis `hyp`: P3P (Grunert) on the first three sampled correspondences, the solution
// that best reprojects the fourth.  False when no solution exists.
__device__ bool hypothesis_pose(unsigned long long seed, int problem, int hyp, int n, const double* sX,
                                const double* sx, const double* K4, Pose& best) {
  int ids[kModelPoints];
  sample_indices(seed, problem, hyp, n, ids);
  Vec3 X[3], f[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    X[k] = {sX[ids[k] * 3 + 0], sX[ids[k] * 3 + 1], sX[ids[k] * 3 + 2]};
    Vec3 ray = {(sx[ids[k] * 2 + 0] - K4[2]) / K4[0], (sx[ids[k] * 2 + 1] - K4[3]) / K4[1], 1.0};
    f[k] = (1.0 / sqrt(dot(ray, ray))) * ray;
  }
  const double a2 = dot(X[1] - X[2], X[1] - X[2]);
  const double b2 = dot(X[0] - X[2], X[0] - X[2]);
  const double c2 = dot(X[0] - X[1], X[0] - X[1]);
  if (a2 == 0.0 || b2 == 0.0 || c2 == 0.0) return false;
  const double ca = dot(f[1], f[2]), cb = dot(f[0], f[2]), cg = dot(f[0], f[1]);
  const double k1 = (a2 - c2) / b2;
  const double k2 = (a2 + c2) / b2;
  const double A4 = (k1 - 1.0) * (k1 - 1.0) - 4.0 * c2 / b2 * ca * ca;
  const double A3 = 4.0 * (k1 * (1.0 - k1) * cb - (1.0 - k2) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
  const double A2 = 2.0 * (k1 * k1 - 1.0 + 2.0 * k1 * k1 * cb * cb + 2.0 * (b2 - c2) / b2 * ca * ca -
                           4.0 * k2 * ca * cb * cg + 2.0 * (b2 - a2) / b2 * cg * cg);
  const double A1 = 4.0 * (-k1 * (1.0 + k1) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - k2) * ca * cg);
  const double A0 = (1.0 + k1) * (1.0 + k1) - 4.0 * a2 / b2 * cg * cg;
  double Fw[9];
  if (!tri_frame(X[0], X[1], X[2], Fw)) return false;
  double roots[4];
  const int nroots = solve_quartic_real(A4, A3, A2, A1, A0, roots);
  const Vec3 X4 = {sX[ids[3] * 3 + 0], sX[ids[3] * 3 + 1], sX[ids[3] * 3 + 2]};
  const double u4 = sx[ids[3] * 2 + 0], v4 = sx[ids[3] * 2 + 1];
  bool found = false;
  double best_e = INFINITY;
  for (int i = 0; i < nroots; ++i) {
    const double v = roots[i];
    if (!(v > 0.0) || !isfinite(v)) continue;
    const double den = 2.0 * (cg - v * ca);
    if (den == 0.0) continue;
    const double u = ((k1 - 1.0) * v * v - 2.0 * k1 * cb * v + 1.0 + k1) / den;
    if (!(u > 0.0)) continue;
    const double d1 = 1.0 + v * v - 2.0 * v * cb;
    if (!(d1 > 0.0)) continue;
    const double s1 = sqrt(b2 / d1);
    const Vec3 P1 = s1 * f[0], P2 = (u * s1) * f[1], P3 = (v * s1) * f[2];
    double Fc[9];
    if (!tri_frame(P1, P2, P3, Fc)) continue;
    Pose cand;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        cand.R[r * 3 + c] = Fc[r * 3 + 0] * Fw[c * 3 + 0] + Fc[r * 3 + 1] * Fw[c * 3 + 1] + Fc[r * 3 + 2] * Fw[c * 3 + 2];
    cand.t[0] = P1.x - (cand.R[0] * X[0].x + cand.R[1] * X[0].y + cand.R[2] * X[0].z);
    cand.t[1] = P1.y - (cand.R[3] * X[0].x + cand.R[4] * X[0].y + cand.R[5] * X[0].z);
    cand.t[2] = P1.z - (cand.R[6] * X[0].x + cand.R[7] * X[0].y + cand.R[8] * X[0].z);
    const double e = reproj_err2(cand, K4, X4, u4, v4);
    if (e < best_e) {
      best_e = e;
      best = cand;
      found = true;
    }
  }
  return found;
}

// OpenCV RANSACUpdateNumIters (modules/calib3d/src/ptsetreg.cpp).
__device__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.0), 1.0);
  ep = fmin(fmax(ep, 0.0), 1.0);
  double num = fmax(1.0 - p, 2.2250738585072014e-308);
  double denom = 1.0 - pow(1.0 - ep, static_cast<double>(model_points));
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  if (denom >= 0 || -num >= max_iters * (-denom)) return max_iters;
  return static_cast<int>(rint(num / denom));
}

__device__ void rodrigues(const double* w, double (&Rm)[9]) {
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double Kx[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
  double a, b;
  if (th < 1e-12) {
    a = 1.0;
    b = 0.0;
  } else {
    a = sin(th) / th;
    b = (1.0 - cos(th)) / (th * th);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double kk = Kx[r * 3 + 0] * Kx[0 * 3 + c] + Kx[r * 3 + 1] * Kx[1 * 3 + c] + Kx[r * 3 + 2] * Kx[2 * 3 + c];
      Rm[r * 3 + c] = (r == c ? 1.0 : 0.0) + a * Kx[r * 3 + c] + b * kk;
    }
}

// Block-wide normal equations of the reprojection error over the inliers at pose P:
// N[0..20] = upper triangle of J^T J (row-major), N[21..26] = J^T r, N[27] = r^T r.
// Every thread returns the same values (fixed-order sum of the warp partials).
__device__ void normal_equations(const Pose& P, const double* K4, int n, const double* sX, const double* sx,
                                 const unsigned char* sMask, double (*sPart)[kNormalTerms], double (&N)[kNormalTerms]) {
  double acc[kNormalTerms];
#pragma unroll
  for (int k = 0; k < kNormalTerms; ++k) acc[k] = 0.0;
  for (int i = threadIdx.x; i < n; i += kPnpThreads) {
    if (!sMask[i]) continue;
    const double X = sX[i * 3 + 0], Y = sX[i * 3 + 1], Z = sX[i * 3 + 2];
    const double xc = P.R[0] * X + P.R[1] * Y + P.R[2] * Z + P.t[0];
    const double yc = P.R[3] * X + P.R[4] * Y + P.R[5] * Z + P.t[1];
    const double zc = P.R[6] * X + P.R[7] * Y + P.R[8] * Z + P.t[2];
    const double iz = 1.0 / zc;
    const double ru = K4[0] * xc * iz + K4[2] - sx[i * 2 + 0];
    const double rv = K4[1] * yc * iz + K4[3] - sx[i * 2 + 1];
    const double du[3] = {K4[0] * iz, 0.0, -K4[0] * xc * iz * iz};
    const double dv[3] = {0.0, K4[1] * iz, -K4[1] * yc * iz * iz};
    double Ju[6], Jv[6];
    Ju[0] = du[2] * yc - du[1] * zc; Ju[1] = du[0] * zc - du[2] * xc; Ju[2] = du[1] * xc - du[0] * yc;
    Ju[3] = du[0]; Ju[4] = du[1]; Ju[5] = du[2];
    Jv[0] = dv[2] * yc - dv[1] * zc; Jv[1] = dv[0] * zc - dv[2] * xc; Jv[2] = dv[1] * xc - dv[0] * yc;
    Jv[3] = dv[0]; Jv[4] = dv[1]; Jv[5] = dv[2];
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) acc[k++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += Ju[a] * ru + Jv[a] * rv;
    acc[27] += ru * ru + rv * rv;
  }
#pragma unroll
  for (int k = 0; k < kNormalTerms; ++k) {
    double v = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    acc[k] = v;
  }
  __syncthreads();   // previous readers of sPart are done
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < kNormalTerms; ++k) sPart[threadIdx.x >> 5][k] = acc[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kNormalTerms; ++k) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kPnpWarps; ++w) v += sPart[w][k];
    N[k] = v;
  }
}

// Solves A d = -g for the damped 6x6 system (LU with partial pivoting).  False when singular.
__device__ bool solve6(const double (&N)[kNormalTerms], double lam, double (&delta)[6]) {
  double A[6][7];
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) {
      A[a][b] = N[k];
      A[b][a] = N[k];
      ++k;
    }
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    A[a][a] += lam * A[a][a];
    A[a][6] = -N[21 + a];
  }
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    double big = fabs(A[c][c]);
    for (int r = c + 1; r < 6; ++r)
      if (fabs(A[r][c]) > big) { big = fabs(A[r][c]); piv = r; }
    if (!(big > 0.0) || !isfinite(big)) return false;
    if (piv != c)
      for (int j = 0; j < 7; ++j) { const double tmp = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = tmp; }
    for (int r = c + 1; r < 6; ++r) {
      const double m = A[r][c] / A[c][c];
      for (int j = c; j < 7; ++j) A[r][j] -= m * A[c][j];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = A[r][6];
    for (int j = r + 1; j < 6; ++j) s -= A[r][j] * delta[j];
    delta[r] = s / A[r][r];
  }
  return true;
}

__global__ void __launch_bounds__(kPnpThreads)
pnp_ransac_kernel(const float* __restrict__ coord_2d, const float* __restrict__ coord_3d,
                  const int* __restrict__ counts, const double* __restrict__ intrinsics, int M, int iters,
                  double thresh2, double confidence, unsigned long long seed, int problem_offset,
                  int* __restrict__ success, double* __restrict__ out_R, double* __restrict__ out_t,
                  unsigned char* __restrict__ inlier_mask, int* __restrict__ num_inliers,
                  int* __restrict__ iters_run, int* __restrict__ best_hyp) {
  extern __shared__ __align__(16) unsigned char pnp_smem[];
  const int p = blockIdx.x;
  double* sX = reinterpret_cast<double*>(pnp_smem);           // [M][3]
  double* sx = sX + static_cast<size_t>(M) * 3;               // [M][2]
  double(*sPart)[kNormalTerms] = reinterpret_cast<double(*)[kNormalTerms]>(sx + static_cast<size_t>(M) * 2);
  int* sCount = reinterpret_cast<int*>(sPart + kPnpWarps);    // [iters]
  unsigned char* sMask = reinterpret_cast<unsigned char*>(sCount + iters);   // [M]
  __shared__ int sBest[3];                                    // best_h, best_count, iterations run
  __shared__ double sK[4];

  int n = counts[p];
  n = n < 0 ? 0 : (n > M ? M : n);
  for (int i = threadIdx.x; i < n; i += kPnpThreads) {
    sX[i * 3 + 0] = coord_3d[(static_cast<size_t>(p) * M + i) * 3 + 0];
    sX[i * 3 + 1] = coord_3d[(static_cast<size_t>(p) * M + i) * 3 + 1];
    sX[i * 3 + 2] = coord_3d[(static_cast<size_t>(p) * M + i) * 3 + 2];
    sx[i * 2 + 0] = coord_2d[(static_cast<size_t>(p) * M + i) * 2 + 0];
    sx[i * 2 + 1] = coord_2d[(static_cast<size_t>(p) * M + i) * 2 + 1];
  }
  if (threadIdx.x < 4) sK[threadIdx.x] = intrinsics[p * 4 + threadIdx.x];
  for (int i = threadIdx.x; i < M; i += kPnpThreads) {
    sMask[i] = 0;
    inlier_mask[static_cast<size_t>(p) * M + i] = 0;
  }
  __syncthreads();
  const double K4[4] = {sK[0], sK[1], sK[2], sK[3]};
  const int problem = problem_offset + p;

  // ---- phase 1: inlier count of every hypothesis ----
  if (n >= kModelPoints) {
    for (int h = threadIdx.x; h < iters; h += kPnpThreads) {
      Pose pose;
      int count = 0;
      if (hypothesis_pose(seed, problem, h, n, sX, sx, K4, pose)) {
        for (int i = 0; i < n; ++i) {
          const Vec3 X = {sX[i * 3 + 0], sX[i * 3 + 1], sX[i * 3 + 2]};
          count += reproj_err2(pose, K4, X, sx[i * 2 + 0], sx[i * 2 + 1]) <= thresh2 ? 1 : 0;
        }
      }
      sCount[h] = count;
    }
  }
  __syncthreads();

  // ---- phase 2: the order-dependent part of RANSAC, one sequential scan over the counts ----
  if (threadIdx.x == 0) {
    int best_count = 0, best_h = -1, niters = iters, h = 0;
    if (n >= kModelPoints) {
      for (; h < niters; ++h) {
        const int c = sCount[h];
        if (c > (best_count > kModelPoints - 1 ? best_count : kModelPoints - 1)) {
          best_count = c;
          best_h = h;
          niters = ransac_update_num_iters(confidence, static_cast<double>(n - c) / n, kModelPoints, niters);
        }
      }
    }
    sBest[0] = best_h;
    sBest[1] = best_count;
    sBest[2] = h;
  }
  __syncthreads();
  const int bh = sBest[0];
  const bool ok = bh >= 0 && sBest[1] >= kMinInliersForPose;
  if (threadIdx.x == 0) {
    success[p] = ok ? 1 : 0;
    num_inliers[p] = ok ? sBest[1] : 0;
    iters_run[p] = sBest[2];
    best_hyp[p] = ok ? bh : -1;
  }
  if (!ok) {
    if (threadIdx.x < 9) out_R[p * 9 + threadIdx.x] = (threadIdx.x % 4 == 0) ? 1.0 : 0.0;
    if (threadIdx.x < 3) out_t[p * 3 + threadIdx.x] = 0.0;
    return;
  }

  // ---- phase 3: inliers of the winning hypothesis (every thread recomputes its pose: identical) ----
  Pose pose;
  hypothesis_pose(seed, problem, bh, n, sX, sx, K4, pose);
  for (int i = threadIdx.x; i < n; i += kPnpThreads) {
    const Vec3 X = {sX[i * 3 + 0], sX[i * 3 + 1], sX[i * 3 + 2]};
    const unsigned char in = reproj_err2(pose, K4, X, sx[i * 2 + 0], sx[i * 2 + 1]) <= thresh2 ? 1 : 0;
    sMask[i] = in;
    inlier_mask[static_cast<size_t>(p) * M + i] = in;
  }
  __syncthreads();

  // ---- phase 4: Levenberg-Marquardt on the inliers (oracle/pnp.py refine_lm) ----
  double lam = 1e-3;
  double N[kNormalTerms];
  normal_equations(pose, K4, n, sX, sx, sMask, sPart, N);
  double cost = N[27];
  for (int it = 0; it < kLmMaxIters; ++it) {
    double delta[6];
    if (!solve6(N, lam, delta)) {
      lam *= 10.0;
      continue;
    }
    double dR[9];
    rodrigues(delta, dR);
    Pose cand;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        cand.R[r * 3 + c] = dR[r * 3 + 0] * pose.R[0 * 3 + c] + dR[r * 3 + 1] * pose.R[1 * 3 + c] + dR[r * 3 + 2] * pose.R[2 * 3 + c];
      cand.t[r] = dR[r * 3 + 0] * pose.t[0] + dR[r * 3 + 1] * pose.t[1] + dR[r * 3 + 2] * pose.t[2] + delta[3 + r];
    }
    double Nn[kNormalTerms];
    normal_equations(cand, K4, n, sX, sx, sMask, sPart, Nn);
    const double cost_n = Nn[27];
    if (isfinite(cost_n) && cost_n < cost) {
      const bool small = cost - cost_n <= 1e-12 * cost;
      pose = cand;
#pragma unroll
      for (int k = 0; k < kNormalTerms; ++k) N[k] = Nn[k];
      cost = cost_n;
      lam = fmax(lam * 0.1, 1e-12);
      if (small) break;
    } else {
      lam *= 10.0;
      if (lam > 1e12) break;
    }
  }
  if (threadIdx.x < 9) out_R[p * 9 + threadIdx.x] = pose.R[threadIdx.x];
  if (threadIdx.x < 3) out_t[p * 3 + threadIdx.x] = pose.t[threadIdx.x];
}

}  // namespace

int pnp_ransac(const float* coord_2d, const float* coord_3d, const int* counts, const double* intrinsics, int P,
               int M, int iters, double thresh, double confidence, unsigned long long seed, int problem_offset,
               int* success, double* out_R, double* out_t, unsigned char* inlier_mask, int* num_inliers,
               int* iters_run, int* best_hyp, cudaStream_t stream) {
  if (P == 0) return 0;
  FP_REQUIRE(P > 0 && M > 0 && M <= 2048, "pnp_ransac: need P > 0 and 0 < M <= 2048 (got P=%d M=%d)", P, M);
  FP_REQUIRE(iters > 0 && iters <= 8192, "pnp_ransac: iterations must be in [1, 8192] (got %d)", iters);
  FP_REQUIRE(thresh > 0.0, "pnp_ransac: the inlier threshold must be positive");
  const size_t smem = static_cast<size_t>(M) * 5 * sizeof(double) + kPnpWarps * kNormalTerms * sizeof(double) +
                      static_cast<size_t>(iters) * sizeof(int) + static_cast<size_t>(M) + 16;
  static bool configured[64] = {};
  if (per_device_once(configured)) {
    FP_CUDA_CHECK(cudaFuncSetAttribute(pnp_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  ProfScope prof(PROF_RETRIEVAL, stream, static_cast<double>(P) * M * 20.0);
  pnp_ransac_kernel<<<P, kPnpThreads, smem, stream>>>(coord_2d, coord_3d, counts, intrinsics, M, iters,
                                                      thresh * thresh, confidence, seed, problem_offset, success,
                                                      out_R, out_t, inlier_mask, num_inliers, iters_run, best_hyp);
  FP_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fp
