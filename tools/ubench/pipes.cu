// Instruction-throughput probe for the softmax inner loop on sm_100a: cycles per warp-instruction per SM
// sub-partition at saturation, for the ops the attention softmax issues.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 16
#define ITERS 512

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[CHAINS];
  uint64_t p[CHAINS / 2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) a[i] = seed * (i + 1) + threadIdx.x * 1e-3f;
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  uint64_t c2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(seed), "f"(seed));
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(a[(i + 1) % CHAINS]));
      if (OP == 2 && i < CHAINS / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(p[(i + 1) % (CHAINS / 2)]));
      if (OP == 3 && i < CHAINS / 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 4 && i < CHAINS / 2) {
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        a[2 * i] = __uint_as_float(r);
      }
      if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 6) { uint32_t u = __float_as_uint(a[i]); asm volatile("{ .reg .b32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1; }" : "+r"(u) : "r"(__float_as_uint(seed))); a[i] = __uint_as_float(u); }
      if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[i]) : "f"(seed));
      if (OP == 8 && i < CHAINS / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 9) {   // softmax mix per pair: 1 FFMA2, 2 MUFU, 1 F2FP
        if (i < CHAINS / 2) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
          float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(lo));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(hi));
          uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
          a[2 * i] = __uint_as_float(r);
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
  printf("%-28s", name);
  for (int threads : {128, 256, 512, 1024}) {
    k<OP><<<1, threads>>>(out, cyc, 0.5f);
    k<OP><<<1, threads>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    int wps = threads / 128;
    printf("  %dw/smsp: %6.2f", wps, double(c) / (double(ITERS) * instr_per_iter * wps));
  }
  printf("   (cycles per warp-instruction per SMSP)\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", CHAINS);
  run<1>("FFMA 3-reg", CHAINS);
  run<7>("FFMA reg,reg,imm", CHAINS);
  run<2>("FFMA2 3-reg", CHAINS / 2);
  run<8>("FFMA2 a,b,b", CHAINS / 2);
  run<3>("FADD2", CHAINS / 2);
  run<4>("F2FP.F16.F32.PACK_AB", CHAINS / 2);
  run<5>("FMNMX", CHAINS);
  run<6>("SHL+ADD (LEA)", CHAINS);
  run<9>("mix FFMA2+2MUFU+F2FP /pair", CHAINS / 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
