// Microbenchmark: does FFMA2 (fma.rn.f32x2, sm_100) retire two FP32 FMAs per issue slot, and does it share the slot with
// integer work?  Four loops of the same FP32 work: scalar FFMA, packed FFMA2, and each mixed with LOP3/IADD3 chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned iop(unsigned a, unsigned b) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>   // 0 scalar, 1 packed, 2 scalar + int, 3 packed + int, 4 int only
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float a[8]; u64 p[4]; unsigned q[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; q[i] = threadIdx.x + i; }
  for (int i = 0; i < 4; i++) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const u64 ps = pk(s, s);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      if (MODE == 0 || MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma1(a[i], s, s);
      }
      if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fma2(p[i], ps, ps);
      }
      if (MODE >= 2) {
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = iop(q[i], q[(i + 1) & 7]);
      }
    }
  }
  float acc = 0;
  for (int i = 0; i < 8; i++) acc += a[i] + (float)q[i];
  for (int i = 0; i < 4; i++) acc += (float)(p[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> float run(float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, 16, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 4096;
  const double fmas = 148.0 * 8 * 256 * iters * 64;   // FP32 FMAs per launch (modes 0-3)
  float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters), t4 = run<4>(d, iters);
  printf("scalar FFMA        %.3f ms  %.1f TFMA/s\n", t0, fmas / t0 * 1e-9);
  printf("packed FFMA2       %.3f ms  %.1f TFMA/s\n", t1, fmas / t1 * 1e-9);
  printf("scalar + 64 LOP3   %.3f ms\n", t2);
  printf("packed + 64 LOP3   %.3f ms\n", t3);
  printf("64 LOP3 only       %.3f ms\n", t4);
  return 0;
}
