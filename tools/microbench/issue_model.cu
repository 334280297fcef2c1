// issue_model.cu — how do FP64 instructions share the issue slots of a B200 SMSP with other work?
// Each kernel runs, per loop iteration and per thread, D independent DFMA + A ALU ops (LOP3/IADD3
// chains) + F FP32 FMA-pipe ops + M MUFU ops, 8 warps per SMSP, and reports cycles per
// warp-iteration per SMSP. If t ~= max(2D, D + A + F + ...) the FP64 pipe is half rate but leaves
// the issue port free; if t ~= 2D + A + F the DFMA holds the dispatch port for both cycles.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_model issue_model.cu && ./issue_model
#include <cstdio>
#include <cuda_runtime.h>

template <int D, int A, int F, int M>
__global__ void __launch_bounds__(256) k_mix(int iters, double seed, double* sink) {
  double d[8];
  unsigned a[8];
  float f[8];
  float m[4];
#pragma unroll
  for (int c = 0; c < 8; ++c) { d[c] = seed + c * 1e-9 + threadIdx.x * 1e-12; a[c] = threadIdx.x * 2654435761u + c; f[c] = 1.0f + c * 1e-3f; }
#pragma unroll
  for (int c = 0; c < 4; ++c) m[c] = 1.5f + c;
  const double mm = 1.0000000001 + seed * 1e-30, bb = 1e-12;
  const unsigned k1 = 0x9E3779B9u + (unsigned)iters, k2 = 0x85EBCA6Bu;
  const float fm = 1.0000001f, fb = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < D; ++c) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(mm), "d"(bb));
#pragma unroll
      for (int c = 0; c < A; ++c) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(k1), "r"(k2));  // one ALU op
#pragma unroll
      for (int c = 0; c < F; ++c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fm), "f"(fb));
#pragma unroll
      for (int c = 0; c < M; ++c) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
    }
  }
  double s = 0;
  unsigned t = 0;
  float g = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) { s += d[c]; t += a[c]; g += f[c]; }
#pragma unroll
  for (int c = 0; c < 4; ++c) g += m[c];
  if (s == 12345.678 || t == 0x12345u || g == 3.25f) sink[0] = s + t + g;
}

template <int D, int A, int F, int M>
void run(const char* name, int sms, double clock_ghz, double* sink) {
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 4;  // 4 x 256 threads per SM = 8 warps per SMSP
  k_mix<D, A, F, M><<<blocks, 256>>>(10, 1.0, sink);
  cudaEventRecord(e0);
  k_mix<D, A, F, M><<<blocks, 256>>>(iters, 1.0, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  // per SMSP: 8 warps x iters x 8 unrolled groups
  const double cyc = ms * 1e-3 * clock_ghz * 1e9 / (8.0 * iters * 8.0);
  printf("%-28s D=%d A=%d F=%d M=%d : %.2f cycles per warp-group per SMSP\n", name, D, A, F, M, cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz (max); cycles assume the max clock\n", p.name, p.multiProcessorCount, ghz);
  double* sink;
  cudaMalloc(&sink, 64);
  run<8, 0, 0, 0>("DFMA only", p.multiProcessorCount, ghz, sink);
  run<0, 8, 0, 0>("ALU only", p.multiProcessorCount, ghz, sink);
  run<0, 0, 8, 0>("FFMA only", p.multiProcessorCount, ghz, sink);
  run<0, 0, 0, 4>("MUFU only", p.multiProcessorCount, ghz, sink);
  run<8, 8, 0, 0>("DFMA + ALU", p.multiProcessorCount, ghz, sink);
  run<8, 0, 8, 0>("DFMA + FFMA", p.multiProcessorCount, ghz, sink);
  run<8, 4, 4, 0>("DFMA + ALU/2 + FFMA/2", p.multiProcessorCount, ghz, sink);
  run<4, 8, 8, 0>("DFMA/2 + ALU + FFMA", p.multiProcessorCount, ghz, sink);
  run<8, 8, 8, 0>("DFMA + ALU + FFMA", p.multiProcessorCount, ghz, sink);
  run<8, 0, 0, 2>("DFMA + MUFU/4", p.multiProcessorCount, ghz, sink);
  run<8, 8, 8, 2>("DFMA + ALU + FFMA + MUFU/4", p.multiProcessorCount, ghz, sink);
  run<4, 8, 8, 4>("DFMA/2 + ALU + FFMA + MUFU/2", p.multiProcessorCount, ghz, sink);
  return 0;
}
