// issue_model2.cu — per-class issue cost of the non-FP64 instructions of the timeline kernel on a
// B200 SMSP, alone and next to DFMA / LOP3 / FFMA streams. 8 warps per SMSP, 8 independent chains
// per thread; every op is pinned with inline PTX. Output: cycles per op per SMSP (warp-instruction).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_model2 issue_model2.cu && ./issue_model2
#include <cstdio>
#include <cuda_runtime.h>

enum Op { NONE, DFMA, LOP3, FFMA, FMUL, IMADW, IMAD, FSEL, ISETP_SEL, I2FP, F2F, MUFU, IADD3, SHF, DSETP_FSEL, DADD, DMUL };

template <int OP>
__device__ __forceinline__ void op(double& d, unsigned& a, float& f, unsigned long long& w, unsigned k1, unsigned k2,
                                   double mm, double bb, float fm, float fb) {
  if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(mm), "d"(bb));
  if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(bb));
  if (OP == DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(mm));
  if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(k1), "r"(k2));
  if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(fm), "f"(fb));
  if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(fm));
  if (OP == IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(a), "r"(k1));
  if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(k1), "r"(k2));
  if (OP == FSEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(f) : "f"(fb), "r"(k2));
  if (OP == ISETP_SEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %2, p;}" : "+r"(a) : "r"(k1), "r"(k2));
  if (OP == I2FP) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a));
  if (OP == F2F) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f));
  if (OP == MUFU) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
  if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(k1));
  if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %0, 5;" : "+r"(a));
  if (OP == DSETP_FSEL) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %0, %1, p;}" : "+d"(d) : "d"(bb));
}

template <int OP1, int N1, int OP2, int N2>
__global__ void __launch_bounds__(256) k_mix(int iters, double seed, double* sink) {
  double d[8];
  unsigned a[8];
  float f[8];
  unsigned long long w[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { d[c] = seed + c * 1e-9 + threadIdx.x * 1e-12; a[c] = threadIdx.x * 2654435761u + c; f[c] = 1.0f + c * 1e-3f; w[c] = a[c]; }
  const double mm = 1.0000000001 + seed * 1e-30, bb = 1e-12;
  const unsigned k1 = 0x9E3779B9u + (unsigned)iters, k2 = 0x85EBCA6Bu;
  const float fm = 1.0000001f, fb = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < N1) op<OP1>(d[c], a[c], f[c], w[c], k1, k2, mm, bb, fm, fb);
        // the second stream works on the other half of the registers so the two never depend on each other
        if (c < N2) op<OP2>(d[7 - c], a[7 - c], f[7 - c], w[7 - c], k1, k2, mm, bb, fm, fb);
      }
    }
  }
  double s = 0;
  unsigned t = 0;
  float g = 0;
  unsigned long long ww = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) { s += d[c]; t += a[c]; g += f[c]; ww += w[c]; }
  if (s == 12345.678 || t == 0x12345u || g == 3.25f || ww == 77) sink[0] = s + t + g;
}

static int g_sms;
static double g_ghz;
static double* g_sink;

template <int OP1, int N1, int OP2, int N2>
double run(const char* name) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = g_sms * 4;
  k_mix<OP1, N1, OP2, N2><<<blocks, 256>>>(10, 1.0, g_sink);
  cudaEventRecord(e0);
  k_mix<OP1, N1, OP2, N2><<<blocks, 256>>>(iters, 1.0, g_sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * g_ghz * 1e9 / (8.0 * iters * 4.0);  // per warp per unrolled group, per SMSP
  printf("%-34s %d + %d ops: %6.2f cycles per group\n", name, N1, N2, cyc);
  return cyc;
}

#define ALONE(OPX) run<OPX, 4, NONE, 0>(#OPX " x4 alone")
#define WITH(OPX, OPY) run<OPX, 4, OPY, 4>(#OPX " x4 + " #OPY " x4")

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  g_sms = p.multiProcessorCount;
  g_ghz = khz * 1e-6;
  cudaMalloc(&g_sink, 64);
  printf("%s, %d SMs, %.3f GHz; cycles per SMSP for a group of ops issued by each of 8 resident warps\n", p.name, g_sms, g_ghz);
  ALONE(DFMA); ALONE(DADD); ALONE(DMUL); ALONE(DSETP_FSEL); ALONE(LOP3); ALONE(IADD3); ALONE(SHF); ALONE(FSEL); ALONE(ISETP_SEL);
  ALONE(FFMA); ALONE(FMUL); ALONE(IMAD); ALONE(IMADW); ALONE(I2FP); ALONE(F2F); ALONE(MUFU);
  WITH(DFMA, LOP3); WITH(DFMA, FFMA); WITH(DFMA, IMADW); WITH(DFMA, FSEL); WITH(DFMA, MUFU); WITH(DFMA, F2F); WITH(DFMA, I2FP);
  WITH(DFMA, DSETP_FSEL);
  WITH(LOP3, FFMA); WITH(LOP3, IMADW); WITH(LOP3, FSEL); WITH(LOP3, MUFU); WITH(LOP3, I2FP); WITH(LOP3, IADD3);
  WITH(FFMA, IMADW); WITH(FFMA, FMUL); WITH(FFMA, MUFU); WITH(IMADW, MUFU); WITH(IMADW, I2FP); WITH(F2F, MUFU); WITH(I2FP, MUFU);
  return 0;
}
