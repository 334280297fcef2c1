// mcr_portable.h — the handful of device intrinsics the path headers use, with host equivalents.
//
// The product is CUDA-only (there is no CPU fallback and libmcr_b200.so contains none of this
// host code). The host equivalents exist for ONE purpose: tests/host_model compiles the very same
// path headers for the host, so that the `-m "not gpu"` suite can check the arithmetic of the
// fast / lean month steps against the CPU oracle (<= 1e-9, identical flags) before a GPU is
// available. Nothing under monte_carlo_retirement_b200/ loads that test library.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define MCR_DEV __host__ __device__ __forceinline__
#else
#define MCR_DEV inline
#endif

namespace mcr {

MCR_DEV float u32_to_float(uint32_t a) {
#ifdef __CUDA_ARCH__
  return __uint2float_rn(a);
#else
  return (float)a;  // round-to-nearest-even, as cvt.rn.f32.u32
#endif
}

MCR_DEV void sincospi_f(float turns, float& s, float& c) {
#ifdef __CUDA_ARCH__
  sincospif(turns, &s, &c);
#else
  s = (float)std::sin(3.14159265358979323846 * (double)turns);
  c = (float)std::cos(3.14159265358979323846 * (double)turns);
#endif
}

MCR_DEV double load_stream(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldcs(p);
#else
  return *p;
#endif
}

MCR_DEV void store_stream(double* p, double v) {
#ifdef __CUDA_ARCH__
  __stcs(p, v);
#else
  *p = v;
#endif
}

MCR_DEV int32_t hi_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t b;
  std::memcpy(&b, &x, 8);
  return (int32_t)(b >> 32);
#endif
}

MCR_DEV double nan_value() {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(0xfff8000000000000ULL);  // CUDART_NAN
#else
  return std::nan("");
#endif
}

// MUFU.RCP64H seed of the fast reciprocals: ~2^-20 relative (the low mantissa word is ignored).
// Host model: the same information content — both low words dropped.
MCR_DEV double rcp_seed(double b) {
#ifdef __CUDA_ARCH__
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(b));
  return x0;
#else
  uint64_t u;
  std::memcpy(&u, &b, 8);
  u &= 0xffffffff00000000ull;
  double t;
  std::memcpy(&t, &u, 8);
  t = 1.0 / t;
  std::memcpy(&u, &t, 8);
  u &= 0xffffffff00000000ull;
  std::memcpy(&t, &u, 8);
  return t;
#endif
}

}  // namespace mcr
