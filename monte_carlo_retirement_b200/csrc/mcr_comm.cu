// mcr_comm.cu — all-reduce over the GPUs of ONE process, written against NVLink / NVSwitch peer
// memory: every rank's kernel publishes its contribution in its own staging buffer, signals the
// peers with a flag in THEIR memory, waits for their flags in its own, and then reduces all
// contributions by loading them straight out of the peers' HBM. One kernel per rank and call, no
// host synchronisation and no NCCL: the callers of the reference are a plain CLI and a FastAPI
// worker thread (backend/main.py:68-106, backend/server.py:309,405), so the multi-GPU form they
// can reach is one process driving G devices, and what crosses NVLink on this path
// is small and latency-bound (success counts, digit histograms of the distributed radix select,
// pooled select candidates, final-balance histograms: bytes to ~10 MB per call).
//
// Protocol (the one-shot scheme of custom all-reduce kernels; all ranks call with the same
// element count, in the same order, like any collective):
//   seq-th call, block b of rank r:
//     1. stage_r[parity][slice_b] = buf[slice_b]          parity = seq & 1
//     2. __threadfence_system(); flags_p[b][r] = seq for every peer p          (release)
//     3. spin until flags_r[b][p] >= seq for every peer p                        (acquire)
//     4. buf[slice_b] = op over p of stage_p[parity][slice_b]                    (peer loads)
//   Re-use of a staging half two calls later is safe without an exit barrier: a rank can only
//   start call seq + 2 after its call seq + 1 completed, whose step 3 saw every peer's flag of
//   seq + 1 — and a peer raises that flag only from a kernel that runs after its own call seq
//   (stream order), i.e. after it finished reading.
//   A peer that never arrives (a host exception on its thread) would hang the spin: step 3 gives up
//   after ~2 s of GPU time and raises a status word in mapped host memory that the Python side
//   checks after every call.
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda_runtime.h>

#include "../../include/mcr.h"

namespace mcr {

constexpr int kCommMaxWorld = 16;
constexpr int kCommMaxBlocks = 128;
constexpr int kCommThreads = 256;
constexpr long long kCommSpinLimit = 4000000000ll;   // clock64 ticks (~2 s)

struct CommDev {
  int32_t rank, world;
  char* stage[kCommMaxWorld];        // peer pointers: [2][max_bytes] each
  uint32_t* flags[kCommMaxWorld];    // peer pointers: [kCommMaxBlocks][kCommMaxWorld] each
  uint32_t* status;                  // mapped host word: non-zero after a spin timeout
  int64_t max_bytes;
};

template <typename T, int OP>
__device__ __forceinline__ T comm_op(T a, T b) {
  if (OP == 0) return a + b;
  if (OP == 1) return b < a ? b : a;
  return b > a ? b : a;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T, int OP>
__global__ void __launch_bounds__(kCommThreads) k_peer_all_reduce(const __grid_constant__ CommDev c, T* __restrict__ buf,
                                                                  int64_t n, uint32_t seq) {
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = lo + per < n ? lo + per : n;
  const int64_t off = (int64_t)(seq & 1u) * c.max_bytes;
  T* mine = (T*)(c.stage[c.rank] + off);
  for (int64_t i = lo + threadIdx.x; i < hi; i += kCommThreads) mine[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < c.world) {
    const int p = threadIdx.x;
    st_release_sys(c.flags[p] + blockIdx.x * kCommMaxWorld + c.rank, seq);
    const uint32_t* f = c.flags[c.rank] + blockIdx.x * kCommMaxWorld + p;
    const long long t0 = clock64();
    // (seq - flag) as a signed distance, so the 32-bit counter may wrap
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
      if (clock64() - t0 > kCommSpinLimit) {
        *c.status = seq ? seq : 1u;
        break;
      }
    }
  }
  __syncthreads();
  for (int64_t i = lo + threadIdx.x; i < hi; i += kCommThreads) {
    // L2-only loads: the staging halves are rewritten every second call and L1 is not coherent
    T acc = __ldcg((const T*)(c.stage[0] + off) + i);
    for (int p = 1; p < c.world; ++p) acc = comm_op<T, OP>(acc, __ldcg((const T*)(c.stage[p] + off) + i));
    buf[i] = acc;
  }
}

}  // namespace mcr

using namespace mcr;

struct mcr_comm {
  int device = 0;
  int rank = -1, world = 0;
  int64_t max_bytes = 0;
  char* stage = nullptr;
  uint32_t* flags = nullptr;
  uint32_t* status_host = nullptr;   // cudaHostAlloc(mapped)
  uint32_t* status_dev = nullptr;
  uint32_t seq = 0;
  CommDev dev;
  std::string err;
};

namespace {

struct Guard {
  int prev = -1;
  explicit Guard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~Guard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int comm_fail(mcr_comm* c, int code, const std::string& m) {
  if (c) c->err = m;
  return code;
}

template <typename T, int OP>
cudaError_t launch(const mcr_comm* c, void* buf, int64_t n, uint32_t seq, cudaStream_t st) {
  int64_t blocks = (n + kCommThreads * 8 - 1) / (kCommThreads * 8);
  blocks = blocks < 1 ? 1 : (blocks > kCommMaxBlocks ? kCommMaxBlocks : blocks);
  k_peer_all_reduce<T, OP><<<(unsigned)blocks, kCommThreads, 0, st>>>(c->dev, (T*)buf, n, seq);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int mcr_comm_device_of(const mcr_ctx* ctx);  // mcr_api.cu

int mcr_comm_create(mcr_ctx* ctx, int64_t max_bytes, mcr_comm** out) {
  if (!ctx || !out || max_bytes <= 0) return MCR_EINVAL;
  *out = nullptr;
  mcr_comm* c = new mcr_comm();
  c->device = mcr_comm_device_of(ctx);
  c->max_bytes = (max_bytes + 255) / 256 * 256;
  Guard g(c->device);
  const size_t flag_bytes = sizeof(uint32_t) * kCommMaxBlocks * kCommMaxWorld;
  if (cudaMalloc(&c->stage, (size_t)c->max_bytes * 2) != cudaSuccess || cudaMalloc(&c->flags, flag_bytes) != cudaSuccess ||
      cudaHostAlloc(&c->status_host, 64, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&c->status_dev, c->status_host, 0) != cudaSuccess) {
    cudaGetLastError();
    if (c->stage) cudaFree(c->stage);
    if (c->flags) cudaFree(c->flags);
    if (c->status_host) cudaFreeHost(c->status_host);
    delete c;
    return MCR_ENOMEM;
  }
  cudaMemset(c->flags, 0, flag_bytes);
  *c->status_host = 0;
  cudaDeviceSynchronize();
  *out = c;
  return MCR_OK;
}

int mcr_comm_connect(mcr_comm* c, int32_t rank, int32_t world, mcr_comm* const* all) {
  if (!c || !all || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || all[rank] != c)
    return comm_fail(c, MCR_EINVAL, "bad rank / world / communicator list");
  Guard g(c->device);
  c->rank = rank;
  c->world = world;
  c->dev.rank = rank;
  c->dev.world = world;
  c->dev.status = c->status_dev;
  c->dev.max_bytes = c->max_bytes;
  for (int p = 0; p < world; ++p) {
    if (!all[p] || all[p]->max_bytes != c->max_bytes) return comm_fail(c, MCR_EINVAL, "communicators differ in size");
    if (all[p]->device != c->device) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, c->device, all[p]->device);
      if (!can) return comm_fail(c, MCR_ECUDA, "devices " + std::to_string(c->device) + " and " +
                                                  std::to_string(all[p]->device) + " have no peer access");
      cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return comm_fail(c, MCR_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
    c->dev.stage[p] = all[p]->stage;
    c->dev.flags[p] = all[p]->flags;
  }
  return MCR_OK;
}

int mcr_comm_all_reduce(mcr_comm* c, int32_t op, void* buf_dev, int64_t n, void* stream) {
  if (!c || c->rank < 0) return comm_fail(c, MCR_EINVAL, "communicator is not connected");
  if (n < 0 || (!buf_dev && n > 0)) return comm_fail(c, MCR_EINVAL, "bad buffer");
  if (n == 0) return MCR_OK;
  const int64_t width = (op == MCR_COMM_SUM_I32) ? 4 : 8;
  if (n * width > c->max_bytes) return comm_fail(c, MCR_EINVAL, "all-reduce larger than the communicator's staging buffer");
  Guard g(c->device);
  const uint32_t seq = ++c->seq;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  switch (op) {
    case MCR_COMM_SUM_I32: e = launch<int32_t, 0>(c, buf_dev, n, seq, st); break;
    case MCR_COMM_SUM_I64: e = launch<long long, 0>(c, buf_dev, n, seq, st); break;
    case MCR_COMM_MIN_I64: e = launch<long long, 1>(c, buf_dev, n, seq, st); break;
    case MCR_COMM_MAX_I64: e = launch<long long, 2>(c, buf_dev, n, seq, st); break;
    case MCR_COMM_SUM_F64: e = launch<double, 0>(c, buf_dev, n, seq, st); break;
    case MCR_COMM_MIN_F64: e = launch<double, 1>(c, buf_dev, n, seq, st); break;
    case MCR_COMM_MAX_F64: e = launch<double, 2>(c, buf_dev, n, seq, st); break;
    default: return comm_fail(c, MCR_EINVAL, "bad all-reduce op");
  }
  if (e != cudaSuccess) return comm_fail(c, MCR_ECUDA, std::string("k_peer_all_reduce: ") + cudaGetErrorString(e));
  return MCR_OK;
}

int32_t mcr_comm_status(const mcr_comm* c) { return c ? (int32_t)*(volatile uint32_t*)c->status_host : -1; }
int64_t mcr_comm_calls(const mcr_comm* c) { return c ? (int64_t)c->seq : 0; }
const char* mcr_comm_last_error(const mcr_comm* c) { return c ? c->err.c_str() : ""; }

int mcr_comm_destroy(mcr_comm* c) {
  if (!c) return MCR_OK;
  {
    Guard g(c->device);
    cudaDeviceSynchronize();
    if (c->stage) cudaFree(c->stage);
    if (c->flags) cudaFree(c->flags);
    if (c->status_host) cudaFreeHost(c->status_host);
  }
  delete c;
  return MCR_OK;
}

}  // extern "C"
