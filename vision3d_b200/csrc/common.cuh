// Shared device/host helpers for the sm_100a kernels behind include/v3d_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/v3d_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && (__CUDA_ARCH__ != 1000)
#error "vision3d_b200 kernels are written for sm_100a only"
#endif

namespace v3d {

constexpr int kNumSMs = 148;  // B200

void set_cuda_error(cudaError_t e);

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return V3D_ERR_CUDA;
  }
  return V3D_OK;
}

#define V3D_CUDA_TRY(expr)                 \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) {               \
      ::v3d::set_cuda_error(_e);           \
      return V3D_ERR_CUDA;                 \
    }                                      \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property and the wrappers accept tensors on
// any device of the process: remember per (call site, device) whether it was applied. Setting the attribute is
// idempotent, so two threads racing on the first call both set it and both mark it (no lock needed).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0ull};
  static unsigned long long bit() {
    int d = 0;
    cudaGetDevice(&d);
    return 1ull << (d & 63);
  }
  bool needed() const { return (mask.load(std::memory_order_acquire) & bit()) == 0ull; }
  void done() { mask.fetch_or(bit(), std::memory_order_release); }
};

inline cudaStream_t as_stream(v3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Epoch-tagged open-addressing hash keyed by a <=40-bit integer.
// Entry = [epoch:24 | key:40]. Entries written under an older epoch read as empty, so a table
// is "cleared" by bumping the epoch: no per-call memset traffic. Linear probing.
// The tables keep a 32-bit call counter; the 24-bit key epoch is derived from it (epoch24()), and every
// payload word carries the full 32-bit counter as its tag. When the 24-bit epoch recurs (every 2^24-1
// calls) a never-overwritten entry from that era can look live in `keys`; that is harmless by
// construction: it either holds the same key (claiming is idempotent, payload tags still mismatch)
// or acts as a tombstone (tables are sized >= 4x the rows, so load stays <= 50 %).
// ---------------------------------------------------------------------------------------------
constexpr int kKeyBits = 40;
constexpr unsigned long long kKeyMask = (1ull << kKeyBits) - 1ull;

__host__ __device__ __forceinline__ unsigned int epoch24(unsigned int calls) {
  return calls % 0xFFFFFFu + 1u;
}

__device__ __forceinline__ unsigned int hash_key(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (unsigned int)k;
}

// Find-or-claim the slot of `key` for this epoch. Returns the slot index.
__device__ __forceinline__ unsigned int table_claim(unsigned long long* __restrict__ keys,
                                                    unsigned int mask, unsigned int epoch,
                                                    unsigned long long key) {
  const unsigned long long mine = ((unsigned long long)epoch << kKeyBits) | key;
  unsigned int s = hash_key(key) & mask;
  unsigned long long cur = __ldcg(&keys[s]);
  while (true) {
    if (cur == mine) return s;
    if ((cur >> kKeyBits) != epoch) {
      // stale entry: try to claim. On failure continue from the value the CAS observed (never
      // re-read through L1, which may hold the stale line).
      unsigned long long old = atomicCAS(&keys[s], cur, mine);
      if (old == cur) return s;
      cur = old;
      continue;
    }
    s = (s + 1) & mask;
    cur = __ldcg(&keys[s]);
  }
}

// Lookup only. Returns slot or 0xFFFFFFFF.
__device__ __forceinline__ unsigned int table_find(const unsigned long long* __restrict__ keys,
                                                   unsigned int mask, unsigned int epoch,
                                                   unsigned long long key) {
  const unsigned long long mine = ((unsigned long long)epoch << kKeyBits) | key;
  unsigned int s = hash_key(key) & mask;
  while (true) {
    unsigned long long cur = __ldg(&keys[s]);
    if (cur == mine) return s;
    if ((cur >> kKeyBits) != epoch) return 0xFFFFFFFFu;
    s = (s + 1) & mask;
  }
}

inline unsigned int next_pow2(unsigned int v) {
  unsigned int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024), returns exclusive prefix;
// total (sum over block) returned through `total`. smem must hold 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    smem[lane] = winc - w;  // exclusive prefix of warp sums
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int res = smem[warp] + inc - v;
  total = smem[32];
  __syncthreads();
  return res;
}

}  // namespace v3d
