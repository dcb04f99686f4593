// Micro-probe: cost of one tcgen05.mma (cta_group::1, M=128) as a function of N, operand kind and A source
// (TMEM vs shared memory). One CTA, one elected thread issues REPS back-to-back MMAs into the same
// accumulator, commits to an mbarrier and waits; clock64 brackets issue..completion.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_probe scripts/mma_probe.cu && /tmp/mma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// kind: 0 = tf32 (a=b=2), 1 = f16 (a=b=0), 2 = bf16 (a=b=1)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int kind) {
  const uint32_t ab = kind == 0 ? 2u : (kind == 1 ? 0u : 1u);
  return (1u << 4) | (ab << 7) | (ab << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int KIND, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (KIND == 0) {
    if (TS)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
  } else {
    if (TS)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
  }
}

// bf16 SS-mode MMA with the disable-output-lane operand (4 x 32-bit lane mask)
__device__ __forceinline__ void mma_masked(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t m0,
                                           uint32_t m1, uint32_t m2, uint32_t m3) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n\t}" ::"r"(d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
               : "memory");
}

template <int KIND, bool TS, int N, int REPS>
__global__ void probe(long long* out, unsigned int mask) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint64_t bdesc = make_desc(smem_u32(base));
    const uint64_t adesc = make_desc(smem_u32(base) + 32768);
    constexpr uint32_t idesc = make_idesc(128, N, KIND == 3 ? 2 : KIND);
    const long long t0 = clock64();
#pragma unroll 8
    for (int r = 0; r < REPS; r++) {
      if (KIND == 3)
        mma_masked(tm, adesc + 2 * (r & 3), bdesc + 2 * (r & 3), make_idesc(128, N, 2), mask, mask, mask, mask);
      else
        mma<KIND, TS>(tm, tm + 256 + 8 * (r & 7), adesc + 2 * (r & 3), bdesc + 2 * (r & 3), idesc);
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int KIND, bool TS, int N>
void run(const char* name, unsigned int mask = 0u) {
  constexpr int REPS = 2048;
  long long* d;
  cudaMalloc(&d, 16);
  auto k = probe<KIND, TS, N, REPS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  long long h[2] = {0, 0};
  for (int it = 0; it < 2; it++) {
    k<<<1, 128, 80 * 1024>>>(d, mask);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-28s N=%3d  ERROR %s\n", name, N, cudaGetErrorString(e));
      return;
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  }
  printf("%-28s N=%3d  issue %.1f clk/mma   complete %.1f clk/mma\n", name, N, (double)h[0] / REPS, (double)h[1] / REPS);
  cudaFree(d);
}

template <int KIND, bool TS>
void sweep(const char* name) {
  run<KIND, TS, 16>(name);
  run<KIND, TS, 32>(name);
  run<KIND, TS, 64>(name);
  run<KIND, TS, 128>(name);
  run<KIND, TS, 256>(name);
}

int main() {
  sweep<0, true>("tf32 K=8  A=TMEM");
  sweep<0, false>("tf32 K=8  A=smem");
  sweep<1, true>("f16  K=16 A=TMEM");
  sweep<1, false>("f16  K=16 A=smem");
  sweep<2, true>("bf16 K=16 A=TMEM");
  sweep<2, false>("bf16 K=16 A=smem");
  run<3, false, 64>("bf16 SS masked, mask=0", 0u);
  run<3, false, 128>("bf16 SS masked, mask=0", 0u);
  run<3, false, 64>("bf16 SS masked, half off", 0x0F0F3355u);
  run<3, false, 128>("bf16 SS masked, half off", 0x0F0F3355u);
  run<3, false, 128>("bf16 SS masked, all off", 0xFFFFFFFFu);
  return 0;
}
