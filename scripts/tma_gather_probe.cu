// Probe: throughput and layout of the Blackwell TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) as the
// A-operand fetch of the sparse convolution: 4 arbitrary rows of a [rows x 128] bf16 table per instruction, 64
// columns (128 B) each, written into a SWIZZLE_128B shared-memory tile; out-of-range row indices read as zeros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_probe scripts/tma_gather_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = smem_u32(b);
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, uint64_t* bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

constexpr int kStageRows = 128, kStageBytes = kStageRows * 128, kStages = 4;

// mode 0: layout check (one stage, dump smem). mode 1: throughput, `warps` issuing warps share each stage.
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ rows, int n_idx,
                                             int iters, int warps, int mode, unsigned char* dump, long long* clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode == 0) {
    if (warp == 0) {
      if (lane == 0) mbar_expect(&full[0], kStageBytes);
      __syncwarp();
      const int4 r = *reinterpret_cast<const int4*>(rows + 4 * lane);
      gather4(smem + lane * 512, &tmap, &full[0], 64, r.x, r.y, r.z, r.w);  // column 64: the second half of the row
      mbar_wait(&full[0], 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kStageBytes; i += blockDim.x) dump[i] = smem[i];
    return;
  }
  const long long t0 = clock64();
  if (warp < warps) {
    const int per_warp = 32 / warps;  // gather4 ops per warp per stage (lanes < per_warp issue)
    const int base = blockIdx.x * 7919;
    for (int it = 0; it < iters; it++) {
      const int s = it % kStages;
      if (it >= kStages) mbar_wait(&full[s], ((it / kStages) - 1) & 1);
      if (warp == 0 && lane == 0) mbar_expect(&full[s], kStageBytes);
      // (the expect may race with completions of the other warps' copies of the same phase: tx counts are signed, fine)
      if (lane < per_warp) {
        const int g = warp * per_warp + lane;  // which 4-row group of the stage
        const int4 r = *reinterpret_cast<const int4*>(rows + ((base + (it * 32 + g)) % (n_idx / 4)) * 4);
        gather4(smem + s * kStageBytes + g * 512, &tmap, &full[s], 0, r.x, r.y, r.z, r.w);
      }
    }
    for (int s = 0; s < kStages; s++) {
      const int last = (iters - 1 - s) / kStages;  // last use of stage s ... wait for everything outstanding
      if (iters > s) mbar_wait(&full[s], last & 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

int main() {
  const int n_rows = 300000, cols = 128;
  std::vector<uint16_t> h((size_t)n_rows * cols);
  for (int r = 0; r < n_rows; r++)
    for (int c = 0; c < cols; c++) h[(size_t)r * cols + c] = (uint16_t)((r * 131 + c) & 0xffff);
  uint16_t* d_tab;
  CK(cudaMalloc(&d_tab, h.size() * 2));
  CK(cudaMemcpy(d_tab, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  const int n_idx = 1 << 20;
  std::vector<int> idx(n_idx);
  uint32_t s = 12345;
  for (int i = 0; i < n_idx; i++) { s = s * 1664525u + 1013904223u; idx[i] = (int)((s >> 8) % n_rows); }
  idx[1] = -1; idx[6] = n_rows + 5; idx[9] = 0x7fffffff;  // out-of-range rows must read as zeros
  int* d_idx;
  CK(cudaMalloc(&d_idx, n_idx * 4));
  CK(cudaMemcpy(d_idx, idx.data(), n_idx * 4, cudaMemcpyHostToDevice));

  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)n_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = ((Encode)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_tab, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)cr);
  if (cr != CUDA_SUCCESS) return 1;

  unsigned char* d_dump;
  long long* d_clk;
  CK(cudaMalloc(&d_dump, kStageBytes));
  CK(cudaMalloc(&d_clk, 148 * 8));
  const int smem = kStages * kStageBytes;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  // ---- layout: row i of group g lands at g*512 + i*128, 16-byte chunk c at ((c ^ (row_in_tile & 7)) * 16)
  probe<<<1, 128, smem>>>(tmap, d_idx, n_idx, 0, 1, 0, d_dump, d_clk);
  CK(cudaDeviceSynchronize());
  std::vector<unsigned char> dump(kStageBytes);
  CK(cudaMemcpy(dump.data(), d_dump, kStageBytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int t = 0; t < 128; t++) {
    const int r = idx[t];
    const bool oob = r < 0 || r >= n_rows;
    for (int c = 0; c < 8; c++)
      for (int e = 0; e < 8; e++) {
        const uint16_t got = *reinterpret_cast<uint16_t*>(&dump[t * 128 + ((c ^ (t & 7)) * 16) + e * 2]);
        const uint16_t want = oob ? 0 : h[(size_t)r * cols + 64 + c * 8 + e];
        if (got != want && bad++ < 5) printf("  mismatch tile row %d (src %d) chunk %d elem %d: got %u want %u\n", t, r, c, e, got, want);
      }
  }
  printf("layout check (swizzle-128B, row t at t*128, OOB rows zero): %s (%d mismatches)\n", bad ? "FAIL" : "ok", bad);

  // ---- throughput
  for (int warps : {1, 2, 4}) {
    for (int grid : {1, 148}) {
      const int iters = 2000;
      probe<<<grid, 128, smem>>>(tmap, d_idx, n_idx, iters, warps, 1, d_dump, d_clk);
      CK(cudaDeviceSynchronize());
      std::vector<long long> c(grid);
      CK(cudaMemcpy(c.data(), d_clk, grid * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (long long v : c) mx = v > mx ? v : mx;
      printf("gather4 x32 per 16 KB stage, %d issuing warp(s), %3d CTAs: %.0f clk per stage = %.1f B/clk/SM, %.1f clk per gather4\n",
             warps, grid, (double)mx / iters, (double)kStageBytes * iters / mx, (double)mx / iters / 32);
    }
  }
  return 0;
}
