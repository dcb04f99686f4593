// a3: SparseConvTensor.dense() -- (N,C) active rows -> zero-filled (B,C,D,H,W)
// (vision3d/detector/sparse_cnn.py:128-133; upstream: zero-fill + scatter_nd + permute).
//
// HBM-write bound: 4*B*C*D*H*W bytes must be written whatever the sparsity (18 MB per SECOND frame).
// Instead of memset + scattered 4-byte writes (each output line touched twice, the second time one
// float at a time), a small cell->row map is built first and the dense tensor is then written exactly
// once, in full 128-byte lines: a CTA owns 64 consecutive cells of one batch item, stages the active
// rows of those cells in shared memory (coalesced row reads), and each warp streams one channel's
// 64-cell segment at a time.
#include "common.cuh"

namespace v3d {
namespace {

constexpr int kCells = 64;

__global__ void __launch_bounds__(256) dense_map_kernel(const int4* __restrict__ idx, const int* __restrict__ n_rows,
                                                        int cap_rows, int D, int H, int W, int B,
                                                        int* __restrict__ cellmap) {
  const int n = min(*n_rows, cap_rows);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = idx[i];
    if ((unsigned)c.x < (unsigned)B && (unsigned)c.y < (unsigned)D && (unsigned)c.z < (unsigned)H &&
        (unsigned)c.w < (unsigned)W)
      cellmap[(((size_t)c.x * D + c.y) * H + c.z) * W + c.w] = i;
  }
}

__global__ void __launch_bounds__(256) dense_write_kernel(const float* __restrict__ feat,
                                                          const int* __restrict__ cellmap, int C, int vol,
                                                          float* __restrict__ out) {
  extern __shared__ float tile[];  // [kCells][C+1]
  __shared__ int rows[kCells];
  const int b = blockIdx.y;
  const int cell0 = blockIdx.x * kCells;
  const int ncell = min(kCells, vol - cell0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int any = 0;
  if (tid < kCells) {
    int r = tid < ncell ? __ldg(&cellmap[(size_t)b * vol + cell0 + tid]) : -1;
    rows[tid] = r;
    any = r >= 0;
  }
  any = __syncthreads_or(any);
  float* obase = out + (size_t)b * C * vol + cell0;
  if (!any) {  // warp-uniform for the whole CTA: stream zeros
    for (int c = warp; c < C; c += 8) {
      float* o = obase + (size_t)c * vol;
      if (lane < ncell) o[lane] = 0.f;
      if (lane + 32 < ncell) o[lane + 32] = 0.f;
    }
    return;
  }
  const int ld = C + 1;
  // stage: one warp per cell row, lanes along channels (coalesced reads of the C-float rows)
  for (int cidx = warp; cidx < kCells; cidx += 8) {
    const int r = rows[cidx];
    for (int c = lane; c < C; c += 32) tile[cidx * ld + c] = r >= 0 ? __ldg(&feat[(size_t)r * C + c]) : 0.f;
  }
  __syncthreads();
  for (int c = warp; c < C; c += 8) {
    float* o = obase + (size_t)c * vol;
    if (lane < ncell) o[lane] = tile[lane * ld + c];
    if (lane + 32 < ncell) o[lane + 32] = tile[(lane + 32) * ld + c];
  }
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" size_t v3d_sparse_to_dense_workspace_bytes(int B, const int* shape_host) {
  if (B <= 0 || !shape_host) return 0;
  return align_up(sizeof(int) * (size_t)B * shape_host[0] * shape_host[1] * shape_host[2], 256);
}

extern "C" int v3d_sparse_to_dense(const float* feat, const int* indices, const int* n_rows, int capacity_rows,
                                   int C, int B, const int* shape_host, float* out, void* workspace,
                                   size_t workspace_bytes, v3d_stream_t stream) {
  if (!feat || !indices || !n_rows || !shape_host || !out || !workspace) return V3D_ERR_INVALID_ARGUMENT;
  if (C <= 0 || B <= 0 || capacity_rows < 0) return V3D_ERR_INVALID_ARGUMENT;
  const long long vol = (long long)shape_host[0] * shape_host[1] * shape_host[2];
  if (vol <= 0 || vol * B >= (1ll << 31)) return V3D_ERR_INVALID_ARGUMENT;
  const size_t need = v3d_sparse_to_dense_workspace_bytes(B, shape_host);
  if (workspace_bytes < need) return V3D_ERR_WORKSPACE_TOO_SMALL;
  if (B > 65535) return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  int* cellmap = static_cast<int*>(workspace);
  V3D_CUDA_TRY(cudaMemsetAsync(cellmap, 0xFF, sizeof(int) * (size_t)B * vol, st));
  int blocks = ceil_div(capacity_rows > 0 ? capacity_rows : 1, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  dense_map_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const int4*>(indices), n_rows, capacity_rows,
                                          shape_host[0], shape_host[1], shape_host[2], B, cellmap);
  const size_t smem = sizeof(float) * kCells * (C + 1);
  if (smem > 96 * 1024) return V3D_ERR_INVALID_ARGUMENT;
  static bool attr_set = false;
  if (!attr_set) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(dense_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  dense_write_kernel<<<dim3((unsigned)ceil_div((int)vol, kCells), B), 256, smem, st>>>(feat, cellmap, C, (int)vol, out);
  return check_launch();
}
