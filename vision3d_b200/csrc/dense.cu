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

// BEV map in channels-last memory: out[b][y][x][c*D + d] = feat[row(b,d,y,x)][c] (0 when the cell is
// inactive) == the memory of `dense().view(B, C*D, H, W)` (sparse_cnn.py:128-133) in torch's
// channels_last format, which is what cuDNN's sm_100 implicit-GEMM kernels consume natively (the NCHW
// tensor costs a 130 us NCHW->NHWC transpose per RPN layer at batch 16). One warp covers 32 consecutive
// channels of one pixel: coalesced 128 B row reads, contiguous D*128 B writes.
template <int D>
__global__ void __launch_bounds__(256) dense_nhwc_kernel(const float* __restrict__ feat,
                                                         const int* __restrict__ cellmap, int C, int HW,
                                                         int pixels_per_block, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = threadIdx.x % C;
  const int pl = threadIdx.x / C;
  const int per_iter = blockDim.x / C;
  for (int p = blockIdx.x * pixels_per_block + pl; p < min(HW, (int)(blockIdx.x + 1) * pixels_per_block);
       p += per_iter) {
    float v[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
      const int r = __ldg(&cellmap[((size_t)b * D + d) * HW + p]);
      v[d] = r >= 0 ? __ldg(&feat[(size_t)r * C + c]) : 0.f;
    }
    float* o = out + (((size_t)b * HW + p) * C + c) * D;
    if (D == 2) {
      *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
      for (int d = 0; d < D; d++) o[d] = v[d];
    }
  }
}

// D == 2 (SECOND's final level): one thread owns two channels of one pixel and writes (c,d0)(c,d1)(c+1,d0)(c+1,d1)
// as one 16-byte store -> a warp streams 512 contiguous bytes.
__global__ void __launch_bounds__(256) dense_nhwc_d2_kernel(const float* __restrict__ feat,
                                                            const int* __restrict__ cellmap, int C, int HW,
                                                            int pixels_per_block, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int half_c = C >> 1;
  const int c2 = threadIdx.x % half_c;
  const int pl = threadIdx.x / half_c;
  const int per_iter = blockDim.x / half_c;
  const int p_end = min(HW, (int)(blockIdx.x + 1) * pixels_per_block);
  for (int p = blockIdx.x * pixels_per_block + pl; p < p_end; p += per_iter) {
    const int r0 = __ldg(&cellmap[((size_t)b * 2 + 0) * HW + p]);
    const int r1 = __ldg(&cellmap[((size_t)b * 2 + 1) * HW + p]);
    const float2 f0 = r0 >= 0 ? __ldg(reinterpret_cast<const float2*>(feat + (size_t)r0 * C) + c2) : make_float2(0.f, 0.f);
    const float2 f1 = r1 >= 0 ? __ldg(reinterpret_cast<const float2*>(feat + (size_t)r1 * C) + c2) : make_float2(0.f, 0.f);
    reinterpret_cast<float4*>(out + ((size_t)b * HW + p) * C * 2)[c2] = make_float4(f0.x, f1.x, f0.y, f1.y);
  }
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" int v3d_sparse_to_dense_nhwc(const float* feat, const int* indices, const int* n_rows,
                                        int capacity_rows, int C, int B, const int* shape_host, float* out,
                                        void* workspace, size_t workspace_bytes, v3d_stream_t stream) {
  if (!feat || !indices || !n_rows || !shape_host || !out || !workspace) return V3D_ERR_INVALID_ARGUMENT;
  if (C <= 0 || C > 256 || (256 % C) != 0 || B <= 0 || B > 65535 || capacity_rows < 0) return V3D_ERR_INVALID_ARGUMENT;
  const int D = shape_host[0];
  const long long HW = (long long)shape_host[1] * shape_host[2];
  const long long vol = HW * D;
  if (vol <= 0 || vol * B >= (1ll << 31) || D > 8) return V3D_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < v3d_sparse_to_dense_workspace_bytes(B, shape_host)) return V3D_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = as_stream(stream);
  int* cellmap = static_cast<int*>(workspace);
  V3D_CUDA_TRY(cudaMemsetAsync(cellmap, 0xFF, sizeof(int) * (size_t)B * vol, st));
  int blocks = ceil_div(capacity_rows > 0 ? capacity_rows : 1, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  dense_map_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const int4*>(indices), n_rows, capacity_rows,
                                          shape_host[0], shape_host[1], shape_host[2], B, cellmap);
  const int ppb = 16 * (256 / C);  // pixels per block
  dim3 grid((unsigned)ceil_div((int)HW, ppb), B);
  if (D == 2 && (C % 2) == 0 && (256 % (C / 2)) == 0) {
    dense_nhwc_d2_kernel<<<grid, 256, 0, st>>>(feat, cellmap, C, (int)HW, ppb, out);
    return check_launch();
  }
  switch (D) {
#define V3D_NHWC_CASE(DD) \
  case DD:                \
    dense_nhwc_kernel<DD><<<grid, 256, 0, st>>>(feat, cellmap, C, (int)HW, ppb, out); \
    break;
    V3D_NHWC_CASE(1)
    V3D_NHWC_CASE(2)
    V3D_NHWC_CASE(3)
    V3D_NHWC_CASE(4)
    V3D_NHWC_CASE(5)
    V3D_NHWC_CASE(6)
    V3D_NHWC_CASE(7)
    V3D_NHWC_CASE(8)
#undef V3D_NHWC_CASE
  }
  return check_launch();
}

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
  static PerDeviceOnce attr_once;
  if (attr_once.needed()) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(dense_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_once.done();
  }
  dense_write_kernel<<<dim3((unsigned)ceil_div((int)vol, kCells), B), 256, smem, st>>>(feat, cellmap, C, (int)vol, out);
  return check_launch();
}
