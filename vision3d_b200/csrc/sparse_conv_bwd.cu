// backward of the sparse convolution (spconv SparseConvFunction / SubMConvFunction.backward, the
// torch.autograd.Functions behind SubMConv3d / SparseConv3d in vision3d/detector/sparse_cnn.py:15-30) so that the
// reference's training step (vision3d/train.py:57-72) runs on the drop-in. Exact fp32 (SIMT), reusing the forward
// rule table nbr[k][o] = input row:
//   dX[i] = sum_k dY[inv[k][i]] * W[k]^T   -- the FORWARD kernel on the inverted table inv[k][i] = o (for a fixed
//           offset the map o -> i is injective) with per-offset transposed weights; for SubM layers the inverse is
//           the mirror inv[k] = nbr[KV-1-k], for strided layers v3d_rulebook_invert scatters it;
//   dW[k] = sum_o X[nbr[k][o]]^T dY[o]      -- v3d_sparse_conv_bwd_weight: per (offset, 1024-row chunk) CTA, gathered
//           rows and dY rows staged through shared memory, a (Cin x Cout) register-tiled outer-product accumulation,
//           one atomicAdd per weight element per CTA.
#include "common.cuh"

namespace v3d {
namespace {

__global__ void __launch_bounds__(256) rule_invert_kernel(const int* __restrict__ nbr, int nbr_stride,
                                                          const int* __restrict__ n_out_ptr, int out_cap, int KV,
                                                          int* __restrict__ inv, int inv_stride) {
  const int n = min(*n_out_ptr, out_cap);
  const long long total = (long long)KV * n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / n), o = (int)(e % n);
    const int i = __ldg(&nbr[(size_t)k * nbr_stride + o]);
    if (i >= 0 && i < inv_stride) inv[(size_t)k * inv_stride + i] = o;
  }
}

constexpr int kBwRows = 32;      // rows staged per step
constexpr int kBwChunk = 1024;   // output rows per CTA

// thread t owns output channels co4 .. co4+3 (co4 = 4 * (t % (Cout/4))) and input channels ci0 + j * ci_step
template <int J>
__global__ void __launch_bounds__(256) conv_bwd_weight_kernel(const float* __restrict__ feat, const float* __restrict__ gout,
                                                              const int* __restrict__ nbr, int nbr_stride,
                                                              const int* __restrict__ n_out_ptr, int out_cap, int Cin,
                                                              int Cout, float* __restrict__ gw) {
  extern __shared__ __align__(16) float sm[];
  float* sF = sm;                        // [kBwRows][Cin]
  float* sG = sF + kBwRows * Cin;        // [kBwRows][Cout]
  __shared__ int sIdx[kBwRows];
  const int n = min(*n_out_ptr, out_cap);
  const int k = blockIdx.y;
  const int r_begin = blockIdx.x * kBwChunk, r_end = min(r_begin + kBwChunk, n);
  if (r_begin >= n) return;
  const int tid = threadIdx.x;
  const int tpr = Cout / 4, ci_step = 256 / tpr;
  const int co4 = 4 * (tid % tpr), ci0 = tid / tpr;
  float acc[J][4];
#pragma unroll
  for (int j = 0; j < J; j++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[j][c] = 0.f;
  const int* nb = nbr + (size_t)k * nbr_stride;
  bool any = false;
  for (int r0 = r_begin; r0 < r_end; r0 += kBwRows) {
    const int nr = min(kBwRows, r_end - r0);
    __syncthreads();
    if (tid < kBwRows) sIdx[tid] = tid < nr ? __ldg(&nb[r0 + tid]) : -1;
    __syncthreads();
    bool live = false;
    for (int r = 0; r < nr; r++) live = live || sIdx[r] >= 0;
    if (!live) continue;  // (block-uniform)
    any = true;
    for (int e = tid; e < kBwRows * Cin; e += 256) {
      const int r = e / Cin, c = e % Cin;
      const int src = sIdx[r];
      sF[e] = src >= 0 ? __ldg(&feat[(size_t)src * Cin + c]) : 0.f;
    }
    for (int e = tid; e < kBwRows * Cout; e += 256) {
      const int r = e / Cout, c = e % Cout;
      sG[e] = (r < nr && sIdx[r] >= 0) ? __ldg(&gout[(size_t)(r0 + r) * Cout + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kBwRows; r++) {
      const float4 g = *reinterpret_cast<const float4*>(&sG[r * Cout + co4]);
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int ci = ci0 + j * ci_step;
        const float f = ci < Cin ? sF[r * Cin + ci] : 0.f;
        acc[j][0] = fmaf(f, g.x, acc[j][0]);
        acc[j][1] = fmaf(f, g.y, acc[j][1]);
        acc[j][2] = fmaf(f, g.z, acc[j][2]);
        acc[j][3] = fmaf(f, g.w, acc[j][3]);
      }
    }
  }
  if (!any) return;
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int ci = ci0 + j * ci_step;
    if (ci < Cin) {
      float* w = gw + ((size_t)k * Cin + ci) * Cout + co4;
#pragma unroll
      for (int c = 0; c < 4; c++) atomicAdd(&w[c], acc[j][c]);
    }
  }
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" int v3d_rulebook_invert(const int* nbr, int nbr_stride, const int* n_out, int out_capacity, int kernel_volume,
                                   int* inv, int inv_stride, v3d_stream_t stream) {
  if (!nbr || !n_out || !inv || out_capacity <= 0 || kernel_volume <= 0 || nbr_stride < out_capacity || inv_stride <= 0)
    return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  V3D_CUDA_TRY(cudaMemsetAsync(inv, 0xFF, sizeof(int) * (size_t)kernel_volume * inv_stride, st));
  const long long total = (long long)kernel_volume * out_capacity;
  const long long want = (total + 255) / 256;
  rule_invert_kernel<<<(int)(want < kNumSMs * 8 ? want : kNumSMs * 8), 256, 0, st>>>(nbr, nbr_stride, n_out, out_capacity,
                                                                                   kernel_volume, inv, inv_stride);
  return check_launch();
}

extern "C" int v3d_sparse_conv_bwd_weight(const float* feat, const float* grad_out, const int* nbr, int nbr_stride,
                                          const int* n_out, int out_capacity, int kernel_volume, int Cin, int Cout,
                                          float* grad_weight, v3d_stream_t stream) {
  if (!feat || !grad_out || !nbr || !n_out || !grad_weight) return V3D_ERR_INVALID_ARGUMENT;
  if (out_capacity <= 0 || kernel_volume <= 0 || kernel_volume > 65535 || nbr_stride < out_capacity) return V3D_ERR_INVALID_ARGUMENT;
  if (Cin <= 0 || Cout <= 0 || (Cout & 3) || Cout > 256 || 256 % (Cout / 4)) return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  V3D_CUDA_TRY(cudaMemsetAsync(grad_weight, 0, sizeof(float) * (size_t)kernel_volume * Cin * Cout, st));
  const int ci_step = 256 / (Cout / 4);
  const int J = ceil_div(Cin, ci_step);
  const size_t smem = sizeof(float) * kBwRows * (size_t)(Cin + Cout);
  if (smem > 48 * 1024 || J > 8) return V3D_ERR_INVALID_ARGUMENT;
  dim3 grid(ceil_div(out_capacity, kBwChunk), kernel_volume);
#define V3D_BW(JJ)                                                                                                  \
  conv_bwd_weight_kernel<JJ><<<grid, 256, smem, st>>>(feat, grad_out, nbr, nbr_stride, n_out, out_capacity, Cin, Cout, \
                                                      grad_weight)
  if (J <= 1) V3D_BW(1);
  else if (J <= 2) V3D_BW(2);
  else if (J <= 4) V3D_BW(4);
  else V3D_BW(8);
#undef V3D_BW
  return check_launch();
}
