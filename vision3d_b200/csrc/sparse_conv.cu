// a5 + a6: sparse convolution forward (SubMConv3d / SparseConv3d) with the eval-mode BatchNorm1d +
// ReLU that follows it in every layer of vision3d/detector/sparse_cnn.py:15-30 folded in.
//
// Upstream (spconv v1.x indice_conv) runs, per kernel offset, gather -> cuBLAS SGEMM -> scatter-add:
// ~3 launches x 27 offsets per layer. Here one launch per layer: each CTA owns a tile of output rows
// (output-stationary), walks the kernel offsets, gathers the neighbour rows named by the rule table
// into shared memory, accumulates in registers and writes every output row exactly once with the
// per-channel affine + ReLU applied -- no atomics, no zero-init of the output, deterministic.
//
// This file holds the exact-fp32 SIMT path (FFMA). It is the numerical reference for, and the
// small-channel (Cin < 16) companion of, the tcgen05 3xTF32 path in sparse_conv_tc.cu.
#include "common.cuh"

namespace v3d {
namespace {

// Thread tile: 4 output rows x 4 output channels. COUT/4 threads span a row group.
template <int COUT>
struct SimtCfg {
  static constexpr int kThreads = 256;
  static constexpr int kTPR = COUT / 4;           // threads per row group
  static constexpr int kGroups = kThreads / kTPR;  // row groups per CTA
  static constexpr int kRows = kGroups * 4;        // output rows per tile: 64 / 128 / 256
};

template <int COUT>
__global__ void __launch_bounds__(256) sparse_conv_simt_kernel(
    const float* __restrict__ feat, const float* __restrict__ weight, const int* __restrict__ nbr,
    int nbr_stride, const int* __restrict__ n_out_ptr, int out_cap, int KV, int Cin,
    const float* __restrict__ scale, const float* __restrict__ shift, int relu, float* __restrict__ out) {
  using C = SimtCfg<COUT>;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                         // [Cin][kRows]  (k-major: 4 consecutive rows = one float4)
  float* sW = sA + (size_t)Cin * C::kRows;  // [Cin][COUT]
  int* sIdx = reinterpret_cast<int*>(sW + (size_t)Cin * COUT);  // [kRows]

  const int n_out = min(*n_out_ptr, out_cap);
  const int n_tiles = ceil_div(n_out, C::kRows);
  const int tid = threadIdx.x;
  const int tc = tid % C::kTPR;  // channel quad
  const int rg = tid / C::kTPR;  // row group
  const int cin4 = Cin >> 2;

  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    sc[j] = scale ? scale[tc * 4 + j] : 1.0f;
    sh[j] = shift ? shift[tc * 4 + j] : 0.0f;
  }

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * C::kRows;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[r][j] = 0.f;

    for (int kk = 0; kk < KV; kk++) {
      // rule tile for this offset (coalesced) + block-wide emptiness vote
      int any = 0;
      for (int r = tid; r < C::kRows; r += C::kThreads) {
        int o = row0 + r;
        int v = o < n_out ? __ldg(&nbr[(size_t)kk * nbr_stride + o]) : -1;
        sIdx[r] = v;
        any |= (v >= 0);
      }
      if (!__syncthreads_or(any)) continue;  // nobody in the tile has a neighbour at this offset
      // gather neighbour rows (float4 chunks), transposed into the k-major tile; zero when absent
      for (int e = tid; e < C::kRows * cin4; e += C::kThreads) {
        const int r = e % C::kRows, c4 = e / C::kRows;
        const int src = sIdx[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) v = __ldg(reinterpret_cast<const float4*>(feat + (size_t)src * Cin) + c4);
        sA[(c4 * 4 + 0) * C::kRows + r] = v.x;
        sA[(c4 * 4 + 1) * C::kRows + r] = v.y;
        sA[(c4 * 4 + 2) * C::kRows + r] = v.z;
        sA[(c4 * 4 + 3) * C::kRows + r] = v.w;
      }
      {
        const float4* wsrc = reinterpret_cast<const float4*>(weight + (size_t)kk * Cin * COUT);
        float4* wdst = reinterpret_cast<float4*>(sW);
        for (int e = tid; e < Cin * COUT / 4; e += C::kThreads) wdst[e] = __ldg(wsrc + e);
      }
      __syncthreads();
#pragma unroll 4
      for (int ci = 0; ci < Cin; ci++) {
        const float4 a = *reinterpret_cast<const float4*>(&sA[ci * C::kRows + rg * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&sW[ci * COUT + tc * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[r][j] = fmaf(av[r], bv[j], acc[r][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int o = row0 + rg * 4 + r;
      if (o < n_out) {
        float4 v;
        v.x = fmaf(acc[r][0], sc[0], sh[0]);
        v.y = fmaf(acc[r][1], sc[1], sh[1]);
        v.z = fmaf(acc[r][2], sc[2], sh[2]);
        v.w = fmaf(acc[r][3], sc[3], sh[3]);
        if (relu) {
          v.x = fmaxf(v.x, 0.f);
          v.y = fmaxf(v.y, 0.f);
          v.z = fmaxf(v.z, 0.f);
          v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4*>(out + (size_t)o * COUT + tc * 4) = v;
      }
    }
  }
}

template <int COUT>
int launch_simt(const float* feat, const float* weight, const int* nbr, int nbr_stride, const int* n_out,
                int out_cap, int KV, int Cin, const float* scale, const float* shift, int relu, float* out,
                cudaStream_t st) {
  using C = SimtCfg<COUT>;
  const size_t smem = sizeof(float) * ((size_t)Cin * C::kRows + (size_t)Cin * COUT) + sizeof(int) * C::kRows;
  if (smem > 200 * 1024) return V3D_ERR_INVALID_ARGUMENT;
  static PerDeviceOnce attr_once;
  if (attr_once.needed()) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(sparse_conv_simt_kernel<COUT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.done();
  }
  const int tiles_cap = ceil_div(out_cap, C::kRows);
  const int grid = tiles_cap < kNumSMs * 2 ? (tiles_cap > 0 ? tiles_cap : 1) : kNumSMs * 2;
  sparse_conv_simt_kernel<COUT><<<grid, C::kThreads, smem, st>>>(feat, weight, nbr, nbr_stride, n_out,
                                                                out_cap, KV, Cin, scale, shift, relu, out);
  return check_launch();
}

}  // namespace

int sparse_conv_simt(const float* feat, const float* weight, const int* nbr, int nbr_stride, const int* n_out,
                     int out_cap, int KV, int Cin, int Cout, const float* scale, const float* shift, int relu,
                     float* out, cudaStream_t st) {
  if (Cin <= 0 || (Cin & 3)) return V3D_ERR_INVALID_ARGUMENT;
  switch (Cout) {
    case 4: return launch_simt<4>(feat, weight, nbr, nbr_stride, n_out, out_cap, KV, Cin, scale, shift, relu, out, st);
    case 8: return launch_simt<8>(feat, weight, nbr, nbr_stride, n_out, out_cap, KV, Cin, scale, shift, relu, out, st);
    case 16: return launch_simt<16>(feat, weight, nbr, nbr_stride, n_out, out_cap, KV, Cin, scale, shift, relu, out, st);
    case 32: return launch_simt<32>(feat, weight, nbr, nbr_stride, n_out, out_cap, KV, Cin, scale, shift, relu, out, st);
    case 64: return launch_simt<64>(feat, weight, nbr, nbr_stride, n_out, out_cap, KV, Cin, scale, shift, relu, out, st);
    case 128: return launch_simt<128>(feat, weight, nbr, nbr_stride, n_out, out_cap, KV, Cin, scale, shift, relu, out, st);
    default: return V3D_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace v3d

using namespace v3d;

extern "C" int v3d_sparse_conv_fwd(const float* feat, const float* weight, const int* nbr, int nbr_stride,
                                   const int* n_out, int out_capacity, int kernel_volume, int Cin, int Cout,
                                   const float* scale, const float* shift, int relu, float* out,
                                   v3d_stream_t stream) {
  if (!feat || !weight || !nbr || !n_out || !out) return V3D_ERR_INVALID_ARGUMENT;
  if (out_capacity <= 0 || kernel_volume <= 0 || nbr_stride < out_capacity) return V3D_ERR_INVALID_ARGUMENT;
  if ((scale == nullptr) != (shift == nullptr)) return V3D_ERR_INVALID_ARGUMENT;
  return sparse_conv_simt(feat, weight, nbr, nbr_stride, n_out, out_capacity, kernel_volume, Cin, Cout, scale,
                          shift, relu, out, as_stream(stream));
}
