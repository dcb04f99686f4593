// Glue kernels between the RPN heads and rotated NMS for the SECOND inference path
// (vision3d/detector/proposal.py:47-80 + core/box_encode.py:13-23 + ops/iou_nms.py:121-132), which the
// reference runs as ~55 tiny torch kernels per batch:
//   v3d_second_head_decode : gather the regression deltas and anchors of the top-k candidates, VoxelNet
//                            decode, BEV boxes, and the per-group coordinate-offset trick of
//                            batched_nms_rotated -- one single-CTA launch (a few thousand boxes).
//   v3d_pack_detections    : gather the kept boxes into the packed result rows the host reads back.
// Arithmetic mirrors the torch expressions operation by operation (fp32, no FMA contraction in the decode).
#include "common.cuh"

namespace v3d {
namespace {

struct HeadGeom {
  int B, n_cls, n_yaw, ny, nx, topk, dof;
  long long reg_sb, reg_sc, reg_sy, reg_sx;  // strides (elements) of the conv_reg output (B, CH, ny, nx)
};

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, v, d);
    v = is_max ? fmaxf(v, o) : fminf(v, o);
  }
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    float w = lane < nw ? sm[lane] : (is_max ? -INFINITY : INFINITY);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, w, d);
      w = is_max ? fmaxf(w, o) : fminf(w, o);
    }
    if (lane == 0) sm[32] = w;
  }
  __syncthreads();
  const float r = sm[32];
  __syncthreads();
  return r;
}

// candidate n = (b, c, k); a_idx indexes the flattened (n_yaw, ny, nx) anchor grid of class c
__global__ void __launch_bounds__(1024) head_decode_kernel(const float* __restrict__ reg, const float* __restrict__ anchors,
                                                           const long long* __restrict__ a_idx, HeadGeom G,
                                                           float* __restrict__ boxes, float* __restrict__ nms_in) {
  __shared__ float sm[33];
  const int N = G.B * G.n_cls * G.topk;
  float mx = -INFINITY, mn = INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int c = (n / G.topk) % G.n_cls, b = n / (G.topk * G.n_cls);
    long long a = a_idx[n];
    const int x = (int)(a % G.nx);
    a /= G.nx;
    const int y = (int)(a % G.ny);
    const int yaw = (int)(a / G.ny);
    const float* an = anchors + ((((size_t)c * G.n_yaw + yaw) * G.ny + y) * G.nx + x) * 7;
    float d[7];
#pragma unroll
    for (int k = 0; k < 7; k++)
      d[k] = reg[b * G.reg_sb + ((long long)(c * G.dof + k) * G.n_yaw + yaw) * G.reg_sc + y * G.reg_sy + x * G.reg_sx];
    // core/box_encode.py:13-23: xyz * [diag, diag, h] + xyz_a ; exp(wlh) * wlh_a ; yaw + yaw_a
    const float diag = sqrtf(__fadd_rn(__fmul_rn(an[3], an[3]), __fmul_rn(an[4], an[4])));
    float bx[7];
    bx[0] = __fadd_rn(__fmul_rn(d[0], diag), an[0]);
    bx[1] = __fadd_rn(__fmul_rn(d[1], diag), an[1]);
    bx[2] = __fadd_rn(__fmul_rn(d[2], an[5]), an[2]);
    bx[3] = __fmul_rn(expf(d[3]), an[3]);
    bx[4] = __fmul_rn(expf(d[4]), an[4]);
    bx[5] = __fmul_rn(expf(d[5]), an[5]);
    bx[6] = __fadd_rn(d[6], an[6]);
#pragma unroll
    for (int k = 0; k < 7; k++) boxes[(size_t)n * 7 + k] = bx[k];
    // BEV box (x, y, w, l, yaw) and the extrema of ops/iou_nms.py:121-126
    mx = fmaxf(mx, __fadd_rn(fmaxf(bx[0], bx[1]), fmaxf(bx[3], bx[4]) / 2.0f));
    mn = fminf(mn, __fsub_rn(fminf(bx[0], bx[1]), fminf(bx[3], bx[4]) / 2.0f));
  }
  const float gmax = block_reduce(mx, true, sm);
  const float gmin = block_reduce(mn, false, sm);
  const float span = __fadd_rn(__fsub_rn(gmax, gmin), 1.0f);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int c = (n / G.topk) % G.n_cls, b = n / (G.topk * G.n_cls);
    const float off = __fmul_rn((float)(c + G.n_cls * b), span);  // group id = class + n_cls * frame
    const float* bx = boxes + (size_t)n * 7;
    float* o = nms_in + (size_t)n * 5;
    o[0] = __fadd_rn(bx[0], off);
    o[1] = __fadd_rn(bx[1], off);
    o[2] = bx[3];
    o[3] = bx[4];
    o[4] = bx[6];
  }
}

// result rows: [7 box | score | frame | class | valid], row N = counters (kept count, then `n_counters` ints)
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                   const long long* __restrict__ keep, const int* __restrict__ count,
                                                   const float* __restrict__ thr, int N, int n_cls, int topk,
                                                   const int* const* __restrict__ counters, int n_counters,
                                                   float* __restrict__ result) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int kept = *count;
  if (i < N) {
    float* r = result + (size_t)i * 11;
    if (i < kept) {
      const long long k = keep[i];
      const int c = (int)((k / topk) % n_cls), b = (int)(k / ((long long)topk * n_cls));
      const float s = scores[k];
#pragma unroll
      for (int j = 0; j < 7; j++) r[j] = boxes[(size_t)k * 7 + j];
      r[7] = s;
      r[8] = (float)b;
      r[9] = (float)c;
      r[10] = s > thr[c] ? 1.0f : 0.0f;  // per-class score threshold (proposal.py:41-45,57-58)
    } else {
#pragma unroll
      for (int j = 0; j < 11; j++) r[j] = 0.f;
    }
  }
  if (i == 0) {
    float* r = result + (size_t)N * 11;
    r[0] = (float)kept;
    for (int j = 0; j < n_counters && j < 10; j++) r[1 + j] = (float)(*counters[j]);
  }
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" int v3d_second_head_decode(const float* reg_map, const long long* reg_strides_host, const float* anchors,
                                      const int64_t* anchor_idx, int B, int n_cls, int n_yaw, int ny, int nx,
                                      int topk, float* boxes, float* nms_in, v3d_stream_t stream) {
  if (!reg_map || !reg_strides_host || !anchors || !anchor_idx || !boxes || !nms_in) return V3D_ERR_INVALID_ARGUMENT;
  if (B <= 0 || n_cls <= 0 || n_yaw <= 0 || ny <= 0 || nx <= 0 || topk <= 0) return V3D_ERR_INVALID_ARGUMENT;
  HeadGeom G{B, n_cls, n_yaw, ny, nx, topk, 7, reg_strides_host[0], reg_strides_host[1], reg_strides_host[2],
             reg_strides_host[3]};
  head_decode_kernel<<<1, 1024, 0, as_stream(stream)>>>(reg_map, anchors,
                                                        reinterpret_cast<const long long*>(anchor_idx), G, boxes, nms_in);
  return check_launch();
}

extern "C" int v3d_pack_detections(const float* boxes, const float* scores, const int64_t* keep, const int* count,
                                   const float* score_thresh, int N, int n_cls, int topk,
                                   const int* const* counters_dev, int n_counters, float* result,
                                   v3d_stream_t stream) {
  if (!boxes || !scores || !keep || !count || !score_thresh || !result || N <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (n_counters > 0 && !counters_dev) return V3D_ERR_INVALID_ARGUMENT;
  pack_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(boxes, scores, reinterpret_cast<const long long*>(keep),
                                                              count, score_thresh, N, n_cls, topk, counters_dev,
                                                              n_counters, result);
  return check_launch();
}
