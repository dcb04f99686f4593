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
__global__ void __launch_bounds__(1024) head_decode_kernel(const float* __restrict__ reg, const float* __restrict__ deltas,
                                                           const float* __restrict__ anchors,
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
    for (int k = 0; k < 7; k++)  // regression deltas: compact (N,7) if given, else gathered from the conv_reg map
      d[k] = deltas ? deltas[(size_t)n * 7 + k]
                    : reg[b * G.reg_sb + ((long long)(c * G.dof + k) * G.n_yaw + yaw) * G.reg_sc + y * G.reg_sy + x * G.reg_sx];
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


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// 1x1 classification head on the channels_last RPN map (detector/proposal.py:61-63, nn.Conv2d(128, n, 1)):
// a warp reads one pixel's 128 channels as one coalesced 512-byte row (lane = 4 channels), NO dot products,
// butterfly reduction. Memory bound: the map is read exactly once, only NO floats per pixel are written.
// logits layout (B, NO, hw) == conv output NCHW.
template <int NO>
__global__ void __launch_bounds__(256) cls_logits_kernel(const float* __restrict__ fmap, long long pixels, int hw,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  float4 wr[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) wr[o] = __ldg(reinterpret_cast<const float4*>(w + o * 128) + lane);
  const float my_bias = (lane < NO && bias) ? __ldg(&bias[lane]) : 0.f;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int U = 4;  // pixels in flight per warp
  for (long long p0 = warp0 * U; p0 < pixels; p0 += n_warps * U) {
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; u++)
      x[u] = p0 + u < pixels ? __ldg(reinterpret_cast<const float4*>(fmap + (p0 + u) * 128) + lane) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; u++) {
      float mine = 0.f;
#pragma unroll
      for (int o = 0; o < NO; o++) {
        float s = x[u].x * wr[o].x;
        s = fmaf(x[u].y, wr[o].y, s);
        s = fmaf(x[u].z, wr[o].z, s);
        s = fmaf(x[u].w, wr[o].w, s);
        s = warp_sum(s);
        if (lane == o) mine = s;
      }
      const long long p = p0 + u;
      if (lane < NO && p < pixels) {
        const long long b = p / hw, pos = p - b * hw;
        logits[(b * NO + lane) * hw + pos] = mine + my_bias;
      }
    }
  }
}

// Row-wise top-k (values sorted descending, ties -> lower index first): one CTA per row, MSB radix select on the
// order-preserving key with warp-aggregated shared-memory histograms, tie resolution on the index bits only when
// the k-th value is not unique, then a bitonic sort of the k selected (key, index) pairs.
constexpr int kTopkMax = 256;

__device__ __forceinline__ unsigned int order_key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// `segs` > 1: the launch treats every row as `segs` consecutive segments of L elements (one CTA each) and
// reports indices relative to the whole row; `idx_map` != null: the reported index is idx_map[row * L + i]
// (second stage over the first stage's candidates).
__global__ void __launch_bounds__(1024) topk_rows_kernel(const float* __restrict__ values, int L, int k,
                                                         float* __restrict__ out_v, long long* __restrict__ out_i, int segs,
                                                         const long long* __restrict__ idx_map) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_sel[3];   // [0] selected digit, [1] remaining, [2] 1 = the selection is already exact
  __shared__ unsigned long long s_pairs[kTopkMax];
  __shared__ unsigned int s_count;
  const float* row = values + (size_t)blockIdx.x * L;
  const int tid = threadIdx.x, lane = tid & 31;

  // select over the 64-bit composite (key << 32 | ~index): all composites are distinct, so exactly k survive.
  // Digits are taken from the 32 key bits first; index bits only if the k-th key is tied.
  unsigned long long prefix = 0ull, mask = 0ull;
  unsigned int remaining = (unsigned int)k;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    constexpr int U = 8;  // loads in flight per thread: one CTA streams its whole row every pass
    for (int i0 = 0; i0 < L; i0 += U * blockDim.x) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = i0 + u * blockDim.x + tid;
        v[u] = i < L ? __ldg(row + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = i0 + u * blockDim.x + tid;
        bool in = false;
        unsigned int digit = 0;
        if (i < L) {
          const unsigned long long comp = ((unsigned long long)order_key(v[u]) << 32) | (unsigned int)(~i);
          in = (comp & mask) == prefix;
          digit = (unsigned int)(comp >> shift) & 255u;
        }
        // warp-aggregated histogram update (values cluster: without it one bin takes tens of thousands of atomics)
        const unsigned int act = __ballot_sync(0xffffffffu, in);
        if (in) {
          const unsigned int peers = __match_any_sync(act, digit);
          if (lane == (__ffs(peers) - 1)) atomicAdd(&hist[digit], __popc(peers));
        }
      }
    }
    __syncthreads();
    if (tid < 32) {  // find the digit d with  count(> d) < remaining <= count(>= d)
      unsigned int acc = 0, sel = 0, rem = remaining;
      bool found = false;
      for (int base = 224; base >= 0 && !found; base -= 32) {
        const unsigned int c = hist[base + lane];
        // suffix sums inside the 32-bin group, from the top lane down
        unsigned int suf = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned int o = __shfl_down_sync(0xffffffffu, suf, d);
          if (lane + d < 32) suf += o;
        }
        const unsigned int tot = __shfl_sync(0xffffffffu, suf, 0);
        if (acc + tot >= remaining) {
          const unsigned int above = acc + suf - c;  // count of digits strictly greater than this lane's digit
          const bool hit = above < remaining && above + c >= remaining;
          const unsigned int m = __ballot_sync(0xffffffffu, hit);
          const int l = 31 - __clz(m);  // highest such lane (unique in fact)
          sel = (unsigned int)(base + l);
          rem = remaining - __shfl_sync(0xffffffffu, above, l);
          found = true;
        } else {
          acc += tot;
        }
      }
      if (lane == 0) {
        s_sel[0] = sel;
        s_sel[1] = rem;
        s_sel[2] = hist[sel] == rem ? 1u : 0u;  // every candidate left in the chosen bin is taken: stop refining
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)s_sel[0] << shift;
    mask |= 255ull << shift;
    remaining = s_sel[1];
    const bool exact = s_sel[2] != 0u;
    __syncthreads();
    if (exact) break;
  }
  // every composite whose refined digits are >= the boundary's is selected: exactly k of them
  if (tid == 0) s_count = 0u;
  __syncthreads();
  for (int i0 = 0; i0 < L; i0 += 8 * blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = i0 + u * blockDim.x + tid;
      v[u] = i < L ? __ldg(row + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = i0 + u * blockDim.x + tid;
      const unsigned long long comp = ((unsigned long long)order_key(v[u]) << 32) | (unsigned int)(~i);
      if (i < L && (comp & mask) >= prefix) {
        const unsigned int slot = atomicAdd(&s_count, 1u);
        if (slot < (unsigned int)kTopkMax) s_pairs[slot] = comp;
      }
    }
  }
  __syncthreads();
  for (int i = (int)s_count + tid; i < kTopkMax; i += blockDim.x) s_pairs[i] = 0ull;
  __syncthreads();
  // bitonic sort, descending, kTopkMax elements
  for (int size = 2; size <= kTopkMax; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < kTopkMax / 2) {
        const int lo = 2 * tid - (tid & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = s_pairs[lo], b = s_pairs[hi];
        if ((a < b) == desc) {
          s_pairs[lo] = b;
          s_pairs[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  if (tid < k) {
    const unsigned long long comp = s_pairs[tid];
    const unsigned int idx = ~(unsigned int)(comp & 0xffffffffull);
    out_i[(size_t)blockIdx.x * k + tid] = idx_map ? idx_map[(size_t)blockIdx.x * L + idx]
                                                  : (long long)idx + (long long)(blockIdx.x % segs) * L;
    out_v[(size_t)blockIdx.x * k + tid] = __ldg(row + idx);
  }
}

// Regression head evaluated ONLY at the top-k candidates (the reference convolves the whole map and gathers
// k rows out of 70 400): one warp per candidate reads the pixel's 128 channels and forms the 7 deltas; the same
// warp turns the candidate's logit into its score, 1 / (1 + exp(-x)) like torch.sigmoid.
__global__ void __launch_bounds__(256) reg_gather_kernel(const float* __restrict__ fmap, const float* __restrict__ w_reg,
                                                         const float* __restrict__ b_reg,
                                                         const float* __restrict__ top_logits,
                                                         const long long* __restrict__ a_idx, int N, int n_cls, int n_yaw,
                                                         int ny, int nx, int topk, float* __restrict__ deltas,
                                                         float* __restrict__ scores) {
  const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const int c = (n / topk) % n_cls, b = n / (topk * n_cls);
  long long a = a_idx[n];
  const int x = (int)(a % nx);
  a /= nx;
  const int y = (int)(a % ny);
  const int yaw = (int)(a / ny);
  const float4 f = __ldg(reinterpret_cast<const float4*>(fmap + (((size_t)b * ny + y) * nx + x) * 128) + lane);
  float mine = 0.f;
#pragma unroll
  for (int k = 0; k < 7; k++) {
    const int ch = (c * 7 + k) * n_yaw + yaw;  // conv_reg channel (proposal.py:20-26 reshape order)
    const float4 w = __ldg(reinterpret_cast<const float4*>(w_reg + (size_t)ch * 128) + lane);
    float s = f.x * w.x;
    s = fmaf(f.y, w.y, s);
    s = fmaf(f.z, w.z, s);
    s = fmaf(f.w, w.w, s);
    s = warp_sum(s);
    if (lane == k) mine = s + (b_reg ? __ldg(&b_reg[ch]) : 0.f);
  }
  if (lane < 7) deltas[(size_t)n * 7 + lane] = mine;
  if (lane == 7) scores[n] = 1.0f / (1.0f + expf(-top_logits[n]));
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
  head_decode_kernel<<<1, 1024, 0, as_stream(stream)>>>(reg_map, nullptr, anchors,
                                                        reinterpret_cast<const long long*>(anchor_idx), G, boxes, nms_in);
  return check_launch();
}

extern "C" int v3d_second_head_decode_compact(const float* deltas, const float* anchors, const int64_t* anchor_idx, int B,
                                              int n_cls, int n_yaw, int ny, int nx, int topk, float* boxes,
                                              float* nms_in, v3d_stream_t stream) {
  if (!deltas || !anchors || !anchor_idx || !boxes || !nms_in) return V3D_ERR_INVALID_ARGUMENT;
  if (B <= 0 || n_cls <= 0 || n_yaw <= 0 || ny <= 0 || nx <= 0 || topk <= 0) return V3D_ERR_INVALID_ARGUMENT;
  HeadGeom G{B, n_cls, n_yaw, ny, nx, topk, 7, 0, 0, 0, 0};
  head_decode_kernel<<<1, 1024, 0, as_stream(stream)>>>(nullptr, deltas, anchors,
                                                        reinterpret_cast<const long long*>(anchor_idx), G, boxes, nms_in);
  return check_launch();
}

extern "C" int v3d_head_cls_logits(const float* fmap_nhwc, int B, int hw, int C, const float* weight, const float* bias,
                                   int n_out, float* logits, v3d_stream_t stream) {
  if (!fmap_nhwc || !weight || !logits || B <= 0 || hw <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (C != 128 || n_out <= 0 || n_out > 8) return V3D_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(fmap_nhwc) & 15) return V3D_ERR_INVALID_ARGUMENT;
  const long long pixels = (long long)B * hw;
  const int blocks = (int)((pixels / 16 + 7) / 8 < (long long)kNumSMs * 8 ? (pixels / 16 + 7) / 8 : (long long)kNumSMs * 8);
  cudaStream_t st = as_stream(stream);
#define V3D_CLS_CASE(NO)                                                                                   \
  if (n_out == NO) {                                                                                       \
    cls_logits_kernel<NO><<<blocks > 0 ? blocks : 1, 256, 0, st>>>(fmap_nhwc, pixels, hw, weight, bias, logits); \
    return check_launch();                                                                                 \
  }
  V3D_CLS_CASE(1) V3D_CLS_CASE(2) V3D_CLS_CASE(3) V3D_CLS_CASE(4) V3D_CLS_CASE(5) V3D_CLS_CASE(6) V3D_CLS_CASE(7)
  V3D_CLS_CASE(8)
#undef V3D_CLS_CASE
  return V3D_ERR_INVALID_ARGUMENT;
}

constexpr int kTopkSegs = 8;

extern "C" size_t v3d_topk_rows_workspace_bytes(int rows, int k) {
  if (rows <= 0 || k <= 0) return 0;
  return (size_t)rows * kTopkSegs * k * (sizeof(float) + sizeof(long long)) + 256;
}

extern "C" int v3d_topk_rows(const float* values, int rows, int row_len, int k, float* out_values, int64_t* out_index,
                             void* workspace, size_t workspace_bytes, v3d_stream_t stream) {
  if (!values || !out_values || !out_index || rows <= 0 || row_len <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (k <= 0 || k > kTopkMax || k > row_len) return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  long long* out_i = reinterpret_cast<long long*>(out_index);
  const int seg_len = row_len / kTopkSegs;
  // long rows: 8 CTAs per row select k candidates each, a second launch picks k of the 8 k (a single CTA per row
  // streams the whole row several times: 70 us for 16 rows of 70 400). Ties still go to the lower index: the
  // candidate array is ordered by (segment, rank inside the segment).
  if (workspace && row_len % kTopkSegs == 0 && seg_len >= 4 * k && kTopkSegs * k >= k &&
      workspace_bytes >= v3d_topk_rows_workspace_bytes(rows, k)) {
    long long* cand_i = static_cast<long long*>(workspace);
    float* cand_v = reinterpret_cast<float*>(cand_i + (size_t)rows * kTopkSegs * k);
    topk_rows_kernel<<<rows * kTopkSegs, 1024, 0, st>>>(values, seg_len, k, cand_v, cand_i, kTopkSegs, nullptr);
    topk_rows_kernel<<<rows, 1024, 0, st>>>(cand_v, kTopkSegs * k, k, out_values, out_i, 1, cand_i);
    return check_launch();
  }
  topk_rows_kernel<<<rows, 1024, 0, st>>>(values, row_len, k, out_values, out_i, 1, nullptr);
  return check_launch();
}

extern "C" int v3d_head_reg_gather(const float* fmap_nhwc, int C, const float* w_reg, const float* b_reg,
                                   const float* top_logits, const int64_t* anchor_idx, int B, int n_cls, int n_yaw,
                                   int ny, int nx, int topk, float* deltas, float* scores, v3d_stream_t stream) {
  if (!fmap_nhwc || !w_reg || !top_logits || !anchor_idx || !deltas || !scores) return V3D_ERR_INVALID_ARGUMENT;
  if (C != 128 || B <= 0 || n_cls <= 0 || n_yaw <= 0 || topk <= 0) return V3D_ERR_INVALID_ARGUMENT;
  const int N = B * n_cls * topk;
  reg_gather_kernel<<<ceil_div(N, 8), 256, 0, as_stream(stream)>>>(fmap_nhwc, w_reg, b_reg, top_logits,
                                                                  reinterpret_cast<const long long*>(anchor_idx), N,
                                                                  n_cls, n_yaw, ny, nx, topk, deltas, scores);
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
