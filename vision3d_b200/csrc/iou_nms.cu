// Rotated BEV IoU (a13/a14) and rotated NMS (a12) for sm_100a.
//
// Replaces vision3d/ops/csrc/box_iou_rotated/box_iou_rotated_cuda.cu:14-121 and
// vision3d/ops/csrc/nms_rotated/nms_rotated_cuda.cu:14-134 (reference tree). Arithmetic follows
// box_iou_rotated_utils.h:56-340 as nvcc compiles it (exchange-sort hull, :197-214), operation by
// operation, with fp64 kept at the sites the reference evaluates in fp64 (:61-63, :97, :283,
// :318-319). THIS FILE IS COMPILED WITH -fmad=false so that no multiply-add is fused: results
// are bit-identical to oracle variant 1 (the same header built on the host).
//
// What is different from the reference kernels (design, not arithmetic):
//   * per-box work (fp64 sin/cos, the four half-extent products, area, bounding radius) is done
//     once per box, not once per pair;
//   * an exact disjointness pre-test (centre distance vs sum of bounding radii, 5 % margin)
//     short-circuits pairs whose IoU the reference would compute as exactly 0 -- with the
//     wrapper's per-group coordinate offsets (ops/iou_nms.py:124-132) that is almost every pair;
//   * NMS never leaves the device: O(N^2) counting rank instead of sort+index_select, only the
//     upper-triangular 64x64 mask tiles, and a chunked greedy scan replacing the reference's
//     blocking D2H copy + serial host loop (nms_rotated_cuda.cu:106-128).
#include "common.cuh"

namespace v3d {
namespace {

struct V2 {
  float x, y;
};
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float vcross(V2 a, V2 b) { return a.x * b.y - b.x * a.y; }

// Per-box invariants of get_rotated_vertices (utils.h:56-74) and single_box_iou_rotated (:329-330).
struct alignas(16) BoxPre {
  float x, y, w, h;      // raw centre / size
  float sh, cw, ch, sw;  // (sin/2)*h, (cos/2)*w, (cos/2)*h, (sin/2)*w
  float area, rad, pad0, pad1;
};

__device__ __forceinline__ BoxPre make_pre(const float* __restrict__ b) {
  BoxPre p;
  p.x = b[0];
  p.y = b[1];
  p.w = b[2];
  p.h = b[3];
  double theta = b[4] * 0.01745329251;  // utils.h:61 (fp64)
  float c2 = (float)cos(theta) * 0.5f;  // :62
  float s2 = (float)sin(theta) * 0.5f;  // :63
  p.sh = s2 * p.h;
  p.cw = c2 * p.w;
  p.ch = c2 * p.h;
  p.sw = s2 * p.w;
  p.area = p.w * p.h;
  p.rad = 0.5f * sqrtf(p.w * p.w + p.h * p.h);
  p.pad0 = p.pad1 = 0.f;
  return p;
}

__device__ __forceinline__ void corners(float cx, float cy, const BoxPre& b, V2 (&o)[4]) {
  o[0].x = cx - b.sh - b.cw;
  o[0].y = cy + b.ch - b.sw;
  o[1].x = cx + b.sh - b.cw;
  o[1].y = cy - b.ch - b.sw;
  o[2].x = 2 * cx - o[0].x;
  o[2].y = 2 * cy - o[0].y;
  o[3].x = 2 * cx - o[1].x;
  o[3].y = 2 * cy - o[1].y;
}

// true  => the reference computes exactly 0 for this pair (no edge crossing, no contained corner)
__device__ __forceinline__ bool surely_disjoint(const BoxPre& a, const BoxPre& b) {
  float dx = a.x - b.x, dy = a.y - b.y;
  float rr = a.rad + b.rad;
  return dx * dx + dy * dy > rr * rr * 1.05f;  // NaN/inf anywhere -> false -> full evaluation
}

__device__ __noinline__ float iou_full(const BoxPre& A, const BoxPre& B) {
  // utils.h:318-321: centre shift in fp64, rounded to fp32 when stored in the RotatedBox
  double sx = (A.x + B.x) / 2.0;
  double sy = (A.y + B.y) / 2.0;
  float ax = (float)(A.x - sx), ay = (float)(A.y - sy);
  float bx = (float)(B.x - sx), by = (float)(B.y - sy);
  if ((double)A.area < 1e-14 || (double)B.area < 1e-14) return 0.f;  // :331-333

  V2 p1[4], p2[4];
  corners(ax, ay, A, p1);
  corners(bx, by, B, p2);

  // ---- get_intersection_points, utils.h:76-155 ----
  V2 pts[24];
  int n = 0;
  V2 e1[4], e2[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    e1[i] = vsub(p1[(i + 1) & 3], p1[i]);
    e2[i] = vsub(p2[(i + 1) & 3], p2[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float det = vcross(e2[j], e1[i]);
      if (fabs((double)det) <= 1e-14) continue;
      V2 d = vsub(p2[j], p1[i]);
      float t1 = vcross(e2[j], d) / det;
      float t2 = vcross(e1[i], d) / det;
      if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) {
        pts[n].x = p1[i].x + e1[i].x * t1;
        pts[n].y = p1[i].y + e1[i].y * t1;
        n++;
      }
    }
  }
  {
    const V2 AB = e2[0], DA = e2[3];
    const float ABAB = vdot(AB, AB), ADAD = vdot(DA, DA);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      V2 AP = vsub(p1[i], p2[0]);
      float pAB = vdot(AP, AB);
      float pAD = -vdot(AP, DA);
      if (pAB >= 0 && pAD >= 0 && pAB <= ABAB && pAD <= ADAD) pts[n++] = p1[i];
    }
  }
  {
    const V2 AB = e1[0], DA = e1[3];
    const float ABAB = vdot(AB, AB), ADAD = vdot(DA, DA);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      V2 AP = vsub(p2[i], p1[0]);
      float pAB = vdot(AP, AB);
      float pAD = -vdot(AP, DA);
      if (pAB >= 0 && pAD >= 0 && pAB <= ABAB && pAD <= ADAD) pts[n++] = p2[i];
    }
  }

  float inter = 0.0f;
  if (n > 2) {  // utils.h:301-303
    // ---- convex_hull_graham(shift_to_zero = true), utils.h:157-270, nvcc branch ----
    int t = 0;
    for (int i = 1; i < n; i++)
      if (pts[i].y < pts[t].y || (pts[i].y == pts[t].y && pts[i].x < pts[t].x)) t = i;
    const V2 org = pts[t];
    V2 q[24];
    float dist[24];
    for (int i = 0; i < n; i++) q[i] = vsub(pts[i], org);
    {
      V2 tmp = q[0];
      q[0] = q[t];
      q[t] = tmp;
    }
    for (int i = 0; i < n; i++) dist[i] = vdot(q[i], q[i]);
    for (int i = 1; i < n - 1; i++) {
      for (int j = i + 1; j < n; j++) {
        float cp = vcross(q[i], q[j]);
        if (((double)cp < -1e-6) || (fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
          V2 tq = q[i];
          q[i] = q[j];
          q[j] = tq;
          float td = dist[i];
          dist[i] = dist[j];
          dist[j] = td;
        }
      }
    }
    int k;
    for (k = 1; k < n; k++)
      if ((double)dist[k] > 1e-8) break;
    int m;
    if (k == n) {
      m = 1;
    } else {
      q[1] = q[k];
      m = 2;
      for (int i = k + 1; i < n; i++) {
        while (m > 1 && vcross(vsub(q[i], q[m - 2]), vsub(q[m - 1], q[m - 2])) >= 0) m--;
        q[m++] = q[i];
      }
    }
    // ---- polygon_area, utils.h:272-284 ----
    if (m > 2) {
      float area = 0;
      for (int i = 1; i < m - 1; i++) area += fabsf(vcross(vsub(q[i], q[0]), vsub(q[i + 1], q[0])));
      inter = (float)(area / 2.0);
    }
  }
  return inter / (A.area + B.area - inter);  // utils.h:336
}

__device__ __forceinline__ float iou_pair(const BoxPre& A, const BoxPre& B) {
  if (surely_disjoint(A, B)) {
    // the reference would still return 0 through the area guard or 0/(a1+a2); a1+a2 == 0 cannot
    // pass the guard, so the value is exactly +0.0f
    return 0.0f;
  }
  return iou_full(A, B);
}

// -------------------------------------------------------------------------------------------
// a13: pairwise IoU. Tile = 16 rows x 128 cols, 256 threads, 8 pairs per thread; threads run
// along the column (contiguous output) dimension.
// -------------------------------------------------------------------------------------------
constexpr int kIouTR = 16, kIouTC = 128, kIouThreads = 256;

__global__ void __launch_bounds__(kIouThreads) box_iou_kernel(const float* __restrict__ b1, int M,
                                                              const float* __restrict__ b2, int N,
                                                              float* __restrict__ out) {
  __shared__ BoxPre rows[kIouTR];
  __shared__ BoxPre cols[kIouTC];
  const int r0 = blockIdx.y * kIouTR, c0 = blockIdx.x * kIouTC;
  const int tid = threadIdx.x;
  if (tid < kIouTC) {
    if (c0 + tid < N) cols[tid] = make_pre(b2 + (size_t)(c0 + tid) * 5);
  } else if (tid < kIouTC + kIouTR) {
    int r = tid - kIouTC;
    if (r0 + r < M) rows[r] = make_pre(b1 + (size_t)(r0 + r) * 5);
  }
  __syncthreads();
  const int c = tid & (kIouTC - 1);
  if (c0 + c >= N) return;
  const BoxPre cb = cols[c];
#pragma unroll 1
  for (int r = tid / kIouTC; r < kIouTR; r += kIouThreads / kIouTC) {
    if (r0 + r >= M) break;
    out[(size_t)(r0 + r) * N + c0 + c] = iou_pair(rows[r], cb);
  }
}

// -------------------------------------------------------------------------------------------
// Training-side consumer of a13: ProposalTargetAssigner.match_class_i
// (core/proposal_targets.py:53-60) = box_iou_rotated(M gt boxes, N anchors) followed by Matcher.__call__
// (ops/matcher.py:86-107: max over the gt axis, threshold strata -> label). Fused: one thread per anchor walks the
// gt boxes staged in shared memory, the M x N matrix (M x 70 400 per class) is never written.
// Same IoU arithmetic as box_iou_kernel; max keeps the FIRST gt index on ties (what torch.max returns for the
// row-major (M, N) matrix on CUDA); strata are evaluated in the reference's order (later strata overwrite).
// -------------------------------------------------------------------------------------------
constexpr int kMatchMaxGt = 256;
constexpr int kMatchMaxStrata = 8;
struct MatchStrata {
  int n;
  float low[kMatchMaxStrata], high[kMatchMaxStrata];
  int label[kMatchMaxStrata];
};

__global__ void __launch_bounds__(256) match_anchors_kernel(const float* __restrict__ gt, int M,
                                                            const float* __restrict__ anchors, int N, MatchStrata S,
                                                            long long* __restrict__ matches,
                                                            signed char* __restrict__ labels,
                                                            float* __restrict__ matched_vals) {
  __shared__ BoxPre g[kMatchMaxGt];
  for (int i = threadIdx.x; i < M; i += blockDim.x) g[i] = make_pre(gt + (size_t)i * 5);
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  const BoxPre A = make_pre(anchors + (size_t)a * 5);
  float best = -1.0f;
  int bi = 0;
  for (int i = 0; i < M; i++) {
    const float v = iou_pair(g[i], A);  // row = gt box, column = anchor: box_iou_rotated(boxes, anchors)
    if (v > best) {
      best = v;
      bi = i;
    }
  }
  int lab = 1;  // matcher.py:96
  for (int t = 0; t < S.n; t++)
    if (best >= S.low[t] && best < S.high[t]) lab = S.label[t];
  matches[a] = bi;
  labels[a] = (signed char)lab;
  if (matched_vals) matched_vals[a] = best;
}

// -------------------------------------------------------------------------------------------
// a12: NMS, three launches, nothing leaves the device.
// -------------------------------------------------------------------------------------------
constexpr int kTile = 64;

// (1) counting rank: position of box i in descending-score order, ties -> lower index first; also writes
//     the per-box invariants at the sorted position. One warp per box, lanes stride over the N scores.
__device__ __forceinline__ float rank_key(float s) { return s != s ? __int_as_float(0x7f800000) : s; }  // NaN -> +inf

__global__ void __launch_bounds__(256) nms_rank_kernel(const float* __restrict__ dets,
                                                       const float* __restrict__ scores, int N,
                                                       int* __restrict__ order, BoxPre* __restrict__ pre) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;  // warp-uniform
  const float si = rank_key(__ldg(&scores[i]));
  int cnt = 0;
  for (int j = lane; j < N; j += 32) {
    const float sj = rank_key(__ldg(&scores[j]));
    cnt += (sj > si) || (sj == si && j < i);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if (lane == 0) {
    order[cnt] = i;
    pre[cnt] = make_pre(dets + (size_t)i * 5);
  }
}

// (2) upper-triangular 64x64 mask tiles, ONE THREAD PER PAIR: block = (64 columns, 16 rows), 4 blocks per
//     tile; a warp is half a row, __ballot_sync gives a 32-bit half of the row's mask word.
constexpr int kMaskRows = 16;
__global__ void __launch_bounds__(kTile * kMaskRows) nms_mask_kernel(const BoxPre* __restrict__ pre, int N,
                                                                     float thr, int col_blocks,
                                                                     unsigned int* __restrict__ mask32) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;  // lower triangle is never read by the scan
  __shared__ BoxPre rbox[kMaskRows];
  __shared__ BoxPre cbox[kTile];
  const int c = threadIdx.x, rl = threadIdx.y;
  const int r = blockIdx.z * kMaskRows + rl;  // row inside the tile
  const int gi = rb * kTile + r, gj = cb * kTile + c;
  if (rl == 0 && gj < N) cbox[c] = pre[gj];
  if (rl == 1 && c < kMaskRows) {
    const int g = rb * kTile + blockIdx.z * kMaskRows + c;
    if (g < N) rbox[c] = pre[g];
  }
  __syncthreads();
  bool hit = false;
  if (gi < N && gj < N && (rb != cb || c > r)) hit = iou_pair(rbox[rl], cbox[c]) > thr;  // nms_rotated_cuda.cu:62-63
  const unsigned int w = __ballot_sync(0xffffffffu, hit);
  if ((c & 31) == 0 && gi < N) mask32[((size_t)gi * col_blocks + cb) * 2 + (c >> 5)] = w;  // little endian halves
}

// (3a) greedy scan (nms_rotated_cuda.cu:115-128), N <= 8192: the 64 mask rows of chunk b+1 are prefetched
//      into shared memory with cp.async while chunk b is resolved; the 64-step dependent chain runs on
//      registers of one thread; kept rows are OR-ed into the removed-words from shared memory.
constexpr int kScanThreads = 256;
__global__ void __launch_bounds__(kScanThreads) nms_scan_smem_kernel(const unsigned long long* __restrict__ mask,
                                                                     const int* __restrict__ order, int N,
                                                                     int cb, long long* __restrict__ keep,
                                                                     int* __restrict__ num_keep) {
  extern __shared__ __align__(16) unsigned long long sm[];
  unsigned long long* remv = sm;                       // [cb]
  unsigned long long* buf0 = sm + ((cb + 1) & ~1);     // [64*cb] x 2, 16-byte aligned
  unsigned long long* buf1 = buf0 + (size_t)kTile * cb;
  __shared__ unsigned long long kept_bits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < cb; j += kScanThreads) remv[j] = 0ull;

  auto prefetch = [&](int b, unsigned long long* dst) {
    const int rows = min(kTile, N - b * kTile);
    const size_t words = (size_t)rows * cb;            // contiguous in the mask array
    const unsigned long long* src = mask + (size_t)b * kTile * cb;
    for (size_t e = (size_t)tid * 2; e + 1 < words + 1; e += (size_t)kScanThreads * 2) {
      if (e + 1 < words) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + e)),
                     "l"(src + e));
      } else if (e < words) {
        dst[e] = src[e];
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, buf0);
  int base = 0;
  for (int b = 0; b < cb; b++) {
    unsigned long long* cur = (b & 1) ? buf1 : buf0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // chunk b rows visible; previous OR phase done
    if (b + 1 < cb) prefetch(b + 1, (b & 1) ? buf0 : buf1);
    const int rows = min(kTile, N - b * kTile);
    if (tid == 0) {
      unsigned long long removed = remv[b], kept = 0ull;
#pragma unroll
      for (int i = 0; i < kTile; i++) {
        const unsigned long long d = i < rows ? cur[(size_t)i * cb + b] : 0ull;  // loads do not depend on the chain
        const bool alive = i < rows && !((removed >> i) & 1ull);
        kept |= alive ? (1ull << i) : 0ull;
        removed |= alive ? d : 0ull;
      }
      kept_bits = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_bits;
    if (tid < rows && ((kept >> tid) & 1ull))
      keep[base + __popcll(kept & ((1ull << tid) - 1ull))] = (long long)order[b * kTile + tid];
    base += __popcll(kept);
    for (int i = warp; i < rows; i += kScanThreads / 32) {
      if (!((kept >> i) & 1ull)) continue;
      const unsigned long long* row = cur + (size_t)i * cb;
      for (int j = b + 1 + lane; j < cb; j += 32) {
        const unsigned long long m = row[j];
        if (m) atomicOr(&remv[j], m);
      }
    }
  }
  if (tid == 0) *num_keep = base;
}

// (3b) greedy scan for large N: one block, per 64-row chunk thread 0 resolves the chunk against its diagonal
//      word, then one warp per kept row ORs the row's later words (all loads of the chunk in flight at once).
__global__ void __launch_bounds__(1024) nms_scan_kernel(const unsigned long long* __restrict__ mask,
                                                        const int* __restrict__ order, int N,
                                                        int col_blocks, long long* __restrict__ keep,
                                                        int* __restrict__ num_keep) {
  extern __shared__ unsigned long long remv[];  // col_blocks words
  __shared__ unsigned long long diag[2][kTile];
  __shared__ unsigned long long kept_bits;
  __shared__ int kept_base;
  const int tid = threadIdx.x;
  for (int j = tid; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  if (tid == 0) kept_base = 0;
  if (tid < kTile && tid < N) diag[0][tid] = mask[(size_t)tid * col_blocks];
  __syncthreads();
  for (int b = 0; b < col_blocks; b++) {
    const int rows = min(kTile, N - b * kTile);
    const int cur = b & 1;
    if (tid >= 64 && tid < 64 + kTile && b + 1 < col_blocks) {
      int r = (b + 1) * kTile + (tid - 64);
      if (r < N) diag[cur ^ 1][tid - 64] = mask[(size_t)r * col_blocks + (b + 1)];
    }
    if (tid == 0) {
      unsigned long long removed = remv[b], kept = 0ull;
      for (int i = 0; i < rows; i++) {
        if (!((removed >> i) & 1ull)) {
          kept |= 1ull << i;
          removed |= diag[cur][i];
        }
      }
      kept_bits = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_bits;
    const int base = kept_base;
    if (tid < rows && ((kept >> tid) & 1ull)) {
      int pos = base + __popcll(kept & ((1ull << tid) - 1ull));
      keep[pos] = (long long)order[b * kTile + tid];
    }
    {
      const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
      for (int i = warp; i < rows; i += nwarps) {
        if (!((kept >> i) & 1ull)) continue;
        const unsigned long long* row = mask + (size_t)(b * kTile + i) * col_blocks;
        for (int j = b + 1 + lane; j < col_blocks; j += 32) {
          const unsigned long long m = row[j];
          if (m) atomicOr(&remv[j], m);
        }
      }
    }
    __syncthreads();
    if (tid == 0) kept_base = base + __popcll(kept);
    __syncthreads();
  }
  if (tid == 0) *num_keep = kept_base;
}

// -------------------------------------------------------------------------------------------
// Grouped NMS: the N boxes are G consecutive groups of `gs` boxes that cannot overlap ACROSS groups -- exactly
// what batched_nms_rotated's per-group coordinate offsets (ops/iou_nms.py:121-132) construct before calling
// nms_rotated. Cross-group IoU is then exactly 0, so the global greedy NMS decomposes into one independent
// greedy NMS per group with the same comparator (score descending, ties -> lower index) and the same IoU
// arithmetic on the same (offset) coordinates: identical keep set, identical order. One CTA per group: counting
// rank + invariants, upper-triangular pair tests into a shared-memory bit matrix, 1-thread greedy chain.
// A tiny second kernel compacts the survivors in GLOBAL score order.
// -------------------------------------------------------------------------------------------
constexpr int kGroupMax = 128;
constexpr int kGroupSlices = 8;  // CTAs per group for the pair tests (rows r = slice, slice + 8, ...)

// local order of group g (score descending, ties -> lower index) + per-box invariants at the sorted position
__device__ __forceinline__ void group_rank(const float* __restrict__ dets, const float* __restrict__ scores, int g0, int n,
                                           float* sc, int* ord, BoxPre* pre) {
  const int tid = threadIdx.x;
  if (tid < n) sc[tid] = rank_key(__ldg(&scores[g0 + tid]));
  __syncthreads();
  if (tid < n) {
    const float si = sc[tid];
    int cnt = 0;
    for (int j = 0; j < n; j++) cnt += (sc[j] > si) || (sc[j] == si && j < tid);
    ord[cnt] = tid;
    if (pre) pre[cnt] = make_pre(dets + (size_t)(g0 + tid) * 5);
  }
  __syncthreads();
}

// pair tests of the rows one slice owns -> bits[g][r][2] (global, every row written by exactly one CTA)
__global__ void __launch_bounds__(1024) nms_group_pairs_kernel(const float* __restrict__ dets,
                                                               const float* __restrict__ scores, int N, int gs, float thr,
                                                               unsigned long long* __restrict__ gbits) {
  __shared__ BoxPre pre[kGroupMax];
  __shared__ float sc[kGroupMax];
  __shared__ int ord[kGroupMax];
  __shared__ unsigned long long bits[kGroupMax][2];
  const int g = blockIdx.y, slice = blockIdx.x, g0 = g * gs, n = min(gs, N - g0), tid = threadIdx.x;
  if (tid < kGroupMax) bits[tid][0] = bits[tid][1] = 0ull;
  group_rank(dets, scores, g0, n, sc, ord, pre);
  const int my_rows = (n - slice + kGroupSlices - 1) / kGroupSlices;  // rows slice, slice + S, ...
  for (int p = tid; p < my_rows * n; p += blockDim.x) {
    const int r = slice + (p / n) * kGroupSlices, c = p % n;
    if (c > r && iou_pair(pre[r], pre[c]) > thr) atomicOr(&bits[r][c >> 6], 1ull << (c & 63));  // nms_rotated_cuda.cu:62-63
  }
  __syncthreads();
  for (int e = tid; e < my_rows * 2; e += blockDim.x) {
    const int r = slice + (e >> 1) * kGroupSlices;
    gbits[((size_t)g0 + r) * 2 + (e & 1)] = bits[r][e & 1];
  }
}

// greedy chain of one group over its bit matrix
__global__ void __launch_bounds__(kGroupMax) nms_group_chain_kernel(const float* __restrict__ scores, int N, int gs,
                                                                    const unsigned long long* __restrict__ gbits,
                                                                    unsigned char* __restrict__ kept) {
  __shared__ float sc[kGroupMax];
  __shared__ int ord[kGroupMax];
  __shared__ unsigned long long bits[kGroupMax][2];
  const int g = blockIdx.x, g0 = g * gs, n = min(gs, N - g0), tid = threadIdx.x;
  group_rank(nullptr, scores, g0, n, sc, ord, nullptr);
  if (tid < n) {
    bits[tid][0] = gbits[((size_t)g0 + tid) * 2];
    bits[tid][1] = gbits[((size_t)g0 + tid) * 2 + 1];
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long rem0 = 0ull, rem1 = 0ull;
    for (int r = 0; r < n; r++) {
      const bool dead = r < 64 ? ((rem0 >> r) & 1ull) : ((rem1 >> (r - 64)) & 1ull);
      kept[g0 + ord[r]] = dead ? 0 : 1;
      if (!dead) {
        rem0 |= bits[r][0];
        rem1 |= bits[r][1];
      }
    }
  }
}

// survivors in global descending-score order (order[] from nms_rank_kernel)
__global__ void __launch_bounds__(1024) nms_compact_kernel(const int* __restrict__ order, const unsigned char* __restrict__ kept,
                                                           int N, long long* __restrict__ keep, int* __restrict__ num_keep) {
  __shared__ int sm[33];
  int base = 0;
  for (int r0 = 0; r0 < N; r0 += blockDim.x) {
    const int r = r0 + threadIdx.x;
    const int i = r < N ? order[r] : 0;
    const int flag = (r < N && kept[i]) ? 1 : 0;
    int total;
    const int ex = block_exclusive_scan(flag, sm, total);
    if (flag) keep[base + ex] = (long long)i;
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = base;
}

struct NmsLayout {
  size_t order_off, pre_off, mask_off, total;
};
inline NmsLayout nms_layout(int N) {
  NmsLayout l;
  size_t cb = (size_t)ceil_div(N, kTile);
  l.order_off = 0;
  l.pre_off = align_up(l.order_off + sizeof(int) * (size_t)N, 256);
  l.mask_off = align_up(l.pre_off + sizeof(BoxPre) * (size_t)N, 256);
  // N x cb mask words; the grouped variant needs 2 words per box + N flags instead
  const size_t mask_bytes = sizeof(unsigned long long) * (size_t)N * cb, grouped_bytes = (size_t)17 * N + 256;
  l.total = align_up(l.mask_off + (mask_bytes > grouped_bytes ? mask_bytes : grouped_bytes), 256);
  return l;
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" int v3d_box_iou_rotated(const float* boxes1, int M, const float* boxes2, int N, float* ious,
                                   v3d_stream_t stream) {
  if (M < 0 || N < 0) return V3D_ERR_INVALID_ARGUMENT;
  if (M == 0 || N == 0) return V3D_OK;  // box_iou_rotated_cuda.cu:80
  if (!boxes1 || !boxes2 || !ious) return V3D_ERR_INVALID_ARGUMENT;
  // any M: the reference transposes the problem when one grid dimension would overflow
  // (box_iou_rotated_cuda.cu:84-95); here the rows are simply cut into launches of <= 65535 row tiles
  constexpr int kRowsPerLaunch = 65535 * kIouTR;
  for (int m0 = 0; m0 < M; m0 += kRowsPerLaunch) {
    const int m = M - m0 < kRowsPerLaunch ? M - m0 : kRowsPerLaunch;
    dim3 grid(ceil_div(N, kIouTC), ceil_div(m, kIouTR));
    box_iou_kernel<<<grid, kIouThreads, 0, as_stream(stream)>>>(boxes1 + (size_t)m0 * 5, m, boxes2, N,
                                                                ious + (size_t)m0 * N);
  }
  return check_launch();
}

extern "C" size_t v3d_nms_rotated_workspace_bytes(int N) {
  if (N <= 0) return 256;
  return nms_layout(N).total;
}

extern "C" int v3d_nms_rotated_grouped(const float* dets, const float* scores, int N, int group_size,
                                       float iou_threshold, int64_t* keep, int* num_keep, void* workspace,
                                       size_t workspace_bytes, v3d_stream_t stream) {
  if (N < 0 || !num_keep || group_size <= 0 || group_size > kGroupMax) return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    V3D_CUDA_TRY(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
    return V3D_OK;
  }
  if (!dets || !scores || !keep || !workspace || N > 65536) return V3D_ERR_INVALID_ARGUMENT;
  NmsLayout l = nms_layout(N);
  if (workspace_bytes < l.total) return V3D_ERR_WORKSPACE_TOO_SMALL;
  char* ws = static_cast<char*>(workspace);
  int* order = reinterpret_cast<int*>(ws + l.order_off);
  BoxPre* pre = reinterpret_cast<BoxPre*>(ws + l.pre_off);
  // the mask area holds the per-group bit matrices (2 words per box) followed by the N kept flags
  const int n_groups = ceil_div(N, group_size);
  unsigned long long* gbits = reinterpret_cast<unsigned long long*>(ws + l.mask_off);
  unsigned char* kept = reinterpret_cast<unsigned char*>(ws + l.mask_off + sizeof(unsigned long long) * 2 * (size_t)N);
  nms_rank_kernel<<<ceil_div(N, 8), 256, 0, st>>>(dets, scores, N, order, pre);
  nms_group_pairs_kernel<<<dim3(kGroupSlices, n_groups), 1024, 0, st>>>(dets, scores, N, group_size, iou_threshold,
                                                                        gbits);
  nms_group_chain_kernel<<<n_groups, kGroupMax, 0, st>>>(scores, N, group_size, gbits, kept);
  nms_compact_kernel<<<1, 1024, 0, st>>>(order, kept, N, reinterpret_cast<long long*>(keep), num_keep);
  return check_launch();
}

extern "C" int v3d_nms_rotated(const float* dets, const float* scores, int N, float iou_threshold,
                               int64_t* keep, int* num_keep, void* workspace, size_t workspace_bytes,
                               v3d_stream_t stream) {
  if (N < 0 || !num_keep) return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  if (N == 0) {  // nms_rotated_cpu.cpp:21-23: empty in, empty out
    V3D_CUDA_TRY(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
    return V3D_OK;
  }
  if (!dets || !scores || !keep || !workspace) return V3D_ERR_INVALID_ARGUMENT;
  if (N > 65536) return V3D_ERR_INVALID_ARGUMENT;  // mask words per row <= 1024
  NmsLayout l = nms_layout(N);
  if (workspace_bytes < l.total) return V3D_ERR_WORKSPACE_TOO_SMALL;
  char* ws = static_cast<char*>(workspace);
  int* order = reinterpret_cast<int*>(ws + l.order_off);
  BoxPre* pre = reinterpret_cast<BoxPre*>(ws + l.pre_off);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + l.mask_off);
  const int cb = ceil_div(N, kTile);

  nms_rank_kernel<<<ceil_div(N, 8), 256, 0, st>>>(dets, scores, N, order, pre);
  nms_mask_kernel<<<dim3(cb, cb, kTile / kMaskRows), dim3(kTile, kMaskRows), 0, st>>>(
      pre, N, iou_threshold, cb, reinterpret_cast<unsigned int*>(mask));
  const size_t smem_fast = sizeof(unsigned long long) * (((size_t)cb + 1) / 2 * 2 + 2 * (size_t)kTile * cb);
  if (smem_fast <= 200 * 1024) {
    static PerDeviceOnce attr_once;
    if (attr_once.needed()) {
      V3D_CUDA_TRY(cudaFuncSetAttribute(nms_scan_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_once.done();
    }
    nms_scan_smem_kernel<<<1, kScanThreads, smem_fast, st>>>(mask, order, N, cb, reinterpret_cast<long long*>(keep),
                                                            num_keep);
  } else {
    nms_scan_kernel<<<1, 1024, sizeof(unsigned long long) * cb, st>>>(mask, order, N, cb,
                                                                     reinterpret_cast<long long*>(keep), num_keep);
  }
  return check_launch();
}

extern "C" int v3d_match_anchors(const float* gt_boxes, int M, const float* anchors, int N, int n_strata,
                                 const float* low_host, const float* high_host, const int* label_host,
                                 int64_t* matches, signed char* labels, float* matched_vals, v3d_stream_t stream) {
  if (M <= 0 || M > kMatchMaxGt || N <= 0 || n_strata <= 0 || n_strata > kMatchMaxStrata) return V3D_ERR_INVALID_ARGUMENT;
  if (!gt_boxes || !anchors || !low_host || !high_host || !label_host || !matches || !labels) return V3D_ERR_INVALID_ARGUMENT;
  MatchStrata S;
  S.n = n_strata;
  for (int t = 0; t < n_strata; t++) {
    S.low[t] = low_host[t];
    S.high[t] = high_host[t];
    S.label[t] = label_host[t];
  }
  match_anchors_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(gt_boxes, M, anchors, N, S,
                                                                        reinterpret_cast<long long*>(matches), labels,
                                                                        matched_vals);
  return check_launch();
}
