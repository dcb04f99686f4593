// a7-a10: PV-RCNN point ops for sm_100a -- furthest point sampling, gather, ball query, grouping.
// Replace pointnet2_utils.furthest_point_sample / gather_operation (vision3d/detector/model.py:53-54)
// and the ball_query / grouping_operation / QueryAndGroup inside PointnetSAModuleMSG
// (detector/model.py:39-43,64; detector/roi_grid_pool.py:28-32,68).
//
// FPS is a 2047-step dependent chain, not a bandwidth problem. Upstream runs one CTA per cloud with the
// running-min array in global memory (8 of 148 SMs busy at B=8, 2 global round trips per step). Here a
// thread-block CLUSTER of 8 CTAs owns one cloud: every point and its running min distance live in
// registers for the whole kernel, each step is a register pass + warp-shuffle arg-max + one exchange of
// the 8 CTA candidates through distributed shared memory, and nothing touches HBM inside the loop.
// Distances are evaluated as (dx*dx + dy*dy) + dz*dz with explicit round-to-nearest mul/add (no FMA);
// ties resolve to the LOWEST point index -- the rule the oracle documents.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace v3d {
namespace {

constexpr int kFpsCluster = 8;
constexpr int kFpsThreads = 512;

struct alignas(16) Cand {
  unsigned int dbits;  // distance as ordered bits (distances are >= 0, so uint order == float order)
  unsigned int nidx;   // ~index: larger = lower index, so max over (dbits, nidx) = farthest point, ties -> lowest index
  float x, y, z;
  float pad[3];
};

__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// arg-max over a warp of (dbits, nidx) pairs in lexicographic order: two REDUX instructions
__device__ __forceinline__ void warp_argmax(unsigned int& dbits, unsigned int& nidx) {
  const unsigned int m = __reduce_max_sync(0xffffffffu, dbits);
  const unsigned int c = __reduce_max_sync(0xffffffffu, dbits == m ? nidx : 0u);
  dbits = m;
  nidx = c;
}

template <int PPT>
__global__ void __cluster_dims__(kFpsCluster, 1, 1) __launch_bounds__(kFpsThreads)
    fps_cluster_kernel(const float* __restrict__ xyz, int stride, int N, int m, int* __restrict__ out,
                       float* __restrict__ out_xyz) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kFpsCluster;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kFpsThreads / 32;
  const float* P = xyz + (size_t)b * N * stride;

  __shared__ Cand cand[2][kFpsCluster];  // written by every CTA of the cluster through DSMEM
  __shared__ unsigned int wd[2][kWarps], wi[2][kWarps];

  // CTA `rank` owns the contiguous slice [s0, s0 + S); thread t owns s0 + t + k*blockDim, k < PPT
  const int S = PPT * kFpsThreads;
  const int s0 = rank * S;
  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int k = 0; k < PPT; k++) {
    const int i = s0 + tid + k * kFpsThreads;
    if (i < N) {
      if (stride == 4) {  // (x, y, z, intensity) rows as the preprocessor emits them: one 128-bit load per point
        const float4 q = __ldg(reinterpret_cast<const float4*>(P) + i);
        px[k] = q.x;
        py[k] = q.y;
        pz[k] = q.z;
      } else {
        px[k] = P[(size_t)stride * i];
        py[k] = P[(size_t)stride * i + 1];
        pz[k] = P[(size_t)stride * i + 2];
      }
      md[k] = 1e10f;
    } else {
      px[k] = py[k] = pz[k] = 0.f;
      md[k] = -1.f;  // never wins (distances are >= 0)
    }
  }
  float cx = P[0], cy = P[1], cz = P[2];
  if (rank == 0 && tid == 0) {
    out[(size_t)b * m] = 0;
    if (out_xyz) {  // gather_operation fused (detector/model.py:54): the winner's coordinates are at hand
      out_xyz[(size_t)b * m * 3] = cx;
      out_xyz[(size_t)b * m * 3 + 1] = cy;
      out_xyz[(size_t)b * m * 3 + 2] = cz;
    }
  }
  cluster.sync();

  for (int j = 1; j < m; j++) {
    const int par = j & 1;
    // 1. running min distance of my points to the selected set, my farthest point (lowest index on ties)
    float bd = -1.f;
    int bk = 0;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
      const float d = dist2(px[k], py[k], pz[k], cx, cy, cz);
      const float d2 = fminf(d, md[k]);  // md = -1 (no point) stays -1
      md[k] = d2;
      if (d2 > bd) {  // ascending index inside the thread: strict > keeps the lowest index
        bd = d2;
        bk = k;
      }
    }
    const unsigned int my_d = bd < 0.f ? 0u : __float_as_uint(bd);
    const unsigned int my_n = bd < 0.f ? 0u : ~(unsigned int)(s0 + tid + bk * kFpsThreads);
    // 2. warp arg-max (2 REDUX), warp leaders publish, every warp reduces the kWarps candidates redundantly
    unsigned int d = my_d, n = my_n;
    warp_argmax(d, n);
    if (lane == 0) {
      wd[par][warp] = d;
      wi[par][warp] = n;
    }
    __syncthreads();
    d = lane < kWarps ? wd[par][lane] : 0u;
    n = lane < kWarps ? wi[par][lane] : 0u;
    warp_argmax(d, n);
    // 3. the thread that owns the CTA's winner delivers it (with its coordinates) to all CTAs of the cluster
    if (d == my_d && n == my_n && bd >= 0.f) {
      float wx = px[0], wy = py[0], wz = pz[0];
#pragma unroll
      for (int k = 1; k < PPT; k++)
        if (bk == k) {
          wx = px[k];
          wy = py[k];
          wz = pz[k];
        }
#pragma unroll
      for (int r = 0; r < kFpsCluster; r++) {
        Cand* remote = cluster.map_shared_rank(&cand[par][rank], r);
        *reinterpret_cast<uint4*>(remote) = make_uint4(d, n, __float_as_uint(wx), __float_as_uint(wy));
        remote->z = wz;
      }
    } else if (d == 0u && n == 0u && tid == 0) {  // this CTA holds no point at all (N < s0): deliver "nothing"
#pragma unroll
      for (int r = 0; r < kFpsCluster; r++) {
        Cand* remote = cluster.map_shared_rank(&cand[par][rank], r);
        *reinterpret_cast<uint4*>(remote) = make_uint4(0u, 0u, 0u, 0u);
        remote->z = 0.f;
      }
    }
    cluster.sync();  // candidates of step j visible everywhere; parity buffers make one sync per step enough
    // 4. cluster arg-max over the 8 candidates (every thread, broadcast reads)
    unsigned int gd = cand[par][0].dbits, gn = cand[par][0].nidx;
    int gr = 0;
#pragma unroll
    for (int r = 1; r < kFpsCluster; r++) {
      const unsigned int rd = cand[par][r].dbits, rn = cand[par][r].nidx;
      if (rd > gd || (rd == gd && rn > gn)) {
        gd = rd;
        gn = rn;
        gr = r;
      }
    }
    cx = cand[par][gr].x;
    cy = cand[par][gr].y;
    cz = cand[par][gr].z;
    if (rank == 0 && tid == 0) {
      out[(size_t)b * m + j] = (int)~gn;
      if (out_xyz) {
        float* o = out_xyz + ((size_t)b * m + j) * 3;
        o[0] = cx;
        o[1] = cy;
        o[2] = cz;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// FPS, direct exchange (clouds of <= 16 384 points, i.e. every configuration the reference runs): the step above
// costs ~2 100 clk, most of it the block barrier + cluster barrier pair around the candidate exchange, not the
// distance updates. Here a step has NO barrier at all:
//   * every CTA of the cluster keeps a copy of the whole cloud's coordinates in shared memory (3 x 64 KB), so a
//     candidate is fully described by (distance bits, point index) and fits, with a 16-bit step tag, in ONE 64-bit
//     word -- a single-copy-atomic store, no ordering between a payload and a flag to enforce;
//   * every WARP reduces its points (2 REDUX) and stores its word into its slot in all 8 CTAs (lanes 0-7, one
//     remote store each); every warp then polls the local slots (4 warps x 8 CTAs = 32: one per lane) until all
//     carry this step's tag, reduces them (2 REDUX) and reads the winner's coordinates from its CTA's copy of the
//     cloud. Measured: 760 clk per step against 2 100 (B = 8, N = 16 384, m = 2 048: 0.82 ms against 2.30 ms).
// Slots are double-buffered by step parity. A slot of parity p is rewritten at step j+2 by a warp that has seen
// ALL step j+1 words, and a warp sends its step j+1 word only after it has read the step j slots -- so nobody can
// still be reading what is overwritten. Results are identical to the kernel above (same distances, same
// lowest-index tie rule).
// ---------------------------------------------------------------------------------------------------------
constexpr int kFpsDirectMaxN = 16384;

template <int PPT, int WARPS, int CLUSTER>
__global__ void __launch_bounds__(WARPS * 32)
    fps_direct_kernel(const float* __restrict__ xyz, int stride, int N, int n_pad, int m, int* __restrict__ out,
                      float* __restrict__ out_xyz) {
  constexpr int kThreads = WARPS * 32;
  constexpr int kSlots = WARPS * CLUSTER;  // one per warp of the cluster
  constexpr int SL = kSlots / 32;          // slots polled per lane
  static_assert(kSlots % 32 == 0 && (SL == 1 || SL == 2 || SL == 4), "slot count");
  extern __shared__ __align__(16) unsigned char fps_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / CLUSTER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = xyz + (size_t)b * N * stride;
  unsigned long long* slots = reinterpret_cast<unsigned long long*>(fps_smem);  // [2][kSlots]
  float* sx = reinterpret_cast<float*>(fps_smem + 2 * kSlots * sizeof(unsigned long long));
  float* sy = sx + n_pad;
  float* sz = sy + n_pad;

  for (int i = tid; i < N; i += kThreads) {
    if (stride == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(P) + i);
      sx[i] = q.x;
      sy[i] = q.y;
      sz[i] = q.z;
    } else {
      sx[i] = P[(size_t)3 * i];
      sy[i] = P[(size_t)3 * i + 1];
      sz[i] = P[(size_t)3 * i + 2];
    }
  }
  for (int i = tid; i < 2 * kSlots; i += kThreads) slots[i] = 0ull;  // tag 0 is never waited for (j starts at 1)
  __syncthreads();
  const int s0 = rank * PPT * kThreads;
  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int k = 0; k < PPT; k++) {
    const int i = s0 + tid + k * kThreads;
    const bool ok = i < N;
    px[k] = ok ? sx[i] : 0.f;
    py[k] = ok ? sy[i] : 0.f;
    pz[k] = ok ? sz[i] : 0.f;
    md[k] = ok ? 1e10f : -1.f;
  }
  float cx = sx[0], cy = sy[0], cz = sz[0];
  if (rank == 0 && tid == 0) {
    out[(size_t)b * m] = 0;
    if (out_xyz) {
      out_xyz[(size_t)b * m * 3] = cx;
      out_xyz[(size_t)b * m * 3 + 1] = cy;
      out_xyz[(size_t)b * m * 3 + 2] = cz;
    }
  }
  cluster.sync();  // every CTA's slots are zeroed before the first remote store can land

  unsigned long long* my_remote = cluster.map_shared_rank(slots, lane & (CLUSTER - 1));
  const int my_slot = rank * WARPS + warp;
  for (int j = 1; j < m; j++) {
    const int par = j & 1;
    const unsigned int tag = (unsigned int)j & 0xffffu;
    float bd = -1.f;
    int bk = 0;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
      const float d = dist2(px[k], py[k], pz[k], cx, cy, cz);
      const float d2 = fminf(d, md[k]);
      md[k] = d2;
      if (d2 > bd) {
        bd = d2;
        bk = k;
      }
    }
    // (distance bits, 0xffff - index): max = farthest point, lowest index on ties; "no point" = (0, 0) never beats a point
    unsigned int d = bd < 0.f ? 0u : __float_as_uint(bd);
    unsigned int n = bd < 0.f ? 0u : 0xffffu - (unsigned int)(s0 + tid + bk * kThreads);
    warp_argmax(d, n);
    if (lane < CLUSTER)
      *reinterpret_cast<volatile unsigned long long*>(my_remote + par * kSlots + my_slot) =
          ((unsigned long long)d << 32) | (unsigned long long)((n << 16) | tag);
    // all words of this step, SL per lane
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(slots + par * kSlots + SL * lane);
    unsigned long long w0, w1 = 0ull, w2 = 0ull, w3 = 0ull;
    unsigned int spins = 0;
    for (;;) {
      bool ready;
      if (SL == 1) {
        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w0) : "r"(a));
        ready = ((unsigned int)w0 & 0xffffu) == tag;
      } else {
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(a));
        ready = (((unsigned int)w0 & 0xffffu) == tag) & (((unsigned int)w1 & 0xffffu) == tag);
        if (SL == 4) {
          asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "r"(a + 16));
          ready = ready & (((unsigned int)w2 & 0xffffu) == tag) & (((unsigned int)w3 & 0xffffu) == tag);
        }
      }
      if (__all_sync(0xffffffffu, ready)) break;
      if (++spins > (1u << 26)) __trap();  // a lost word would otherwise hang the GPU: fail loudly instead
    }
    w0 >>= 16;
    if (SL >= 2) {
      w1 >>= 16;
      w0 = w0 > w1 ? w0 : w1;
    }
    if (SL == 4) {
      w2 >>= 16;
      w3 >>= 16;
      w2 = w2 > w3 ? w2 : w3;
      w0 = w0 > w2 ? w0 : w2;
    }
    d = (unsigned int)(w0 >> 16);
    n = (unsigned int)w0 & 0xffffu;
    warp_argmax(d, n);
    const int win = (int)(0xffffu - n);
    cx = sx[win];
    cy = sy[win];
    cz = sz[win];
    if (rank == 0 && tid == 0) {
      out[(size_t)b * m + j] = win;
      if (out_xyz) {
        float* o = out_xyz + ((size_t)b * m + j) * 3;
        o[0] = cx;
        o[1] = cy;
        o[2] = cz;
      }
    }
  }
  cluster.sync();  // no CTA may exit while a peer can still store into its slots
}

// a8: out[b,c,j] = feat[b,c,idx[b,j]]
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ feat, const int* __restrict__ idx,
                                                     int C, int N, int m, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  out[((size_t)b * C + c) * m + j] = __ldg(&feat[((size_t)b * C + c) * N + idx[(size_t)b * m + j]]);
}

// a9: one thread per query; sources streamed through shared memory in tiles of 1024.
constexpr int kBqThreads = 256, kBqTile = 1024;
__global__ void __launch_bounds__(kBqThreads) ball_query_kernel(const float* __restrict__ xyz,
                                                               const float* __restrict__ new_xyz, int N, int M,
                                                               float r2, int nsample, int* __restrict__ out) {
  __shared__ float sx[kBqTile], sy[kBqTile], sz[kBqTile];
  const int b = blockIdx.y;
  const int q = blockIdx.x * kBqThreads + threadIdx.x;
  const bool live = q < M;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  int* o = nullptr;
  if (live) {
    const float* c = new_xyz + ((size_t)b * M + q) * 3;
    qx = c[0];
    qy = c[1];
    qz = c[2];
    o = out + ((size_t)b * M + q) * nsample;
    for (int l = 0; l < nsample; l++) o[l] = 0;  // no hit -> zeros (upstream zero-initialises idx)
  }
  int cnt = 0;
  const float* P = xyz + (size_t)b * N * 3;
  for (int t0 = 0; t0 < N; t0 += kBqTile) {
    const int nt = min(kBqTile, N - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt; e += kBqThreads) {
      sx[e] = P[3 * (size_t)(t0 + e)];
      sy[e] = P[3 * (size_t)(t0 + e) + 1];
      sz[e] = P[3 * (size_t)(t0 + e) + 2];
    }
    __syncthreads();
    const bool done = !live || cnt >= nsample;
    if (__syncthreads_and(done)) break;
    if (!done) {
      for (int e = 0; e < nt; e++) {
        const float d2 = dist2(qx, qy, qz, sx[e], sy[e], sz[e]);
        if (d2 < r2) {
          const int k = t0 + e;
          if (cnt == 0)
            for (int l = 0; l < nsample; l++) o[l] = k;
          o[cnt++] = k;
          if (cnt >= nsample) break;
        }
      }
    }
  }
}

// a10: out[b,c,j,l] = feat[b,c,idx[b,j,l]]
__global__ void __launch_bounds__(256) group_kernel(const float* __restrict__ feat, const int* __restrict__ idx,
                                                    int C, int N, int MS, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= MS) return;
  out[((size_t)b * C + c) * MS + e] = __ldg(&feat[((size_t)b * C + c) * N + idx[(size_t)b * MS + e]]);
}

// QueryAndGroup(use_xyz=True): channels 0..2 = xyz[idx] - new_xyz (plain fp32 subtraction), then feat[idx]
__global__ void __launch_bounds__(256) query_group_kernel(const float* __restrict__ xyz,
                                                          const float* __restrict__ new_xyz,
                                                          const float* __restrict__ feat,
                                                          const int* __restrict__ idx, int C, int N, int M, int ns,
                                                          float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int MS = M * ns;
  if (e >= MS) return;
  const int CT = 3 + (feat ? C : 0);
  const int k = idx[(size_t)b * MS + e];
  float v;
  if (c < 3) {
    const int q = e / ns;
    v = __fsub_rn(__ldg(&xyz[((size_t)b * N + k) * 3 + c]), __ldg(&new_xyz[((size_t)b * M + q) * 3 + c]));
  } else {
    v = __ldg(&feat[((size_t)b * C + (c - 3)) * N + k]);
  }
  out[((size_t)b * CT + c) * MS + e] = v;
}


// ---------------------------------------------------------------------------------------------------------
// a9, multi-scale: one WARP per query, all radii of a PointnetSAModuleMSG in one pass over the sources.
// The thread-per-query kernel above keeps 64 CTAs busy at B=8, M=2048 (5 % of the machine's thread slots) and
// walks the sources once per radius; here 8 queries share a CTA, the lanes test 32 consecutive sources per step
// (ballot -> hits appended in ascending index order, exactly the sequential rule), and the scan stops as soon as
// every radius has its nsample hits. Sources may be RAGGED: frame b owns rows [row_offsets[b], row_offsets[b+1])
// of a packed (rows, stride) array -- the layout the sparse levels already have -- so the reference's
// pad_batch (random duplicate rows appended to make the batch dense, sparse_cnn.py:118-126) is not needed:
// duplicates appended AFTER the real rows can only occupy slots the first hit would have filled, and they carry
// the features of real hits of the same ball, so the max-pooled result is unchanged.
// ---------------------------------------------------------------------------------------------------------
constexpr int kBqWarps = 8, kBqTileW = 2048, kBqMaxR = 4;
struct BqArgs {
  float r2[kBqMaxR];
  int ns[kBqMaxR];
  int* out[kBqMaxR];
};

template <int R>
__global__ void __launch_bounds__(kBqWarps * 32) ball_query_warp_kernel(const float* __restrict__ xyz, int stride, int N,
                                                                        const int* __restrict__ row_offsets,
                                                                        const float* __restrict__ new_xyz, int M, BqArgs A) {
  __shared__ float sx[kBqTileW], sy[kBqTileW], sz[kBqTileW];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kBqWarps + warp;
  int base = b * N, n = N;
  if (row_offsets) {
    base = __ldg(&row_offsets[b]);
    n = __ldg(&row_offsets[b + 1]) - base;
  }
  const float* P = xyz + (size_t)base * stride;
  const bool live = q < M;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const float* c = new_xyz + ((size_t)b * M + q) * 3;
    qx = __ldg(c);
    qy = __ldg(c + 1);
    qz = __ldg(c + 2);
  }
  int cnt[R], first[R];
#pragma unroll
  for (int r = 0; r < R; r++) cnt[r] = 0, first[r] = 0;
  bool done = !live;
  const unsigned lt = (1u << lane) - 1u;
  for (int t0 = 0; t0 < n; t0 += kBqTileW) {
    const int nt = min(kBqTileW, n - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt; e += kBqWarps * 32) {
      const float* s = P + (size_t)(t0 + e) * stride;
      sx[e] = __ldg(s);
      sy[e] = __ldg(s + 1);
      sz[e] = __ldg(s + 2);
    }
    if (__syncthreads_and(done)) break;
    if (done) continue;
    for (int e0 = 0; e0 < nt; e0 += 32) {
      const int e = e0 + lane;
      const float d2 = e < nt ? dist2(qx, qy, qz, sx[e], sy[e], sz[e]) : 3.0e38f;
      bool all_full = true;
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (cnt[r] >= A.ns[r]) continue;
        const bool hit = d2 < A.r2[r];
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          if (cnt[r] == 0) first[r] = t0 + e0 + __ffs(m) - 1;
          const int pos = cnt[r] + __popc(m & lt);
          if (hit && pos < A.ns[r]) A.out[r][((size_t)b * M + q) * A.ns[r] + pos] = t0 + e;
          cnt[r] += __popc(m);
        }
        all_full = all_full && cnt[r] >= A.ns[r];
      }
      if (all_full) {
        done = true;
        break;
      }
    }
  }
  if (live) {
#pragma unroll
    for (int r = 0; r < R; r++)  // unused slots = first hit; no hit at all = zeros (upstream zero-initialises idx)
      for (int l = min(cnt[r], A.ns[r]) + lane; l < A.ns[r]; l += 32) A.out[r][((size_t)b * M + q) * A.ns[r] + l] = first[r];
  }
}

// ---------------------------------------------------------------------------------------------------------
// a9 with hierarchical culling. Every 32 consecutive source rows ("chunk") get an axis-aligned bounding box
// (bq_bounds_kernel, one warp per chunk). A query warp tests 32 chunk boxes at once (one per lane): the squared
// distance from the query to a box is a lower bound -- evaluated with the same rounded operations in the same
// order as dist2, so it is a lower bound in floating point too -- of the distance to every point inside it, so a
// chunk whose bound is >= r^2 of the largest radius contains no hit and is skipped. Surviving chunks are tested
// point by point in ascending row order exactly as in ball_query_warp_kernel: the result is bit-identical.
// Rows of the conv-produced sparse levels are in ascending (b, z, y, x) order, so a chunk is a thin slab (one z,
// a few y lines) and a ball of 1-5 m meets ~5 % of them; for sources in random order (raw points, level 0) every
// box spans the scene and the plain kernel is used instead.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bq_bounds_kernel(const float* __restrict__ xyz, int stride, int N,
                                                        const int* __restrict__ row_offsets, int max_chunks,
                                                        float4* __restrict__ lo, float4* __restrict__ hi) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= max_chunks) return;
  int base = b * N, n = N;
  if (row_offsets) {
    base = __ldg(&row_offsets[b]);
    n = __ldg(&row_offsets[b + 1]) - base;
  }
  const int r = c * 32 + lane;
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (r < n) {
    const float* p = xyz + (size_t)(base + r) * stride;
#pragma unroll
    for (int d = 0; d < 3; d++) mn[d] = mx[d] = __ldg(&p[d]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int d = 0; d < 3; d++) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if (lane == 0) {
    lo[(size_t)b * max_chunks + c] = make_float4(mn[0], mn[1], mn[2], 0.f);
    hi[(size_t)b * max_chunks + c] = make_float4(mx[0], mx[1], mx[2], 0.f);
  }
}

template <int R>
__global__ void __launch_bounds__(kBqWarps * 32) ball_query_cull_kernel(const float* __restrict__ xyz, int stride, int N,
                                                                        const int* __restrict__ row_offsets,
                                                                        const float4* __restrict__ lo,
                                                                        const float4* __restrict__ hi, int max_chunks,
                                                                        const float* __restrict__ new_xyz, int M, BqArgs A) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kBqWarps + warp;
  if (q >= M) return;  // (no block-wide barrier in this kernel)
  int base = b * N, n = N;
  if (row_offsets) {
    base = __ldg(&row_offsets[b]);
    n = __ldg(&row_offsets[b + 1]) - base;
  }
  const float* P = xyz + (size_t)base * stride;
  const float* c = new_xyz + ((size_t)b * M + q) * 3;
  const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
  float r2max = 0.f;
#pragma unroll
  for (int r = 0; r < R; r++) r2max = fmaxf(r2max, A.r2[r]);
  int cnt[R], first[R];
#pragma unroll
  for (int r = 0; r < R; r++) cnt[r] = 0, first[r] = 0;
  const unsigned lt = (1u << lane) - 1u;
  const int n_chunks = (n + 31) >> 5;
  bool done = false;
  for (int sc = 0; sc < n_chunks && !done; sc += 32) {
    const int ci = sc + lane;
    bool alive = false;
    if (ci < n_chunks) {
      const float4 l = __ldg(&lo[(size_t)b * max_chunks + ci]), h = __ldg(&hi[(size_t)b * max_chunks + ci]);
      // per-axis gap between the query and the box: the same rounded subtraction dist2 performs on a point's
      // coordinate, applied to the nearest face -> a lower bound of |q - p| on that axis for every p in the box
      const float gx = fmaxf(fmaxf(__fsub_rn(l.x, qx), __fsub_rn(qx, h.x)), 0.f);
      const float gy = fmaxf(fmaxf(__fsub_rn(l.y, qy), __fsub_rn(qy, h.y)), 0.f);
      const float gz = fmaxf(fmaxf(__fsub_rn(l.z, qz), __fsub_rn(qz, h.z)), 0.f);
      const float bound = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
      alive = bound < r2max;
    }
    unsigned live = __ballot_sync(0xffffffffu, alive);
    while (live && !done) {
      const int cc = sc + __ffs(live) - 1;
      live &= live - 1;
      const int e = cc * 32 + lane;
      float d2 = 3.0e38f;
      if (e < n) {
        const float* s = P + (size_t)e * stride;
        d2 = dist2(qx, qy, qz, __ldg(s), __ldg(s + 1), __ldg(s + 2));
      }
      bool all_full = true;
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (cnt[r] >= A.ns[r]) continue;
        const bool hit = d2 < A.r2[r];
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          if (cnt[r] == 0) first[r] = cc * 32 + __ffs(m) - 1;
          const int pos = cnt[r] + __popc(m & lt);
          if (hit && pos < A.ns[r]) A.out[r][((size_t)b * M + q) * A.ns[r] + pos] = e;
          cnt[r] += __popc(m);
        }
        all_full = all_full && cnt[r] >= A.ns[r];
      }
      done = all_full;
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++)  // unused slots = first hit; no hit at all = zeros (upstream zero-initialises idx)
    for (int l = min(cnt[r], A.ns[r]) + lane; l < A.ns[r]; l += 32) A.out[r][((size_t)b * M + q) * A.ns[r] + l] = first[r];
}

// ---------------------------------------------------------------------------------------------------------
// a9 for sources in RANDOM row order (raw points are shuffled, level-0 voxels are in first-appearance order): chunk
// boxes of consecutive rows span the whole scene, so the culled kernel above cannot skip anything. Here the rows of
// every frame are first bucketed along x (counting sort into 0.25 m slabs: histogram, scan, scatter -- the order
// INSIDE a slab is whatever the atomics produce, which does not matter, see below) into float4 rows
// {x, y, z, original index}; chunk boxes of the bucketed rows are thin in x and the same box test skips ~97 % of
// them. The reference semantics "the first nsample hits in ascending ORIGINAL index order" is then a selection
// problem: a query warp keeps, per radius, the nsample smallest original indices among all hits in a sorted
// register array (one entry per lane), inserting a new hit only when it beats the current worst. The hit set is
// the same (same rounded distance test) and the nsample smallest indices of a set do not depend on the order the
// set is enumerated in, so the result is bit-identical to the sequential scan -- and deterministic although the
// bucketing is not.
// ---------------------------------------------------------------------------------------------------------
constexpr int kBsMaxBuckets = 1024;

__global__ void __launch_bounds__(256) bs_hist_kernel(const float* __restrict__ xyz, int stride, int N,
                                                      const int* __restrict__ row_offsets, float x0, float inv_w,
                                                      int n_buckets, int* __restrict__ cnt) {
  const int b = blockIdx.y;
  int base = b * N, n = N;
  if (row_offsets) {
    base = __ldg(&row_offsets[b]);
    n = __ldg(&row_offsets[b + 1]) - base;
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const float x = __ldg(&xyz[(size_t)(base + r) * stride]);
    const int k = min(max((int)floorf((x - x0) * inv_w), 0), n_buckets - 1);
    atomicAdd(&cnt[b * kBsMaxBuckets + k], 1);
  }
}

__global__ void __launch_bounds__(kBsMaxBuckets) bs_scan_kernel(int* __restrict__ cnt, int* __restrict__ cursor) {
  __shared__ int sm[33];
  const int b = blockIdx.x;
  const int v = cnt[b * kBsMaxBuckets + threadIdx.x];
  int total;
  const int ex = block_exclusive_scan(v, sm, total);
  cursor[b * kBsMaxBuckets + threadIdx.x] = ex;
  cnt[b * kBsMaxBuckets + threadIdx.x] = 0;  // ready for the next call
}

__global__ void __launch_bounds__(256) bs_scatter_kernel(const float* __restrict__ xyz, int stride, int N,
                                                         const int* __restrict__ row_offsets, float x0, float inv_w,
                                                         int n_buckets, int* __restrict__ cursor,
                                                         float4* __restrict__ sorted) {
  const int b = blockIdx.y;
  int base = b * N, n = N;
  if (row_offsets) {
    base = __ldg(&row_offsets[b]);
    n = __ldg(&row_offsets[b + 1]) - base;
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const float* p = xyz + (size_t)(base + r) * stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const int k = min(max((int)floorf((x - x0) * inv_w), 0), n_buckets - 1);
    const int pos = atomicAdd(&cursor[b * kBsMaxBuckets + k], 1);
    sorted[(size_t)base + pos] = make_float4(x, y, z, __int_as_float(r));
  }
}

template <int R>
__global__ void __launch_bounds__(kBqWarps * 32) ball_query_select_kernel(const float4* __restrict__ sorted, int N,
                                                                          const int* __restrict__ row_offsets,
                                                                          const float4* __restrict__ lo,
                                                                          const float4* __restrict__ hi, int max_chunks,
                                                                          const float* __restrict__ new_xyz, int M, BqArgs A) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kBqWarps + warp;
  if (q >= M) return;
  int base = b * N, n = N;
  if (row_offsets) {
    base = __ldg(&row_offsets[b]);
    n = __ldg(&row_offsets[b + 1]) - base;
  }
  const float4* P = sorted + base;
  const float* c = new_xyz + ((size_t)b * M + q) * 3;
  const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
  float r2max = 0.f;
#pragma unroll
  for (int r = 0; r < R; r++) r2max = fmaxf(r2max, A.r2[r]);
  constexpr int kNone = 0x7fffffff;
  int best[R];  // lane l: the l-th smallest original index among the hits of radius r seen so far (l < ns[r])
#pragma unroll
  for (int r = 0; r < R; r++) best[r] = kNone;
  const int n_chunks = (n + 31) >> 5;
  for (int sc = 0; sc < n_chunks; sc += 32) {
    const int ci = sc + lane;
    bool alive = false;
    if (ci < n_chunks) {
      const float4 l = __ldg(&lo[(size_t)b * max_chunks + ci]), h = __ldg(&hi[(size_t)b * max_chunks + ci]);
      const float gx = fmaxf(fmaxf(__fsub_rn(l.x, qx), __fsub_rn(qx, h.x)), 0.f);
      const float gy = fmaxf(fmaxf(__fsub_rn(l.y, qy), __fsub_rn(qy, h.y)), 0.f);
      const float gz = fmaxf(fmaxf(__fsub_rn(l.z, qz), __fsub_rn(qz, h.z)), 0.f);
      alive = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz)) < r2max;
    }
    unsigned live = __ballot_sync(0xffffffffu, alive);
    while (live) {
      const int cc = sc + __ffs(live) - 1;
      live &= live - 1;
      const int e = cc * 32 + lane;
      float d2 = 3.0e38f;
      int oi = kNone;
      if (e < n) {
        const float4 p = __ldg(&P[e]);
        d2 = dist2(qx, qy, qz, p.x, p.y, p.z);
        oi = __float_as_int(p.w);
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        unsigned m = __ballot_sync(0xffffffffu, d2 < A.r2[r]);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int v = __shfl_sync(0xffffffffu, oi, src);
          const int worst = __shfl_sync(0xffffffffu, best[r], A.ns[r] - 1);
          if (v < worst) {  // (warp-uniform) insert v, dropping the current worst
            const int pos = __popc(__ballot_sync(0xffffffffu, best[r] < v));  // entries are sorted: a prefix is smaller
            const int up = __shfl_up_sync(0xffffffffu, best[r], 1);
            if (lane >= pos) best[r] = lane == pos ? v : up;
            if (lane >= A.ns[r]) best[r] = kNone;
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int cnt = __popc(__ballot_sync(0xffffffffu, best[r] != kNone));
    const int first = __shfl_sync(0xffffffffu, best[r], 0);
    if (lane < A.ns[r])  // unused slots = first hit; no hit at all = zeros (upstream zero-initialises idx)
      A.out[r][((size_t)b * M + q) * A.ns[r] + lane] = lane < cnt ? best[r] : (cnt > 0 ? first : 0);
  }
}

// QueryAndGroup on ROW-major sources (what the sparse levels are): out[b, c, q, l] for c < 3 is
// xyz[row][c] - new_xyz[b, q, c] and feat[row][c - 3] after, row = row_offsets[b] + idx[b, q, l] (or b*N + idx).
// One thread per (b, q, l) walks the channels: the source row is read once, contiguously.
__global__ void __launch_bounds__(256) query_group_rows_kernel(const float* __restrict__ xyz, int xyz_stride,
                                                               const float* __restrict__ feat, int feat_stride, int C, int N,
                                                               const int* __restrict__ row_offsets,
                                                               const float* __restrict__ new_xyz,
                                                               const int* __restrict__ idx, int M, int ns,
                                                               float* __restrict__ out) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int MS = M * ns;
  if (e >= MS) return;
  const int base = row_offsets ? __ldg(&row_offsets[b]) : b * N;
  const size_t row = (size_t)base + idx[(size_t)b * MS + e];
  const int q = e / ns;
  const int CT = 3 + C;
  float* o = out + (size_t)b * CT * MS + e;
#pragma unroll
  for (int c = 0; c < 3; c++)
    o[(size_t)c * MS] = __fsub_rn(__ldg(&xyz[row * xyz_stride + c]), __ldg(&new_xyz[((size_t)b * M + q) * 3 + c]));
  const float* f = feat + row * feat_stride;
  for (int c = 0; c < C; c++) o[(size_t)(3 + c) * MS] = __ldg(&f[c]);
}

// a15 on the device: start row of every frame in a (b,z,y,x)-sorted index list = searchsorted(batch column, 0..B)
// (compute_pad_amounts, detector/sparse_cnn.py:107-116, without its .cpu().numpy() round trip)
__global__ void batch_offsets_kernel(const int4* __restrict__ idx, const int* __restrict__ n_rows, int cap, int B,
                                     int* __restrict__ offsets) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  const int n = min(*n_rows, cap);
  int lo = 0, hi = n;  // first row whose batch index is >= b
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (idx[mid].x < b) lo = mid + 1; else hi = mid;
  }
  offsets[b] = lo;
}

// to_global (detector/sparse_cnn.py:91-105): voxel index (b,z,y,x) -> metric xyz = float(x,y,z) * voxel_size + offset,
// evaluated as torch does (one rounded multiply, one rounded add)
__global__ void __launch_bounds__(256) to_global_kernel(const int4* __restrict__ idx, const int* __restrict__ n_rows, int cap,
                                                        float3 vs, float3 off, float* __restrict__ xyz) {
  const int n = min(*n_rows, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 c = idx[i];
    xyz[(size_t)i * 3] = __fadd_rn(__fmul_rn((float)c.w, vs.x), off.x);
    xyz[(size_t)i * 3 + 1] = __fadd_rn(__fmul_rn((float)c.z, vs.y), off.y);
    xyz[(size_t)i * 3 + 2] = __fadd_rn(__fmul_rn((float)c.y, vs.z), off.z);
  }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// pad_batch / pad_for_batch (sparse_cnn.py:118-126, core/preprocess.py:35-45): ragged rows -> dense (B, cap, C),
// frames shorter than `cap` are filled with uniformly chosen duplicates of their own rows (counter-based RNG:
// row j of frame b picks splitmix64(seed, b, j) % count -- the same pick for every tensor padded with the same
// seed, so padded xyz and padded features stay paired).
__global__ void __launch_bounds__(256) pad_batch_kernel(const float* __restrict__ src, int C,
                                                        const int* __restrict__ row_offsets, int cap,
                                                        unsigned long long seed, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int start = __ldg(&row_offsets[b]), cnt = __ldg(&row_offsets[b + 1]) - start;
  const long long total = (long long)cap * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e / C), c = (int)(e % C);
    float v = 0.f;
    if (cnt > 0) {
      const int r = j < cnt ? j : (int)(splitmix64(seed ^ (((unsigned long long)b << 32) | (unsigned)j)) % (unsigned long long)cnt);
      v = __ldg(&src[(size_t)(start + r) * C + c]);
    }
    out[((size_t)b * cap + j) * C + c] = v;
  }
}

// BEVFeatureGatherer.forward (detector/layers.py:30-50): bilinear F.grid_sample(align_corners=True, zero padding) of
// the channels-last BEV map at the keypoints, index arithmetic restated op for op in fp32 (including the reference's
// (size - 2) normaliser and its x <-> W swap) so that positions agree with the torch expression to the last bit or
// two. A CTA handles 32 keypoints: a warp reads the four corner rows of a keypoint with 128-bit loads (lane = 4
// channels), results are transposed through shared memory and written as 128-byte rows of the channel-major output.
__global__ void __launch_bounds__(256) bev_gather_kernel(const float* __restrict__ map, int H, int W, int C,
                                                         const float* __restrict__ kp, int M, float x_off, float y_off,
                                                         float px, float py, float* __restrict__ out, int c_total,
                                                         int c_off) {
  __shared__ float tile[128][33];
  const int b = blockIdx.y, m0 = blockIdx.x * 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = 0; cb < C; cb += 128) {
    for (int q = warp; q < 32; q += 8) {
      const int m = m0 + q;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = cb + 4 * lane;
      if (m < M && c < C) {
        const float* k = kp + ((size_t)b * M + m) * 3;
        // indices = (xy - pixel_offset) / (base_pixel * stride); clamp to [0, dims]; 2 * (i / (dims - 1)) - 1
        float i0 = __fdiv_rn(__fsub_rn(__ldg(k), x_off), px);       // along x, clamped with W - 1
        float i1 = __fdiv_rn(__fsub_rn(__ldg(k + 1), y_off), py);   // along y, clamped with H - 1
        i0 = fminf(fmaxf(i0, 0.f), (float)(W - 1));
        i1 = fminf(fmaxf(i1, 0.f), (float)(H - 1));
        const float n0 = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(i0, (float)(W - 2))), 1.f);
        const float n1 = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(i1, (float)(H - 2))), 1.f);
        // after .flip(3): grid x = n1 (samples along W), grid y = n0 (samples along H)
        const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(n1, 1.f), 2.f), (float)(W - 1));
        const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(n0, 1.f), 2.f), (float)(H - 1));
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.f), ix);
        const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.f), iy);
        const float wgt[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy1)};
#pragma unroll
        for (int t = 0; t < 4; t++) {   // nw, ne, sw, se: the order torch accumulates in
          const int x = x0 + (t & 1), y = y0 + (t >> 1);
          if (x >= 0 && x < W && y >= 0 && y < H) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(map + (((size_t)b * H + y) * W + x) * C + c));
            acc.x = fmaf(v.x, wgt[t], acc.x);
            acc.y = fmaf(v.y, wgt[t], acc.y);
            acc.z = fmaf(v.z, wgt[t], acc.z);
            acc.w = fmaf(v.w, wgt[t], acc.w);
          }
        }
      }
      tile[4 * lane][q] = acc.x;
      tile[4 * lane + 1][q] = acc.y;
      tile[4 * lane + 2][q] = acc.z;
      tile[4 * lane + 3][q] = acc.w;
    }
    __syncthreads();
    for (int cc = warp; cc < 128; cc += 8)
      if (cb + cc < C && m0 + lane < M) out[((size_t)b * c_total + c_off + cb + cc) * M + m0 + lane] = tile[cc][lane];
    __syncthreads();
  }
}

template <int PPT>
int launch_fps(const float* xyz, int stride, int B, int N, int m, int* idx, float* out_xyz, cudaStream_t st) {
  fps_cluster_kernel<PPT><<<B * kFpsCluster, kFpsThreads, 0, st>>>(xyz, stride, N, m, idx, out_xyz);
  return check_launch();
}

template <int PPT, int WARPS, int CLUSTER>
int launch_fps_direct(const float* xyz, int stride, int B, int N, int m, int* idx, float* out_xyz, cudaStream_t st) {
  const int n_pad = (N + 3) & ~3;
  const size_t slot_bytes = (size_t)2 * WARPS * CLUSTER * sizeof(unsigned long long);
  auto kern = fps_direct_kernel<PPT, WARPS, CLUSTER>;
  static PerDeviceOnce once;
  if (once.needed()) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(slot_bytes + 3 * kFpsDirectMaxN * sizeof(float))));
    if (CLUSTER > 8) V3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    once.done();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * CLUSTER);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = slot_bytes + (size_t)3 * n_pad * sizeof(float);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  V3D_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, xyz, stride, N, n_pad, m, idx, out_xyz));
  return check_launch();
}

template <int WARPS, int CLUSTER>
int launch_fps_direct_n(const float* xyz, int stride, int B, int N, int m, int* idx, float* out_xyz, cudaStream_t st) {
  const int ppt = ceil_div(N, WARPS * 32 * CLUSTER);
  if (ppt <= 1) return launch_fps_direct<1, WARPS, CLUSTER>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 2) return launch_fps_direct<2, WARPS, CLUSTER>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 4) return launch_fps_direct<4, WARPS, CLUSTER>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 8) return launch_fps_direct<8, WARPS, CLUSTER>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 16) return launch_fps_direct<16, WARPS, CLUSTER>(xyz, stride, B, N, m, idx, out_xyz, st);
  return V3D_ERR_INVALID_ARGUMENT;
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" size_t v3d_fps_workspace_bytes(int B, int N) {
  (void)B;
  (void)N;
  return 256;  // everything lives in registers / distributed shared memory
}

static int fps_dispatch(const float* xyz, int stride, int B, int N, int m, int* idx, float* out_xyz, cudaStream_t st) {
  if (!xyz || !idx || B <= 0 || N <= 0 || m <= 0 || (stride != 3 && stride != 4)) return V3D_ERR_INVALID_ARGUMENT;
  if (stride == 4 && (reinterpret_cast<uintptr_t>(xyz) & 15)) return V3D_ERR_INVALID_ARGUMENT;
  const int per_cta = ceil_div(N, kFpsCluster);
  const int ppt = ceil_div(per_cta, kFpsThreads);
  // 4 warps x 8 CTAs: measured best of {16, 8, 4 warps} x {8, 16 CTAs} (profiles/r02_fps_sweep.json): fewer warps =
  // fewer words to exchange and poll per step; 16-CTA clusters do not all fit on the chip at once for B = 8
  if (N <= kFpsDirectMaxN) return launch_fps_direct_n<4, 8>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 1) return launch_fps<1>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 2) return launch_fps<2>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 4) return launch_fps<4>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 8) return launch_fps<8>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 16) return launch_fps<16>(xyz, stride, B, N, m, idx, out_xyz, st);
  if (ppt <= 32) return launch_fps<32>(xyz, stride, B, N, m, idx, out_xyz, st);
  return V3D_ERR_INVALID_ARGUMENT;  // > 65536 points per cloud
}

extern "C" int v3d_fps(const float* xyz, int B, int N, int m, int* idx, void* workspace, size_t workspace_bytes,
                       v3d_stream_t stream) {
  (void)workspace;
  (void)workspace_bytes;
  return fps_dispatch(xyz, 3, B, N, m, idx, nullptr, as_stream(stream));
}

extern "C" int v3d_fps_keypoints(const float* points, int point_stride, int B, int N, int m, int* idx,
                                 float* keypoints, v3d_stream_t stream) {
  return fps_dispatch(points, point_stride, B, N, m, idx, keypoints, as_stream(stream));
}

extern "C" int v3d_gather(const float* feat, const int* idx, int B, int C, int N, int m, float* out,
                          v3d_stream_t stream) {
  if (!feat || !idx || !out || B <= 0 || C <= 0 || N <= 0 || m <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (C > 65535 || B > 65535) return V3D_ERR_INVALID_ARGUMENT;
  gather_kernel<<<dim3(ceil_div(m, 256), C, B), 256, 0, as_stream(stream)>>>(feat, idx, C, N, m, out);
  return check_launch();
}

extern "C" int v3d_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M, float radius,
                              int nsample, int* idx, v3d_stream_t stream) {
  if (!xyz || !new_xyz || !idx || B <= 0 || N <= 0 || M <= 0 || nsample <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (B > 65535) return V3D_ERR_INVALID_ARGUMENT;
  const float r2 = radius * radius;
  ball_query_kernel<<<dim3(ceil_div(M, kBqThreads), B), kBqThreads, 0, as_stream(stream)>>>(xyz, new_xyz, N, M, r2,
                                                                                       nsample, idx);
  return check_launch();
}

extern "C" int v3d_group(const float* feat, const int* idx, int B, int C, int N, int M, int nsample, float* out,
                         v3d_stream_t stream) {
  if (!feat || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M <= 0 || nsample <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (C > 65535 || B > 65535) return V3D_ERR_INVALID_ARGUMENT;
  const int MS = M * nsample;
  group_kernel<<<dim3(ceil_div(MS, 256), C, B), 256, 0, as_stream(stream)>>>(feat, idx, C, N, MS, out);
  return check_launch();
}

extern "C" int v3d_query_and_group(const float* xyz, const float* new_xyz, const float* feat, const int* idx, int B,
                                   int C, int N, int M, int nsample, float* out, v3d_stream_t stream) {
  if (!xyz || !new_xyz || !idx || !out || B <= 0 || N <= 0 || M <= 0 || nsample <= 0) return V3D_ERR_INVALID_ARGUMENT;
  const int CT = 3 + (feat ? C : 0);
  if (CT > 65535 || B > 65535) return V3D_ERR_INVALID_ARGUMENT;
  const int MS = M * nsample;
  query_group_kernel<<<dim3(ceil_div(MS, 256), CT, B), 256, 0, as_stream(stream)>>>(xyz, new_xyz, feat, idx, C, N, M,
                                                                               nsample, out);
  return check_launch();
}

extern "C" int v3d_ball_query_msg(const float* xyz, int point_stride, const int* row_offsets, const float* new_xyz, int B,
                                  int N, int M, int n_radii, const float* radii_host, const int* nsamples_host,
                                  int* const* idx_host, v3d_stream_t stream) {
  if (!xyz || !new_xyz || !radii_host || !nsamples_host || !idx_host) return V3D_ERR_INVALID_ARGUMENT;
  if (B <= 0 || B > 65535 || M <= 0 || n_radii <= 0 || n_radii > kBqMaxR || point_stride < 3) return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && N <= 0) return V3D_ERR_INVALID_ARGUMENT;
  BqArgs A;
  for (int r = 0; r < kBqMaxR; r++) {
    const int s = r < n_radii ? r : 0;
    if (nsamples_host[s] <= 0 || !idx_host[s]) return V3D_ERR_INVALID_ARGUMENT;
    A.r2[r] = radii_host[s] * radii_host[s];
    A.ns[r] = nsamples_host[s];
    A.out[r] = idx_host[s];
  }
  dim3 grid(ceil_div(M, kBqWarps), B);
  cudaStream_t st = as_stream(stream);
  switch (n_radii) {
    case 1: ball_query_warp_kernel<1><<<grid, kBqWarps * 32, 0, st>>>(xyz, point_stride, N, row_offsets, new_xyz, M, A); break;
    case 2: ball_query_warp_kernel<2><<<grid, kBqWarps * 32, 0, st>>>(xyz, point_stride, N, row_offsets, new_xyz, M, A); break;
    case 3: ball_query_warp_kernel<3><<<grid, kBqWarps * 32, 0, st>>>(xyz, point_stride, N, row_offsets, new_xyz, M, A); break;
    default: ball_query_warp_kernel<4><<<grid, kBqWarps * 32, 0, st>>>(xyz, point_stride, N, row_offsets, new_xyz, M, A); break;
  }
  return check_launch();
}

extern "C" int v3d_query_and_group_rows(const float* xyz, int xyz_stride, const float* feat, int feat_stride, int C, int N,
                                        const int* row_offsets, const float* new_xyz, const int* idx, int B, int M,
                                        int nsample, float* out, v3d_stream_t stream) {
  if (!xyz || !new_xyz || !idx || !out || B <= 0 || B > 65535 || M <= 0 || nsample <= 0 || xyz_stride < 3 || C < 0)
    return V3D_ERR_INVALID_ARGUMENT;
  if (C > 0 && (!feat || feat_stride < C)) return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && N <= 0) return V3D_ERR_INVALID_ARGUMENT;
  query_group_rows_kernel<<<dim3(ceil_div(M * nsample, 256), B), 256, 0, as_stream(stream)>>>(
      xyz, xyz_stride, feat, feat_stride, C, N, row_offsets, new_xyz, idx, M, nsample, out);
  return check_launch();
}

extern "C" int v3d_batch_offsets(const int* indices, const int* n_rows, int capacity_rows, int B, int* offsets,
                                 v3d_stream_t stream) {
  if (!indices || !n_rows || !offsets || B <= 0 || capacity_rows < 0) return V3D_ERR_INVALID_ARGUMENT;
  batch_offsets_kernel<<<ceil_div(B + 1, 128), 128, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(indices), n_rows,
                                                                          capacity_rows, B, offsets);
  return check_launch();
}

extern "C" int v3d_to_global(const int* indices, const int* n_rows, int capacity_rows, const float* voxel_size_host,
                             const float* offset_host, float* xyz, v3d_stream_t stream) {
  if (!indices || !n_rows || !voxel_size_host || !offset_host || !xyz || capacity_rows <= 0) return V3D_ERR_INVALID_ARGUMENT;
  const int want = ceil_div(capacity_rows, 256), cap_blocks = kNumSMs * 8;
  to_global_kernel<<<want < cap_blocks ? want : cap_blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const int4*>(indices), n_rows, capacity_rows,
      make_float3(voxel_size_host[0], voxel_size_host[1], voxel_size_host[2]),
      make_float3(offset_host[0], offset_host[1], offset_host[2]), xyz);
  return check_launch();
}

extern "C" int v3d_pad_batch(const float* src, int C, const int* row_offsets, int B, int frame_capacity,
                             unsigned long long seed, float* out, v3d_stream_t stream) {
  if (!src || !row_offsets || !out || C <= 0 || B <= 0 || B > 65535 || frame_capacity <= 0) return V3D_ERR_INVALID_ARGUMENT;
  const long long total = (long long)frame_capacity * C;
  const long long want = (total + 255) / 256;
  const int blocks = (int)(want < kNumSMs * 4 ? want : kNumSMs * 4);
  pad_batch_kernel<<<dim3(blocks, B), 256, 0, as_stream(stream)>>>(src, C, row_offsets, frame_capacity, seed, out);
  return check_launch();
}

extern "C" size_t v3d_ball_query_bounds_bytes(int B, int max_rows_per_frame) {
  if (B <= 0 || max_rows_per_frame <= 0) return 0;
  return (size_t)2 * B * ((max_rows_per_frame + 31) / 32) * sizeof(float4);
}

extern "C" int v3d_ball_query_bounds(const float* xyz, int point_stride, const int* row_offsets, int B, int N,
                                     int max_rows_per_frame, void* bounds, v3d_stream_t stream) {
  if (!xyz || !bounds || B <= 0 || B > 65535 || point_stride < 3 || max_rows_per_frame <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && (N <= 0 || N > max_rows_per_frame)) return V3D_ERR_INVALID_ARGUMENT;
  const int max_chunks = (max_rows_per_frame + 31) / 32;
  float4* lo = static_cast<float4*>(bounds);
  float4* hi = lo + (size_t)B * max_chunks;
  bq_bounds_kernel<<<dim3(ceil_div(max_chunks, 8), B), 256, 0, as_stream(stream)>>>(xyz, point_stride, N, row_offsets,
                                                                                  max_chunks, lo, hi);
  return check_launch();
}

extern "C" int v3d_ball_query_msg_culled(const float* xyz, int point_stride, const int* row_offsets, const void* bounds,
                                         int max_rows_per_frame, const float* new_xyz, int B, int N, int M, int n_radii,
                                         const float* radii_host, const int* nsamples_host, int* const* idx_host,
                                         v3d_stream_t stream) {
  if (!xyz || !new_xyz || !bounds || !radii_host || !nsamples_host || !idx_host) return V3D_ERR_INVALID_ARGUMENT;
  if (B <= 0 || B > 65535 || M <= 0 || n_radii <= 0 || n_radii > kBqMaxR || point_stride < 3 || max_rows_per_frame <= 0)
    return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && (N <= 0 || N > max_rows_per_frame)) return V3D_ERR_INVALID_ARGUMENT;
  BqArgs A;
  for (int r = 0; r < kBqMaxR; r++) {
    const int s = r < n_radii ? r : 0;
    if (nsamples_host[s] <= 0 || !idx_host[s]) return V3D_ERR_INVALID_ARGUMENT;
    A.r2[r] = radii_host[s] * radii_host[s];
    A.ns[r] = nsamples_host[s];
    A.out[r] = idx_host[s];
  }
  const int max_chunks = (max_rows_per_frame + 31) / 32;
  const float4* lo = static_cast<const float4*>(bounds);
  const float4* hi = lo + (size_t)B * max_chunks;
  dim3 grid(ceil_div(M, kBqWarps), B);
  cudaStream_t st = as_stream(stream);
#define V3D_BQC(RR) \
  ball_query_cull_kernel<RR><<<grid, kBqWarps * 32, 0, st>>>(xyz, point_stride, N, row_offsets, lo, hi, max_chunks, new_xyz, M, A)
  switch (n_radii) {
    case 1: V3D_BQC(1); break;
    case 2: V3D_BQC(2); break;
    case 3: V3D_BQC(3); break;
    default: V3D_BQC(4); break;
  }
#undef V3D_BQC
  return check_launch();
}

// Bucketed copy of a source for v3d_ball_query_msg_select: `sorted` = float4 rows {x, y, z, original row index within
// the frame} (same frame ranges as the source), `workspace` = 2 * B * 1024 ints (zero-initialised once by the caller).
extern "C" size_t v3d_ball_query_sort_workspace_bytes(int B) { return B > 0 ? (size_t)2 * B * kBsMaxBuckets * sizeof(int) : 0; }

extern "C" int v3d_ball_query_sort_x(const float* xyz, int point_stride, const int* row_offsets, int B, int N,
                                     int max_rows_per_frame, float x_min, float x_max, void* sorted, void* workspace,
                                     v3d_stream_t stream) {
  if (!xyz || !sorted || !workspace || B <= 0 || B > 65535 || point_stride < 3 || max_rows_per_frame <= 0 || !(x_max > x_min))
    return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && (N <= 0 || N > max_rows_per_frame)) return V3D_ERR_INVALID_ARGUMENT;
  int* cnt = static_cast<int*>(workspace);
  int* cursor = cnt + (size_t)B * kBsMaxBuckets;
  const float inv_w = (float)kBsMaxBuckets / (x_max - x_min);
  cudaStream_t st = as_stream(stream);
  const int blocks = ceil_div(max_rows_per_frame, 256) < 64 ? ceil_div(max_rows_per_frame, 256) : 64;
  bs_hist_kernel<<<dim3(blocks, B), 256, 0, st>>>(xyz, point_stride, N, row_offsets, x_min, inv_w, kBsMaxBuckets, cnt);
  bs_scan_kernel<<<B, kBsMaxBuckets, 0, st>>>(cnt, cursor);
  bs_scatter_kernel<<<dim3(blocks, B), 256, 0, st>>>(xyz, point_stride, N, row_offsets, x_min, inv_w, kBsMaxBuckets, cursor,
                                                     static_cast<float4*>(sorted));
  return check_launch();
}

extern "C" int v3d_ball_query_msg_select(const void* sorted, const int* row_offsets, const void* bounds,
                                         int max_rows_per_frame, const float* new_xyz, int B, int N, int M, int n_radii,
                                         const float* radii_host, const int* nsamples_host, int* const* idx_host,
                                         v3d_stream_t stream) {
  if (!sorted || !new_xyz || !bounds || !radii_host || !nsamples_host || !idx_host) return V3D_ERR_INVALID_ARGUMENT;
  if (B <= 0 || B > 65535 || M <= 0 || n_radii <= 0 || n_radii > kBqMaxR || max_rows_per_frame <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && (N <= 0 || N > max_rows_per_frame)) return V3D_ERR_INVALID_ARGUMENT;
  BqArgs A;
  for (int r = 0; r < kBqMaxR; r++) {
    const int s = r < n_radii ? r : 0;
    if (nsamples_host[s] <= 0 || nsamples_host[s] > 32 || !idx_host[s]) return V3D_ERR_INVALID_ARGUMENT;
    A.r2[r] = radii_host[s] * radii_host[s];
    A.ns[r] = nsamples_host[s];
    A.out[r] = idx_host[s];
  }
  const int max_chunks = (max_rows_per_frame + 31) / 32;
  const float4* lo = static_cast<const float4*>(bounds);
  const float4* hi = lo + (size_t)B * max_chunks;
  dim3 grid(ceil_div(M, kBqWarps), B);
  cudaStream_t st = as_stream(stream);
#define V3D_BQS(RR)                                                                                              \
  ball_query_select_kernel<RR><<<grid, kBqWarps * 32, 0, st>>>(static_cast<const float4*>(sorted), N, row_offsets, lo, hi, \
                                                               max_chunks, new_xyz, M, A)
  switch (n_radii) {
    case 1: V3D_BQS(1); break;
    case 2: V3D_BQS(2); break;
    case 3: V3D_BQS(3); break;
    default: V3D_BQS(4); break;
  }
#undef V3D_BQS
  return check_launch();
}

extern "C" int v3d_bev_gather(const float* map_nhwc, int B, int H, int W, int C, const float* keypoints, int M,
                              float x_offset, float y_offset, float pixel_x, float pixel_y, float* out, int c_total,
                              int c_off, v3d_stream_t stream) {
  if (!map_nhwc || !keypoints || !out || B <= 0 || B > 65535 || H < 3 || W < 3 || C <= 0 || (C & 3) || M <= 0 ||
      c_off < 0 || c_off + C > c_total || !(pixel_x > 0.f) || !(pixel_y > 0.f))
    return V3D_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(map_nhwc) & 15) return V3D_ERR_INVALID_ARGUMENT;
  bev_gather_kernel<<<dim3(ceil_div(M, 32), B), 256, 0, as_stream(stream)>>>(map_nhwc, H, W, C, keypoints, M, x_offset,
                                                                             y_offset, pixel_x, pixel_y, out, c_total, c_off);
  return check_launch();
}
