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
constexpr int kFpsThreads = 256;

struct Cand {
  float d;
  int i;
  float x, y, z;
  float pad[3];
};

__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// better = larger distance, ties -> smaller index
__device__ __forceinline__ bool better(float d1, int i1, float d2, int i2) {
  return d1 > d2 || (d1 == d2 && i1 < i2);
}

template <int PPT>
__global__ void __cluster_dims__(kFpsCluster, 1, 1) __launch_bounds__(kFpsThreads)
    fps_cluster_kernel(const float* __restrict__ xyz, int N, int m, int* __restrict__ out) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kFpsCluster;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = xyz + (size_t)b * N * 3;

  __shared__ Cand cand[2][kFpsCluster];  // written by every CTA of the cluster through DSMEM
  __shared__ float wd[kFpsThreads / 32];
  __shared__ int wi[kFpsThreads / 32];

  // CTA `rank` owns the contiguous slice [s0, s0 + S); thread t owns s0 + t + k*blockDim, k < PPT
  const int S = PPT * kFpsThreads;
  const int s0 = rank * S;
  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int k = 0; k < PPT; k++) {
    const int i = s0 + tid + k * kFpsThreads;
    if (i < N) {
      px[k] = P[3 * i];
      py[k] = P[3 * i + 1];
      pz[k] = P[3 * i + 2];
      md[k] = 1e10f;
    } else {
      px[k] = py[k] = pz[k] = 0.f;
      md[k] = -1.f;  // never wins (distances are >= 0)
    }
  }
  float cx = P[0], cy = P[1], cz = P[2];
  if (rank == 0 && tid == 0) out[(size_t)b * m] = 0;
  cluster.sync();

  for (int j = 1; j < m; j++) {
    float bd = -1.f;
    int bi = 0x7fffffff;
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
      const float d = dist2(px[k], py[k], pz[k], cx, cy, cz);
      const float d2 = md[k] < 0.f ? -1.f : fminf(d, md[k]);
      md[k] = d2;
      if (d2 > bd) {  // ascending index inside the thread: strict > keeps the lowest index
        bd = d2;
        bi = s0 + tid + k * kFpsThreads;
        bx = px[k];
        by = py[k];
        bz = pz[k];
      }
    }
    // warp arg-max
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const float ox = __shfl_xor_sync(0xffffffffu, bx, o);
      const float oy = __shfl_xor_sync(0xffffffffu, by, o);
      const float oz = __shfl_xor_sync(0xffffffffu, bz, o);
      if (better(od, oi, bd, bi)) {
        bd = od;
        bi = oi;
        bx = ox;
        by = oy;
        bz = oz;
      }
    }
    __shared__ float wx[kFpsThreads / 32], wy[kFpsThreads / 32], wz[kFpsThreads / 32];
    if (lane == 0) {
      wd[warp] = bd;
      wi[warp] = bi;
      wx[warp] = bx;
      wy[warp] = by;
      wz[warp] = bz;
    }
    __syncthreads();
    const int par = j & 1;
    if (warp == 0) {
      float d = lane < kFpsThreads / 32 ? wd[lane] : -2.f;
      int i = lane < kFpsThreads / 32 ? wi[lane] : 0x7fffffff;
      float x = lane < kFpsThreads / 32 ? wx[lane] : 0.f;
      float y = lane < kFpsThreads / 32 ? wy[lane] : 0.f;
      float z = lane < kFpsThreads / 32 ? wz[lane] : 0.f;
#pragma unroll
      for (int o = 4; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        const float ox = __shfl_xor_sync(0xffffffffu, x, o);
        const float oy = __shfl_xor_sync(0xffffffffu, y, o);
        const float oz = __shfl_xor_sync(0xffffffffu, z, o);
        if (better(od, oi, d, i)) {
          d = od;
          i = oi;
          x = ox;
          y = oy;
          z = oz;
        }
      }
      // lanes 0..7 each deliver this CTA's candidate to one CTA of the cluster
      d = __shfl_sync(0xffffffffu, d, 0);
      i = __shfl_sync(0xffffffffu, i, 0);
      x = __shfl_sync(0xffffffffu, x, 0);
      y = __shfl_sync(0xffffffffu, y, 0);
      z = __shfl_sync(0xffffffffu, z, 0);
      if (lane < kFpsCluster) {
        Cand* remote = cluster.map_shared_rank(&cand[par][rank], lane);
        remote->d = d;
        remote->i = i;
        remote->x = x;
        remote->y = y;
        remote->z = z;
      }
    }
    cluster.sync();  // candidates of step j visible everywhere; parity buffers make one sync enough
    float gd = cand[par][0].d;
    int gi = cand[par][0].i;
    int gr = 0;
#pragma unroll
    for (int r = 1; r < kFpsCluster; r++) {
      const float d = cand[par][r].d;
      const int i = cand[par][r].i;
      if (better(d, i, gd, gi)) {
        gd = d;
        gi = i;
        gr = r;
      }
    }
    cx = cand[par][gr].x;
    cy = cand[par][gr].y;
    cz = cand[par][gr].z;
    if (rank == 0 && tid == 0) out[(size_t)b * m + j] = gi;
  }
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

template <int PPT>
int launch_fps(const float* xyz, int B, int N, int m, int* idx, cudaStream_t st) {
  fps_cluster_kernel<PPT><<<B * kFpsCluster, kFpsThreads, 0, st>>>(xyz, N, m, idx);
  return check_launch();
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" size_t v3d_fps_workspace_bytes(int B, int N) {
  (void)B;
  (void)N;
  return 256;  // everything lives in registers / distributed shared memory
}

extern "C" int v3d_fps(const float* xyz, int B, int N, int m, int* idx, void* workspace, size_t workspace_bytes,
                       v3d_stream_t stream) {
  (void)workspace;
  (void)workspace_bytes;
  if (!xyz || !idx || B <= 0 || N <= 0 || m <= 0) return V3D_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  const int per_cta = ceil_div(N, kFpsCluster);
  const int ppt = ceil_div(per_cta, kFpsThreads);
  if (ppt <= 1) return launch_fps<1>(xyz, B, N, m, idx, st);
  if (ppt <= 2) return launch_fps<2>(xyz, B, N, m, idx, st);
  if (ppt <= 4) return launch_fps<4>(xyz, B, N, m, idx, st);
  if (ppt <= 8) return launch_fps<8>(xyz, B, N, m, idx, st);
  if (ppt <= 16) return launch_fps<16>(xyz, B, N, m, idx, st);
  if (ppt <= 32) return launch_fps<32>(xyz, B, N, m, idx, st);
  return V3D_ERR_INVALID_ARGUMENT;  // > 65536 points per cloud
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
