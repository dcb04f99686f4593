// fused set abstraction for PointnetSAModuleMSG (detector/model.py:58-66, detector/roi_grid_pool.py:26-33,68):
//   grouping (a10)  ->  shared MLP (two 1x1 Conv2d + folded BatchNorm + ReLU)  ->  max over the nsample axis
// in ONE kernel on the 5th-generation tensor cores. The reference materialises the grouped tensor
// (B, 3 + C, M, nsample) -- 844 MB for the RoI-grid pool at config C3 -- runs two cuDNN convolutions over it and a
// max-pool; here a CTA owns 128 consecutive (query, sample) rows (= 8 queries x 16 samples or 4 x 32), gathers
// their source rows straight into the swizzled A tiles of a tcgen05 GEMM, keeps the layer-1 activations on chip
// (TMEM -> registers -> shared memory as the A operand of layer 2) and writes only the pooled (B, Cout, M) result.
//
// Number format: as in sparse_conv_tc.cu (bf16 3-term split, fp32 accumulation): a value x travels as
// h1 = bf16(x), h2 = bf16(x - h1); a product is h1*g1 + h1*g2 + h2*g1. Source features arrive as packed rows
// [h1(0..Cp-1) | h2(0..Cp-1)] (Cp = channels rounded up to 8), so the gather is a pure 16-byte copy.
//
// K layout of layer 1: [features 0..Cp-1 | dx dy dz | zero padding to a multiple of 64] -- the reference
// concatenates [xyz - centre ; features] (QueryAndGroup, use_xyz=True); the order of the input channels of a 1x1
// convolution is immaterial, so the weight rows are permuted accordingly on the host (ops.PreparedSaMlp) and the
// copies stay 16-byte aligned.
//
// One persistent CTA per SM, warp specialised like the sparse convolution:
//   warps 0-3  epilogue : (1) layer-1 accumulator (TMEM) + bias -> ReLU -> bf16 split -> A tiles of layer 2 in shared
//                         memory, 64 columns at a time; (2) layer-2 accumulator + bias -> ReLU -> max over the
//                         nsample rows of each query (warp shuffles: lanes are rows) -> out[b, c, q]
//   warp  4    MMA      : layer 1: per K=64 chunk 4 x {A_h1*G1, A_h1*G2, A_h2*G1} into acc1 (N = N1);
//                         layer 2: the same on the chunks of the activations into acc2 (N = N2)
//   warps 5-12 fetchers : per tile the (source row, dx, dy, dz) of its 128 rows (next tile's prefetched), per
//                         layer-1 chunk the gather (cp.async, 8 consecutive rows per lane) + the chunk's weight image
//                         by 1-D TMA; layer-2 weight chunks travel through the same ring.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace v3d {
namespace {

constexpr int kSaTileM = 128;
constexpr int kSaEpiWarps = 4, kSaWarpMma = 4, kSaWarpFetch0 = 5, kSaFetchWarps = 8;
constexpr int kSaATile = kSaTileM * 128;  // one half (h1 or h2) of a K = 64 chunk
constexpr int kSaABytes = 2 * kSaATile;

constexpr int sa_max(int a, int b) { return a > b ? a : b; }

template <int N1, int N2>
struct SaCfg {
  static constexpr int kW1Bytes = 2 * N1 * 128;  // [G1 rows | G2 rows] of one K = 64 chunk
  static constexpr int kW2Bytes = 2 * N2 * 128;
  static constexpr int kBBytes = sa_max(kW1Bytes, kW2Bytes);
  static constexpr int kStageBytes = kSaABytes + kBBytes;
  static constexpr int kStages = N1 > 64 ? 2 : 3;
  static constexpr int kAcc2Col = 256;
  static constexpr size_t kSmemBytes = 1024 /*align*/ + (size_t)kStages * kStageBytes + kSaABytes /*layer-2 A*/ +
                                       kSaTileM * 4 /*source rows*/ + kSaTileM * 16 /*xyz split*/ + 512 /*barriers*/ +
                                       (N1 + N2) * sizeof(float);
  static_assert(kStageBytes % 1024 == 0, "swizzle atoms");
  static_assert(N1 % 16 == 0 && N2 % 16 == 0 && N1 <= 256 && N2 <= 256, "MMA N");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct SaArgs {
  const unsigned char* featp;  // packed source rows, 4 * Cp bytes each
  int Cp;                      // padded feature channels (multiple of 8)
  const float* xyz;            // source coordinates, rows of xyz_stride floats
  int xyz_stride;
  const int* row_offsets;      // ragged sources: frame b owns rows [row_offsets[b], row_offsets[b+1]); or NULL
  int N;                       // dense sources: rows per frame
  const float* new_xyz;        // (B, M, 3)
  const int* idx;              // (B, M, ns)
  int B, M, ns;
  const unsigned char* w1;     // nc1 chunk images
  const unsigned char* w2;     // nc2 chunk images
  const float* b1;             // N1
  const float* b2;             // N2
  int nc1, nc2;
  float* out;                  // (B, c_total, M), written at channel c_off
  int c_total, c_off;
};

template <int N1, int N2>
__global__ void __launch_bounds__(32 * (kSaWarpFetch0 + kSaFetchWarps), 1) sa_fused_kernel(SaArgs P) {
  using C = SaCfg<N1, N2>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = base;
  unsigned char* a2 = ring + (size_t)C::kStages * C::kStageBytes;  // layer-2 A operand: [h1 tile | h2 tile]
  int* rowsrc = reinterpret_cast<int*>(a2 + kSaABytes);            // [128] global source row or -1
  uint4* xyzh = reinterpret_cast<uint4*>(rowsrc + kSaTileM);       // [128] {h1(dx,dy), h1(dz,0), h2(dx,dy), h2(dz,0)}
  uint64_t* full = reinterpret_cast<uint64_t*>(xyzh + kSaTileM);   // [4]
  uint64_t* empty = full + 4;                                      // [4]
  uint64_t* acc1_full = empty + 4;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* a2_full = acc1_empty + 1;
  uint64_t* a2_empty = a2_full + 1;
  uint64_t* acc2_full = a2_empty + 1;
  uint64_t* acc2_empty = acc2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);
  float* s_b1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full) + 512);
  float* s_b2 = s_b1 + N1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total_rows = (long long)P.B * P.M * P.ns;
  const int n_tiles = (int)((total_rows + kSaTileM - 1) / kSaTileM);

  if (tid == 0) {
    for (int s = 0; s < C::kStages; s++) {
      mbar_init(&full[s], kSaFetchWarps * 32 + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, kSaEpiWarps * 32);
    mbar_init(a2_full, kSaEpiWarps * 32);
    mbar_init(a2_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kSaEpiWarps * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = tid; c < N1; c += (int)blockDim.x) s_b1[c] = P.b1 ? __ldg(&P.b1[c]) : 0.f;
  for (int c = tid; c < N2; c += (int)blockDim.x) s_b2[c] = P.b2 ? __ldg(&P.b2[c]) : 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nc1 = P.nc1, nc2 = P.nc2;

  if (warp < kSaEpiWarps) {
    // =========================== epilogue ===========================
    const int r = warp * 32 + lane;  // row of the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a2_row = smem_u32(a2) + (uint32_t)(r * 128);
    uint32_t u2 = 0;  // uses of the layer-2 A buffer so far
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      // ---- (1) layer-1 activations -> A operand of layer 2, 64 columns per chunk
      mbar_wait(acc1_full, it & 1);
      tc_fence_after();
      for (int j = 0; j < nc2; j++, u2++) {
        mbar_wait(a2_empty, (u2 & 1u) ^ 1u);  // the MMAs that read the previous contents have completed
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int c0 = 64 * j + 16 * t;
          uint32_t v[16];
          if (c0 < N1) {
            tmem_ld16(lane_addr + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          uint32_t h1[8], h2[8];
#pragma unroll
          for (int e = 0; e < 8; e++) {
            float x0 = 0.f, x1 = 0.f;
            if (c0 < N1) {  // (N1 is a multiple of 16: a 16-column group is entirely inside or outside)
              x0 = fmaxf(__uint_as_float(v[2 * e]) + s_b1[c0 + 2 * e], 0.f);
              x1 = fmaxf(__uint_as_float(v[2 * e + 1]) + s_b1[c0 + 2 * e + 1], 0.f);
            }
            split2(x0, x1, h1[e], h2[e]);
          }
#pragma unroll
          for (int hu = 0; hu < 2; hu++) {  // two 16-byte units (8 K-elements each) per 16 columns
            const uint32_t uo = (uint32_t)(((2 * t + hu) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2_row + uo), "r"(h1[4 * hu]), "r"(h1[4 * hu + 1]),
                         "r"(h1[4 * hu + 2]), "r"(h1[4 * hu + 3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2_row + kSaATile + uo), "r"(h2[4 * hu]),
                         "r"(h2[4 * hu + 1]), "r"(h2[4 * hu + 2]), "r"(h2[4 * hu + 3])
                         : "memory");
          }
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(a2_full);
      }
      tc_fence_before();
      mbar_arrive(acc1_empty);  // layer-1 accumulator fully read: the next tile's layer 1 may start
      // ---- (2) layer-2 accumulator -> bias, ReLU, max over the nsample rows of each query
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      const long long R = (long long)tile * kSaTileM + r;
      const long long qg = R / P.ns;  // global query index b * M + q (same for the lanes of a sample group)
      const bool row_ok = R < total_rows;
#pragma unroll
      for (int c0 = 0; c0 < N2; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + (uint32_t)(C::kAcc2Col + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 16 >= N2) {
          tc_fence_before();
          mbar_arrive(acc2_empty);
        }
        float keep0 = 0.f, keep1 = 0.f;  // lane l keeps column c0 + (l & 15) of sample group 0 / 1 of the warp
#pragma unroll
        for (int e = 0; e < 16; e++) {
          float x = row_ok ? fmaxf(__uint_as_float(v[e]) + s_b2[c0 + e], 0.f) : 0.f;  // post-ReLU values are >= 0
          x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
          x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
          x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
          x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
          if (P.ns == 32) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
          const float g0 = __shfl_sync(0xffffffffu, x, 0), g1 = __shfl_sync(0xffffffffu, x, 16);
          if ((lane & 15) == e) {
            keep0 = g0;
            keep1 = g1;
          }
        }
        // lanes 0..15 write group 0's 16 columns, lanes 16..31 group 1's (ns = 16 only)
        const long long q0 = ((long long)tile * kSaTileM + warp * 32) / P.ns;  // first query of this warp
        const int grp = lane >> 4;
        const long long qq = q0 + (P.ns == 16 ? grp : 0);
        const bool writer = P.ns == 16 || grp == 0;
        if (writer && qq < (long long)P.B * P.M) {
          const int b = (int)(qq / P.M), q = (int)(qq % P.M);
          P.out[((size_t)b * P.c_total + P.c_off + c0 + (lane & 15)) * P.M + q] = grp ? keep1 : keep0;
        }
      }
      (void)qg;
    }
  } else if (warp == kSaWarpMma) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc1 = make_idesc(kSaTileM, N1), idesc2 = make_idesc(kSaTileM, N2);
    const uint32_t ring_u32 = smem_u32(ring);
    const uint64_t d2a1 = make_desc(smem_u32(a2)), d2a2 = make_desc(smem_u32(a2) + kSaATile);
    uint32_t q = 0, u2 = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      // ---- layer 1 into acc1
      mbar_wait(acc1_empty, (it & 1) ^ 1);
      for (int c = 0; c < nc1; c++, q++) {
        const uint32_t s = q % C::kStages;
        mbar_wait(&full[s], (q / C::kStages) & 1u);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t st = ring_u32 + s * (uint32_t)C::kStageBytes;
        const uint64_t da1 = make_desc(st), da2 = make_desc(st + kSaATile);
        const uint64_t dg1 = make_desc(st + kSaABytes), dg2 = make_desc(st + kSaABytes + N1 * 128);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            umma_bf16_ss(tmem_base, da1 + 2u * ks, dg1 + 2u * ks, idesc1, (c | ks) ? 1u : 0u);
            umma_bf16_ss(tmem_base, da1 + 2u * ks, dg2 + 2u * ks, idesc1, 1u);
            umma_bf16_ss(tmem_base, da2 + 2u * ks, dg1 + 2u * ks, idesc1, 1u);
          }
          umma_commit(&empty[s]);
          if (c == nc1 - 1) umma_commit(acc1_full);
        }
        __syncwarp();
      }
      // ---- layer 2 into acc2: A = the activations the epilogue converts chunk by chunk
      mbar_wait(acc2_empty, (it & 1) ^ 1);
      for (int j = 0; j < nc2; j++, q++, u2++) {
        const uint32_t s = q % C::kStages;
        mbar_wait(&full[s], (q / C::kStages) & 1u);  // weight chunk j
        mbar_wait(a2_full, u2 & 1u);                 // activation chunk j
        tc_fence_after();
        const uint32_t st = ring_u32 + s * (uint32_t)C::kStageBytes;
        const uint64_t dg1 = make_desc(st + kSaABytes), dg2 = make_desc(st + kSaABytes + N2 * 128);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            umma_bf16_ss(tmem_base + C::kAcc2Col, d2a1 + 2u * ks, dg1 + 2u * ks, idesc2, (j | ks) ? 1u : 0u);
            umma_bf16_ss(tmem_base + C::kAcc2Col, d2a1 + 2u * ks, dg2 + 2u * ks, idesc2, 1u);
            umma_bf16_ss(tmem_base + C::kAcc2Col, d2a2 + 2u * ks, dg1 + 2u * ks, idesc2, 1u);
          }
          umma_commit(&empty[s]);
          umma_commit(a2_empty);
          if (j == nc2 - 1) umma_commit(acc2_full);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== fetchers ===========================
    constexpr int NF = kSaFetchWarps * 32;
    const int fw = warp - kSaWarpFetch0, gt = fw * 32 + lane;
    const int rsub = lane >> 4, part = (lane >> 3) & 1, unit = lane & 7;
    const int row0 = 16 * fw + 8 * rsub;  // this lane copies rows row0 .. row0 + 7
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t dst_lane = ring_u32 + (uint32_t)(part * kSaATile + row0 * 128);
    const int Cp = P.Cp;
    const size_t row_bytes = (size_t)4 * Cp;
    const unsigned char* feat_lane = P.featp + (size_t)part * 2 * Cp;

    // per-row staging (threads 0..127: one tile row each); the NEXT tile's values are prefetched into registers
    int pre_src = -1;
    float pre_d[3] = {0.f, 0.f, 0.f};
    auto prefetch = [&](int tile) {
      pre_src = -1;
      pre_d[0] = pre_d[1] = pre_d[2] = 0.f;
      if (gt < kSaTileM && tile < n_tiles) {
        const long long R = (long long)tile * kSaTileM + gt;
        if (R < total_rows) {
          const long long qg = R / P.ns;
          const int b = (int)(qg / P.M);
          const int basei = P.row_offsets ? __ldg(&P.row_offsets[b]) : b * P.N;
          pre_src = basei + __ldg(&P.idx[R]);
          const float* sx = P.xyz + (size_t)pre_src * P.xyz_stride;
          const float* qx = P.new_xyz + (size_t)qg * 3;
#pragma unroll
          for (int d = 0; d < 3; d++) pre_d[d] = __fsub_rn(__ldg(&sx[d]), __ldg(&qx[d]));
        }
      }
    };
    prefetch(blockIdx.x);
    uint32_t q = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      asm volatile("bar.sync 1, %0;" ::"n"(NF) : "memory");  // copies of the previous tile that read rowsrc are issued
      if (gt < kSaTileM) {
        rowsrc[gt] = pre_src;
        uint32_t a1, a2w, b1w, b2w;
        split2(pre_d[0], pre_d[1], a1, b1w);
        split2(pre_d[2], 0.f, a2w, b2w);
        xyzh[gt] = make_uint4(a1, a2w, b1w, b2w);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NF) : "memory");
      prefetch(tile + gridDim.x);
      const int4 sa = *reinterpret_cast<const int4*>(rowsrc + row0), sb = *reinterpret_cast<const int4*>(rowsrc + row0 + 4);
      const int src[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
      for (int c = 0; c < nc1; c++, q++) {
        const uint32_t s = q % C::kStages;
        mbar_wait(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
        const uint32_t st = s * (uint32_t)C::kStageBytes;
        if (gt == 0) {
          mbar_arrive_expect_tx(&full[s], (uint32_t)C::kW1Bytes);
          bulk_g2s(ring_u32 + st + kSaABytes, P.w1 + (size_t)c * C::kW1Bytes, (uint32_t)C::kW1Bytes, &full[s]);
        }
        const int ch = 64 * c + 8 * unit;  // first K element (= channel) of this lane's 16-byte unit
        if (ch == Cp) {  // the unit that carries (dx, dy, dz): computed values, plain shared stores
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint4 xv = xyzh[row0 + i];
            const uint32_t w0 = part ? xv.z : xv.x, w1 = part ? xv.w : xv.y;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_lane + st + (uint32_t)(i * 128 + ((unit ^ i) << 4))),
                         "r"(w0), "r"(w1), "r"(0u), "r"(0u)
                         : "memory");
          }
          __threadfence_block();  // ordered before this thread's arrival below
        } else {
          const bool in_feat = ch + 8 <= Cp;  // else: zero padding beyond the last channel
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const bool ok = in_feat && src[i] >= 0;
            const unsigned char* p = feat_lane + (size_t)(ok ? src[i] : 0) * row_bytes + (size_t)(in_feat ? ch : 0) * 2;
            cp_async16(dst_lane + st + (uint32_t)(i * 128 + ((unit ^ i) << 4)), p, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(&full[s]);
      }
      for (int j = 0; j < nc2; j++, q++) {  // layer-2 weight chunks ride the same ring (no gather)
        const uint32_t s = q % C::kStages;
        mbar_wait(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
        if (gt == 0) {
          mbar_arrive_expect_tx(&full[s], (uint32_t)C::kW2Bytes);
          bulk_g2s(ring_u32 + s * (uint32_t)C::kStageBytes + kSaABytes, P.w2 + (size_t)j * C::kW2Bytes,
                   (uint32_t)C::kW2Bytes, &full[s]);
        }
        cp_async_arrive_noinc(&full[s]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// (G, 64, N) fp32 chunked weights -> per chunk the shared-memory image [G1 rows | G2 rows] x 128 B, K-major,
// 128B-swizzled, bf16 split (the B operand layout of sparse_conv_tc.cu with Cin = 64)
__global__ void sa_prepare_kernel(const float* __restrict__ w, int G, int N, unsigned char* __restrict__ img) {
  const size_t per_chunk = (size_t)2 * N * 128;
  const int total = G * N * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kl = e % 64;
    int t = e / 64;
    const int n = t % N, g = t / N;
    const float v = w[((size_t)g * 64 + kl) * N + n];
    const __nv_bfloat16 g1 = __float2bfloat16_rn(v);
    const __nv_bfloat16 g2 = __float2bfloat16_rn(v - __bfloat162float(g1));
    const size_t off = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)((((kl >> 3) ^ (n & 7)) << 4) + (kl & 7) * 2);
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)g * per_chunk + off) = g1;
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)g * per_chunk + (size_t)N * 128 + off) = g2;
  }
}

// channel-major fp32 (B, C, N) -> packed rows (B * N, 2 * Cp) bf16 [h1 | h2], channels C..Cp-1 zero
// (keypoint features (B, 512, 2048) -> the gather source of the RoI-grid pool; point intensity -> 8-channel rows)
__global__ void __launch_bounds__(256) sa_pack_cmajor_kernel(const float* __restrict__ f, long long f_bstride,
                                                             long long f_cstride, long long f_nstride, int B, int C, int N,
                                                             int Cp, unsigned char* __restrict__ packed) {
  const long long total = (long long)B * N * (Cp / 2);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    // consecutive threads -> consecutive n (coalesced reads along N), channel pair cp
    const int n = (int)(e % N);
    long long t = e / N;
    const int cp = (int)(t % (Cp / 2));
    const int b = (int)(t / (Cp / 2));
    const int c = 2 * cp;
    const float x0 = c < C ? f[b * f_bstride + c * f_cstride + n * f_nstride] : 0.f;
    const float x1 = c + 1 < C ? f[b * f_bstride + (c + 1) * f_cstride + n * f_nstride] : 0.f;
    uint32_t h1, h2;
    split2(x0, x1, h1, h2);
    unsigned char* row = packed + ((size_t)b * N + n) * (size_t)(4 * Cp);
    *reinterpret_cast<uint32_t*>(row + 4 * cp) = h1;
    *reinterpret_cast<uint32_t*>(row + 2 * Cp + 4 * cp) = h2;
  }
}

template <int N1, int N2>
int launch_sa(const SaArgs& P, cudaStream_t st) {
  using C = SaCfg<N1, N2>;
  static PerDeviceOnce attr_once;
  if (attr_once.needed()) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(sa_fused_kernel<N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    attr_once.done();
  }
  const long long total_rows = (long long)P.B * P.M * P.ns;
  const long long tiles = (total_rows + kSaTileM - 1) / kSaTileM;
  const int grid = (int)(tiles < kNumSMs ? (tiles > 0 ? tiles : 1) : kNumSMs);
  sa_fused_kernel<N1, N2><<<grid, 32 * (kSaWarpFetch0 + kSaFetchWarps), C::kSmemBytes, st>>>(P);
  return check_launch();
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" size_t v3d_sa_mlp_prepared_bytes(int n_chunks, int N) {
  if (n_chunks <= 0 || N <= 0 || (N & 15)) return 0;
  return (size_t)n_chunks * 2 * N * 128;
}

extern "C" int v3d_sa_mlp_prepare(const float* weight_chunks, int n_chunks, int N, void* prepared, size_t prepared_bytes,
                                  v3d_stream_t stream) {
  if (!weight_chunks || !prepared || n_chunks <= 0 || N <= 0 || (N & 15)) return V3D_ERR_INVALID_ARGUMENT;
  if (prepared_bytes < v3d_sa_mlp_prepared_bytes(n_chunks, N)) return V3D_ERR_WORKSPACE_TOO_SMALL;
  const int total = n_chunks * N * 64;
  sa_prepare_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(weight_chunks, n_chunks, N,
                                                                        static_cast<unsigned char*>(prepared));
  return check_launch();
}

extern "C" int v3d_pack_channel_major(const float* feat, long long b_stride, long long c_stride, long long n_stride, int B,
                                      int C, int N, int Cp, void* packed, v3d_stream_t stream) {
  if (!feat || !packed || B <= 0 || C <= 0 || N <= 0 || Cp < C || (Cp & 7)) return V3D_ERR_INVALID_ARGUMENT;
  const long long total = (long long)B * N * (Cp / 2);
  const long long want = (total + 255) / 256;
  const int blocks = (int)(want < kNumSMs * 8 ? want : kNumSMs * 8);
  sa_pack_cmajor_kernel<<<blocks, 256, 0, as_stream(stream)>>>(feat, b_stride, c_stride, n_stride, B, C, N, Cp,
                                                              static_cast<unsigned char*>(packed));
  return check_launch();
}

extern "C" int v3d_sa_fused(const void* feat_packed, int Cp, const float* xyz, int xyz_stride, const int* row_offsets, int N,
                            const float* new_xyz, const int* idx, int B, int M, int nsample, const void* w1_prepared,
                            const float* b1, int N1, const void* w2_prepared, const float* b2, int N2, float* out,
                            int c_total, int c_off, v3d_stream_t stream) {
  if (!feat_packed || !xyz || !new_xyz || !idx || !w1_prepared || !w2_prepared || !out) return V3D_ERR_INVALID_ARGUMENT;
  if (Cp <= 0 || (Cp & 7) || xyz_stride < 3 || B <= 0 || M <= 0 || (nsample != 16 && nsample != 32)) return V3D_ERR_INVALID_ARGUMENT;
  if (!row_offsets && N <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (c_off < 0 || c_off + N2 > c_total) return V3D_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(feat_packed) & 15) || (reinterpret_cast<uintptr_t>(w1_prepared) & 15) ||
      (reinterpret_cast<uintptr_t>(w2_prepared) & 15))
    return V3D_ERR_INVALID_ARGUMENT;
  SaArgs P;
  P.featp = static_cast<const unsigned char*>(feat_packed);
  P.Cp = Cp;
  P.xyz = xyz;
  P.xyz_stride = xyz_stride;
  P.row_offsets = row_offsets;
  P.N = N;
  P.new_xyz = new_xyz;
  P.idx = idx;
  P.B = B;
  P.M = M;
  P.ns = nsample;
  P.w1 = static_cast<const unsigned char*>(w1_prepared);
  P.w2 = static_cast<const unsigned char*>(w2_prepared);
  P.b1 = b1;
  P.b2 = b2;
  P.nc1 = (Cp + 3 + 63) / 64;
  P.nc2 = (N1 + 63) / 64;
  P.out = out;
  P.c_total = c_total;
  P.c_off = c_off;
  cudaStream_t st = as_stream(stream);
#define V3D_SA_CASE(A, Bq) \
  if (N1 == A && N2 == Bq) return launch_sa<A, Bq>(P, st);
  V3D_SA_CASE(16, 16)
  V3D_SA_CASE(32, 32)
  V3D_SA_CASE(64, 64)
  V3D_SA_CASE(192, 96)
#undef V3D_SA_CASE
  return V3D_ERR_INVALID_ARGUMENT;
}
