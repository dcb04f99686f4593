// a5 + a6 on the 5th-generation tensor cores: output-stationary sparse convolution with tcgen05.mma
// (kind::f16, bf16 operands, 3-term split "bf16x3"), fp32 accumulators in TMEM, folded BN + ReLU epilogue.
//
// Number format. A value x is carried as two bf16 numbers h1 = bf16_rn(x), h2 = bf16_rn(x - h1)
// (|x - h1 - h2| <= 2^-18 |x|). A product x*w is evaluated as h1*g1 + h1*g2 + h2*g1 (every partial product
// is exact in the fp32 accumulator); the dropped terms bound the error of a product by ~3 * 2^-18 |x w|.
// Measured against fp64 on post-ReLU data: 4e-6 relative (Frobenius) per layer, i.e. well inside the 1e-4
// the contract allows, at HALF the tensor-pipe and operand-fetch cost of a 3xTF32 split (K = 16 per
// instruction instead of 8). The exact-fp32 SIMT kernel (sparse_conv.cu) stays available.
//
// "Packed" feature rows: [h1(0..C-1) | h2(0..C-1)] as bf16 = 4*C bytes, the same footprint as fp32. Every
// producer (this kernel's epilogue, v3d_feature_pack) writes that format, so the gather is a pure copy:
// 16-byte cp.async straight into the 128B-swizzled K-major A tiles the MMAs read from shared memory.
// No conversion pass, no register staging, no TMEM operand.
//
// One persistent CTA per SM, warp specialised (416 threads; 448 with the weight warp of schemes 2 / 4):
//   warps 0-3  epilogue   : tcgen05.ld the 128 x COUT fp32 accumulator halves (lane = output row), add
//                           them, apply scale/shift/ReLU, store the row as fp32 and/or packed bf16x2
//   warp  4    MMA issuer : warp-uniform loop, one elected lane issues per pipeline slot (K = 64) 2*4
//                           tcgen05.mma, M=128, K=16, both operands from shared memory:
//                           A_h1 * [G1 | G2] (N = 2*COUT) and A_h2 * G1 (N = COUT); tcgen05.commit frees
//                           the stage and publishes the accumulator through mbarriers
//   warps 5-12 fetchers   : stage the tile's rule rows in shared memory (the next tile's are prefetched
//                           into registers), find the kernel offsets the tile uses, and for every slot
//                           gather the neighbour rows with cp.async into the stage's A tiles. Completion is
//                           asynchronous (cp.async.mbarrier.arrive on the stage's full barrier), so up to kStages
//                           slots of global-load latency are in flight and the fetchers never wait for their own loads.
//   warp  13   weights    : (schemes 2 / 4) per slot the meta record, the expect_tx arrival and the 1-D TMA bulk copy
//                           of the slot's weight image; in scheme 1 thread 0 of the fetchers does this.
// What bounds it (ncu): the SHARED-MEMORY CROSSBAR. With SS operands a Cin = Cout = 64 slot moves
// 104 KB through shared memory (A written 32 KB + read 32 KB, weight image written 16 KB + read 24 KB) = 832
// wavefronts at 128 B/clk, against 444 clk of MMAs (one M=128 MMA costs max(47, N/2) clk, scripts/mma_probe.cu);
// measured 0.92 wavefronts per clk with every tile row copied or zero-filled. The shipped scheme 4 therefore does
// not write absent neighbours at all (57 % of the rows): the first slot of a tile zero-fills and initialises the
// accumulator, later slots copy present rows only and run lane-masked MMAs (tcgen05 disable-output-lane).
// Output rows are written exactly once (no atomics, deterministic).
//
// Operand layout (K-major, SWIZZLE_128B, bf16): a tile row is 64 K-elements = 128 bytes; 8-row groups are
// 1024 B apart (SBO); the 16-byte unit u of row r lives at unit u ^ (r & 7). A slot is always K = 64:
// 64/CIN consecutive kernel offsets are stacked along K. Stage = [A_h1 16 KB | A_h2 16 KB | G1 rows | G2 rows].
// TMEM columns: two accumulator buffers of 2*COUT columns: [h1*g1 + h2*g1 | h1*g2].
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace v3d {
namespace {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 4;
constexpr int kWarpMma = kEpiWarps, kWarpFetch0 = kEpiWarps + 1;
constexpr int kMaxKV = 27;
constexpr int kATileBytes = kTileM * 128;  // 128 rows x 64 bf16
constexpr int kABytes = 2 * kATileBytes;   // h1 tile | h2 tile


// per-offset usage mask -> per-group mask (GK consecutive offsets per group)
template <int GK>
__device__ __forceinline__ uint32_t group_mask(uint32_t m) {
  if (GK == 1) return m;
  uint32_t g = 0;
#pragma unroll
  for (int i = 0; i < 32 / GK; i++)
    if (m & (((1u << GK) - 1u) << (i * GK))) g |= 1u << i;
  return g;
}

// One 16-byte piece of a gathered row: dst + kDstOff <- base[idx * kRowBytes], zeros if idx < 0. Written as one PTX
// block so that the row costs exactly three SASS instructions (ISETP, IMAD.WIDE, LDGSTS with the zero-fill
// predicate); a predicated-off copy does not read its source address.
template <int kRowBytes, int kDstOff, bool kBypassL1>
__device__ __forceinline__ void gather_row16(uint32_t dst, const unsigned char* base, int idx) {
  if (kBypassL1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
        "setp.lt.s32 p, %2, 0;\n\t"
        "mad.wide.s32 a, %2, %3, %1;\n\t"
        "cp.async.cg.shared.global [%0+%4], [a], 16, p;\n\t}" ::"r"(dst),
        "l"(base), "r"(idx), "n"(kRowBytes), "n"(kDstOff)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
        "setp.lt.s32 p, %2, 0;\n\t"
        "mad.wide.s32 a, %2, %3, %1;\n\t"
        "cp.async.ca.shared.global [%0+%4], [a], 16, p;\n\t}" ::"r"(dst),
        "l"(base), "r"(idx), "n"(kRowBytes), "n"(kDstOff)
        : "memory");
}

// The same piece when an absent row may keep its stale shared-memory contents (its output lane is masked off in the
// MMAs): the copy is GUARDED, not zero-filled -- no shared-memory write at all for an absent neighbour.
template <int kRowBytes, int kDstOff, bool kBypassL1>
__device__ __forceinline__ void gather_row16_skip(uint32_t dst, const unsigned char* base, int idx) {
  if (kBypassL1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
        "setp.ge.s32 p, %2, 0;\n\t"
        "mad.wide.s32 a, %2, %3, %1;\n\t"
        "@p cp.async.cg.shared.global [%0+%4], [a], 16;\n\t}" ::"r"(dst),
        "l"(base), "r"(idx), "n"(kRowBytes), "n"(kDstOff)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
        "setp.ge.s32 p, %2, 0;\n\t"
        "mad.wide.s32 a, %2, %3, %1;\n\t"
        "@p cp.async.ca.shared.global [%0+%4], [a], 16;\n\t}" ::"r"(dst),
        "l"(base), "r"(idx), "n"(kRowBytes), "n"(kDstOff)
        : "memory");
}

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512))); }

template <int CIN, int COUT>
struct TcCfg {
  // Every slot is a K = 64 GEMM step: kGK = 64/CIN consecutive kernel offsets are stacked along K (their
  // gathered rows side by side in the A tile, their weights stacked in the B image), so small-channel layers
  // pay the per-slot pipeline handshakes once per 64 K-elements instead of once per 16 or 32.
  static constexpr int kGK = 64 / CIN;
  static constexpr int kKSteps = 4;                       // K = 16 per instruction
  // B image of one slot: rows [0, COUT) hold G1^T, rows [COUT, 2*COUT) hold G2^T, 128 B per row, so that
  // ONE N = 2*COUT MMA computes A_h1*[G1 | G2] (the gathered operand is fetched once for both products)
  static constexpr int kBBytes = 2 * COUT * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;   // multiple of 1024
  static constexpr int kStages = COUT == 64 ? 4 : 5;
  static constexpr int kAccBufCols = 2 * COUT;            // [h1*g1 + h2*g1 | h1*g2]
  static constexpr int kTmemCols = pow2_cols(2 * kAccBufCols);
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kStages * kStageBytes +
                                       sizeof(int) * kMaxKV * kTileM + 1024 /*barriers + meta*/ +
                                       2 * COUT * sizeof(float);
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte swizzle-atom alignment");
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB per CTA)");
};

struct alignas(16) SlotMeta {
  int last;  // 1 = last slot of its output tile
  int end;   // 1 = all tiles done
  int pad[2];
  uint4 off;  // scheme 4: disable-output-lane words of the slot (bit set = the row has no neighbour at this offset)
};

// kVer = scheme + 8 * cg + 16 * spin. scheme: 1 = "lean" (round 2); 2 = "phase-aligned" (4 instructions per copied
// row instead of 9, weight stream on its own warp); 4 = 2 + absent neighbours are not copied at all: every slot but
// the first of a tile runs lane-masked MMAs (tcgen05 disable-output-lane), so an absent row costs no shared-memory
// write (CIN = 64 only); 5 = the same on top of scheme 1's row mapping. cg: copies bypass L1 (cp.async.cg). spin: mbarrier waits poll with test_wait instead of the
// suspending try_wait. All variants produce bit-identical results; V3D_TC_FETCH / V3D_TC_CG / V3D_TC_WAIT pick one.
constexpr bool tc_weight_warp(int kVer) { return (kVer & 7) == 2 || (kVer & 7) == 4; }
constexpr int tc_threads(int kVer) { return 32 * (kWarpFetch0 + 8 + (tc_weight_warp(kVer) ? 1 : 0)); }

template <int CIN, int COUT, int kFetchWarps, int kVer>
__global__ void __launch_bounds__(tc_threads(kVer), 1)
sparse_conv_tc_kernel(const unsigned char* __restrict__ feat, const unsigned char* __restrict__ wprep,
                      const int* __restrict__ nbr, int nbr_stride, const int* __restrict__ n_out_ptr, int out_cap,
                      int KV, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                      float* __restrict__ out, unsigned char* __restrict__ out_packed) {
  using C = TcCfg<CIN, COUT>;
  constexpr int kScheme = kVer & 7;
  constexpr bool kCg = ((kVer >> 3) & 1) != 0, kSpin = ((kVer >> 4) & 1) != 0;
  static_assert(kScheme == 1 || kScheme == 2 || ((kScheme == 4 || kScheme == 5) && CIN == 64), "fetch scheme");
  constexpr bool kSkip = kScheme == 4 || kScheme == 5;  // absent rows are not copied, lane-masked MMAs
  extern __shared__ unsigned char smem_raw[];
  // round up to 1024 B (SWIZZLE_128B atoms) by OFFSETTING the __shared__ array: casting through an integer
  // would make every later access a generic LD/ST instead of LDS/STS
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = base;                                                          // kStages x stage
  int* idx_tile = reinterpret_cast<int*>(ring + (size_t)C::kStages * C::kStageBytes);  // [KV][128]
  unsigned char* tail = reinterpret_cast<unsigned char*>(idx_tile + kMaxKV * kTileM);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);  // [8]
  uint64_t* empty = full + 8;                          // [8]
  uint64_t* acc_full = empty + 8;                      // [2]
  uint64_t* acc_empty = acc_full + 2;                  // [2]
  SlotMeta* meta = reinterpret_cast<SlotMeta*>(acc_empty + 2);  // [8] fetcher -> MMA issuer (per stage)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(meta + 8);
  uint32_t* fmask = tmem_slot + 1;  // [2] offsets used by the tile being staged (parity double buffer)
  int* neg_row = reinterpret_cast<int*>(tail + 512);  // [128] all -1 (scheme 2: rule row of a non-existent offset)
  uint32_t* pres = reinterpret_cast<uint32_t*>(tail + 512);  // scheme 4 (never needs neg_row): [KV][4] presence words
  float* s_scale = reinterpret_cast<float*>(tail + 1024);
  float* s_shift = s_scale + COUT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_out = min(*n_out_ptr, out_cap);
  const int n_tiles = (n_out + kTileM - 1) / kTileM;

  if (tid == 0) {
    for (int s = 0; s < C::kStages; s++) {
      mbar_init(&full[s], kFetchWarps * 32 + 1);  // every fetch thread's copies + the expect_tx arrival (B image)
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kEpiWarps * 32);
    }
    fmask[0] = fmask[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kTileM) neg_row[tid] = -1;
  for (int c = tid; c < COUT; c += (int)blockDim.x) {
    s_scale[c] = scale ? __ldg(&scale[c]) : 1.0f;
    s_shift[c] = shift ? __ldg(&shift[c]) : 0.0f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // =========================== epilogue ===========================
    const float* sc = s_scale;
    const float* sh = s_shift;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int a = it & 1;
      mbar_wait_as<kSpin>(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      const int row = tile * kTileM + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * C::kAccBufCols);
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t v[16], v2[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld16(taddr + (uint32_t)(COUT + c0), v2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 16 >= COUT) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[a]);
        }
        float o[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {  // (h1*g1 + h2*g1) + h1*g2, then folded BN + ReLU
          const float acc = __uint_as_float(v[e]) + __uint_as_float(v2[e]);
          o[e] = fmaf(acc, sc[c0 + e], sh[c0 + e]);
          if (relu) o[e] = fmaxf(o[e], 0.f);
        }
        if (row < n_out) {
          if (out != nullptr) {
            float4* orow = reinterpret_cast<float4*>(out + (size_t)row * COUT + c0);
#pragma unroll
            for (int q = 0; q < 4; q++) orow[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
          if (out_packed != nullptr) {
            uint32_t h1[8], h2[8];
#pragma unroll
            for (int e = 0; e < 8; e++) split2(o[2 * e], o[2 * e + 1], h1[e], h2[e]);
            unsigned char* prow = out_packed + (size_t)row * (4 * COUT) + (size_t)c0 * 2;
            uint4* p1 = reinterpret_cast<uint4*>(prow);
            uint4* p2 = reinterpret_cast<uint4*>(prow + 2 * COUT);
            p1[0] = make_uint4(h1[0], h1[1], h1[2], h1[3]);
            p1[1] = make_uint4(h1[4], h1[5], h1[6], h1[7]);
            p2[0] = make_uint4(h2[0], h2[1], h2[2], h2[3]);
            p2[1] = make_uint4(h2[4], h2[5], h2[6], h2[7]);
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer ===========================
    // The WHOLE warp walks the slot sequence with warp-uniform control flow and only the tcgen05
    // instructions are predicated on one elected lane: operands then live in uniform registers.
    constexpr uint32_t idesc_wide = make_idesc(kTileM, 2 * COUT), idesc_g1 = make_idesc(kTileM, COUT);
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t q = 0;
    int it = 0;
    bool done = false;
    while (!done) {
      const int a = it & 1;
      uint32_t accum = 0u;
      bool tile_open = false;
      while (true) {
        const uint32_t s = q % C::kStages;
        mbar_wait_as<kSpin>(&full[s], (q / C::kStages) & 1u);
        const int m_last = meta[s].last, m_end = meta[s].end;
        if (m_end) {
          done = true;
          break;
        }
        if (!tile_open) {  // first slot of a tile: the accumulator buffer must have been drained
          mbar_wait_as<kSpin>(&acc_empty[a], ((it >> 1) & 1) ^ 1);
          tile_open = true;
        }
        fence_proxy_async();  // the A tiles were written by cp.async (generic proxy), the MMAs read them through
        tc_fence_after();     // the async proxy
        const uint32_t st = ring_u32 + s * (uint32_t)C::kStageBytes;
        const uint64_t da1 = make_desc(st), da2 = make_desc(st + kATileBytes), db0 = make_desc(st + kABytes);
        const uint32_t d = tmem_base + (uint32_t)(a * C::kAccBufCols);
        if (kSkip && accum != 0u) {
          // not the first slot of the tile: rows without a neighbour at this offset were NOT copied (their A rows
          // hold stale data) and their output lanes are switched off
          const uint4 off = meta[s].off;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < C::kKSteps; ks++) {
              umma_bf16_ss_masked(d, da1 + 2u * ks, db0 + 2u * ks, idesc_wide, off.x, off.y, off.z, off.w);
              umma_bf16_ss_masked(d, da2 + 2u * ks, db0 + 2u * ks, idesc_g1, off.x, off.y, off.z, off.w);
            }
            umma_commit(&empty[s]);
            if (m_last) umma_commit(&acc_full[a]);
          }
        } else if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < C::kKSteps; ks++) {
            // K = 16 bf16 = 32 bytes along the swizzled row: the start address advances by 2 (16-byte units)
            umma_bf16_ss(d, da1 + 2u * ks, db0 + 2u * ks, idesc_wide, accum);  // A_h1 * [G1 | G2], N = 2*COUT
            umma_bf16_ss(d, da2 + 2u * ks, db0 + 2u * ks, idesc_g1, 1u);       // A_h2 * G1 into the first COUT columns
            accum = 1u;
          }
          umma_commit(&empty[s]);  // stage reusable once these MMAs have read it
          if (m_last) umma_commit(&acc_full[a]);
        }
        accum = 1u;
        __syncwarp();
        q++;
        if (m_last) break;
      }
      it++;
    }
  } else if (kScheme == 1 || kScheme == 5) {
    // =========================== fetchers (scheme 1) ===========================
    // The round-2 "lean" loop: a lane's 8 rows are CONSECUTIVE, their rule entries arrive as two 128-bit shared
    // loads, one LDGSTS form, every tile row is copied or zero-filled (~137 SASS instructions per slot per warp, a
    // dependent chain of ~5 clk per instruction: profiles/r02_conv_fetch_bisect.md). It ships for the K-stacked layers
    // (CIN = 16 / 32); scheme 5 = the same with absent rows skipped (measured slower: the per-row branches lengthen
    // the chain, profiles/r02x_conv_variants.txt).
    constexpr int NF = kFetchWarps * 32;
    static_assert(kFetchWarps == 8, "16 tile rows per fetch warp");
    constexpr int kRowBytes = 4 * CIN;  // packed source row: [h1 (2*CIN bytes) | h2 (2*CIN bytes)]
    const int fw = warp - kWarpFetch0, gt = fw * 32 + lane;
    // copy geometry: 16 consecutive lanes cover one tile row (8 units of h1, 8 units of h2), two rows per warp
    // instruction; unit u holds K-elements [8u, 8u+8) of the slot = channels ch0.. of stacked offset `off`
    const int rsub = lane >> 4, part = (lane >> 3) & 1, unit = lane & 7;
    const int off = (unit * 8) / CIN;
    const int src_byte = part * (2 * CIN) + ((unit * 8) % CIN) * 2;
    constexpr int kRowsPerLane = 8;
    const int row0 = 16 * fw + 8 * rsub;  // this lane copies rows row0 .. row0 + 7 (row0 % 8 == 0)
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t dst_lane = ring_u32 + (uint32_t)(part * kATileBytes + row0 * 128);
    const unsigned char* feat_lane = feat + src_byte;

    // rule rows of the NEXT tile, in flight while the current tile is fetched: thread -> row gt % 128, offsets
    // kpart + kParts j (kpart = gt / 128 is warp-uniform)
    constexpr int kParts = NF / kTileM, kPre = (kMaxKV + kParts - 1) / kParts;
    const int srow = gt & (kTileM - 1), kpart = gt >> 7;
    int pre[kPre];
    auto prefetch = [&](int tile) {
      const int o = tile * kTileM + srow;
      const bool ok = tile < n_tiles && o < n_out;
#pragma unroll
      for (int j = 0; j < kPre; j++) {
        const int k = kpart + kParts * j;
        pre[j] = (ok && k < KV) ? __ldg(nbr + (size_t)k * nbr_stride + o) : -1;
      }
    };
    prefetch(blockIdx.x);
    uint32_t q = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      asm volatile("bar.sync 1, %0;" ::"n"(NF) : "memory");  // all copies that read idx_tile are issued
      if (gt == 0) fmask[(it + 1) & 1] = 0u;
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < kPre; j++) {
        const int k = kpart + kParts * j;
        if (k < KV) idx_tile[k * kTileM + srow] = pre[j];
        const uint32_t present = __ballot_sync(0xffffffffu, pre[j] >= 0);  // bit i = tile row 32 * (srow / 32) + i
        if (present) mine |= 1u << k;
        if (kSkip && lane == 0 && k < KV) pres[k * 4 + (srow >> 5)] = present;
      }
      if (lane == 0 && mine) atomicOr(&fmask[it & 1], mine);
      asm volatile("bar.sync 1, %0;" ::"n"(NF) : "memory");
      uint32_t mask = fmask[it & 1];
      if (mask == 0) mask = 1u;  // cannot happen for a well-formed rule book; keeps the protocol total
      mask = group_mask<C::kGK>(mask);
      prefetch(tile + gridDim.x);
      bool first = true;
      while (mask) {
        const int g = __ffs(mask) - 1;  // offset GROUP index = prepared image index
        mask &= mask - 1;
        const uint32_t s = q % C::kStages;
        mbar_wait_as<kSpin>(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
        const uint32_t st = s * (uint32_t)C::kStageBytes;
        if (gt == 0) {
          meta[s].last = (mask == 0);
          meta[s].end = 0;
          if (kSkip) {  // rows without a neighbour at offset g (GK = 1: group = offset) -> output lanes switched off
            const uint4 pm = *reinterpret_cast<const uint4*>(pres + g * 4);
            meta[s].off = make_uint4(~pm.x, ~pm.y, ~pm.z, ~pm.w);
          }
          mbar_arrive_expect_tx(&full[s], (uint32_t)C::kBBytes);
          bulk_g2s(ring_u32 + st + kABytes, wprep + (size_t)g * C::kBBytes, (uint32_t)C::kBBytes, &full[s]);
        }
        const int kk = g * C::kGK + off;
        const bool kv_ok = kk < KV;  // the last group of a layer may be padded with non-existent offsets
        const int4* idx_row = reinterpret_cast<const int4*>(idx_tile + (kv_ok ? kk : 0) * kTileM + row0);
        const int4 sa = idx_row[0], sb = idx_row[1];
        const int src[kRowsPerLane] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
        for (int i = 0; i < kRowsPerLane; i++) {  // (row0 + i) & 7 == i: the swizzle phase is a compile-time constant
          const bool ok = kv_ok && src[i] >= 0;
          const unsigned char* p = feat_lane + (size_t)(ok ? src[i] : 0) * kRowBytes;
          if (kSkip && !first && !ok) continue;  // lane-masked slot: an absent row is not written at all
          if (kCg)
            cp_async16_cg(dst_lane + st + (uint32_t)(i * 128 + ((unit ^ i) << 4)), p, ok ? 16u : 0u);
          else
            cp_async16(dst_lane + st + (uint32_t)(i * 128 + ((unit ^ i) << 4)), p, ok ? 16u : 0u);
        }
        first = false;
        cp_async_arrive_noinc(&full[s]);  // arrives when this thread's copies have landed
        q++;
      }
    }
    {  // termination slot
      const uint32_t s = q % C::kStages;
      mbar_wait_as<kSpin>(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
      if (gt == 0) {
        meta[s].last = 1;
        meta[s].end = 1;
        mbar_arrive(&full[s]);  // stands in for the expect_tx arrival of a real slot
      }
      mbar_arrive(&full[s]);
    }
  } else if (warp < kWarpFetch0 + kFetchWarps) {
    // =========================== fetchers (scheme 2: phase-aligned rows) ===========================
    // Scheme 1 spends ~140 SASS instructions per slot in every fetch warp, 72 of them in the eight row copies
    // (compare, 64-bit multiply, two selects, lane-offset OR, 64-bit add, destination add, LDGSTS), and the measured
    // cost of that dependent chain is ~5 clk per instruction (profiles/r02_conv_fetch_bisect.md). Here a lane copies
    // the eight tile rows r_j = 64 h + 8 j + p that share ONE swizzle phase p (= r & 7): the destination of row j is
    // a constant register + j * 1024 (an LDGSTS immediate), the source is one IMAD.WIDE (index * row bytes + lane
    // base), and an absent neighbour is the LDGSTS zero-fill predicate on the raw index (no address select: a
    // predicated-off copy does not read its source, the CUTLASS zfill convention). Rule tile rows are staged
    // permuted -- entry of tile row r at ((r & 7) * 2 + (r >> 6)) * 8 + ((r >> 3) & 7) -- so that the lane's eight
    // entries are still two 128-bit shared loads. The weight image / slot meta / expect_tx arrival moved to a warp
    // of their own (below), off the critical chain of fetch warp 0.
    constexpr int NF = kFetchWarps * 32;
    static_assert(kFetchWarps == 8, "one swizzle phase per fetch warp");
    constexpr int kRowBytes = 4 * CIN;
    const int fw = warp - kWarpFetch0, gt = fw * 32 + lane;
    const int half = lane >> 4, part = (lane >> 3) & 1, unit = lane & 7;
    const int off = (unit * 8) / CIN;
    const int src_byte = part * (2 * CIN) + ((unit * 8) % CIN) * 2;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t dst_lane =
        ring_u32 + (uint32_t)(part * kATileBytes + (64 * half + fw) * 128 + ((unit ^ fw) << 4));
    const unsigned char* feat_lane = feat + src_byte;
    const int idx_pos = (fw * 2 + half) * 8;  // first of this lane's 8 consecutive (permuted) rule entries

    constexpr int kParts = NF / kTileM, kPre = (kMaxKV + kParts - 1) / kParts;
    const int srow = gt & (kTileM - 1), kpart = gt >> 7;
    const int spos = ((srow & 7) * 2 + (srow >> 6)) * 8 + ((srow >> 3) & 7);
    int pre[kPre];
    auto prefetch = [&](int tile) {
      const int o = tile * kTileM + srow;
      const bool ok = tile < n_tiles && o < n_out;
#pragma unroll
      for (int j = 0; j < kPre; j++) {
        const int k = kpart + kParts * j;
        pre[j] = (ok && k < KV) ? __ldg(nbr + (size_t)k * nbr_stride + o) : -1;
      }
    };
    prefetch(blockIdx.x);
    uint32_t q = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      asm volatile("bar.sync 1, %0;" ::"n"(NF + 32) : "memory");  // all copies that read idx_tile are issued
      if (gt == 0) fmask[(it + 1) & 1] = 0u;
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < kPre; j++) {
        const int k = kpart + kParts * j;
        if (k < KV) idx_tile[k * kTileM + spos] = pre[j];
        const uint32_t present = __ballot_sync(0xffffffffu, pre[j] >= 0);  // bit i = tile row 32 * (srow / 32) + i
        if (present) mine |= 1u << k;
        if (kSkip && lane == 0 && k < KV) pres[k * 4 + (srow >> 5)] = present;
      }
      if (lane == 0 && mine) atomicOr(&fmask[it & 1], mine);
      asm volatile("bar.sync 1, %0;" ::"n"(NF + 32) : "memory");
      uint32_t mask = fmask[it & 1];
      if (mask == 0) mask = 1u;
      mask = group_mask<C::kGK>(mask);
      prefetch(tile + gridDim.x);
      bool first = true;
      while (mask) {
        const int g = __ffs(mask) - 1;
        mask &= mask - 1;
        // rule entries first: their shared-memory latency overlaps the wait for the stage
        const int kk = g * C::kGK + off;
        // a padded (non-existent) offset of the last group reads the all -1 row
        const int4* idx_row = reinterpret_cast<const int4*>((kk < KV ? idx_tile + kk * kTileM : neg_row) + idx_pos);
        const int4 sa = idx_row[0], sb = idx_row[1];
        const uint32_t s = q % C::kStages;
        mbar_wait_as<kSpin>(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
        const uint32_t dst = dst_lane + s * (uint32_t)C::kStageBytes;
        if (kSkip && !first) {  // lane-masked slot: absent rows are simply not copied
          gather_row16_skip<kRowBytes, 0 * 1024, kCg>(dst, feat_lane, sa.x);
          gather_row16_skip<kRowBytes, 1 * 1024, kCg>(dst, feat_lane, sa.y);
          gather_row16_skip<kRowBytes, 2 * 1024, kCg>(dst, feat_lane, sa.z);
          gather_row16_skip<kRowBytes, 3 * 1024, kCg>(dst, feat_lane, sa.w);
          gather_row16_skip<kRowBytes, 4 * 1024, kCg>(dst, feat_lane, sb.x);
          gather_row16_skip<kRowBytes, 5 * 1024, kCg>(dst, feat_lane, sb.y);
          gather_row16_skip<kRowBytes, 6 * 1024, kCg>(dst, feat_lane, sb.z);
          gather_row16_skip<kRowBytes, 7 * 1024, kCg>(dst, feat_lane, sb.w);
        } else {
          gather_row16<kRowBytes, 0 * 1024, kCg>(dst, feat_lane, sa.x);
          gather_row16<kRowBytes, 1 * 1024, kCg>(dst, feat_lane, sa.y);
          gather_row16<kRowBytes, 2 * 1024, kCg>(dst, feat_lane, sa.z);
          gather_row16<kRowBytes, 3 * 1024, kCg>(dst, feat_lane, sa.w);
          gather_row16<kRowBytes, 4 * 1024, kCg>(dst, feat_lane, sb.x);
          gather_row16<kRowBytes, 5 * 1024, kCg>(dst, feat_lane, sb.y);
          gather_row16<kRowBytes, 6 * 1024, kCg>(dst, feat_lane, sb.z);
          gather_row16<kRowBytes, 7 * 1024, kCg>(dst, feat_lane, sb.w);
        }
        first = false;
        cp_async_arrive_noinc(&full[s]);
        q++;
      }
    }
    {  // termination slot: the fetchers' share of the arrivals
      const uint32_t s = q % C::kStages;
      mbar_wait_as<kSpin>(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
      mbar_arrive(&full[s]);
    }
  } else {
    // =========================== weight streamer (scheme 2) ===========================
    // Walks the same slot sequence as the fetchers (same barriers, same usage mask) and per slot publishes the
    // slot meta, posts the expect_tx arrival and launches the 1-D TMA bulk copy of the slot's weight image.
    constexpr int NF = kFetchWarps * 32;
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t q = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      asm volatile("bar.sync 1, %0;" ::"n"(NF + 32) : "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(NF + 32) : "memory");
      uint32_t mask = fmask[it & 1];
      if (mask == 0) mask = 1u;
      mask = group_mask<C::kGK>(mask);
      while (mask) {
        const int g = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t s = q % C::kStages;
        mbar_wait_as<kSpin>(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
        if (lane == 0) {
          meta[s].last = (mask == 0);
          meta[s].end = 0;
          if (kSkip) {  // rows without a neighbour at offset g (GK = 1: group = offset) -> lanes switched off
            const uint4 pm = *reinterpret_cast<const uint4*>(pres + g * 4);
            meta[s].off = make_uint4(~pm.x, ~pm.y, ~pm.z, ~pm.w);
          }
          mbar_arrive_expect_tx(&full[s], (uint32_t)C::kBBytes);
          bulk_g2s(ring_u32 + s * (uint32_t)C::kStageBytes + kABytes, wprep + (size_t)g * C::kBBytes,
                   (uint32_t)C::kBBytes, &full[s]);
        }
        __syncwarp();
        q++;
      }
    }
    {  // termination slot
      const uint32_t s = q % C::kStages;
      mbar_wait_as<kSpin>(&empty[s], ((q / C::kStages) & 1u) ^ 1u);
      if (lane == 0) {
        meta[s].last = 1;
        meta[s].end = 1;
        mbar_arrive(&full[s]);  // stands in for the expect_tx arrival of a real slot
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

// One-time weight preparation: (KV, Cin, Cout) fp32 -> per offset group the exact shared-memory image the
// kernel consumes: [G1 rows | G2 rows] x 128 B, K-major, 128B-swizzled, bf16 split.
__global__ void prepare_weights_kernel(const float* __restrict__ w, int KV, int Cin, int Cout,
                                       unsigned char* __restrict__ img) {
  const int gk = 64 / Cin;  // offsets stacked along K per image
  const int n_groups = (KV + gk - 1) / gk;
  const size_t per_group = (size_t)2 * Cout * 128;
  const int total = n_groups * Cout * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kl = e % 64;  // K index inside the group image
    int t = e / 64;
    const int n = t % Cout;
    const int g = t / Cout;
    const int kk = g * gk + kl / Cin, ci = kl % Cin;
    const float v = kk < KV ? w[((size_t)kk * Cin + ci) * Cout + n] : 0.f;
    const __nv_bfloat16 g1 = __float2bfloat16_rn(v);
    const __nv_bfloat16 g2 = __float2bfloat16_rn(v - __bfloat162float(g1));
    // rows [0,Cout) = g1, rows [Cout, 2*Cout) = g2 (Cout is a multiple of 8, so the g2 rows start on an 8-row
    // group boundary and keep the same (row & 7) swizzle phase)
    const size_t off = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)((((kl >> 3) ^ (n & 7)) << 4) + (kl & 7) * 2);
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)g * per_group + off) = g1;
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)g * per_group + (size_t)Cout * 128 + off) = g2;
  }
}

// fp32 feature rows (Csrc channels) -> packed rows [h1 | h2] of C >= Csrc channels, zero padded (one thread per
// 8 output channels). Padding lets narrow inputs (the 4-channel voxel means) use the K = 16 tensor-core path.
__global__ void feature_pack_kernel(const float* __restrict__ feat, const int* __restrict__ n_rows_ptr, int cap, int Csrc,
                                    int C, unsigned char* __restrict__ packed) {
  const int n_rows = min(*n_rows_ptr, cap);
  const int upr = C / 8;
  const long long total = (long long)n_rows * upr;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / upr;
    const int u = (int)(e % upr), c0 = u * 8;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (c0 < Csrc) a = __ldg(reinterpret_cast<const float4*>(feat + r * Csrc + c0));
    if (c0 + 4 < Csrc) b = __ldg(reinterpret_cast<const float4*>(feat + r * Csrc + c0 + 4));
    uint32_t h1[4], h2[4];
    split2(a.x, a.y, h1[0], h2[0]);
    split2(a.z, a.w, h1[1], h2[1]);
    split2(b.x, b.y, h1[2], h2[2]);
    split2(b.z, b.w, h1[3], h2[3]);
    unsigned char* prow = packed + r * (4 * C) + u * 16;
    *reinterpret_cast<uint4*>(prow) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
    *reinterpret_cast<uint4*>(prow + 2 * C) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
  }
}

// Variant of the tensor-core kernel (see the kernel's header): V3D_TC_FETCH = 1 | 2 | 4 | 5, V3D_TC_CG = 0 | 1,
// V3D_TC_WAIT = 0 | 1. Read once per process.
// Defaults = the fastest parity-green variant of the matrix measured on B200 (profiles/r02x_conv_variants.txt):
// CIN = 64 layers: scheme 4 + cg (conv total 1615 us against 1787 us for scheme 1); other layers: scheme 1 + cg.
constexpr int kDefaultFetchScheme = 4, kDefaultCg = 1, kDefaultSpin = 0;
inline int tc_env_digit(const char* name, int dflt, const char* allowed) {
  const char* e = getenv(name);
  if (e && e[0] && !e[1])
    for (const char* a = allowed; *a; a++)
      if (*a == e[0]) return e[0] - '0';
  return dflt;
}
inline int tc_variant() {
  static const int v = tc_env_digit("V3D_TC_FETCH", kDefaultFetchScheme, "1245") +
                       8 * tc_env_digit("V3D_TC_CG", kDefaultCg, "01") +
                       16 * tc_env_digit("V3D_TC_WAIT", kDefaultSpin, "01");
  return v;
}

template <int CIN, int COUT, int kVer>
int launch_tc_ver(const unsigned char* feat, const unsigned char* wprep, const int* nbr, int nbr_stride, const int* n_out,
                  int out_cap, int KV, const float* scale, const float* shift, int relu, float* out,
                  unsigned char* out_packed, cudaStream_t st) {
  using C = TcCfg<CIN, COUT>;
  static PerDeviceOnce attr_once;
  if (attr_once.needed()) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(sparse_conv_tc_kernel<CIN, COUT, 8, kVer>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    attr_once.done();
  }
  const int tiles_cap = ceil_div(out_cap, kTileM);
  const int grid = tiles_cap < kNumSMs ? (tiles_cap > 0 ? tiles_cap : 1) : kNumSMs;
  sparse_conv_tc_kernel<CIN, COUT, 8, kVer><<<grid, tc_threads(kVer), C::kSmemBytes, st>>>(
      feat, wprep, nbr, nbr_stride, n_out, out_cap, KV, scale, shift, relu, out, out_packed);
  return check_launch();
}

template <int CIN, int COUT>
int launch_tc(const unsigned char* feat, const unsigned char* wprep, const int* nbr, int nbr_stride, const int* n_out,
              int out_cap, int KV, const float* scale, const float* shift, int relu, float* out,
              unsigned char* out_packed, cudaStream_t st) {
  int v = tc_variant();
  if (((v & 7) == 4 || (v & 7) == 5) && CIN != 64) v = (v & ~7) | 1;  // schemes 4, 5: CIN = 64 (one offset per slot) only
#define V3D_TC_VER(VER)                                                                                           \
  if (v == (VER))                                                                                                 \
    return launch_tc_ver<CIN, COUT, (VER)>(feat, wprep, nbr, nbr_stride, n_out, out_cap, KV, scale, shift, relu, out, \
                                           out_packed, st);
  V3D_TC_VER(1)
  V3D_TC_VER(1 + 8)
  V3D_TC_VER(1 + 16)
  V3D_TC_VER(1 + 8 + 16)
  V3D_TC_VER(2)
  V3D_TC_VER(2 + 8)
  V3D_TC_VER(2 + 16)
  V3D_TC_VER(2 + 8 + 16)
  if constexpr (CIN == 64) {
    V3D_TC_VER(4)
    V3D_TC_VER(4 + 8)
    V3D_TC_VER(4 + 16)
    V3D_TC_VER(4 + 8 + 16)
    V3D_TC_VER(5)
    V3D_TC_VER(5 + 8)
    V3D_TC_VER(5 + 16)
    V3D_TC_VER(5 + 8 + 16)
  }
#undef V3D_TC_VER
  return V3D_ERR_INVALID_ARGUMENT;
}

inline bool tc_supported(int KV, int Cin, int Cout) {
  return KV <= kMaxKV && (Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 16 || Cout == 32 || Cout == 64);
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" int v3d_sparse_conv_tc_variant(void) { return tc_variant(); }

extern "C" size_t v3d_sparse_conv_prepared_bytes(int kernel_volume, int Cin, int Cout) {
  if (!tc_supported(kernel_volume, Cin, Cout)) return 0;  // 0 = this shape runs on the exact-fp32 SIMT path
  const int gk = 64 / Cin;  // offsets per image (K = 64 per pipeline slot)
  return (size_t)((kernel_volume + gk - 1) / gk) * (2 * Cout * 128);
}

extern "C" int v3d_sparse_conv_prepare(const float* weight, int kernel_volume, int Cin, int Cout, void* prepared,
                                       size_t prepared_bytes, v3d_stream_t stream) {
  if (!weight || !prepared) return V3D_ERR_INVALID_ARGUMENT;
  const size_t need = v3d_sparse_conv_prepared_bytes(kernel_volume, Cin, Cout);
  if (need == 0) return V3D_ERR_INVALID_ARGUMENT;
  if (prepared_bytes < need) return V3D_ERR_WORKSPACE_TOO_SMALL;
  const int total = ((kernel_volume + 64 / Cin - 1) / (64 / Cin)) * Cout * 64;
  prepare_weights_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      weight, kernel_volume, Cin, Cout, static_cast<unsigned char*>(prepared));
  return check_launch();
}

extern "C" int v3d_feature_pack(const float* feat, const int* n_rows, int capacity, int C_src, int C, void* packed,
                                v3d_stream_t stream) {
  if (!feat || !n_rows || !packed || capacity <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (C != 16 && C != 32 && C != 64) return V3D_ERR_INVALID_ARGUMENT;
  if (C_src <= 0 || C_src > C || (C_src & 3)) return V3D_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(feat) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15)) return V3D_ERR_INVALID_ARGUMENT;
  const long long units = (long long)capacity * (C / 8);
  const long long want = (units + 255) / 256, cap_blocks = (long long)kNumSMs * 8;
  const int blocks = (int)(want < cap_blocks ? want : cap_blocks);
  feature_pack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(feat, n_rows, capacity, C_src, C,
                                                             static_cast<unsigned char*>(packed));
  return check_launch();
}

extern "C" int v3d_sparse_conv_fwd_tc(const void* feat_packed, const void* prepared, const int* nbr, int nbr_stride,
                                      const int* n_out, int out_capacity, int kernel_volume, int Cin, int Cout,
                                      const float* scale, const float* shift, int relu, float* out, void* out_packed,
                                      v3d_stream_t stream) {
  if (!feat_packed || !prepared || !nbr || !n_out || (!out && !out_packed)) return V3D_ERR_INVALID_ARGUMENT;
  if (out_capacity <= 0 || kernel_volume <= 0 || nbr_stride < out_capacity) return V3D_ERR_INVALID_ARGUMENT;
  if ((scale == nullptr) != (shift == nullptr)) return V3D_ERR_INVALID_ARGUMENT;
  if (!tc_supported(kernel_volume, Cin, Cout)) return V3D_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(prepared) & 15) || (reinterpret_cast<uintptr_t>(feat_packed) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(out_packed) & 15))
    return V3D_ERR_INVALID_ARGUMENT;
  const unsigned char* fp = static_cast<const unsigned char*>(feat_packed);
  const unsigned char* wp = static_cast<const unsigned char*>(prepared);
  unsigned char* op = static_cast<unsigned char*>(out_packed);
  cudaStream_t st = as_stream(stream);
#define V3D_TC_CASE(CI, CO)                                                                                      \
  if (Cin == CI && Cout == CO)                                                                                   \
    return launch_tc<CI, CO>(fp, wp, nbr, nbr_stride, n_out, out_capacity, kernel_volume, scale, shift, relu, out, op, \
                             st);
  V3D_TC_CASE(16, 16)
  V3D_TC_CASE(16, 32)
  V3D_TC_CASE(16, 64)
  V3D_TC_CASE(32, 32)
  V3D_TC_CASE(32, 64)
  V3D_TC_CASE(64, 64)
  V3D_TC_CASE(32, 16)
  V3D_TC_CASE(64, 32)
  V3D_TC_CASE(64, 16)
#undef V3D_TC_CASE
  return V3D_ERR_INVALID_ARGUMENT;
}
