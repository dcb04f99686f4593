// a5 + a6 on the 5th-generation tensor cores: output-stationary sparse convolution with tcgen05.mma
// (kind::tf32, 3xTF32 split for fp32-grade accuracy), accumulators AND the gathered A operand in TMEM,
// folded BN + ReLU epilogue.
//
// One persistent CTA per SM, warp specialised (576 threads):
//   warps 0-3   epilogue   : tcgen05.ld the 128 x COUT fp32 accumulator (lane = output row), apply
//                            scale/shift/ReLU, store each output row once
//   warp  4     MMA issuer : warp-uniform loop, one elected lane issues per pipeline slot 2*8 tcgen05.mma
//                            with A from TMEM and B from shared memory, K=8: A_hi*[B_hi|B_lo] (M=128,
//                            N=2*COUT) and A_lo*B_hi (N=COUT); tcgen05.commit frees the A stage, the B
//                            stage and publishes the accumulator through mbarriers
//   warp  5     weight TMA : finds the kernel offsets each tile uses and streams the pre-swizzled,
//                            pre-split W images (hi+lo) with cp.async.bulk (1-D TMA) into a shared-memory
//                            ring, running up to kBStages slots ahead of the MMAs
//   warps 6-9   fetchers   : stage the tile's rule rows in shared memory (the next tile's are prefetched
//                            into registers), then for every slot issue 16-byte cp.async copies of the
//                            gathered neighbour rows (zero-fill for missing neighbours) into a ring of raw
//                            fp32 tiles; completion is tracked by mbarriers (cp.async.mbarrier.arrive), so
//                            the global-load latency of kSStages slots is in flight without holding a
//                            single register
//   warps 10-17 converters : two groups of 4 warps, slot q belongs to group q % 2; thread = output row. A
//                            thread reads its 256-byte row from the ring (conflict-free: 272-byte pitch),
//                            splits every value into tf32 hi / lo parts and tcgen05.st's them into the
//                            TMEM A stage.
// Why A goes through TMEM: with both operands in shared memory every K=8 step re-reads 4 KB of A three
// times; TMEM-resident A leaves only the B reads on the shared-memory port.
// Why fetch and convert are separate roles: with register-staged gathers (LDG -> split -> TMEM in one
// thread) the gather warps were issue/latency bound at ~3000 clk per slot (measured: tensor pipe 50 %
// active, gather warps never waiting on a barrier).
// Output rows are written exactly once (no atomics, deterministic). Arithmetic: every product is
// exact in fp32 (11-bit x 11-bit mantissas); dropping only a_lo*b_lo bounds the relative error of a
// product by ~2^-21, far inside the 1e-4 the contract allows.
//
// B operand layout (K-major, SWIZZLE_128B, fp32 elements): a "chunk" is 2*COUT rows x 128 bytes
// (32 K-elements; rows [0,COUT) = W_hi^T, rows [COUT,2*COUT) = W_lo^T); 8-row groups are 1024 B apart
// (SBO); the 16-byte unit u of row r lives at unit u ^ (r & 7). A slot is always K = 64 (two chunks):
// 64/CIN consecutive kernel offsets are stacked along K.
// TMEM columns: [0, 4*COUT) two accumulator buffers; then kAStages x (64 hi | 64 lo) A stages.
#include "common.cuh"

namespace v3d {
namespace {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 4, kFetchWarps = 4, kConvWarps = 4, kConvGroups = 2;
constexpr int kWarpMma = kEpiWarps, kWarpB = kEpiWarps + 1, kWarpFetch0 = kEpiWarps + 2;
constexpr int kWarpConv0 = kWarpFetch0 + kFetchWarps;
constexpr int kThreads = 32 * (kWarpConv0 + kConvWarps * kConvGroups);  // 576
constexpr int kMaxKV = 27;
constexpr int kRowPitch = 256 + 16;              // bytes between rows of a raw A tile (bank-conflict-free)
constexpr int kStageBytes = kTileM * kRowPitch;  // one raw fp32 A tile: 128 rows x 64 K-elements

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 16-byte asynchronous copy global -> shared; src_bytes = 0 writes zeros (missing neighbour)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) | version=1 [46,48)
// | base_offset=0 [49,52) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
// a,b K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A operand from TMEM, B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

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

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512))); }

template <int CIN, int COUT>
struct TcCfg {
  // Every slot is a K = 64 GEMM step: kGK = 64/CIN consecutive kernel offsets are stacked along K (their
  // gathered rows side by side in the A stage, their weights stacked in the B image), so small-channel layers
  // pay the per-slot pipeline handshakes once per 64 K-elements instead of once per 16 or 32.
  static constexpr int kGK = 64 / CIN;
  static constexpr int kChunks = 2;
  static constexpr int kKSteps = 8;
  // one B chunk = 2*COUT rows x 128 B: rows [0, COUT) hold W_hi^T, rows [COUT, 2*COUT) hold W_lo^T, so that
  // ONE N = 2*COUT MMA computes A_hi*[B_hi | B_lo] (the gathered operand is fetched once for both products)
  static constexpr int kBChunkBytes = 2 * COUT * 128;
  static constexpr int kBBytes = kChunks * kBChunkBytes;  // == one prepared image (one offset group)
  static constexpr int kBStages = COUT == 64 ? 3 : 4;
  static constexpr int kSStages = COUT == 64 ? 3 : 4;     // raw A tiles in flight (fetch -> convert ring)
  static constexpr int kAccBufCols = 2 * COUT;            // [A_hi*B_hi + A_lo*B_hi | A_hi*B_lo]
  static constexpr int kAccCols = 2 * kAccBufCols;        // double buffered
  static constexpr int kAStageCols = 128;                 // 64 hi | 64 lo
  static constexpr int kAStagesFit = (512 - kAccCols) / kAStageCols;
  static constexpr int kAStages = kAStagesFit > 4 ? 4 : kAStagesFit;
  static constexpr int kTmemCols = pow2_cols(kAccCols + kAStages * kAStageCols);
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kBStages * kBBytes +
                                       (size_t)kSStages * kStageBytes + sizeof(int) * kMaxKV * kTileM +
                                       1024 /*barriers + meta*/ + 2 * COUT * sizeof(float);
  static_assert(kAStages >= 2 && kBStages >= 2 && kSStages >= 2, "pipeline needs two stages");
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB per CTA)");
};

struct SlotMeta {
  int last;  // 1 = last slot of its output tile
  int end;   // != 0: all tiles done (1 = forward the termination to the MMA warp, 2 = just stop)
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
sparse_conv_tc_kernel(const float* __restrict__ feat, const unsigned char* __restrict__ wprep,
                      const int* __restrict__ nbr, int nbr_stride, const int* __restrict__ n_out_ptr, int out_cap,
                      int KV, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                      float* __restrict__ out) {
  using C = TcCfg<CIN, COUT>;
  extern __shared__ unsigned char smem_raw[];
  // round up to 1024 B (SWIZZLE_128B atoms) by OFFSETTING the __shared__ array: casting through an integer
  // would make every later access a generic LD/ST instead of LDS/STS
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* bring = base;                                         // kBStages x [B_hi | B_lo] images
  unsigned char* sring = bring + (size_t)C::kBStages * C::kBBytes;     // kSStages x raw fp32 A tiles
  int* idx_tile = reinterpret_cast<int*>(sring + (size_t)C::kSStages * kStageBytes);  // [KV][128]
  unsigned char* tail = reinterpret_cast<unsigned char*>(idx_tile + kMaxKV * kTileM);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(tail);  // [4]
  uint64_t* empty_a = full_a + 4;                        // [4]
  uint64_t* full_b = empty_a + 4;                        // [4]
  uint64_t* empty_b = full_b + 4;                        // [4]
  uint64_t* acc_full = empty_b + 4;                      // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint64_t* sfull = acc_empty + 2;                       // [4]
  uint64_t* sempty = sfull + 4;                          // [4]
  SlotMeta* meta = reinterpret_cast<SlotMeta*>(sempty + 4);  // [4]  converter -> MMA issuer (per A stage)
  SlotMeta* smeta = meta + 4;                                // [4]  fetcher -> converter (per raw stage)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smeta + 4);
  uint32_t* fmask = tmem_slot + 1;  // [2] offsets used by the tile being staged (parity double buffer)
  float* s_scale = reinterpret_cast<float*>(tail + 1024);
  float* s_shift = s_scale + COUT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_out = min(*n_out_ptr, out_cap);
  const int n_tiles = (n_out + kTileM - 1) / kTileM;

  if (tid == 0) {
    for (int s = 0; s < C::kAStages; s++) {
      mbar_init(&full_a[s], kConvWarps * 32);  // the one converter group that owns the slot
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < C::kBStages; s++) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kEpiWarps * 32);
    }
    for (int s = 0; s < C::kSStages; s++) {
      mbar_init(&sfull[s], kFetchWarps * 32 + 1);  // every fetch thread's cp.async group + the meta writer
      mbar_init(&sempty[s], kConvWarps * 32);
    }
    fmask[0] = fmask[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = tid; c < COUT; c += kThreads) {
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
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      const int row = tile * kTileM + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * C::kAccBufCols);
      float* orow = out + (size_t)row * COUT;
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t v[16], v2[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
            "%14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
            "%14, %15}, [%16];"
            : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]),
              "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]),
              "=r"(v2[15])
            : "r"(taddr + (uint32_t)(COUT + c0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 16; e++)  // (A_hi*B_hi + A_lo*B_hi) + A_hi*B_lo
          v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
        if (c0 + 16 >= COUT) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[a]);
        }
        if (row < n_out) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float4 o;
            o.x = fmaf(__uint_as_float(v[4 * q + 0]), sc[c0 + 4 * q + 0], sh[c0 + 4 * q + 0]);
            o.y = fmaf(__uint_as_float(v[4 * q + 1]), sc[c0 + 4 * q + 1], sh[c0 + 4 * q + 1]);
            o.z = fmaf(__uint_as_float(v[4 * q + 2]), sc[c0 + 4 * q + 2], sh[c0 + 4 * q + 2]);
            o.w = fmaf(__uint_as_float(v[4 * q + 3]), sc[c0 + 4 * q + 3], sh[c0 + 4 * q + 3]);
            if (relu) {
              o.x = fmaxf(o.x, 0.f);
              o.y = fmaxf(o.y, 0.f);
              o.z = fmaxf(o.z, 0.f);
              o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(orow + c0 + 4 * q) = o;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer ===========================
    // The WHOLE warp walks the slot sequence with warp-uniform control flow and only the tcgen05
    // instructions are predicated on one elected lane: operands then live in uniform registers. (With a
    // single divergent thread every UTCHMMA needed ELECT + R2UR moves: ~380 SASS instructions and ~2500
    // clk of issue time per slot, three times the 768 clk the tensor pipe needs.)
    constexpr uint32_t idesc_wide = make_idesc(kTileM, 2 * COUT), idesc_hi = make_idesc(kTileM, COUT);
    const uint32_t b_ring = smem_u32(bring);
    uint32_t q = 0;
    int it = 0;
    bool done = false;
    while (!done) {
      const int a = it & 1;
      uint32_t accum = 0u;
      bool tile_open = false;
      while (true) {
        const uint32_t as = q % C::kAStages, bs = q % C::kBStages;
        mbar_wait(&full_a[as], (q / C::kAStages) & 1u);
        const int m_last = meta[as].last, m_end = meta[as].end;
        if (m_end) {
          done = true;
          break;
        }
        if (!tile_open) {  // first slot of a tile: the accumulator buffer must have been drained
          mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
          tile_open = true;
        }
        mbar_wait(&full_b[bs], (q / C::kBStages) & 1u);
        tc_fence_after();
        const uint64_t db0 = make_desc(b_ring + bs * (uint32_t)C::kBBytes);
        const uint32_t a_hi = tmem_base + (uint32_t)(C::kAccCols + as * C::kAStageCols), a_lo = a_hi + 64;
        const uint32_t d = tmem_base + (uint32_t)(a * C::kAccBufCols);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < C::kKSteps; ks++) {
            // descriptor start address advances in 16-byte units (low 14 bits of the descriptor)
            const uint64_t db = db0 + (uint64_t)(((ks >> 2) * C::kBChunkBytes + (ks & 3) * 32) >> 4);
            umma_tf32_ts(d, a_hi + 8u * ks, db, idesc_wide, accum);  // A_hi * [B_hi | B_lo], N = 2*COUT
            umma_tf32_ts(d, a_lo + 8u * ks, db, idesc_hi, 1u);       // A_lo * B_hi into the first COUT columns
            accum = 1u;
          }
          umma_commit(&empty_a[as]);  // stages reusable once these MMAs have read them
          umma_commit(&empty_b[bs]);
          if (m_last) umma_commit(&acc_full[a]);
        }
        accum = 1u;
        __syncwarp();
        q++;
        if (m_last) break;
      }
      it++;
    }
  } else if (warp == kWarpB) {
    // =========================== weight loader (1-D TMA), whole warp ===========================
    uint32_t q = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // which kernel offsets does this tile use? lanes cover rows lane + 32 j; 9 offsets per round trip
      const int o0 = tile * kTileM + lane;
      uint32_t mask = 0;
      for (int k0 = 0; k0 < KV; k0 += 9) {
        int v[9][4];
#pragma unroll
        for (int u = 0; u < 9; u++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int o = o0 + 32 * j;
            v[u][j] = (k0 + u < KV && o < n_out) ? __ldg(&nbr[(size_t)(k0 + u) * nbr_stride + o]) : -1;
          }
#pragma unroll
        for (int u = 0; u < 9; u++) {
          const bool any = (v[u][0] >= 0) | (v[u][1] >= 0) | (v[u][2] >= 0) | (v[u][3] >= 0);
          if (__any_sync(0xffffffffu, any)) mask |= 1u << (k0 + u);
        }
      }
      if (mask == 0) mask = 1u;  // must mirror the fetchers' rule
      mask = group_mask<C::kGK>(mask);
      while (mask) {
        const int kk = __ffs(mask) - 1;  // offset GROUP index = prepared image index
        mask &= mask - 1;
        const uint32_t bs = q % C::kBStages;
        if (lane == 0) {
          mbar_wait(&empty_b[bs], ((q / C::kBStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full_b[bs], (uint32_t)C::kBBytes);
          bulk_g2s(bring + (size_t)bs * C::kBBytes, wprep + (size_t)kk * C::kBBytes, (uint32_t)C::kBBytes,
                   &full_b[bs]);
        }
        q++;
      }
      __syncwarp();
    }
  } else if (warp < kWarpConv0) {
    // =========================== fetchers ===========================
    constexpr int NF = kFetchWarps * 32;
    const int fw = warp - kWarpFetch0, gt = fw * 32 + lane;
    // copy geometry: 16 consecutive lanes cover one 256-byte stage row (two rows per warp instruction)
    const int unit = lane & 15, rsub = lane >> 4;
    const int off = (unit * 4) / CIN;   // which offset of the slot's group this 16-byte unit belongs to
    const int col = (unit * 4) % CIN;   // first channel of the unit inside that offset's feature row
    const uint32_t dst_lane = smem_u32(sring) + (uint32_t)((fw * 2 + rsub) * kRowPitch + unit * 16);

    int pre[kMaxKV];  // rule rows of the NEXT tile (row = gt), in flight while the current tile is fetched
    auto prefetch = [&](int tile) {
      const int o = tile * kTileM + gt;
      const bool ok = tile < n_tiles && o < n_out;
#pragma unroll
      for (int k = 0; k < kMaxKV; k++) pre[k] = (ok && k < KV) ? __ldg(nbr + (size_t)k * nbr_stride + o) : -1;
    };
    prefetch(blockIdx.x);
    uint32_t q = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      asm volatile("bar.sync 1, %0;" ::"n"(NF) : "memory");  // all copies that read idx_tile are issued
      if (gt == 0) fmask[(it + 1) & 1] = 0u;
      uint32_t mine = 0;
#pragma unroll
      for (int k = 0; k < kMaxKV; k++) {
        if (k < KV) idx_tile[k * kTileM + gt] = pre[k];
        if (__any_sync(0xffffffffu, pre[k] >= 0)) mine |= 1u << k;
      }
      if (lane == 0 && mine) atomicOr(&fmask[it & 1], mine);
      asm volatile("bar.sync 1, %0;" ::"n"(NF) : "memory");
      uint32_t mask = fmask[it & 1];
      if (mask == 0) mask = 1u;  // cannot happen for a well-formed rule book; keeps the protocol total
      mask = group_mask<C::kGK>(mask);
      prefetch(tile + gridDim.x);
      while (mask) {
        const int g = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t s = q % C::kSStages;
        mbar_wait(&sempty[s], ((q / C::kSStages) & 1u) ^ 1u);
        const int kk = g * C::kGK + off;
        const bool kv_ok = kk < KV;  // the last group of a layer may be padded with non-existent offsets
        const int* idx_row = idx_tile + (kv_ok ? kk : 0) * kTileM + fw * 2 + rsub;
        const uint32_t dst = dst_lane + s * (uint32_t)kStageBytes;
        int src[16];
#pragma unroll
        for (int i = 0; i < 16; i++) src[i] = idx_row[8 * i];  // rows 8 i + 2 fw + rsub
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const bool ok = kv_ok && src[i] >= 0;
          const float* p = feat + (size_t)(ok ? src[i] : 0) * CIN + col;
          cp_async16(dst + (uint32_t)(8 * i * kRowPitch), p, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(&sfull[s]);
        if (gt == 0) {
          smeta[s].last = (mask == 0);
          smeta[s].end = 0;
          mbar_arrive(&sfull[s]);  // release: publishes the meta write
        }
        q++;
      }
    }
    // two termination slots, one per converter group
    for (int t = 0; t < kConvGroups; t++) {
      const uint32_t s = q % C::kSStages;
      mbar_wait(&sempty[s], ((q / C::kSStages) & 1u) ^ 1u);
      cp_async_arrive_noinc(&sfull[s]);
      if (gt == 0) {
        smeta[s].last = 1;
        smeta[s].end = 1 + t;
        mbar_arrive(&sfull[s]);
      }
      q++;
    }
  } else {
    // =========================== converters ===========================
    const int cw = warp - kWarpConv0;
    const int grp = cw / kConvWarps;
    const bool leader = (cw % kConvWarps) == 0 && lane == 0;
    const int my_row = 32 * (warp & 3) + lane;  // TMEM lane == tile row; a warp may only touch its lane quarter
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    for (uint32_t q = grp;; q += kConvGroups) {
      const uint32_t s = q % C::kSStages, as = q % C::kAStages;
      mbar_wait(&sfull[s], (q / C::kSStages) & 1u);
      const int m_last = smeta[s].last, m_end = smeta[s].end;
      if (m_end) {
        if (m_end == 1) {  // this group owns the slot number the MMA warp will look at next
          mbar_wait(&empty_a[as], ((q / C::kAStages) & 1u) ^ 1u);
          if (leader) {
            meta[as].last = 1;
            meta[as].end = 1;
          }
          mbar_arrive(&full_a[as]);
        }
        break;
      }
      mbar_wait(&empty_a[as], ((q / C::kAStages) & 1u) ^ 1u);
      tc_fence_after();
      if (leader) {
        meta[as].last = m_last;
        meta[as].end = 0;
      }
      const float4* rowp = reinterpret_cast<const float4*>(sring + (size_t)s * kStageBytes + (size_t)my_row * kRowPitch);
      const uint32_t a_hi = tmem_base + lane_base + (uint32_t)(C::kAccCols + as * C::kAStageCols);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float own[16];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const float4 x = rowp[4 * j + u];
          own[4 * u + 0] = x.x;
          own[4 * u + 1] = x.y;
          own[4 * u + 2] = x.z;
          own[4 * u + 3] = x.w;
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const uint32_t h = __float_as_uint(own[e]) & 0xFFFFE000u;
          hi[e] = h;
          lo[e] = __float_as_uint(own[e] - __uint_as_float(h));
        }
        tmem_st16(a_hi + 16u * j, hi);
        tmem_st16(a_hi + 64u + 16u * j, lo);
      }
      mbar_arrive(&sempty[s]);  // the row has been consumed (the TMEM stores depend on every load)
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(&full_a[as]);  // release also orders the leader's meta write
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
// kernel consumes: chunks x ([hi rows | lo rows] x 128 B), K-major, 128B-swizzled, tf32-split.
__global__ void prepare_weights_kernel(const float* __restrict__ w, int KV, int Cin, int Cout,
                                       unsigned char* __restrict__ img) {
  const int gk = 64 / Cin;                      // offsets stacked along K per image
  const int n_groups = (KV + gk - 1) / gk;
  const size_t chunk_bytes = (size_t)2 * Cout * 128, per_group = 2 * chunk_bytes;
  const int total = n_groups * Cout * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kl = e % 64;                      // K index inside the group image
    int t = e / 64;
    const int n = t % Cout;
    const int g = t / Cout;
    const int kk = g * gk + kl / Cin, ci = kl % Cin;
    const float v = kk < KV ? w[((size_t)kk * Cin + ci) * Cout + n] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float lo = v - hi;
    const int ch = kl >> 5, kc = kl & 31;
    // chunk-major; inside a chunk rows [0,Cout) = hi, rows [Cout, 2*Cout) = lo (Cout is a multiple of 8, so the
    // lo rows start on an 8-row group boundary and keep the same (row & 7) swizzle phase)
    const size_t off = (size_t)ch * chunk_bytes + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 +
                       (size_t)((((kc >> 2) ^ (n & 7)) << 4) + (kc & 3) * 4);
    *reinterpret_cast<float*>(img + (size_t)g * per_group + off) = hi;
    *reinterpret_cast<float*>(img + (size_t)g * per_group + (size_t)Cout * 128 + off) = lo;
  }
}

template <int CIN, int COUT>
int launch_tc(const float* feat, const unsigned char* wprep, const int* nbr, int nbr_stride, const int* n_out,
              int out_cap, int KV, const float* scale, const float* shift, int relu, float* out, cudaStream_t st) {
  using C = TcCfg<CIN, COUT>;
  static bool attr_set = false;
  if (!attr_set) {
    V3D_CUDA_TRY(cudaFuncSetAttribute(sparse_conv_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C::kSmemBytes));
    attr_set = true;
  }
  const int tiles_cap = ceil_div(out_cap, kTileM);
  const int grid = tiles_cap < kNumSMs ? (tiles_cap > 0 ? tiles_cap : 1) : kNumSMs;
  sparse_conv_tc_kernel<CIN, COUT><<<grid, kThreads, C::kSmemBytes, st>>>(feat, wprep, nbr, nbr_stride, n_out, out_cap,
                                                                         KV, scale, shift, relu, out);
  return check_launch();
}

inline bool tc_supported(int KV, int Cin, int Cout) {
  return KV <= kMaxKV && (Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 16 || Cout == 32 || Cout == 64);
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" size_t v3d_sparse_conv_prepared_bytes(int kernel_volume, int Cin, int Cout) {
  if (!tc_supported(kernel_volume, Cin, Cout)) return 0;  // 0 = this shape runs on the exact-fp32 SIMT path
  const int gk = 64 / Cin;  // offsets per image (K = 64 per pipeline slot)
  return (size_t)((kernel_volume + gk - 1) / gk) * 2 * (2 * Cout * 128);
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

extern "C" int v3d_sparse_conv_fwd_tc(const float* feat, const void* prepared, const int* nbr, int nbr_stride,
                                      const int* n_out, int out_capacity, int kernel_volume, int Cin, int Cout,
                                      const float* scale, const float* shift, int relu, float* out,
                                      v3d_stream_t stream) {
  if (!feat || !prepared || !nbr || !n_out || !out) return V3D_ERR_INVALID_ARGUMENT;
  if (out_capacity <= 0 || kernel_volume <= 0 || nbr_stride < out_capacity) return V3D_ERR_INVALID_ARGUMENT;
  if ((scale == nullptr) != (shift == nullptr)) return V3D_ERR_INVALID_ARGUMENT;
  if (!tc_supported(kernel_volume, Cin, Cout)) return V3D_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(prepared) & 15) || (reinterpret_cast<uintptr_t>(feat) & 15))
    return V3D_ERR_INVALID_ARGUMENT;
  const unsigned char* wp = static_cast<const unsigned char*>(prepared);
  cudaStream_t st = as_stream(stream);
#define V3D_TC_CASE(CI, CO)                                                                                       \
  if (Cin == CI && Cout == CO)                                                                                    \
    return launch_tc<CI, CO>(feat, wp, nbr, nbr_stride, n_out, out_capacity, kernel_volume, scale, shift, relu, out, st);
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
