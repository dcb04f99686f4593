// a4: rule-book construction for SubMConv3d / SparseConv3d (spconv get_indice_pairs; layers built at
// vision3d/detector/sparse_cnn.py:15-30,151-175) on sm_100a.
//
// Output-stationary rule table: nbr[kk*stride + o] = input row that feeds output row o through
// kernel offset kk (or -1). This is the transpose of upstream's per-offset (in,out) pair lists and
// is what lets the convolution accumulate every offset of an output tile in TMEM/registers and
// write each output row exactly once (no atomics, no zero-init, BN+ReLU fused in the epilogue).
//
// Site table  : epoch-tagged hash (common.cuh) of flat (b,z,y,x) -> row, one per resolution level.
// SubM        : one lookup per (output row, offset).
// Strided conv: candidates (in row, offset) mark a bitmap over the output grid; two-level popcount
//               prefix gives every marked cell its rank in ascending flat order = its output row;
//               then the same lookup kernel fills nbr from the INPUT level's site table.
// All counts stay on the device (n_rows / n_out are device ints): no host synchronisation.
#include "common.cuh"

namespace v3d {
namespace {

struct TableHeader {
  unsigned int epoch;
  unsigned int pad[63];
};

struct SiteTable {
  TableHeader* hdr;
  unsigned long long* keys;
  unsigned long long* vals;  // [calls:32 | row:32]
  unsigned int cap;
  size_t total;
};

inline SiteTable table_layout(void* base, int capacity_rows) {
  SiteTable t;
  char* p = static_cast<char*>(base);
  t.cap = next_pow2((unsigned int)(4 * (size_t)(capacity_rows > 512 ? capacity_rows : 512)));
  size_t off = 0;
  t.hdr = reinterpret_cast<TableHeader*>(p + off);
  off = align_up(off + sizeof(TableHeader), 256);
  t.keys = reinterpret_cast<unsigned long long*>(p + off);
  off = align_up(off + 8ull * t.cap, 256);
  t.vals = reinterpret_cast<unsigned long long*>(p + off);
  off = align_up(off + 8ull * t.cap, 256);
  t.total = off;
  return t;
}

struct Shape3 {
  int d[3];
};

__device__ __forceinline__ unsigned long long flat_key(int b, int z, int y, int x, const Shape3& s) {
  return (((unsigned long long)b * s.d[0] + z) * s.d[1] + y) * (unsigned long long)s.d[2] + x;
}

__global__ void table_bump_kernel(TableHeader* hdr) { hdr->epoch += 1; }

__global__ void __launch_bounds__(256) table_build_kernel(SiteTable T, const int4* __restrict__ idx,
                                                          const int* __restrict__ n_rows, int cap_rows,
                                                          Shape3 shape) {
  const int n = min(*n_rows, cap_rows);
  const unsigned int calls = T.hdr->epoch;
  const unsigned int epoch = epoch24(calls);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = idx[i];
    unsigned int s = table_claim(T.keys, T.cap - 1, epoch, flat_key(c.x, c.y, c.z, c.w, shape));
    T.vals[s] = ((unsigned long long)calls << 32) | (unsigned int)i;
  }
}

struct ConvGeom {
  int ks[3], stride[3], pad[3], dil[3];
  int in_shape[3], out_shape[3];
  int KV;
};

// nbr[kk][o] = row of the input site at out*stride - pad + k*dil (or -1). Used by SubM (out set ==
// in set, stride 1, pad k/2) and by the strided conv once its output rows are known.
__global__ void __launch_bounds__(256) rule_lookup_kernel(SiteTable T, const int4* __restrict__ out_idx,
                                                          const int* __restrict__ n_out, int out_cap,
                                                          ConvGeom G, int identity_kk, int mirror_kv,
                                                          int* __restrict__ nbr, int nbr_stride) {
  // mirror_kv = KV for SubM (else 0): nbr[kk][o] = r  <=>  nbr[KV-1-kk][r] = o, so only the offsets below the
  // centre are looked up and each hit also writes its mirror entry (the upper half is pre-filled with -1)
  const int n = min(*n_out, out_cap);
  const unsigned int calls = T.hdr->epoch;
  const unsigned int epoch = epoch24(calls);
  Shape3 ish{{G.in_shape[0], G.in_shape[1], G.in_shape[2]}};
  // one thread per output row walks all kernel offsets (coordinates loaded once; for a fixed offset
  // consecutive threads write consecutive nbr entries)
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    const int4 c = out_idx[o];
    int kk = 0;
    for (int kz = 0; kz < G.ks[0]; kz++) {
      const int z = c.y * G.stride[0] - G.pad[0] + kz * G.dil[0];
      for (int ky = 0; ky < G.ks[1]; ky++) {
        const int y = c.z * G.stride[1] - G.pad[1] + ky * G.dil[1];
        const bool zy_ok = z >= 0 && z < G.in_shape[0] && y >= 0 && y < G.in_shape[1];
        for (int kx = 0; kx < G.ks[2]; kx++, kk++) {
          int r;
          if (mirror_kv && kk > identity_kk) continue;
          if (kk == identity_kk) {
            r = o;
          } else {
            const int x = c.w * G.stride[2] - G.pad[2] + kx * G.dil[2];
            r = -1;
            if (zy_ok && x >= 0 && x < G.in_shape[2]) {
              unsigned int s = table_find(T.keys, T.cap - 1, epoch, flat_key(c.x, z, y, x, ish));
              if (s != 0xFFFFFFFFu) {
                const unsigned long long v = __ldg(&T.vals[s]);
                if ((unsigned int)(v >> 32) == calls) r = (int)(unsigned int)v;  // else: stale alias, not a site
              }
            }
            if (mirror_kv && r >= 0) nbr[(size_t)(mirror_kv - 1 - kk) * nbr_stride + r] = o;
          }
          nbr[(size_t)kk * nbr_stride + o] = r;
        }
      }
    }
  }
}

// ---- strided conv: discover output sites --------------------------------------------------------
constexpr int kCoarse = 256;   // cells per level-1 counter (8 bitmap words = one 32-byte sector)
constexpr int kWordsPerCoarse = kCoarse / 32;

struct ConvWs {
  unsigned int* bitmap;  // cells/32 words
  int* l1;               // per 1024 cells: count, then exclusive prefix inside its level-2 group
  int* l2;               // per 1024 l1 entries: count, then exclusive prefix
  unsigned int* uniq;    // reserved (was: append list of marked cells; the bitmap is enumerated instead)
  int* n_uniq;           // reserved
  size_t n_words, n_l1, n_l2;
  int cap;            // row capacity of the level this workspace indexes (ranks >= cap are overflow rows)
  size_t zero_bytes;  // prefix of the workspace that must be zeroed per call
  size_t total;
};

inline ConvWs conv_layout(void* base, unsigned long long cells, int out_capacity) {
  ConvWs w;
  char* p = static_cast<char*>(base);
  w.n_words = (size_t)((cells + 31) / 32);
  w.n_l1 = (size_t)((cells + kCoarse - 1) / kCoarse);
  w.n_l2 = (w.n_l1 + 1023) / 1024;
  w.cap = out_capacity;
  size_t off = 0;
  w.bitmap = reinterpret_cast<unsigned int*>(p + off);
  off = align_up(off + 4 * (w.n_l1 * kWordsPerCoarse), 256);  // whole sectors
  w.l1 = reinterpret_cast<int*>(p + off);
  off = align_up(off + 4 * (w.n_l2 * 1024), 256);
  w.l2 = reinterpret_cast<int*>(p + off);
  off = align_up(off + 4 * (w.n_l2 > 1024 ? w.n_l2 : 1024), 256);
  w.n_uniq = reinterpret_cast<int*>(p + off);
  off = align_up(off + 4, 256);
  w.zero_bytes = off;
  w.uniq = reinterpret_cast<unsigned int*>(p + off);
  off = align_up(off + 4ull * (size_t)out_capacity, 256);
  w.total = off;
  return w;
}

// Marks every output cell some input row reaches. Fire-and-forget RED.OR only: no returned value, no global
// append list, no per-group counter (an earlier version appended "fresh" cells to a list through ONE global
// counter -- tens of thousands of same-address atomics, 80-130 us per level; the counts and the cell
// enumeration now come from the bitmap itself).
__global__ void __launch_bounds__(256) conv_mark_kernel(const int4* __restrict__ idx,
                                                        const int* __restrict__ n_rows, int cap_rows,
                                                        ConvGeom G, ConvWs W) {
  const int n = min(*n_rows, cap_rows);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 c = idx[i];
    for (int kz = 0; kz < G.ks[0]; kz++) {
      const int nz = c.y + G.pad[0] - kz * G.dil[0];
      const int oz = nz / G.stride[0];
      if (nz < 0 || (nz - oz * G.stride[0]) != 0 || oz >= G.out_shape[0]) continue;
      for (int ky = 0; ky < G.ks[1]; ky++) {
        const int ny = c.z + G.pad[1] - ky * G.dil[1];
        const int oy = ny / G.stride[1];
        if (ny < 0 || (ny - oy * G.stride[1]) != 0 || oy >= G.out_shape[1]) continue;
        const size_t row_base = (((size_t)c.x * G.out_shape[0] + oz) * G.out_shape[1] + oy) * G.out_shape[2];
        for (int kx = 0; kx < G.ks[2]; kx++) {
          const int nx = c.w + G.pad[2] - kx * G.dil[2];
          const int ox = nx / G.stride[2];
          if (nx < 0 || (nx - ox * G.stride[2]) != 0 || ox >= G.out_shape[2]) continue;
          const unsigned int cell = (unsigned int)(row_base + ox);
          atomicOr(&W.bitmap[cell >> 5], 1u << (cell & 31));
        }
      }
    }
  }
}

// l1[i] = popcount of coarse group i (one 32-byte bitmap sector), then in place -> exclusive prefix inside the
// 1024-group block g; l2[g] = block total
__global__ void __launch_bounds__(1024) conv_scan_l1_kernel(ConvWs W) {
  __shared__ int sm[33];
  const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
  int v = 0;
  if (i < W.n_l1) {
    const uint4* p = reinterpret_cast<const uint4*>(W.bitmap) + 2 * i;
    const uint4 a = p[0], b = p[1];
    v = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
  }
  int total;
  int ex = block_exclusive_scan(v, sm, total);
  if (i < W.n_l1) W.l1[i] = ex;
  if (threadIdx.x == 0) W.l2[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) conv_scan_l2_kernel(ConvWs W, int* __restrict__ n_out) {
  __shared__ int sm[33];
  int carry = 0;
  for (size_t base = 0; base < W.n_l2; base += 1024) {  // one pass for <= 268 M cells, loops beyond
    const size_t i = base + threadIdx.x;
    int v = i < W.n_l2 ? W.l2[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, sm, total);
    if (i < W.n_l2) W.l2[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *n_out = carry;  // un-clamped: caller detects overflow as n_out > capacity
}

// row of an active cell = number of active cells before it in flat order (two-level prefix + <= 1 sector
// of popcounts); -1 if the cell is not active. This IS the site index of every level produced by a
// strided conv (its rows are numbered in ascending flat order), so those levels need no hash table.
__device__ __forceinline__ int rank_of_cell(const ConvWs& W, unsigned int cell) {
  const unsigned int wq = cell >> 5;
  const unsigned int word = __ldg(&W.bitmap[wq]);
  const unsigned int bit = 1u << (cell & 31);
  if (!(word & bit)) return -1;
  const unsigned int g1 = cell / kCoarse;
  int rank = __ldg(&W.l2[g1 >> 10]) + __ldg(&W.l1[g1]) + __popc(word & (bit - 1u));
  for (unsigned int w = g1 * kWordsPerCoarse; w < wq; w++) rank += __popc(__ldg(&W.bitmap[w]));
  // A level that overflowed its row capacity has cells whose rank is >= cap: those rows do not exist in the
  // feature / index / rule buffers (the producers clamp), so they must read as "no neighbour" here. The overflow
  // itself is reported through the un-clamped n_out counter (SecondEngine.finalize raises on it).
  return rank < W.cap ? rank : -1;
}

// Output rows in ascending flat order: one thread per bitmap word walks its set bits; the row number of the
// word's first active cell is the two-level prefix plus the popcounts of the earlier words of its sector.
__global__ void __launch_bounds__(256) conv_rank_kernel(ConvWs W, ConvGeom G, int out_capacity,
                                                        int4* __restrict__ out_idx) {
  const size_t n_words = W.n_l1 * kWordsPerCoarse;
  for (size_t wq = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wq < n_words; wq += (size_t)gridDim.x * blockDim.x) {
    unsigned int word = W.bitmap[wq];
    if (word == 0u) continue;
    const size_t g = wq / kWordsPerCoarse;
    int rank = W.l2[g >> 10] + W.l1[g];
    for (size_t w = g * kWordsPerCoarse; w < wq; w++) rank += __popc(W.bitmap[w]);
    // decode the first cell once, then step along x (cells of one word are consecutive)
    unsigned int r = (unsigned int)(wq * 32);
    int x = r % G.out_shape[2];
    r /= G.out_shape[2];
    int y = r % G.out_shape[1];
    r /= G.out_shape[1];
    int z = r % G.out_shape[0];
    int bidx = (int)(r / G.out_shape[0]);
    int prev = 0;
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      x += bit - prev;
      prev = bit;
      while (x >= G.out_shape[2]) {  // carry into y / z / batch (at most a few times per word)
        x -= G.out_shape[2];
        if (++y == G.out_shape[1]) {
          y = 0;
          if (++z == G.out_shape[0]) {
            z = 0;
            bidx++;
          }
        }
      }
      if (rank < out_capacity) out_idx[rank] = make_int4(bidx, z, y, x);
      rank++;
    }
  }
}

// set cells before `cell` in flat order (two-level prefix + the earlier words of the cell's 32-byte sector, read as two
// 128-bit loads); `word` = the bitmap word of the cell (already loaded by the caller)
__device__ __forceinline__ int rank_before(const ConvWs& W, unsigned int cell, unsigned int word) {
  const unsigned int wq = cell >> 5, g1 = cell / kCoarse, wi = wq & (kWordsPerCoarse - 1);
  int rank = __ldg(&W.l2[g1 >> 10]) + __ldg(&W.l1[g1]) + __popc(word & ((1u << (cell & 31)) - 1u));
  if (wi) {
    const uint4* sec = reinterpret_cast<const uint4*>(W.bitmap) + 2 * (size_t)g1;
    const uint4 a = __ldg(sec);
    rank += __popc(a.x) + (wi > 1 ? __popc(a.y) : 0) + (wi > 2 ? __popc(a.z) : 0) + (wi > 3 ? __popc(a.w) : 0);
    if (wi > 4) {
      const uint4 b = __ldg(sec + 1);
      rank += __popc(b.x) + (wi > 5 ? __popc(b.y) : 0) + (wi > 6 ? __popc(b.z) : 0);
    }
  }
  return rank;
}

// Same contract as rule_lookup_kernel, for an INPUT level whose rows are in ascending flat order and
// indexed by the bitmap/prefix workspace of the strided conv that produced it (no hash probes).
// The kernel offsets of one (kz, ky) LINE address x-consecutive cells, i.e. consecutive bitmap bits, and rows are
// numbered in flat order, so their ranks are consecutive too: one bitmap word (two when the run straddles a word)
// answers "which of the line's neighbours exist", and only if one does, ONE prefix lookup ranks all of them
// (rank of the next cell = rank + bit of this one) -- 9 line lookups per output row instead of 27 cell lookups
// (5 instead of 13 for the mirrored SubM walk).
__global__ void __launch_bounds__(256) rule_lookup_rank_kernel(ConvWs Win, const int4* __restrict__ out_idx,
                                                               const int* __restrict__ n_out, int out_cap,
                                                               ConvGeom G, int identity_kk, int mirror_kv,
                                                               int* __restrict__ nbr, int nbr_stride) {
  const int n = min(*n_out, out_cap);
  const int KX = G.ks[2];
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    const int4 c = out_idx[o];
    const int x0 = c.w * G.stride[2] - G.pad[2];
    int kk0 = 0;
    for (int kz = 0; kz < G.ks[0]; kz++) {
      const int z = c.y * G.stride[0] - G.pad[0] + kz * G.dil[0];
      for (int ky = 0; ky < G.ks[1]; ky++, kk0 += KX) {
        if (mirror_kv && kk0 > identity_kk) continue;  // whole line written by the mirrors of earlier offsets
        const int y = c.z * G.stride[1] - G.pad[1] + ky * G.dil[1];
        const bool zy_ok = z >= 0 && z < G.in_shape[0] && y >= 0 && y < G.in_shape[1];
        const unsigned int row_base =
            (unsigned int)((((size_t)c.x * G.in_shape[0] + z) * G.in_shape[1] + y) * G.in_shape[2]);
        // pass 1: presence bits of the line's cells; pass 2 (only if any): ranks
        unsigned int bits = 0u, w_cur = 0u, wq_cur = 0xFFFFFFFFu, first_word = 0u;
        int first = -1;
        for (int kx = 0; kx < KX; kx++) {
          const int x = x0 + kx * G.dil[2];
          if (!zy_ok || x < 0 || x >= G.in_shape[2]) continue;
          const unsigned int cell = row_base + (unsigned int)x, wq = cell >> 5;
          if (wq != wq_cur) {
            w_cur = __ldg(&Win.bitmap[wq]);
            wq_cur = wq;
          }
          if (w_cur & (1u << (cell & 31))) {
            bits |= 1u << kx;
            if (first < 0) {
              first = kx;
              first_word = w_cur;
            }
          }
        }
        int rank = 0;
        if (first >= 0) rank = rank_before(Win, row_base + (unsigned int)(x0 + first * G.dil[2]), first_word);
        for (int kx = 0; kx < KX; kx++) {
          const int kk = kk0 + kx;
          if (mirror_kv && kk > identity_kk) break;
          int r = -1;
          if (kk == identity_kk) {
            r = o;
          } else if (bits & (1u << kx)) {
            // cells between two present neighbours of a line are x-consecutive only for dilation 1; with a dilated
            // kernel the cells skipped in between may be active, so every present cell is ranked on its own
            if (G.dil[2] == 1 || kx == first) {
              r = rank;
            } else {
              const unsigned int cell = row_base + (unsigned int)(x0 + kx * G.dil[2]);
              r = rank_before(Win, cell, __ldg(&Win.bitmap[cell >> 5]));
            }
            if (r >= Win.cap) r = -1;  // overflow rows do not exist (see rank_of_cell)
            if (mirror_kv && r >= 0) nbr[(size_t)(mirror_kv - 1 - kk) * nbr_stride + r] = o;
          }
          if (bits & (1u << kx)) rank++;
          nbr[(size_t)kk * nbr_stride + o] = r;
        }
      }
    }
  }
}

inline void fill_geom(ConvGeom& G, const int* shape, const int* ks, const int* stride, const int* pad,
                      const int* dil) {
  for (int d = 0; d < 3; d++) {
    G.ks[d] = ks[d];
    G.stride[d] = stride[d];
    G.pad[d] = pad[d];
    G.dil[d] = dil[d];
    G.in_shape[d] = shape[d];
    G.out_shape[d] = (shape[d] + 2 * pad[d] - dil[d] * (ks[d] - 1) - 1) / stride[d] + 1;
  }
  G.KV = ks[0] * ks[1] * ks[2];
}

inline int row_grid(int cap_rows) {
  int blocks = ceil_div(cap_rows > 0 ? cap_rows : 1, 256);
  return blocks < kNumSMs * 8 ? blocks : kNumSMs * 8;
}

}  // namespace
}  // namespace v3d

using namespace v3d;

extern "C" size_t v3d_site_table_bytes(int capacity_rows) {
  if (capacity_rows < 0) return 0;
  return table_layout(nullptr, capacity_rows).total;
}

extern "C" int v3d_site_table_init(void* table, size_t table_bytes, int capacity_rows, v3d_stream_t stream) {
  if (!table || capacity_rows < 0) return V3D_ERR_INVALID_ARGUMENT;
  SiteTable T = table_layout(table, capacity_rows);
  if (table_bytes < T.total) return V3D_ERR_WORKSPACE_TOO_SMALL;
  V3D_CUDA_TRY(cudaMemsetAsync(table, 0, T.total, as_stream(stream)));  // epoch 0 = never written
  return V3D_OK;
}

extern "C" int v3d_site_table_build(void* table, const int* indices, const int* n_rows, int capacity_rows,
                                    const int* shape_host, v3d_stream_t stream) {
  if (!table || !indices || !n_rows || !shape_host || capacity_rows < 0) return V3D_ERR_INVALID_ARGUMENT;
  SiteTable T = table_layout(table, capacity_rows);
  Shape3 s{{shape_host[0], shape_host[1], shape_host[2]}};
  cudaStream_t st = as_stream(stream);
  table_bump_kernel<<<1, 1, 0, st>>>(T.hdr);
  table_build_kernel<<<row_grid(capacity_rows), 256, 0, st>>>(T, reinterpret_cast<const int4*>(indices),
                                                            n_rows, capacity_rows, s);
  return check_launch();
}

extern "C" int v3d_rulebook_subm(const void* table, const int* indices, const int* n_rows,
                                 int capacity_rows, const int* shape_host, const int* ksize_host,
                                 const int* dilation_host, int* nbr, int nbr_stride, v3d_stream_t stream) {
  if (!table || !indices || !n_rows || !shape_host || !ksize_host || !dilation_host || !nbr)
    return V3D_ERR_INVALID_ARGUMENT;
  if (nbr_stride < capacity_rows) return V3D_ERR_INVALID_ARGUMENT;
  for (int d = 0; d < 3; d++)
    if (ksize_host[d] <= 0 || (ksize_host[d] & 1) == 0) return V3D_ERR_INVALID_ARGUMENT;  // SubM: odd
  SiteTable T = table_layout(const_cast<void*>(table), capacity_rows);
  ConvGeom G;
  int one[3] = {1, 1, 1};
  int pad[3] = {ksize_host[0] / 2 * dilation_host[0], ksize_host[1] / 2 * dilation_host[1],
                ksize_host[2] / 2 * dilation_host[2]};
  fill_geom(G, shape_host, ksize_host, one, pad, dilation_host);
  if (G.KV > 65535) return V3D_ERR_INVALID_ARGUMENT;
  const int centre = ((ksize_host[0] / 2) * ksize_host[1] + ksize_host[1] / 2) * ksize_host[2] +
                     ksize_host[2] / 2;
  dim3 grid(row_grid(capacity_rows));
  // entries above the centre offset are produced as mirrors of the ones below it: pre-fill them with -1
  V3D_CUDA_TRY(cudaMemsetAsync(nbr + (size_t)(centre + 1) * nbr_stride, 0xFF,
                               sizeof(int) * (size_t)(G.KV - 1 - centre) * nbr_stride, as_stream(stream)));
  rule_lookup_kernel<<<grid, 256, 0, as_stream(stream)>>>(T, reinterpret_cast<const int4*>(indices), n_rows,
                                                          capacity_rows, G, centre, G.KV, nbr, nbr_stride);
  return check_launch();
}

// SubM on a level produced by a strided conv: `level_index` is that conv's workspace (bitmap + prefix).
extern "C" int v3d_rulebook_subm_ranked(const void* level_index, int B, int index_capacity, const int* indices,
                                        const int* n_rows, int capacity_rows, const int* shape_host,
                                        const int* ksize_host, const int* dilation_host, int* nbr, int nbr_stride,
                                        v3d_stream_t stream) {
  if (!level_index || !indices || !n_rows || !shape_host || !ksize_host || !dilation_host || !nbr)
    return V3D_ERR_INVALID_ARGUMENT;
  if (nbr_stride < capacity_rows || B <= 0) return V3D_ERR_INVALID_ARGUMENT;
  for (int d = 0; d < 3; d++)
    if (ksize_host[d] <= 0 || (ksize_host[d] & 1) == 0) return V3D_ERR_INVALID_ARGUMENT;
  ConvGeom G;
  int one[3] = {1, 1, 1};
  int pad[3] = {ksize_host[0] / 2 * dilation_host[0], ksize_host[1] / 2 * dilation_host[1],
                ksize_host[2] / 2 * dilation_host[2]};
  fill_geom(G, shape_host, ksize_host, one, pad, dilation_host);
  if (G.KV > 65535) return V3D_ERR_INVALID_ARGUMENT;
  const unsigned long long cells = (unsigned long long)B * shape_host[0] * shape_host[1] * shape_host[2];
  ConvWs Win = conv_layout(const_cast<void*>(level_index), cells, index_capacity);
  const int centre = ((ksize_host[0] / 2) * ksize_host[1] + ksize_host[1] / 2) * ksize_host[2] + ksize_host[2] / 2;
  V3D_CUDA_TRY(cudaMemsetAsync(nbr + (size_t)(centre + 1) * nbr_stride, 0xFF,
                               sizeof(int) * (size_t)(G.KV - 1 - centre) * nbr_stride, as_stream(stream)));
  rule_lookup_rank_kernel<<<dim3(row_grid(capacity_rows)), 256, 0, as_stream(stream)>>>(
      Win, reinterpret_cast<const int4*>(indices), n_rows, capacity_rows, G, centre, G.KV, nbr, nbr_stride);
  return check_launch();
}

extern "C" void v3d_conv_out_shape(const int* shape, const int* ksize, const int* stride, const int* pad,
                                   const int* dilation, int* out_shape) {
  ConvGeom G;
  fill_geom(G, shape, ksize, stride, pad, dilation);
  for (int d = 0; d < 3; d++) out_shape[d] = G.out_shape[d];
}

extern "C" size_t v3d_rulebook_conv_workspace_bytes(int B, const int* out_shape_host, int capacity_rows,
                                                    int kernel_volume) {
  (void)kernel_volume;
  if (B <= 0 || !out_shape_host || capacity_rows < 0) return 0;
  unsigned long long cells = (unsigned long long)B * out_shape_host[0] * out_shape_host[1] * out_shape_host[2];
  return conv_layout(nullptr, cells, capacity_rows).total;
}

static int rulebook_conv_impl(const void* in_table, const void* in_level_index, int in_index_capacity,
                              const int* indices, const int* n_rows, int capacity_rows, int B,
                              const int* shape_host, const int* ksize_host, const int* stride_host,
                              const int* pad_host, const int* dilation_host, int* out_indices, int* n_out,
                              int out_capacity, int* nbr, int nbr_stride, void* workspace, size_t workspace_bytes,
                              v3d_stream_t stream) {
  if ((!in_table && !in_level_index) || !indices || !n_rows || !shape_host || !ksize_host || !stride_host || !pad_host ||
      !dilation_host || !out_indices || !n_out || !nbr || !workspace)
    return V3D_ERR_INVALID_ARGUMENT;
  if (B <= 0 || out_capacity <= 0 || nbr_stride < out_capacity) return V3D_ERR_INVALID_ARGUMENT;
  ConvGeom G;
  fill_geom(G, shape_host, ksize_host, stride_host, pad_host, dilation_host);
  for (int d = 0; d < 3; d++)
    if (G.out_shape[d] <= 0 || stride_host[d] <= 0) return V3D_ERR_INVALID_ARGUMENT;
  if (G.KV > 65535) return V3D_ERR_INVALID_ARGUMENT;
  unsigned long long cells = (unsigned long long)B * G.out_shape[0] * G.out_shape[1] * G.out_shape[2];
  if (cells >= (1ull << 32)) return V3D_ERR_INVALID_ARGUMENT;  // cell ids are 32-bit
  ConvWs W = conv_layout(workspace, cells, out_capacity);
  if (workspace_bytes < W.total) return V3D_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = as_stream(stream);
  V3D_CUDA_TRY(cudaMemsetAsync(workspace, 0, W.zero_bytes, st));
  conv_mark_kernel<<<dim3(row_grid(capacity_rows)), 256, 0, st>>>(
      reinterpret_cast<const int4*>(indices), n_rows, capacity_rows, G, W);
  conv_scan_l1_kernel<<<(unsigned int)W.n_l2, 1024, 0, st>>>(W);
  conv_scan_l2_kernel<<<1, 1024, 0, st>>>(W, n_out);
  conv_rank_kernel<<<row_grid((int)(W.n_words < (1u << 30) ? W.n_words : (1u << 30))), 256, 0, st>>>(W, G, out_capacity,
                                                          reinterpret_cast<int4*>(out_indices));
  if (in_level_index) {
    const unsigned long long in_cells = (unsigned long long)B * shape_host[0] * shape_host[1] * shape_host[2];
    ConvWs Win = conv_layout(const_cast<void*>(in_level_index), in_cells, in_index_capacity);
    rule_lookup_rank_kernel<<<dim3(row_grid(out_capacity)), 256, 0, st>>>(
        Win, reinterpret_cast<const int4*>(out_indices), n_out, out_capacity, G, -1, 0, nbr, nbr_stride);
  } else {
    SiteTable T = table_layout(const_cast<void*>(in_table), capacity_rows);
    rule_lookup_kernel<<<dim3(row_grid(out_capacity)), 256, 0, st>>>(
        T, reinterpret_cast<const int4*>(out_indices), n_out, out_capacity, G, -1, 0, nbr, nbr_stride);
  }
  return check_launch();
}

extern "C" int v3d_rulebook_conv(const void* in_table, const int* indices, const int* n_rows,
                                 int capacity_rows, int B, const int* shape_host, const int* ksize_host,
                                 const int* stride_host, const int* pad_host, const int* dilation_host,
                                 int* out_indices, int* n_out, int out_capacity, int* nbr, int nbr_stride,
                                 void* workspace, size_t workspace_bytes, v3d_stream_t stream) {
  return rulebook_conv_impl(in_table, nullptr, 0, indices, n_rows, capacity_rows, B, shape_host, ksize_host,
                            stride_host, pad_host, dilation_host, out_indices, n_out, out_capacity, nbr, nbr_stride,
                            workspace, workspace_bytes, stream);
}

// Strided conv whose INPUT level was itself produced by a strided conv: `in_level_index` is that conv's
// workspace (its capacity = in_index_capacity); no hash table is involved.
extern "C" int v3d_rulebook_conv_ranked(const void* in_level_index, int in_index_capacity, const int* indices,
                                        const int* n_rows, int capacity_rows, int B, const int* shape_host,
                                        const int* ksize_host, const int* stride_host, const int* pad_host,
                                        const int* dilation_host, int* out_indices, int* n_out, int out_capacity,
                                        int* nbr, int nbr_stride, void* workspace, size_t workspace_bytes,
                                        v3d_stream_t stream) {
  return rulebook_conv_impl(nullptr, in_level_index, in_index_capacity, indices, n_rows, capacity_rows, B, shape_host,
                            ksize_host, stride_host, pad_host, dilation_host, out_indices, n_out, out_capacity, nbr,
                            nbr_stride, workspace, workspace_bytes, stream);
}
