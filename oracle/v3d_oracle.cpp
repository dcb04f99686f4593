// TEST INFRASTRUCTURE ONLY -- CPU oracle for the vision3d per-frame LiDAR hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library, and only as the checker / reported CPU baseline. The
// product (vision3d_b200/) never calls into it and has no CPU fallback.
//
// Every function is a scalar restatement of the algorithm the reference runs for one
// row of SURVEY.md section 8(a); file:line citations are relative to /root/reference.
//
// Pin status
//   * orc_iou_* / orc_nms_rotated (a12-a14): PINNED. variant 0 restates the reference's
//     host build of vision3d/ops/csrc/box_iou_rotated/box_iou_rotated_utils.h and is
//     checked bit-for-bit against oracle/_ref (the reference's own CPU sources compiled
//     where they lie) and against tests/golden/iou_nms_*.npz generated from it.
//     variant 1 restates the same header as nvcc sees it (__CUDACC__ branch: exchange
//     sort, dist[] permuted with the points) and is checked against that header compiled
//     on the host with __CUDACC__ defined (oracle/ref_iou_shim.cpp).
//   * voxelize, rule book, sparse conv, dense, FPS, gather, ball query, grouping
//     (a1-a10): PARITY UNPINNED. The arithmetic lives in un-vendored, un-pinned
//     third-party packages (spconv fork jhultman/spconv, sshaoshuai/Pointnet2.PyTorch;
//     install.md:19-37) that are absent from /root/reference and from this image. These
//     functions restate the published algorithms of spconv v1.x and Pointnet2.PyTorch
//     and DEFINE the contract; the sparse conv is additionally cross-checked against a
//     dense torch conv3d in tests/.
//
// Build: g++ -O2 -ffp-contract=off (no -march): no FMA contraction, like the reference's
// x86-64 host build.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

// ---------------------------------------------------------------------------------------
// Rotated-rectangle IoU (a14). Follows box_iou_rotated_utils.h:56-340 operation by
// operation; all arithmetic is fp32 except the sites the reference evaluates in fp64.
// ---------------------------------------------------------------------------------------
struct V2 {
  float x, y;
};
static inline V2 sub(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
static inline float dotp(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }      // utils.h:46-49
static inline float crossp(V2 a, V2 b) { return a.x * b.y - b.x * a.y; }    // utils.h:51-54

struct Rect {
  float cx, cy, w, h, a;
};

// utils.h:56-74. theta and the trig are fp64 (:61-63), everything else fp32.
static void corners(const Rect& r, V2 out[4]) {
  double theta = r.a * 0.01745329251;
  float c2 = (float)std::cos(theta) * 0.5f;
  float s2 = (float)std::sin(theta) * 0.5f;
  out[0].x = r.cx - s2 * r.h - c2 * r.w;
  out[0].y = r.cy + c2 * r.h - s2 * r.w;
  out[1].x = r.cx + s2 * r.h - c2 * r.w;
  out[1].y = r.cy - c2 * r.h - s2 * r.w;
  out[2].x = 2 * r.cx - out[0].x;
  out[2].y = 2 * r.cy - out[0].y;
  out[3].x = 2 * r.cx - out[1].x;
  out[3].y = 2 * r.cy - out[1].y;
}

// utils.h:76-155: <=16 edge/edge crossings, then corners of 1 inside 2, then 2 inside 1.
static int clip_points(const V2 p1[4], const V2 p2[4], V2 out[24]) {
  V2 e1[4], e2[4];
  for (int i = 0; i < 4; i++) {
    e1[i] = sub(p1[(i + 1) % 4], p1[i]);
    e2[i] = sub(p2[(i + 1) % 4], p2[i]);
  }
  int n = 0;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) {
      float det = crossp(e2[j], e1[i]);
      if (std::fabs((double)det) <= 1e-14) continue;  // :97, fp64 compare
      V2 d = sub(p2[j], p1[i]);
      float t1 = crossp(e2[j], d) / det;
      float t2 = crossp(e1[i], d) / det;
      if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) {
        out[n].x = p1[i].x + e1[i].x * t1;
        out[n].y = p1[i].y + e1[i].y * t1;
        n++;
      }
    }
  }
  {
    const V2 AB = e2[0], DA = e2[3];
    float ABAB = dotp(AB, AB), ADAD = dotp(DA, DA);
    for (int i = 0; i < 4; i++) {
      V2 AP = sub(p1[i], p2[0]);
      float pAB = dotp(AP, AB);
      float pAD = -dotp(AP, DA);
      if (pAB >= 0 && pAD >= 0 && pAB <= ABAB && pAD <= ADAD) out[n++] = p1[i];
    }
  }
  {
    const V2 AB = e1[0], DA = e1[3];
    float ABAB = dotp(AB, AB), ADAD = dotp(DA, DA);
    for (int i = 0; i < 4; i++) {
      V2 AP = sub(p2[i], p1[0]);
      float pAB = dotp(AP, AB);
      float pAD = -dotp(AP, DA);
      if (pAB >= 0 && pAD >= 0 && pAB <= ABAB && pAD <= ADAD) out[n++] = p2[i];
    }
  }
  return n;
}

// utils.h:157-270 with shift_to_zero=true (the only way the IoU path calls it, :308).
// variant 0 = host build (std::sort + tolerance comparator :216-225; dist[] is NOT
// permuted by the sort, so step 4 reads pre-sort distances -- preserved).
// variant 1 = nvcc build (:197-214 exchange sort that swaps dist[] with the points).
static int hull(const V2 p[24], int n, V2 q[24], int variant) {
  int t = 0;
  for (int i = 1; i < n; i++)
    if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  const V2 origin = p[t];
  for (int i = 0; i < n; i++) q[i] = sub(p[i], origin);
  std::swap(q[0], q[t]);
  float dist[24];
  for (int i = 0; i < n; i++) dist[i] = dotp(q[i], q[i]);
  if (variant == 1) {
    for (int i = 1; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) {
        float cp = crossp(q[i], q[j]);
        if (((double)cp < -1e-6) || (std::fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
          std::swap(q[i], q[j]);
          std::swap(dist[i], dist[j]);
        }
      }
  } else {
    std::sort(q + 1, q + n, [](const V2& A, const V2& B) -> bool {
      float cp = crossp(A, B);
      if (std::fabs((double)cp) < 1e-6) return dotp(A, A) < dotp(B, B);
      return cp > 0;
    });
  }
  int k;
  for (k = 1; k < n; k++)
    if ((double)dist[k] > 1e-8) break;
  if (k == n) {
    q[0] = p[t];
    return 1;
  }
  q[1] = q[k];
  int m = 2;
  for (int i = k + 1; i < n; i++) {
    while (m > 1 && crossp(sub(q[i], q[m - 2]), sub(q[m - 1], q[m - 2])) >= 0) m--;
    q[m++] = q[i];
  }
  return m;
}

// utils.h:272-284
static float fan_area(const V2 q[24], int m) {
  if (m <= 2) return 0;
  float area = 0;
  for (int i = 1; i < m - 1; i++) area += std::fabs(crossp(sub(q[i], q[0]), sub(q[i + 1], q[0])));
  return (float)(area / 2.0);
}

// utils.h:313-340 (centre shift in fp64 :318-319) + :286-309.
static float iou_one(const float* b1, const float* b2, int variant) {
  double sx = (b1[0] + b2[0]) / 2.0;
  double sy = (b1[1] + b2[1]) / 2.0;
  Rect r1{(float)(b1[0] - sx), (float)(b1[1] - sy), b1[2], b1[3], b1[4]};
  Rect r2{(float)(b2[0] - sx), (float)(b2[1] - sy), b2[2], b2[3], b2[4]};
  const float a1 = r1.w * r1.h, a2 = r2.w * r2.h;
  if ((double)a1 < 1e-14 || (double)a2 < 1e-14) return 0.f;
  V2 p1[4], p2[4], raw[24], ord[24];
  corners(r1, p1);
  corners(r2, p2);
  int n = clip_points(p1, p2, raw);
  float inter = 0.0f;
  if (n > 2) {
    int m = hull(raw, n, ord, variant);
    inter = fan_area(ord, m);
  }
  return inter / (a1 + a2 - inter);
}

static inline int64_t flat3(int64_t b, int64_t z, int64_t y, int64_t x, const int* shp) {
  return ((b * shp[0] + z) * shp[1] + y) * (int64_t)shp[2] + x;
}

}  // namespace

// ---- a13/a14: pairwise IoU (box_iou_rotated_cpu.cpp:7-44) -----------------------------
ORC_API float orc_iou_single(const float* b1, const float* b2, int variant) {
  return iou_one(b1, b2, variant);
}

ORC_API void orc_box_iou_rotated(const float* b1, int M, const float* b2, int N, float* out,
                                 int variant) {
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N; j++) out[(int64_t)i * N + j] = iou_one(b1 + 5 * i, b2 + 5 * j, variant);
}

// ---- a12: greedy rotated NMS ----------------------------------------------------------
// variant 0: nms_rotated_cpu.cpp:7-57 (suppress when iou >= thr, host hull sort).
// variant 1: nms_rotated_cuda.cu:14-134 (bit j set when iou > thr, nvcc hull sort; greedy
//            scan over the score-sorted list).
// Score order: descending, ties broken by lower original index (torch's sort leaves tie
// order unspecified; the product uses the same stable rule).
ORC_API int orc_nms_rotated(const float* dets, const float* scores, int N, float thr, int variant,
                            int64_t* keep) {
  std::vector<int> order(N);
  for (int i = 0; i < N; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return scores[a] > scores[b]; });
  std::vector<uint8_t> dead(N, 0);
  int nk = 0;
  for (int a = 0; a < N; a++) {
    int i = order[a];
    if (dead[i]) continue;
    keep[nk++] = i;
    for (int b = a + 1; b < N; b++) {
      int j = order[b];
      if (dead[j]) continue;
      float v = iou_one(dets + 5 * i, dets + 5 * j, variant);
      bool hit = variant == 0 ? (v >= thr) : (v > thr);
      if (hit) dead[j] = 1;
    }
  }
  return nk;
}

// ---- a1: point -> voxel (spconv v1.x utils.VoxelGenerator.generate =
// points_to_voxel_3d_np; called at core/preprocess.py:17-31). PARITY UNPINNED. ------------
// points (N, C) f32 row-major, first three columns xyz. lo[3] = range minimum (xyz),
// vsize[3] (xyz), grid[3] = cells per axis (xyz). All of lo/vsize are fp32 as upstream's
// VoxelGenerator stores them; cell = floor((p - lo) / vsize) in fp32.
// Voxel id = order of first appearance; first max_pts points of a voxel kept in arrival
// order, rest dropped; voxels are zero padded; coords are (z, y, x) int32.
// cap_policy 0: stop at the first point that would open voxel number max_voxels
// (spconv v1.0/1.1 `break`); 1: skip only that point (`continue`, spconv >= 1.2).
ORC_API int orc_voxelize(const float* points, int N, int C, const float* lo, const float* vsize,
                         const int* grid, int max_pts, int max_voxels, int cap_policy,
                         float* voxels, int* coords, int* num_pts) {
  std::unordered_map<int64_t, int> cell2vox;
  cell2vox.reserve((size_t)N * 2);
  int nv = 0;
  for (int i = 0; i < N; i++) {
    int c[3];
    bool ok = true;
    for (int j = 0; j < 3; j++) {
      float f = std::floor((points[(int64_t)i * C + j] - lo[j]) / vsize[j]);
      if (!(f >= 0.0f) || !(f < (float)grid[j])) {
        ok = false;
        break;
      }
      c[j] = (int)f;
    }
    if (!ok) continue;
    int64_t key = ((int64_t)c[2] * grid[1] + c[1]) * grid[0] + c[0];
    auto it = cell2vox.find(key);
    int v;
    if (it == cell2vox.end()) {
      if (nv >= max_voxels) {
        if (cap_policy == 0) break;
        continue;
      }
      v = nv++;
      cell2vox.emplace(key, v);
      coords[3 * v + 0] = c[2];
      coords[3 * v + 1] = c[1];
      coords[3 * v + 2] = c[0];
      num_pts[v] = 0;
      std::memset(voxels + (int64_t)v * max_pts * C, 0, sizeof(float) * max_pts * C);
    } else {
      v = it->second;
    }
    int k = num_pts[v];
    if (k < max_pts) {
      std::memcpy(voxels + ((int64_t)v * max_pts + k) * C, points + (int64_t)i * C,
                  sizeof(float) * C);
      num_pts[v] = k + 1;
    }
  }
  return nv;
}

// ---- a2: VoxelFeatureExtractor (detector/layers.py:10-17): sum over slots / occupancy ----
ORC_API void orc_vfe_mean(const float* voxels, const int* num_pts, int M, int K, int C, float* out) {
  for (int v = 0; v < M; v++)
    for (int c = 0; c < C; c++) {
      float s = 0.f;
      for (int k = 0; k < K; k++) s += voxels[((int64_t)v * K + k) * C + c];
      out[(int64_t)v * C + c] = s / (float)num_pts[v];
    }
}

// ---- a4: rule book (spconv get_indice_pairs). PARITY UNPINNED. ---------------------------
// indices (N,4) int32 b,z,y,x. Kernel offset id kk = (kz*KS[1] + ky)*KS[2] + kx.
// Correlation convention of torch conv3d: out[o] += in[o*stride - pad + k*dil] * W[k].
// nbr is the output-stationary rule table: nbr[kk*n_out + o] = input row or -1.
//
// SubM (submanifold): outputs = inputs (same rows, same order); stride 1, pad = ks/2.
ORC_API void orc_rulebook_subm(const int* idx, int N, const int* shape, const int* ks,
                               const int* dil, int* nbr) {
  std::unordered_map<int64_t, int> tab;
  tab.reserve((size_t)N * 2);
  for (int i = 0; i < N; i++)
    tab.emplace(flat3(idx[4 * i], idx[4 * i + 1], idx[4 * i + 2], idx[4 * i + 3], shape), i);
  const int KV = ks[0] * ks[1] * ks[2];
  for (int o = 0; o < N; o++) {
    int kk = 0;
    for (int kz = 0; kz < ks[0]; kz++)
      for (int ky = 0; ky < ks[1]; ky++)
        for (int kx = 0; kx < ks[2]; kx++, kk++) {
          int z = idx[4 * o + 1] + (kz - ks[0] / 2) * dil[0];
          int y = idx[4 * o + 2] + (ky - ks[1] / 2) * dil[1];
          int x = idx[4 * o + 3] + (kx - ks[2] / 2) * dil[2];
          int r = -1;
          if (z >= 0 && z < shape[0] && y >= 0 && y < shape[1] && x >= 0 && x < shape[2]) {
            auto it = tab.find(flat3(idx[4 * o], z, y, x, shape));
            if (it != tab.end()) r = it->second;
          }
          nbr[(int64_t)kk * N + o] = r;
        }
    (void)KV;
  }
}

// Strided sparse conv: out_shape[d] = (in + 2*pad - dil*(ks-1) - 1)/stride + 1. The
// output set is every cell that at least one active input reaches; rows are numbered in
// ascending flat (b,z,y,x) order (upstream's GPU path sorts its unique flat indices;
// its CPU path numbers first-come -- implementation-defined upstream, fixed here).
// Returns n_out; out_idx (n_out,4); nbr[kk*n_out_cap + o] with row stride n_out_cap.
ORC_API int orc_rulebook_conv(const int* idx, int N, const int* shape, const int* ks,
                              const int* stride, const int* pad, const int* dil, int* out_shape,
                              int* out_idx, int* nbr, int n_out_cap) {
  for (int d = 0; d < 3; d++)
    out_shape[d] = (shape[d] + 2 * pad[d] - dil[d] * (ks[d] - 1) - 1) / stride[d] + 1;
  const int KV = ks[0] * ks[1] * ks[2];
  struct Hit {
    int64_t key;
    int in_row, kk;
  };
  std::vector<Hit> hits;
  hits.reserve((size_t)N * KV);
  for (int i = 0; i < N; i++) {
    int kk = 0;
    for (int kz = 0; kz < ks[0]; kz++)
      for (int ky = 0; ky < ks[1]; ky++)
        for (int kx = 0; kx < ks[2]; kx++, kk++) {
          int nz = idx[4 * i + 1] + pad[0] - kz * dil[0];
          int ny = idx[4 * i + 2] + pad[1] - ky * dil[1];
          int nx = idx[4 * i + 3] + pad[2] - kx * dil[2];
          if (nz < 0 || ny < 0 || nx < 0) continue;
          if (nz % stride[0] || ny % stride[1] || nx % stride[2]) continue;
          int oz = nz / stride[0], oy = ny / stride[1], ox = nx / stride[2];
          if (oz >= out_shape[0] || oy >= out_shape[1] || ox >= out_shape[2]) continue;
          hits.push_back(Hit{flat3(idx[4 * i], oz, oy, ox, out_shape), i, kk});
        }
  }
  std::vector<int64_t> keys(hits.size());
  for (size_t h = 0; h < hits.size(); h++) keys[h] = hits[h].key;
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  int n_out = (int)keys.size();
  if (n_out > n_out_cap) return -n_out;
  for (int64_t e = 0; e < (int64_t)KV * n_out_cap; e++) nbr[e] = -1;
  for (int o = 0; o < n_out; o++) {
    int64_t k = keys[o];
    int x = (int)(k % out_shape[2]);
    k /= out_shape[2];
    int y = (int)(k % out_shape[1]);
    k /= out_shape[1];
    int z = (int)(k % out_shape[0]);
    k /= out_shape[0];
    out_idx[4 * o] = (int)k;
    out_idx[4 * o + 1] = z;
    out_idx[4 * o + 2] = y;
    out_idx[4 * o + 3] = x;
  }
  for (const Hit& h : hits) {
    int o = (int)(std::lower_bound(keys.begin(), keys.end(), h.key) - keys.begin());
    nbr[(int64_t)h.kk * n_out_cap + o] = h.in_row;
  }
  return n_out;
}

// ---- a5/a6: sparse conv forward (+ folded eval BatchNorm1d + ReLU) -----------------------
// feat (n_in, Cin), W (KV, Cin, Cout) [= spconv weight (k0,k1,k2,Cin,Cout) flattened],
// out (n_out, Cout). Accumulates in fp64 and rounds once, so it is the "true" value the
// <=1e-4 rel tolerance is measured against. scale/shift may be NULL (no BN), relu 0/1.
ORC_API void orc_sparse_conv(const float* feat, const float* W, const int* nbr, int nbr_stride,
                             int n_out, int KV, int Cin, int Cout, const float* scale,
                             const float* shift, int relu, float* out) {
  std::vector<double> acc(Cout);
  for (int o = 0; o < n_out; o++) {
    std::fill(acc.begin(), acc.end(), 0.0);
    for (int kk = 0; kk < KV; kk++) {
      int r = nbr[(int64_t)kk * nbr_stride + o];
      if (r < 0) continue;
      const float* f = feat + (int64_t)r * Cin;
      const float* w = W + (int64_t)kk * Cin * Cout;
      for (int ci = 0; ci < Cin; ci++)
        for (int co = 0; co < Cout; co++) acc[co] += (double)f[ci] * (double)w[ci * Cout + co];
    }
    for (int co = 0; co < Cout; co++) {
      double v = acc[co];
      if (scale) v = v * (double)scale[co] + (double)shift[co];
      if (relu && v < 0) v = 0;
      out[(int64_t)o * Cout + co] = (float)v;
    }
  }
}

// ---- a3: SparseConvTensor.dense() (detector/sparse_cnn.py:128-133) -----------------------
// (N,C) rows -> zero-filled (B,C,D,H,W).
ORC_API void orc_dense(const float* feat, const int* idx, int N, int C, int B, const int* shape,
                       float* out) {
  int64_t vol = (int64_t)shape[0] * shape[1] * shape[2];
  std::memset(out, 0, sizeof(float) * B * C * vol);
  for (int i = 0; i < N; i++) {
    int64_t cell = ((int64_t)idx[4 * i + 1] * shape[1] + idx[4 * i + 2]) * shape[2] + idx[4 * i + 3];
    for (int c = 0; c < C; c++) out[((int64_t)idx[4 * i] * C + c) * vol + cell] = feat[(int64_t)i * C + c];
  }
}

// ---- a7: furthest point sampling (pointnet2_utils.furthest_point_sample, called at
// detector/model.py:53). PARITY UNPINNED. Start at index 0; running min squared distance
// initialised to 1e10; d = (dx*dx + dy*dy) + dz*dz in fp32 without FMA; arg-max picks the
// LOWEST index among equal maxima (upstream's tie rule depends on its thread layout). -------
ORC_API void orc_fps(const float* xyz, int B, int N, int m, int* out) {
  std::vector<float> mind(N);
  for (int b = 0; b < B; b++) {
    const float* p = xyz + (int64_t)b * N * 3;
    std::fill(mind.begin(), mind.end(), 1e10f);
    int cur = 0;
    out[(int64_t)b * m] = 0;
    for (int j = 1; j < m; j++) {
      float cx = p[3 * cur], cy = p[3 * cur + 1], cz = p[3 * cur + 2];
      float best = -1.f;
      int besti = 0;
      for (int k = 0; k < N; k++) {
        float dx = p[3 * k] - cx, dy = p[3 * k + 1] - cy, dz = p[3 * k + 2] - cz;
        float d = (dx * dx + dy * dy) + dz * dz;
        float d2 = std::min(d, mind[k]);
        mind[k] = d2;
        if (d2 > best) {
          best = d2;
          besti = k;
        }
      }
      cur = besti;
      out[(int64_t)b * m + j] = cur;
    }
  }
}

// ---- a8: gather_operation (detector/model.py:54): out[b,c,j] = feat[b,c,idx[b,j]] --------
ORC_API void orc_gather(const float* feat, const int* idx, int B, int C, int N, int m, float* out) {
  for (int b = 0; b < B; b++)
    for (int c = 0; c < C; c++)
      for (int j = 0; j < m; j++)
        out[((int64_t)b * C + c) * m + j] = feat[((int64_t)b * C + c) * N + idx[(int64_t)b * m + j]];
}

// ---- a9: ball_query (inside PointnetSAModuleMSG; detector/model.py:39-43,64,
// roi_grid_pool.py:28-32,68). PARITY UNPINNED. First nsample sources in ascending index with
// d2 < r*r (strict, fp32, no FMA); unused slots repeat the first hit; no hit -> zeros. -------
ORC_API void orc_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M,
                            float radius, int nsample, int* out) {
  float r2 = radius * radius;
  for (int b = 0; b < B; b++)
    for (int q = 0; q < M; q++) {
      const float* c = new_xyz + ((int64_t)b * M + q) * 3;
      int* o = out + ((int64_t)b * M + q) * nsample;
      for (int l = 0; l < nsample; l++) o[l] = 0;
      int cnt = 0;
      for (int k = 0; k < N && cnt < nsample; k++) {
        const float* p = xyz + ((int64_t)b * N + k) * 3;
        float dx = c[0] - p[0], dy = c[1] - p[1], dz = c[2] - p[2];
        float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < r2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; l++) o[l] = k;
          o[cnt++] = k;
        }
      }
    }
}

// ---- a10: grouping_operation: out[b,c,j,l] = feat[b,c,idx[b,j,l]] -------------------------
ORC_API void orc_group(const float* feat, const int* idx, int B, int C, int N, int M, int ns,
                       float* out) {
  for (int b = 0; b < B; b++)
    for (int c = 0; c < C; c++)
      for (int64_t e = 0; e < (int64_t)M * ns; e++)
        out[((int64_t)b * C + c) * M * ns + e] =
            feat[((int64_t)b * C + c) * N + idx[(int64_t)b * M * ns + e]];
}

// QueryAndGroup(use_xyz=True): channels 0..2 = xyz[idx] - new_xyz, then C feature channels.
// xyz (B,N,3), new_xyz (B,M,3), feat (B,C,N) or NULL -> out (B, 3+C, M, ns).
ORC_API void orc_query_and_group(const float* xyz, const float* new_xyz, const float* feat,
                                 const int* idx, int B, int C, int N, int M, int ns, float* out) {
  int CT = 3 + (feat ? C : 0);
  for (int b = 0; b < B; b++)
    for (int q = 0; q < M; q++)
      for (int l = 0; l < ns; l++) {
        int k = idx[((int64_t)b * M + q) * ns + l];
        for (int d = 0; d < 3; d++)
          out[(((int64_t)b * CT + d) * M + q) * ns + l] =
              xyz[((int64_t)b * N + k) * 3 + d] - new_xyz[((int64_t)b * M + q) * 3 + d];
        if (feat)
          for (int c = 0; c < C; c++)
            out[(((int64_t)b * CT + 3 + c) * M + q) * ns + l] = feat[((int64_t)b * C + c) * N + k];
      }
}
