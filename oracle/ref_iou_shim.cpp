// TEST INFRASTRUCTURE ONLY. Thin extern "C" driver around the reference's OWN header
// vision3d/ops/csrc/box_iou_rotated/box_iou_rotated_utils.h, included from where it lies
// under /root/reference (see oracle/Makefile: -I$(REF)/vision3d/ops/csrc). Nothing from the
// reference is copied; this file only loops over pairs exactly like the reference drivers:
//   * box_iou_rotated_cpu.cpp:7-29   (double loop over M x N)
//   * nms_rotated_cpu.cpp:27-57      (greedy scan, `>=`)         -> REF_NVCC_VIEW undefined
//   * nms_rotated_cuda.cu:51-66,106-128 (mask bit when `>`, greedy over sorted list)
//                                                                -> REF_NVCC_VIEW defined
// Built twice into oracle/_ref/: libref_iou_host.so (header as g++ sees it) and
// libref_iou_nvccview.so (header as nvcc sees it: -D__CUDACC__ with the CUDA decorators
// defined away, so the exchange-sort hull branch at utils.h:197-214 is what compiles).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "box_iou_rotated/box_iou_rotated_utils.h"

extern "C" __attribute__((visibility("default"))) void ref_box_iou_rotated(const float* b1, int M,
                                                                           const float* b2, int N,
                                                                           float* out) {
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N; j++)
      out[(int64_t)i * N + j] = detectron2::single_box_iou_rotated<float>(b1 + 5 * i, b2 + 5 * j);
}

extern "C" __attribute__((visibility("default"))) int ref_nms_rotated(const float* dets,
                                                                      const float* scores, int N,
                                                                      float thr, int64_t* keep) {
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
      float v = detectron2::single_box_iou_rotated<float>(dets + 5 * i, dets + 5 * j);
#ifdef REF_NVCC_VIEW
      if (v > thr) dead[j] = 1;
#else
      if (v >= thr) dead[j] = 1;
#endif
    }
  }
  return nk;
}
