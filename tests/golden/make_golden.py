"""Generate tests/golden/iou_nms_golden.npz from the REFERENCE's own compiled sources (oracle/_ref).

Run in the build container only (needs /root/reference to have been compiled by
`make -C oracle` + `python oracle/build_ref.py`):   python tests/golden/make_golden.py
Outputs are the reference's answers, not ours:
  iou_host / keep_host_*   : ref_C_cpu.so  = vision.cpp + box_iou_rotated_cpu.cpp + nms_rotated_cpu.cpp
                             (std::sort hull, NMS `>=`)                      -> pins oracle variant 0
  iou_nvcc / keep_nvcc_*   : box_iou_rotated_utils.h compiled as nvcc sees it (exchange-sort hull,
                             NMS `>`; oracle/ref_iou_shim.cpp)               -> pins oracle variant 1
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def cases():
    rng = np.random.default_rng(20260924)
    out = {}
    # known answers
    out["known"] = np.array([[0, 0, 2, 2, 0], [1, 1, 2, 2, 0], [0, 0, 2, 2, 45], [10, 10, 2, 2, 0]], np.float32)
    n = 160
    out["spread"] = np.stack([rng.uniform(0, 70.4, n), rng.uniform(-40, 40, n), rng.uniform(0.5, 3, n),
                              rng.uniform(1, 6, n), rng.uniform(-180, 180, n)], 1).astype(np.float32)
    out["cluster"] = np.stack([rng.uniform(0, 8, n), rng.uniform(-4, 4, n), np.full(n, 1.6), np.full(n, 3.9),
                               rng.uniform(0, np.pi, n)], 1).astype(np.float32)  # radians fed as degrees
    d = out["cluster"][:40].copy()
    d[::4, 4] = 0.0
    d[1::4, 4] = 90.0
    d[2::4, 4] = 45.0
    d[3::4] = d[2::4]            # exact duplicates
    d[5, 2] = 0.0                # zero-area box
    d[6, 2:4] = 1e-8             # area below the 1e-14 guard
    d[7, :2] = d[8, :2]          # concentric
    d[9] = d[10] + np.array([3.9, 0, 0, 0, 0], np.float32)  # touching edges
    out["degenerate"] = d.astype(np.float32)
    big = out["cluster"].copy()
    big[:, :2] += 28800.0        # magnitude the fp32 group-offset trick produces at 192 groups
    out["offset"] = big
    return out


def main():
    ref = oracle.ref_torch_module(cuda=False)
    rng = np.random.default_rng(7)
    blob = {}
    for name, b in cases().items():
        s = rng.random(len(b)).astype(np.float32)
        blob[name + "_boxes"] = b
        blob[name + "_scores"] = s
        tb, ts = torch.from_numpy(b), torch.from_numpy(s)
        blob[name + "_iou_host"] = ref.box_iou_rotated(tb, tb).numpy()
        blob[name + "_iou_nvcc"] = oracle.ref_shim_iou(b, b, nvcc_view=True)
        for thr in (0.01, 0.1, 0.5):
            tag = "%s_%03d" % (name, int(thr * 100))
            blob["keep_host_" + tag] = ref.nms_rotated(tb, ts, thr).numpy()
            blob["keep_nvcc_" + tag] = oracle.ref_shim_nms(b, s, thr, nvcc_view=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "iou_nms_golden.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), "bytes;", len(blob), "arrays")


if __name__ == "__main__":
    main()
