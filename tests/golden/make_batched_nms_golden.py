"""Golden vectors for batched_nms_rotated produced by the REFERENCE's own Python wrapper
(vision3d/ops/iou_nms.py:90-134) running on the REFERENCE's own compiled CPU ops (oracle/_ref/ref_C_cpu.so):

    python tests/golden/make_batched_nms_golden.py     # build container only (needs /root/reference + oracle/_ref)

The wrapper is imported by file path with `vision3d._C` bound to that compiled module, so nothing of this repo is
on the path that produces the answers. Writes tests/golden/batched_nms_golden.npz."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from vision3d_b200 import synth  # noqa: E402  (box generator only)


def reference_wrapper():
    ref_c = oracle.ref_torch_module(cuda=False)
    pkg = types.ModuleType("vision3d")
    pkg.__path__ = []
    pkg._C = ref_c
    sys.modules["vision3d"] = pkg
    sys.modules["vision3d._C"] = ref_c
    spec = importlib.util.spec_from_file_location("ref_iou_nms", "/root/reference/vision3d/ops/iou_nms.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    m = reference_wrapper()
    out = {}
    for case, (seed, n) in enumerate(((0, 300), (1, 1600), (2, 64))):
        boxes, scores, idxs = synth.make_nms_boxes(seed, n)
        keep = m.batched_nms_rotated(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(idxs), 0.01)
        out["c%d_boxes" % case], out["c%d_scores" % case], out["c%d_idxs" % case] = boxes, scores, idxs
        out["c%d_keep" % case] = keep.numpy()
    np.savez_compressed(os.path.join(HERE, "batched_nms_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
