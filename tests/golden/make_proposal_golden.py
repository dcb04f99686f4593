"""Golden vectors for the whole proposal stage, produced by the REFERENCE's own ProposalLayer.inference
(vision3d/detector/proposal.py:72-80) with the reference's anchor generator, box decode, batched_nms_rotated wrapper
and compiled CPU ops (oracle/_ref/ref_C_cpu.so):

    python tests/golden/make_proposal_golden.py     # build container only (needs /root/reference + oracle/_ref)

The reference modules are imported by file path; `vision3d.ops` / `vision3d._C` / `vision3d.core.box_encode` are bound
to the reference's own files, so nothing of this repo computes the answers (only the config values and the
deterministic input generator below come from here). Writes tests/golden/proposal_golden.npz: the layer's weights,
and (boxes, batch_idx, class_idx, scores) for a 3-class and a car-only configuration."""
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
from vision3d_b200 import second  # noqa: E402  (configuration values only)

REF = "/root/reference/vision3d"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def feature_map(B, C, ny, nx):
    """Deterministic pseudo-random fp32 map from integer arithmetic only (reproducible anywhere)."""
    i = np.arange(B * C * ny * nx, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(12345)) % np.uint64(1 << 32)
    h = (h ^ (h >> np.uint64(15))) * np.uint64(2246822519) % np.uint64(1 << 32)
    return ((h.astype(np.float64) / 2.0 ** 32 - 0.5) * 4.0).astype(np.float32).reshape(B, C, ny, nx)


def main():
    ref_c = oracle.ref_torch_module(cuda=False)
    pkg = types.ModuleType("vision3d")
    pkg.__path__ = []
    pkg._C = ref_c
    sys.modules["vision3d"] = pkg
    sys.modules["vision3d._C"] = ref_c
    iou_nms = load("ref_iou_nms", os.path.join(REF, "ops", "iou_nms.py"))
    ops = types.ModuleType("vision3d.ops")
    ops.batched_nms_rotated = iou_nms.batched_nms_rotated
    ops.sigmoid_focal_loss = None
    sys.modules["vision3d.ops"] = ops
    core = types.ModuleType("vision3d.core")
    core.__path__ = []
    sys.modules["vision3d.core"] = core
    load("vision3d.core.box_encode", os.path.join(REF, "core", "box_encode.py"))
    ag = load("ref_anchor_generator", os.path.join(REF, "core", "anchor_generator.py"))
    proposal = load("ref_proposal", os.path.join(REF, "detector", "proposal.py"))

    out = {}
    for tag, cfg in (("three", second.three_class_config()), ("car", second.car_config())):
        rcfg = types.SimpleNamespace(**{k: getattr(cfg, k) for k in ("NUM_YAW", "BOX_DOF", "ANCHORS", "VOXEL_SIZE",
                                                                    "GRID_BOUNDS", "STRIDES")},
                                     NUM_CLASSES=cfg.NUM_CLASSES,
                                     PROPOSAL=types.SimpleNamespace(C_IN=cfg.PROPOSAL_C_IN, TOPK=cfg.TOPK))
        torch.manual_seed(7)
        layer = proposal.ProposalLayer(rcfg).eval()
        with torch.no_grad():  # spread the scores so that the per-class thresholds keep a non-trivial subset
            layer.conv_cls.weight.mul_(12.0)
            layer.conv_cls.bias.fill_(-1.0)
            layer.conv_reg.weight.mul_(8.0)
        anchors = ag.AnchorGenerator(rcfg).anchors
        fmap = torch.from_numpy(feature_map(2, cfg.PROPOSAL_C_IN, anchors.shape[2], anchors.shape[3]))
        with torch.no_grad():
            boxes, b_idx, c_idx, scores = layer.inference(fmap, anchors)
        for k, v in layer.state_dict().items():
            out["%s_w_%s" % (tag, k)] = v.numpy()
        out[tag + "_boxes"], out[tag + "_scores"] = boxes.numpy(), scores.numpy()
        out[tag + "_batch_idx"], out[tag + "_class_idx"] = b_idx.numpy(), c_idx.numpy()
        print(tag, "kept", len(scores))
    np.savez_compressed(os.path.join(HERE, "proposal_golden.npz"), **out)


if __name__ == "__main__":
    main()
