"""Training-side consumer of the rotated IoU: host mirror of the reference's
`ProposalTargetAssigner` (core/proposal_targets.py:11-90) -- M ground-truth boxes against the dense anchor grid
(70 400 anchors per class), detectron2's `Matcher` (ops/matcher.py:6-107) and the VoxelNet box encoding
(core/box_encode.py:27-38).

Same method names and tensor shapes as the reference class. Two execution paths with identical results:
  fused=True  (default, CUDA): `ops.match_anchors` -- IoU + max over gt + label strata in one kernel per class, the
               M x 70 400 matrix is never written;
  fused=False: the reference's expression sequence on top of an `iou_fn` (default `ops.box_iou_rotated`; the CPU
               tests inject the oracle's restatement of the reference CPU op, the product never imports it).
"""
import math

import torch
from torch import nn

from . import ops
from .second import make_anchors


def encode_boxes(boxes, anchors):
    """VoxelNet encode (core/box_encode.py:27-38); both (*, 7)."""
    g_xyz, g_wlh, g_yaw = boxes.split([3, 3, 1], -1)
    a_xyz, a_wlh, a_yaw = anchors.split([3, 3, 1], -1)
    diag = a_wlh[..., :2].norm(dim=-1, keepdim=True)
    norm = torch.cat((diag, diag, a_wlh[..., 2:3]), dim=-1)
    return torch.cat(((g_xyz - a_xyz) / norm, (g_wlh / a_wlh).log(), (g_yaw - a_yaw) % math.pi), dim=-1)


class Matcher:
    """ops/matcher.py:6-107 (labels per IoU stratum; low-quality matches optional)."""

    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        self.raw_thresholds = list(thresholds)
        th = list(thresholds)
        assert th[0] > 0
        th = [-float("inf")] + th + [float("inf")]
        assert all(lo <= hi for lo, hi in zip(th[:-1], th[1:])) and len(labels) == len(th) - 1
        assert all(l in (-1, 0, 1) for l in labels)
        self.thresholds, self.labels, self.allow_low_quality_matches = th, list(labels), allow_low_quality_matches

    def __call__(self, q):
        if q.numel() == 0:
            return (q.new_full((q.size(1),), 0, dtype=torch.int64),
                    q.new_full((q.size(1),), self.labels[0], dtype=torch.int8))
        vals, matches = q.max(dim=0)
        out = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for l, lo, hi in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            out[(vals >= lo) & (vals < hi)] = l
        if self.allow_low_quality_matches:
            best, _ = q.max(dim=1)
            out[torch.nonzero(q == best[:, None])[:, 1]] = 1
        return matches, out


class ProposalTargetAssignerB200(nn.Module):
    def __init__(self, cfg, device=None, fused=True, iou_fn=None):
        """cfg: SecondConfig whose ANCHORS entries carry 'iou_thresh' (core/config.py:22-48)."""
        super().__init__()
        self.cfg = cfg
        self.anchors = make_anchors(cfg)
        if device is not None:
            self.anchors = self.anchors.to(device)
        self.matchers = [Matcher(a["iou_thresh"], [0, -1, +1], getattr(cfg, "ALLOW_LOW_QUALITY_MATCHES", False))
                         for a in cfg.ANCHORS]
        self.fused = bool(fused) and not any(m.allow_low_quality_matches for m in self.matchers)
        self.iou_fn = iou_fn or ops.box_iou_rotated

    def compute_iou(self, boxes, anchors):
        return self.iou_fn(boxes[:, [0, 1, 3, 4, 6]].contiguous(), anchors[:, [0, 1, 3, 4, 6]].contiguous())

    def match_class_i(self, boxes, class_idx, full_idx, i):
        """proposal_targets.py:53-60."""
        class_mask = class_idx == i
        anchors = self.anchors[i].view(-1, self.cfg.BOX_DOF)
        if self.fused and anchors.is_cuda:
            m = self.matchers[i]
            matches, labels = ops.match_anchors(boxes[class_mask][:, [0, 1, 3, 4, 6]].contiguous(),
                                                anchors[:, [0, 1, 3, 4, 6]].contiguous(), m.raw_thresholds, m.labels)
        else:
            matches, labels = self.matchers[i](self.compute_iou(boxes[class_mask], anchors))
        if class_mask.any():
            matches = full_idx[class_mask][matches]
        return matches, labels

    def match_all_classes(self, boxes, class_idx, box_ignore=None):
        full_idx = torch.arange(boxes.shape[0], device=boxes.device)
        matches, labels = zip(*[self.match_class_i(boxes, class_idx, full_idx, i) for i in range(self.cfg.NUM_CLASSES)])
        shape = self.anchors.shape[:-1]
        return torch.stack(matches).view(shape), torch.stack(labels).view(shape)

    @staticmethod
    def get_cls_targets(G_cls):
        M_cls = G_cls.ne(-1)
        return G_cls.clamp_(min=0), M_cls

    def get_reg_targets(self, boxes, box_idx, G_cls):
        M_reg = G_cls == 1
        G_reg = encode_boxes(boxes[box_idx[M_reg]], self.anchors[M_reg])
        M_reg = M_reg.unsqueeze(-1)
        return torch.zeros_like(self.anchors).masked_scatter_(M_reg, G_reg), M_reg

    def forward(self, item):
        """proposal_targets.py:83-88. item: boxes (M, 7), class_idx (M,), box_ignore (M,)."""
        dev = self.anchors.device
        boxes, class_idx = item["boxes"].to(dev), item["class_idx"].to(dev)
        box_idx, G_cls = self.match_all_classes(boxes, class_idx, item.get("box_ignore"))
        G_cls, M_cls = self.get_cls_targets(G_cls)
        G_reg, M_reg = self.get_reg_targets(boxes, box_idx, G_cls)
        item.update(dict(G_cls=G_cls, G_reg=G_reg, M_cls=M_cls, M_reg=M_reg))
        return item
