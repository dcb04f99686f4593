"""TEST INFRASTRUCTURE ONLY: the SECOND inference path on the host CPU, op for op the way the reference
stack runs it, used (a) as the end-to-end checker of vision3d_b200.second.SecondEngine and (b) as the
reported CPU baseline / `bench.py --impl reference` arm.

  voxelize   : oracle.voxelize per frame + batch prefix          (spconv VoxelGenerator, preprocess.py:26-33)
  VFE        : features.sum(1) / occupancy                       (detector/layers.py:10-17)
  sparse CNN : per layer, rule book (oracle C++) then, per kernel offset, gather -> torch.mm -> scatter-add
               exactly like spconv v1.x indice_conv on CPU; BatchNorm1d(eval) + ReLU in torch
  dense      : torch index_put into zeros                        (sparse_cnn.py:128-133)
  RPN / head : the model's own torch modules on CPU
  NMS        : the REFERENCE's compiled CPU op (oracle/_ref/ref_C_cpu.so, `>=`) when present, else the
               oracle restatement of the CUDA variant; `nms_variant` says which

Takes its weights from a vision3d_b200.second.SecondB200 (plain tensors); never calls a CUDA kernel.
"""
import numpy as np
import torch

import oracle


def _bn_relu(x, bn):
    inv = torch.rsqrt(bn.running_var + bn.eps)
    return torch.relu((x - bn.running_mean) * inv * bn.weight + bn.bias)


def _indice_conv(feat, w, nbr, n_out):
    """spconv v1.x CPU indice_conv: for each kernel offset with pairs: gather rows, mm, scatter-add."""
    out = torch.zeros((n_out, w.shape[-1]), dtype=torch.float32)
    kv = nbr.shape[0]
    w = w.reshape(kv, w.shape[-2], w.shape[-1])
    for kk in range(kv):
        src = nbr[kk]
        o = torch.nonzero(src >= 0).squeeze(1)
        if o.numel() == 0:
            continue
        out.index_add_(0, o, feat[src[o].long()] @ w[kk])
    return out


def sparse_middle(model, feat, coords, batch_size, return_levels=False):
    """feat (N,4) f32 torch CPU, coords (N,4) int32 numpy [b,z,y,x] -> BEV (B, 128, 200, 176)."""
    shape = list(model.cnn.grid_shape)
    idx = np.ascontiguousarray(coords, np.int32)
    x = feat
    levels = [(idx, shape)]
    level_feats = [feat]   # x0 .. x4 as SparseCNNBase.forward names them (sparse_cnn.py:135-146)
    for blk in model.cnn.blocks:
        nbr_subm = None
        for seq in blk:
            conv, bn = seq[0], seq[1]
            w = conv.weight.detach().float()
            if conv.subm:
                if nbr_subm is None:
                    nbr_subm = torch.from_numpy(oracle.rulebook_subm(idx, shape, conv.kernel_size, conv.dilation))
                x = _indice_conv(x, w, nbr_subm, len(idx))
            else:
                out_idx, nbr, shape = oracle.rulebook_conv(idx, shape, conv.kernel_size, conv.stride, conv.padding,
                                                           conv.dilation)
                x = _indice_conv(x, w, torch.from_numpy(nbr), len(out_idx))
                idx = out_idx
                levels.append((idx, shape))
            x = _bn_relu(x, bn)
        level_feats.append(x)
    dense = torch.zeros((batch_size, x.shape[1], *shape), dtype=torch.float32)
    ii = torch.from_numpy(idx.astype(np.int64))
    dense[ii[:, 0], :, ii[:, 1], ii[:, 2], ii[:, 3]] = x
    bev = dense.view(batch_size, -1, shape[1], shape[2])
    if return_levels == "features":
        return bev, levels, level_feats
    if return_levels:
        return bev, levels, x
    return bev


def voxelize_batch(cfg, clouds, cap_policy=0):
    feats, coords, occ = [], [], []
    for i, p in enumerate(clouds):
        v, c, n = oracle.voxelize(p, cfg.VOXEL_SIZE, cfg.GRID_BOUNDS, cfg.MAX_OCCUPANCY, cfg.MAX_VOXELS, cap_policy)
        feats.append(v)
        coords.append(np.concatenate([np.full((len(c), 1), i, np.int32), c], 1))
        occ.append(n)
    return np.concatenate(feats), np.concatenate(coords), np.concatenate(occ)


_REF = {}


def _nms(bev, scores, thr):
    """Returns (keep, variant)."""
    if oracle.ref_available("ref_C_cpu.so"):
        if "m" not in _REF:
            _REF["m"] = oracle.ref_torch_module(cuda=False)
        return _REF["m"].nms_rotated(bev.contiguous(), scores.contiguous(), thr).numpy(), "reference_cpu(>=)"
    return oracle.nms_rotated(bev.numpy(), scores.numpy(), thr, 1), "oracle_nvcc_view(>)"


@torch.no_grad()
def infer(model, clouds, anchors, nms_variant=None, stages=None):
    """model: SecondB200 on CPU in eval mode. Returns (boxes, batch_idx, class_idx, scores) numpy."""
    from vision3d_b200.second import group_offsets
    cfg = model.cfg
    B = len(clouds)
    v, c, n = voxelize_batch(cfg, clouds)
    feat = torch.from_numpy(oracle.vfe_mean(v, n))
    bev = sparse_middle(model, feat, c, B)
    fmap = model.rpn(bev)
    boxes, scores = model.head.candidates(fmap, anchors)
    n_cls = cfg.NUM_CLASSES
    b_idx = torch.arange(B)[:, None, None].expand(-1, n_cls, cfg.TOPK).reshape(-1)
    c_idx = torch.arange(n_cls)[None, :, None].expand(B, -1, cfg.TOPK).reshape(-1)
    scores, boxes = scores.reshape(-1), boxes.reshape(-1, cfg.BOX_DOF)
    nms_in = group_offsets(boxes[:, [0, 1, 3, 4, 6]], c_idx + n_cls * b_idx)
    if nms_variant == "oracle":
        keep = oracle.nms_rotated(nms_in.numpy(), scores.numpy(), cfg.NMS_THRESH, 1)
    else:
        keep, _ = _nms(nms_in, scores, cfg.NMS_THRESH)
    keep = torch.from_numpy(np.asarray(keep, np.int64))
    if stages is not None:
        stages.update(bev=bev, fmap=fmap, cand_boxes=boxes, cand_scores=scores, nms_in=nms_in, keep=keep,
                      n_voxels=len(c))
    boxes, b_idx, c_idx, scores = boxes[keep], b_idx[keep], c_idx[keep], scores[keep]
    thr = scores.new_tensor([a["score_thresh"] for a in cfg.ANCHORS])
    m = scores > thr[c_idx]
    return boxes[m].numpy(), b_idx[m].numpy(), c_idx[m].numpy(), scores[m].numpy()
