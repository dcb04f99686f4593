"""Torch-tensor front end of the C-ABI (include/v3d_b200.h). PyTorch is plumbing only: it owns the
device buffers and the stream; every computation below is a hand-written sm_100a kernel.

Two flavours per op:
  * `*_padded` / capacity-based calls never synchronise: data-dependent sizes stay in device int32
    counters, outputs are allocated at a static capacity (CUDA-graph friendly; used by second.py);
  * the reference-shaped wrappers (same names/arguments/return shapes as the ops the reference
    imports) read the counter back and slice, like the reference ops do.

CPU tensors are rejected loudly -- there is no CPU path (the reference dispatches on
`tensor.device().is_cuda()`, box_iou_rotated.h:23-31; here CUDA is the only branch).
"""
import ctypes

import torch

from . import _lib
from ._lib import V3DError, check, f3, i3

_I32 = torch.int32
_F32 = torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_f32(t, name, shape_last=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise V3DError("%s must be a CUDA tensor (vision3d_b200 has no CPU path)" % name)
    if t.dtype != _F32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if shape_last is not None and (t.dim() == 0 or t.shape[-1] != shape_last):
        raise V3DError("%s must have last dimension %d, got %s" % (name, shape_last, tuple(t.shape)))
    return t


def _cuda_i32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise V3DError("%s must be a CUDA tensor (vision3d_b200 has no CPU path)" % name)
    if t.dtype != _I32:
        t = t.int()
    if not t.is_contiguous():
        t = t.contiguous()
    return t


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return [int(x) for x in v]
    return [int(v)] * 3


# =============================================================================================
# a13 / a12: rotated IoU + NMS  (vision3d._C.box_iou_rotated / nms_rotated, csrc/vision.cpp:63-64)
# =============================================================================================
def box_iou_rotated(boxes1, boxes2):
    b1 = _cuda_f32(boxes1, "boxes1", 5)
    b2 = _cuda_f32(boxes2, "boxes2", 5)
    m, n = b1.shape[0], b2.shape[0]
    out = torch.empty((m, n), dtype=_F32, device=b1.device)
    with torch.cuda.device(b1.device):
        check(_lib.load().v3d_box_iou_rotated(b1.data_ptr(), m, b2.data_ptr(), n, out.data_ptr(), _stream()),
              "v3d_box_iou_rotated")
    return out


def nms_workspace(n, device):
    nbytes = _lib.load().v3d_nms_rotated_workspace_bytes(int(n))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def nms_rotated_padded(dets, scores, iou_threshold, workspace=None, keep=None, count=None, group_size=None):
    """No host sync. Returns (keep[N] int64, count[1] int32); keep[:count] is valid. `group_size`: the boxes are
    consecutive groups of that many boxes that cannot overlap across groups (batched_nms_rotated's offsets)."""
    d = _cuda_f32(dets, "dets", 5)
    s = _cuda_f32(scores, "scores")
    n = d.shape[0]
    if s.numel() != n:
        raise V3DError("dets and scores disagree on N")
    if keep is None:
        keep = torch.empty(max(n, 1), dtype=torch.int64, device=d.device)
    if count is None:
        count = torch.zeros(1, dtype=_I32, device=d.device)
    if workspace is None:
        workspace = nms_workspace(n, d.device)
    with torch.cuda.device(d.device):
        if group_size:
            check(_lib.load().v3d_nms_rotated_grouped(d.data_ptr(), s.data_ptr(), n, int(group_size),
                                                      float(iou_threshold), keep.data_ptr(), count.data_ptr(),
                                                      workspace.data_ptr(), workspace.numel(), _stream()),
                  "v3d_nms_rotated_grouped")
            return keep, count
        check(_lib.load().v3d_nms_rotated(d.data_ptr(), s.data_ptr(), n, float(iou_threshold), keep.data_ptr(),
                                          count.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()),
              "v3d_nms_rotated")
    return keep, count


def nms_rotated(dets, scores, iou_threshold):
    """Reference-shaped: int64 indices of kept boxes, descending score (nms_rotated_cuda.cu:131-133)."""
    if dets.numel() == 0:
        if not dets.is_cuda:
            raise V3DError("dets must be a CUDA tensor (vision3d_b200 has no CPU path)")
        return torch.empty((0,), dtype=torch.int64, device=dets.device)
    keep, count = nms_rotated_padded(dets, scores, iou_threshold)
    return keep[: int(count.item())]


def match_anchors(gt_bev, anchors_bev, thresholds, labels, want_vals=False):
    """Fused box_iou_rotated + Matcher.__call__ (core/proposal_targets.py:53-60, ops/matcher.py:86-107).
    gt_bev (M, 5), anchors_bev (N, 5) -> matches (N,) int64, labels (N,) int8 [, matched IoU (N,) f32].
    `thresholds` / `labels` as given to the reference Matcher (thresholds WITHOUT the +-inf sentinels)."""
    g = _cuda_f32(gt_bev, "gt_bev", 5)
    a = _cuda_f32(anchors_bev, "anchors_bev", 5)
    m, n = g.shape[0], a.shape[0]
    th = [-float("inf")] + [float(t) for t in thresholds] + [float("inf")]
    assert len(labels) == len(th) - 1
    matches = torch.zeros(n, dtype=torch.int64, device=a.device)
    if m == 0:  # matcher.py:73-84: no gt -> match 0, label of the lowest stratum
        return (matches, torch.full((n,), int(labels[0]), dtype=torch.int8, device=a.device)) + (
            (torch.zeros(n, dtype=_F32, device=a.device),) if want_vals else ())
    lab = torch.empty(n, dtype=torch.int8, device=a.device)
    vals = torch.empty(n, dtype=_F32, device=a.device) if want_vals else None
    k = len(labels)
    lo = (ctypes.c_float * k)(*th[:-1])
    hi = (ctypes.c_float * k)(*th[1:])
    lb = (ctypes.c_int * k)(*[int(v) for v in labels])
    with torch.cuda.device(a.device):
        check(_lib.load().v3d_match_anchors(g.data_ptr(), m, a.data_ptr(), n, k, lo, hi, lb, matches.data_ptr(),
                                            lab.data_ptr(), vals.data_ptr() if vals is not None else None, _stream()),
              "v3d_match_anchors")
    return (matches, lab, vals) if want_vals else (matches, lab)


# =============================================================================================
# a1 / a2: voxelize
# =============================================================================================
class Voxelizer:
    """Batched point->voxel generator (replaces spconv.utils.VoxelGenerator + the batching loop of
    vision3d/core/preprocess.py:26-33). Holds the persistent, epoch-tagged hash workspace."""

    def __init__(self, voxel_size, point_cloud_range, max_voxels, max_num_points, batch_capacity,
                 points_capacity, device="cuda", cap_policy=0):
        import numpy as np
        self.voxel_size = np.asarray(voxel_size, dtype=np.float32)
        self.range = np.asarray(point_cloud_range, dtype=np.float32)
        # spconv VoxelGenerator: grid = round((hi - lo) / voxel_size) in fp32
        self.grid = np.round((self.range[3:] - self.range[:3]) / self.voxel_size).astype(np.int64)
        self.max_voxels, self.max_pts = int(max_voxels), int(max_num_points)
        self.B, self.P = int(batch_capacity), int(points_capacity)
        self.cap_policy = int(cap_policy)
        self.device = torch.device(device)
        lib = _lib.load()
        nbytes = lib.v3d_voxelize_workspace_bytes(self.P, self.B)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._calls = 0
        self._init_ws()

    def _init_ws(self):
        with torch.cuda.device(self.device):
            check(_lib.load().v3d_voxelize_workspace_init(self.ws.data_ptr(), self.ws.numel(), self.P, self.B,
                                                          _stream()), "v3d_voxelize_workspace_init")
        self._calls = 0

    def alloc_outputs(self, C, with_mean=True):
        rows = self.B * self.max_voxels
        dev = self.device
        out = dict(voxels=torch.empty((rows, self.max_pts, C), dtype=_F32, device=dev),
                   coords=torch.empty((rows, 4), dtype=_I32, device=dev),
                   num_points=torch.empty((rows,), dtype=_I32, device=dev),
                   voxel_offsets=torch.zeros((self.B + 1,), dtype=_I32, device=dev))
        out["mean"] = torch.empty((rows, C), dtype=_F32, device=dev) if with_mean else None
        return out

    def run(self, points, frame_offsets, max_frame_points, out, batch_size=None):
        """points (total, C) f32 CUDA; frame_offsets (B+1) int32 CUDA. No host sync."""
        pts = _cuda_f32(points, "points")
        off = _cuda_i32(frame_offsets, "frame_offsets")
        B = self.B if batch_size is None else int(batch_size)
        if off.numel() != B + 1 or B > self.B:
            raise V3DError("frame_offsets must have batch_size+1 entries (<= capacity)")
        total, C = pts.shape
        mean = out.get("mean")
        with torch.cuda.device(self.device):
            check(_lib.load().v3d_voxelize_batch(
                pts.data_ptr(), total, int(max_frame_points), C, off.data_ptr(), B,
                f3(self.range[:3]), f3(self.voxel_size), i3(self.grid), self.max_pts, self.max_voxels,
                self.cap_policy, out["voxels"].data_ptr(), out["coords"].data_ptr(),
                out["num_points"].data_ptr(), out["voxel_offsets"].data_ptr(),
                mean.data_ptr() if mean is not None else None, self.ws.data_ptr(), self.ws.numel(), self.P,
                _stream()), "v3d_voxelize_batch")
        return out


# =============================================================================================
# a3-a6: site table, rule book, sparse conv, dense
# =============================================================================================
class SiteTable:
    """Device hash of the active sites of one resolution level (shared by the SubM layers with the
    same indice_key and the strided conv leaving the level)."""

    def __init__(self, capacity_rows, device):
        self.capacity = int(capacity_rows)
        self.device = torch.device(device)
        nbytes = _lib.load().v3d_site_table_bytes(self.capacity)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.load().v3d_site_table_init(self.buf.data_ptr(), self.buf.numel(), self.capacity, _stream()),
                  "v3d_site_table_init")

    def build(self, indices, n_rows, shape):
        with torch.cuda.device(self.device):
            check(_lib.load().v3d_site_table_build(self.buf.data_ptr(), indices.data_ptr(), n_rows.data_ptr(),
                                                   self.capacity, i3(shape), _stream()), "v3d_site_table_build")
        return self


def conv_out_shape(shape, ksize, stride, padding, dilation):
    out = (ctypes.c_int * 3)()
    _lib.load().v3d_conv_out_shape(i3(shape), i3(ksize), i3(stride), i3(padding), i3(dilation), out)
    return [int(x) for x in out]


def rulebook_subm(table, indices, n_rows, shape, ksize, dilation, nbr=None, capacity=None):
    """nbr (KV, capacity) int32: nbr[kk, o] = input row or -1. Outputs == inputs.
    `table` is a SiteTable (hash) or the ConvRulebookWorkspace of the strided conv that produced the level."""
    ks, dl = _triple(ksize), _triple(dilation)
    kv = ks[0] * ks[1] * ks[2]
    if isinstance(table, ConvRulebookWorkspace):
        cap = int(capacity if capacity is not None else table.capacity)
        if nbr is None:
            nbr = torch.empty((kv, cap), dtype=_I32, device=indices.device)
        assert list(shape) == table.shape
        with torch.cuda.device(indices.device):
            check(_lib.load().v3d_rulebook_subm_ranked(table.buf.data_ptr(), table.batch_size, table.capacity,
                                                       indices.data_ptr(), n_rows.data_ptr(), cap, i3(shape),
                                                       i3(ks), i3(dl), nbr.data_ptr(), nbr.shape[1], _stream()),
                  "v3d_rulebook_subm_ranked")
        return nbr
    cap = table.capacity
    if nbr is None:
        nbr = torch.empty((kv, cap), dtype=_I32, device=indices.device)
    with torch.cuda.device(indices.device):
        check(_lib.load().v3d_rulebook_subm(table.buf.data_ptr(), indices.data_ptr(), n_rows.data_ptr(), cap,
                                            i3(shape), i3(ks), i3(dl), nbr.data_ptr(), nbr.shape[1], _stream()),
              "v3d_rulebook_subm")
    return nbr


class ConvRulebookWorkspace:
    """Workspace of one strided rule-book build. After the call it doubles as the SITE INDEX of the output
    level (bitmap + popcount prefix; rows are in ascending flat order), see rulebook_subm(level_index=...)."""

    def __init__(self, batch_size, out_shape, out_capacity, kernel_volume, device):
        self.batch_size, self.shape, self.capacity = int(batch_size), [int(v) for v in out_shape], int(out_capacity)
        nbytes = _lib.load().v3d_rulebook_conv_workspace_bytes(int(batch_size), i3(out_shape), int(out_capacity),
                                                               int(kernel_volume))
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)


def rulebook_conv(in_table, indices, n_rows, batch_size, shape, ksize, stride, padding, dilation, out_capacity,
                  out_indices=None, n_out=None, nbr=None, workspace=None):
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    kv = ks[0] * ks[1] * ks[2]
    dev = indices.device
    out_shape = conv_out_shape(shape, ks, st, pd, dl)
    if out_indices is None:
        out_indices = torch.empty((out_capacity, 4), dtype=_I32, device=dev)
    if n_out is None:
        n_out = torch.zeros(1, dtype=_I32, device=dev)
    if nbr is None:
        nbr = torch.empty((kv, out_capacity), dtype=_I32, device=dev)
    if workspace is None:
        workspace = ConvRulebookWorkspace(batch_size, out_shape, out_capacity, kv, dev)
    if isinstance(in_table, ConvRulebookWorkspace):  # input level indexed by the conv that produced it
        assert list(shape) == in_table.shape and int(batch_size) == in_table.batch_size
        with torch.cuda.device(dev):
            check(_lib.load().v3d_rulebook_conv_ranked(
                in_table.buf.data_ptr(), in_table.capacity, indices.data_ptr(), n_rows.data_ptr(),
                int(indices.shape[0]), int(batch_size), i3(shape), i3(ks), i3(st), i3(pd), i3(dl),
                out_indices.data_ptr(), n_out.data_ptr(), int(out_capacity), nbr.data_ptr(), nbr.shape[1],
                workspace.buf.data_ptr(), workspace.buf.numel(), _stream()), "v3d_rulebook_conv_ranked")
        return out_indices, n_out, nbr, out_shape
    with torch.cuda.device(dev):
        check(_lib.load().v3d_rulebook_conv(
            in_table.buf.data_ptr(), indices.data_ptr(), n_rows.data_ptr(), in_table.capacity, int(batch_size),
            i3(shape), i3(ks), i3(st), i3(pd), i3(dl), out_indices.data_ptr(), n_out.data_ptr(),
            int(out_capacity), nbr.data_ptr(), nbr.shape[1], workspace.buf.data_ptr(), workspace.buf.numel(),
            _stream()), "v3d_rulebook_conv")
    return out_indices, n_out, nbr, out_shape


class PreparedWeights:
    """Per-layer weight image for the tcgen05 path (bf16 3-term split, K-major, 128B-swizzled), built once.
    `.buf` is None when the shape is only supported by the exact-fp32 SIMT path (e.g. Cin = 4)."""

    def __init__(self, weight):
        self.cin, self.cout = int(weight.shape[-2]), int(weight.shape[-1])
        self.kv = weight.numel() // (self.cin * self.cout)
        w = weight.detach().reshape(self.kv, self.cin, self.cout).contiguous().float()
        self.weight = w
        nbytes = _lib.load().v3d_sparse_conv_prepared_bytes(self.kv, self.cin, self.cout)
        self.buf = None
        if nbytes:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
            with torch.cuda.device(w.device):
                check(_lib.load().v3d_sparse_conv_prepare(w.data_ptr(), self.kv, self.cin, self.cout,
                                                          self.buf.data_ptr(), nbytes, _stream()),
                      "v3d_sparse_conv_prepare")


def _rows_of(feat, n_out, nbr):
    """Device int32 row count to convert when packing on the fly: every row of `feat` (the input row count is
    not an argument of the conv; rows past it are never gathered)."""
    return torch.full((1,), feat.shape[0], dtype=torch.int32, device=feat.device)


def pack_features(feat, n_rows, out=None, channels=None):
    """fp32 rows (cap, Csrc) -> packed rows (cap, 2*C) bf16 = [h1 | h2] with h1 = bf16(x), h2 = bf16(x - h1):
    the operand format of the tensor-core sparse conv (v3d_feature_pack). `channels` = C >= Csrc zero-pads."""
    cap, csrc = feat.shape
    c = int(channels or csrc)
    if out is None:
        out = torch.empty((cap, 2 * c), dtype=torch.bfloat16, device=feat.device)
    assert out.shape == (cap, 2 * c) and out.dtype == torch.bfloat16 and feat.dtype == _F32
    with torch.cuda.device(feat.device):
        check(_lib.load().v3d_feature_pack(feat.data_ptr(), n_rows.data_ptr(), int(cap), int(csrc), c, out.data_ptr(),
                                           _stream()), "v3d_feature_pack")
    return out


def unpack_features(packed):
    """packed rows (rows, 2*C) bf16 -> fp32 (rows, C) = h1 + h2 (tests / debugging)."""
    c = packed.shape[1] // 2
    return packed[:, :c].float() + packed[:, c:].float()


def sparse_conv(feat, weight, nbr, n_out, out_capacity, scale=None, shift=None, relu=False, out=None,
                out_packed=None, write_f32=True):
    """feat (rows, Cin); weight (KV, Cin, Cout) / spconv's (k0,k1,k2,Cin,Cout) -> exact-fp32 SIMT kernel,
    or a PreparedWeights -> tcgen05 bf16x3 kernel when the shape supports it; nbr (KV, stride).

    Tensor-core path: `feat` may be fp32 (packed on the fly, one extra kernel) or an already packed bf16
    (rows, 2*Cin) tensor; `out_packed` (out_capacity, 2*Cout) bf16 receives the packed result for the next
    layer, `write_f32=False` skips the fp32 rows (then `out_packed` is returned)."""
    if isinstance(weight, PreparedWeights):
        pw = weight
        if pw.buf is not None:
            if feat.dtype != torch.bfloat16:
                feat = pack_features(feat, _rows_of(feat, n_out, nbr))
            assert feat.shape[1] == 2 * pw.cin
            if write_f32 and out is None:
                out = torch.empty((out_capacity, pw.cout), dtype=_F32, device=feat.device)
            if not write_f32:
                out = None
                assert out_packed is not None
            with torch.cuda.device(feat.device):
                check(_lib.load().v3d_sparse_conv_fwd_tc(
                    feat.data_ptr(), pw.buf.data_ptr(), nbr.data_ptr(), nbr.shape[1], n_out.data_ptr(),
                    int(out_capacity), pw.kv, pw.cin, pw.cout, scale.data_ptr() if scale is not None else None,
                    shift.data_ptr() if shift is not None else None, int(bool(relu)),
                    out.data_ptr() if out is not None else None,
                    out_packed.data_ptr() if out_packed is not None else None, _stream()),
                    "v3d_sparse_conv_fwd_tc")
            return out if out is not None else out_packed
        weight = pw.weight
        assert feat.dtype == _F32, "the exact-fp32 path takes fp32 rows"
    cin, cout = weight.shape[-2], weight.shape[-1]
    kv = weight.numel() // (cin * cout)
    if out is None:
        out = torch.empty((out_capacity, cout), dtype=_F32, device=feat.device)
    with torch.cuda.device(feat.device):
        check(_lib.load().v3d_sparse_conv_fwd(
            feat.data_ptr(), weight.data_ptr(), nbr.data_ptr(), nbr.shape[1], n_out.data_ptr(), int(out_capacity),
            kv, cin, cout, scale.data_ptr() if scale is not None else None,
            shift.data_ptr() if shift is not None else None, int(bool(relu)), out.data_ptr(), _stream()),
            "v3d_sparse_conv_fwd")
    return out


def rulebook_invert(nbr, n_out, out_capacity, in_rows):
    """inv (KV, in_rows) int32 with inv[k][i] = o  <=>  nbr[k][o] = i (else -1): the rule table of the conv's backward
    w.r.t. its input (for a fixed offset the map is injective)."""
    kv = nbr.shape[0]
    inv = torch.empty((kv, max(int(in_rows), 1)), dtype=_I32, device=nbr.device)
    with torch.cuda.device(nbr.device):
        check(_lib.load().v3d_rulebook_invert(nbr.data_ptr(), nbr.shape[1], n_out.data_ptr(), int(out_capacity), kv,
                                              inv.data_ptr(), inv.shape[1], _stream()), "v3d_rulebook_invert")
    return inv


def sparse_conv_backward(feat, weight, nbr, n_out, n_out_host, grad_out, subm, need_input_grad=True, need_weight_grad=True):
    """Gradients of out = sparse_conv(feat, weight, nbr) (no folded BN / ReLU): (dX (rows, Cin) or None,
    dW (KV, Cin, Cout) or None). dX runs the forward kernel on the inverted rule table with transposed weights,
    dW is v3d_sparse_conv_bwd_weight."""
    cin, cout = weight.shape[-2], weight.shape[-1]
    kv = weight.numel() // (cin * cout)
    w = weight.reshape(kv, cin, cout)
    go = _cuda_f32(grad_out, "grad_out")
    n_in = feat.shape[0]
    cap = max(int(n_out_host), 1)
    if go.shape[0] < cap:
        go = torch.cat([go, go.new_zeros((cap - go.shape[0], cout))])
    g_feat = g_w = None
    if need_input_grad:
        if subm:  # nbr[k][o] = i  <=>  nbr[KV-1-k][i] = o
            inv = torch.flip(nbr[:, :n_in], dims=(0,)).contiguous() if nbr.shape[1] >= n_in else None
        else:
            inv = rulebook_invert(nbr, n_out, cap, n_in)
        n_in_dev = torch.full((1,), n_in, dtype=_I32, device=feat.device)
        g_feat = sparse_conv(go, w.transpose(1, 2).contiguous(), inv, n_in_dev, max(n_in, 1))[:n_in]
    if need_weight_grad:
        g_w = torch.empty((kv, cin, cout), dtype=_F32, device=feat.device)
        f = _cuda_f32(feat, "feat")
        with torch.cuda.device(feat.device):
            check(_lib.load().v3d_sparse_conv_bwd_weight(f.data_ptr(), go.data_ptr(), nbr.data_ptr(), nbr.shape[1],
                                                         n_out.data_ptr(), cap, kv, cin, cout, g_w.data_ptr(), _stream()),
                  "v3d_sparse_conv_bwd_weight")
    return g_feat, g_w


def sparse_to_dense(feat, indices, n_rows, capacity_rows, batch_size, shape, out=None, workspace=None):
    C = feat.shape[1]
    dev = feat.device
    if out is None:
        out = torch.empty((batch_size, C, shape[0], shape[1], shape[2]), dtype=_F32, device=dev)
    if workspace is None:
        nbytes = _lib.load().v3d_sparse_to_dense_workspace_bytes(int(batch_size), i3(shape))
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().v3d_sparse_to_dense(feat.data_ptr(), indices.data_ptr(), n_rows.data_ptr(),
                                              int(capacity_rows), C, int(batch_size), i3(shape), out.data_ptr(),
                                              workspace.data_ptr(), workspace.numel(), _stream()),
              "v3d_sparse_to_dense")
    return out


def sparse_to_bev_nhwc(feat, indices, n_rows, capacity_rows, batch_size, shape, out=None, workspace=None):
    """BEV map (B, C*D, H, W) in torch channels_last memory (= (B,H,W,C*D) storage), channel = c*D + d:
    the tensor `dense().view(B, C*D, H, W)` of the reference, laid out the way cuDNN wants it."""
    C = feat.shape[1]
    dev = feat.device
    if out is None:
        out = torch.empty((batch_size, C * shape[0], shape[1], shape[2]), dtype=_F32, device=dev,
                          memory_format=torch.channels_last)
    assert out.is_contiguous(memory_format=torch.channels_last)
    if workspace is None:
        nbytes = _lib.load().v3d_sparse_to_dense_workspace_bytes(int(batch_size), i3(shape))
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().v3d_sparse_to_dense_nhwc(feat.data_ptr(), indices.data_ptr(), n_rows.data_ptr(),
                                                   int(capacity_rows), C, int(batch_size), i3(shape),
                                                   out.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                                   _stream()), "v3d_sparse_to_dense_nhwc")
    return out


# =============================================================================================
# a7-a10: point ops (pointnet2_utils names and argument order)
# =============================================================================================
def furthest_point_sample(xyz, npoint):
    x = _cuda_f32(xyz, "xyz", 3)
    B, N, _ = x.shape
    out = torch.empty((B, int(npoint)), dtype=_I32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_fps(x.data_ptr(), B, N, int(npoint), out.data_ptr(), None, 0, _stream()), "v3d_fps")
    return out


def gather_operation(features, idx):
    f = _cuda_f32(features, "features")
    i = _cuda_i32(idx, "idx")
    B, C, N = f.shape
    m = i.shape[1]
    out = torch.empty((B, C, m), dtype=_F32, device=f.device)
    with torch.cuda.device(f.device):
        check(_lib.load().v3d_gather(f.data_ptr(), i.data_ptr(), B, C, N, m, out.data_ptr(), _stream()),
              "v3d_gather")
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    x = _cuda_f32(xyz, "xyz", 3)
    q = _cuda_f32(new_xyz, "new_xyz", 3)
    B, N, _ = x.shape
    M = q.shape[1]
    out = torch.empty((B, M, int(nsample)), dtype=_I32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_ball_query(x.data_ptr(), q.data_ptr(), B, N, M, float(radius), int(nsample),
                                         out.data_ptr(), _stream()), "v3d_ball_query")
    return out


def grouping_operation(features, idx):
    f = _cuda_f32(features, "features")
    i = _cuda_i32(idx, "idx")
    B, C, N = f.shape
    _, M, ns = i.shape
    out = torch.empty((B, C, M, ns), dtype=_F32, device=f.device)
    with torch.cuda.device(f.device):
        check(_lib.load().v3d_group(f.data_ptr(), i.data_ptr(), B, C, N, M, ns, out.data_ptr(), _stream()),
              "v3d_group")
    return out


def query_and_group(xyz, new_xyz, features, idx):
    """[xyz[idx] - new_xyz ; features[idx]] -> (B, 3+C, M, ns)  (QueryAndGroup, use_xyz=True)."""
    x = _cuda_f32(xyz, "xyz", 3)
    q = _cuda_f32(new_xyz, "new_xyz", 3)
    i = _cuda_i32(idx, "idx")
    B, N, _ = x.shape
    _, M, ns = i.shape
    if features is not None:
        f = _cuda_f32(features, "features")
        C = f.shape[1]
    else:
        f, C = None, 0
    out = torch.empty((B, 3 + C, M, ns), dtype=_F32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_query_and_group(x.data_ptr(), q.data_ptr(), f.data_ptr() if f is not None else None,
                                              i.data_ptr(), B, C, N, M, ns, out.data_ptr(), _stream()),
              "v3d_query_and_group")
    return out


def fps_keypoints(points, npoint, idx=None, keypoints=None):
    """PV_RCNN.sample_keypoints (detector/model.py:46-56) in one kernel: points (B, N, 3 or 4) -> (idx (B, m) int32,
    keypoints (B, m, 3)). Rows of 4 floats (x, y, z, intensity) are read with 128-bit loads; no xyz slice copy,
    no transposes, no separate gather_operation."""
    x = _cuda_f32(points, "points")
    if x.dim() != 3 or x.shape[-1] not in (3, 4):
        raise V3DError("points must be (B, N, 3) or (B, N, 4)")
    B, N, S = x.shape
    m = int(npoint)
    if idx is None:
        idx = torch.empty((B, m), dtype=_I32, device=x.device)
    if keypoints is None:
        keypoints = torch.empty((B, m, 3), dtype=_F32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_fps_keypoints(x.data_ptr(), S, B, N, m, idx.data_ptr(), keypoints.data_ptr(), _stream()),
              "v3d_fps_keypoints")
    return idx, keypoints


class BallQueryBounds:
    """Per-chunk bounding boxes of a source (32 consecutive rows per box) for the culled ball query. Build once per
    source with `.build(xyz, row_offsets)`, pass as `bounds=` to ball_query_msg. `max_rows_per_frame` bounds every
    frame's row count (for ragged sources: the level's total row capacity is always safe)."""

    def __init__(self, batch_size, max_rows_per_frame, device):
        self.B, self.max_rows = int(batch_size), int(max_rows_per_frame)
        nbytes = _lib.load().v3d_ball_query_bounds_bytes(self.B, self.max_rows)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def build(self, xyz, row_offsets=None):
        x = _cuda_f32(xyz, "xyz")
        N = x.shape[1] if row_offsets is None else 0
        with torch.cuda.device(x.device):
            check(_lib.load().v3d_ball_query_bounds(x.data_ptr(), x.shape[-1],
                                                    row_offsets.data_ptr() if row_offsets is not None else None, self.B,
                                                    N, self.max_rows, self.buf.data_ptr(), _stream()),
                  "v3d_ball_query_bounds")
        return self


def bev_gather(feature_map, keypoints, xy_offset, pixel_size, out=None, c_off=0):
    """BEVFeatureGatherer.forward (detector/layers.py:30-50). feature_map: (B, C, H, W) f32 in channels_last memory;
    keypoints (B, M, 3); xy_offset = GRID_BOUNDS[:2]; pixel_size = float32(VOXEL_SIZE[:2]) * STRIDES[-1].
    -> (B, C, M), or channels [c_off, c_off + C) of `out` (B, c_total, M)."""
    B, C, H, W = feature_map.shape
    if not feature_map.is_contiguous(memory_format=torch.channels_last) or feature_map.dtype != _F32 or not feature_map.is_cuda:
        raise V3DError("bev_gather: feature_map must be a channels_last f32 CUDA tensor")
    kp = _cuda_f32(keypoints, "keypoints", 3)
    M = kp.shape[1]
    if out is None:
        out = torch.empty((B, C, M), dtype=_F32, device=kp.device)
    with torch.cuda.device(kp.device):
        check(_lib.load().v3d_bev_gather(feature_map.data_ptr(), B, H, W, C, kp.data_ptr(), M, float(xy_offset[0]),
                                         float(xy_offset[1]), float(pixel_size[0]), float(pixel_size[1]), out.data_ptr(),
                                         out.shape[1], int(c_off), _stream()), "v3d_bev_gather")
    return out


class BallQuerySorted:
    """x-bucketed copy of a source in random row order + its chunk boxes, for the selecting ball query
    (v3d_ball_query_sort_x / _bounds / _msg_select): `.build(xyz, row_offsets)` once per step, then
    `.query(radii, nsamples, new_xyz, out)`; results are bit-identical to ball_query_msg on the original rows."""

    def __init__(self, batch_size, max_rows_per_frame, total_rows, x_range, device):
        self.B, self.max_rows, self.x_range = int(batch_size), int(max_rows_per_frame), (float(x_range[0]), float(x_range[1]))
        self.sorted = torch.zeros((int(total_rows), 4), dtype=_F32, device=device)
        self.ws = torch.zeros(_lib.load().v3d_ball_query_sort_workspace_bytes(self.B), dtype=torch.uint8, device=device)
        self.bounds = BallQueryBounds(self.B, self.max_rows, device)
        self.row_offsets, self.N = None, 0

    def build(self, xyz, row_offsets=None):
        x = _cuda_f32(xyz, "xyz")
        self.row_offsets, self.N = row_offsets, (x.shape[1] if row_offsets is None else 0)
        with torch.cuda.device(x.device):
            check(_lib.load().v3d_ball_query_sort_x(x.data_ptr(), x.shape[-1],
                                                    row_offsets.data_ptr() if row_offsets is not None else None, self.B,
                                                    self.N, self.max_rows, self.x_range[0], self.x_range[1],
                                                    self.sorted.data_ptr(), self.ws.data_ptr(), _stream()),
                  "v3d_ball_query_sort_x")
        view = self.sorted if row_offsets is not None else self.sorted[: self.B * self.N].view(self.B, self.N, 4)
        self.bounds.build(view, row_offsets)
        return self

    def query(self, radii, nsamples, new_xyz, out=None):
        q = _cuda_f32(new_xyz, "new_xyz", 3)
        B, M, _ = q.shape
        R = len(radii)
        if out is None:
            out = [torch.empty((B, M, int(ns)), dtype=_I32, device=q.device) for ns in nsamples]
        rad = (ctypes.c_float * R)(*[float(r) for r in radii])
        nsa = (ctypes.c_int * R)(*[int(n) for n in nsamples])
        ptrs = (ctypes.c_void_p * R)(*[o.data_ptr() for o in out])
        with torch.cuda.device(q.device):
            check(_lib.load().v3d_ball_query_msg_select(
                self.sorted.data_ptr(), self.row_offsets.data_ptr() if self.row_offsets is not None else None,
                self.bounds.buf.data_ptr(), self.max_rows, q.data_ptr(), B, self.N, M, R, rad, nsa, ptrs, _stream()),
                "v3d_ball_query_msg_select")
        return out


def ball_query_msg(radii, nsamples, xyz, new_xyz, row_offsets=None, out=None, bounds=None):
    """All ball queries of one PointnetSAModuleMSG in one pass. xyz: dense (B, N, S>=3) or, with `row_offsets`
    (B+1 int32 device), packed (rows, S) ragged sources (indices relative to the frame start). new_xyz (B, M, 3).
    Returns [idx_r (B, M, ns_r) int32 ...]."""
    x = _cuda_f32(xyz, "xyz")
    q = _cuda_f32(new_xyz, "new_xyz", 3)
    B, M, _ = q.shape
    S = x.shape[-1]
    N = x.shape[1] if row_offsets is None else 0
    R = len(radii)
    if out is None:
        out = [torch.empty((B, M, int(ns)), dtype=_I32, device=x.device) for ns in nsamples]
    rad = (ctypes.c_float * R)(*[float(r) for r in radii])
    nsa = (ctypes.c_int * R)(*[int(n) for n in nsamples])
    ptrs = (ctypes.c_void_p * R)(*[o.data_ptr() for o in out])
    with torch.cuda.device(x.device):
        if bounds is not None:  # culled variant: bit-identical results, skips chunks of 32 rows far from the query
            check(_lib.load().v3d_ball_query_msg_culled(
                x.data_ptr(), S, row_offsets.data_ptr() if row_offsets is not None else None, bounds.buf.data_ptr(),
                bounds.max_rows, q.data_ptr(), B, N, M, R, rad, nsa, ptrs, _stream()), "v3d_ball_query_msg_culled")
            return out
        check(_lib.load().v3d_ball_query_msg(x.data_ptr(), S, row_offsets.data_ptr() if row_offsets is not None else None,
                                             q.data_ptr(), B, N, M, R, rad, nsa, ptrs, _stream()), "v3d_ball_query_msg")
    return out


def query_and_group_rows(xyz, feat, new_xyz, idx, row_offsets=None, out=None):
    """QueryAndGroup(use_xyz=True) on row-major sources: xyz (B, N, S) / (rows, S), feat (B, N, C) / (rows, C) or None
    -> (B, 3 + C, M, ns). `feat` may be a column slice of a wider row-major tensor (e.g. points[..., 3:], the
    intensity the reference splits off at detector/model.py:69): it is read in place through its row stride."""
    x = _cuda_f32(xyz, "xyz")
    q = _cuda_f32(new_xyz, "new_xyz", 3)
    i = _cuda_i32(idx, "idx")
    B, M, ns = i.shape
    S = x.shape[-1]
    N = x.shape[1] if row_offsets is None else 0
    f, C, fs = None, 0, 0
    if feat is not None:
        if not feat.is_cuda or feat.dtype != _F32:
            raise V3DError("feat must be a CUDA float32 tensor")
        C = feat.shape[-1]
        rows_ok = feat.stride(-1) == 1 and (feat.dim() == 2 or feat.stride(0) == feat.shape[1] * feat.stride(1))
        f = feat if rows_ok else feat.contiguous()
        fs = f.stride(-2)
    if out is None:
        out = torch.empty((B, 3 + C, M, ns), dtype=_F32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_query_and_group_rows(x.data_ptr(), S, f.data_ptr() if f is not None else None, fs, C, N,
                                                   row_offsets.data_ptr() if row_offsets is not None else None,
                                                   q.data_ptr(), i.data_ptr(), B, M, ns, out.data_ptr(), _stream()),
              "v3d_query_and_group_rows")
    return out


def batch_offsets(indices, n_rows, batch_size, out=None):
    """offsets (B+1) int32: first row of every frame of a (b,z,y,x)-ordered index list; offsets[B] = n_rows.
    Device-side compute_pad_amounts (detector/sparse_cnn.py:107-116): no host sync."""
    ind = _cuda_i32(indices, "indices")
    if out is None:
        out = torch.empty((int(batch_size) + 1,), dtype=_I32, device=ind.device)
    with torch.cuda.device(ind.device):
        check(_lib.load().v3d_batch_offsets(ind.data_ptr(), n_rows.data_ptr(), int(ind.shape[0]), int(batch_size),
                                            out.data_ptr(), _stream()), "v3d_batch_offsets")
    return out


def to_global(indices, n_rows, voxel_size, offset, out=None):
    """Voxel indices (rows, 4) [b,z,y,x] -> metric xyz (rows, 3) (SparseCNNBase.to_global, sparse_cnn.py:91-105)."""
    ind = _cuda_i32(indices, "indices")
    if out is None:
        out = torch.empty((ind.shape[0], 3), dtype=_F32, device=ind.device)
    with torch.cuda.device(ind.device):
        check(_lib.load().v3d_to_global(ind.data_ptr(), n_rows.data_ptr(), int(ind.shape[0]), f3(voxel_size), f3(offset),
                                        out.data_ptr(), _stream()), "v3d_to_global")
    return out


def pad_batch(src, row_offsets, batch_size, frame_capacity, seed=0, out=None):
    """Ragged rows (rows, C) -> dense (B, frame_capacity, C), short frames filled with random duplicates of their
    own rows (pad_batch, sparse_cnn.py:118-126; pad_for_batch, core/preprocess.py:35-45). Same seed => same picks."""
    x = _cuda_f32(src, "src")
    C = x.shape[1]
    if out is None:
        out = torch.empty((int(batch_size), int(frame_capacity), C), dtype=_F32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_pad_batch(x.data_ptr(), C, row_offsets.data_ptr(), int(batch_size), int(frame_capacity),
                                        int(seed), out.data_ptr(), _stream()), "v3d_pad_batch")
    return out


# =============================================================================================
# 8f-2: fused set abstraction (grouping -> shared MLP -> max) on the tensor cores
# =============================================================================================
_SA_SHAPES = {(16, 16), (32, 32), (64, 64), (192, 96)}


class PreparedSaMlp:
    """Weights of ONE scale of a PointnetSAModuleMSG (two 1x1 conv layers, BatchNorm already folded:
    [(W1 (n1, 3 + C), b1 (n1,)), (W2 (n2, n1), b2 (n2,))], input channel order [xyz ; features] as QueryAndGroup
    emits it) in the layout v3d_sa_fused consumes: K order [features | pad to Cp | xyz | 0], widths padded to the
    supported (N1, N2), bf16 3-term split, swizzled chunk images."""

    def __init__(self, layers, C, Cp=None):
        (w1, b1), (w2, b2) = layers
        dev = w1.device
        n1, n2 = w1.shape[0], w2.shape[0]
        assert w1.shape[1] == 3 + C and w2.shape[1] == n1
        self.C, self.Cp = int(C), int(Cp) if Cp else -(-int(C) // 8) * 8   # Cp: channels of the packed source rows
        assert self.Cp >= C and self.Cp % 8 == 0
        self.N1 = max(16, -(-n1 // 16) * 16)
        self.N2 = n2
        if (self.N1, self.N2) not in _SA_SHAPES:
            raise V3DError("unsupported shared-MLP widths %s" % ((n1, n2),))
        self.nc1, self.nc2 = -(-(self.Cp + 3) // 64), -(-self.N1 // 64)
        k1 = torch.zeros((self.nc1 * 64, self.N1), dtype=_F32, device=dev)
        k1[:C, :n1] = w1[:, 3:].t()
        k1[self.Cp:self.Cp + 3, :n1] = w1[:, :3].t()
        k2 = torch.zeros((self.nc2 * 64, self.N2), dtype=_F32, device=dev)
        k2[:n1] = w2.t()
        self.b1 = torch.zeros(self.N1, dtype=_F32, device=dev)
        self.b1[:n1] = b1
        self.b2 = b2.detach().float().contiguous()
        lib = _lib.load()
        self.w1 = torch.empty(lib.v3d_sa_mlp_prepared_bytes(self.nc1, self.N1), dtype=torch.uint8, device=dev)
        self.w2 = torch.empty(lib.v3d_sa_mlp_prepared_bytes(self.nc2, self.N2), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.v3d_sa_mlp_prepare(k1.contiguous().data_ptr(), self.nc1, self.N1, self.w1.data_ptr(), self.w1.numel(),
                                         _stream()), "v3d_sa_mlp_prepare")
            check(lib.v3d_sa_mlp_prepare(k2.contiguous().data_ptr(), self.nc2, self.N2, self.w2.data_ptr(), self.w2.numel(),
                                         _stream()), "v3d_sa_mlp_prepare")
        torch.cuda.current_stream(dev).synchronize()  # k1 / k2 are temporaries


def pack_channel_major(feat, Cp=None, out=None):
    """fp32 feat (B, C, N) with arbitrary strides -> packed rows (B*N, 2*Cp) bf16 [h1 | h2] (Cp >= C, multiple of 8)."""
    if not feat.is_cuda or feat.dtype != _F32 or feat.dim() != 3:
        raise V3DError("feat must be a CUDA float32 (B, C, N) tensor")
    B, C, N = feat.shape
    Cp = int(Cp or -(-C // 8) * 8)
    if out is None:
        out = torch.empty((B * N, 2 * Cp), dtype=torch.bfloat16, device=feat.device)
    assert out.shape == (B * N, 2 * Cp) and out.dtype == torch.bfloat16
    sb, sc, sn = feat.stride()
    with torch.cuda.device(feat.device):
        check(_lib.load().v3d_pack_channel_major(feat.data_ptr(), sb, sc, sn, B, C, N, Cp, out.data_ptr(), _stream()),
              "v3d_pack_channel_major")
    return out


def sa_fused(feat_packed, xyz, new_xyz, idx, mlp, out, c_off=0, row_offsets=None):
    """One scale of a set-abstraction module, fused (v3d_sa_fused). feat_packed (rows, 2*Cp) bf16; xyz (B, N, S) or
    (rows, S) f32 (S >= 3); new_xyz (B, M, 3); idx (B, M, ns) int32, ns in {16, 32}; mlp: PreparedSaMlp;
    out (B, c_total, M) f32 receives channels [c_off, c_off + mlp.N2)."""
    x = _cuda_f32(xyz, "xyz")
    q = _cuda_f32(new_xyz, "new_xyz", 3)
    i = _cuda_i32(idx, "idx")
    B, M, ns = i.shape
    S = x.shape[-1]
    N = x.shape[1] if row_offsets is None else 0
    assert feat_packed.dtype == torch.bfloat16 and feat_packed.shape[-1] == 2 * mlp.Cp and feat_packed.is_contiguous()
    assert out.dtype == _F32 and out.is_contiguous() and out.shape[0] == B and out.shape[2] == M
    with torch.cuda.device(x.device):
        check(_lib.load().v3d_sa_fused(feat_packed.data_ptr(), mlp.Cp, x.data_ptr(), S,
                                       row_offsets.data_ptr() if row_offsets is not None else None, N, q.data_ptr(),
                                       i.data_ptr(), B, M, ns, mlp.w1.data_ptr(), mlp.b1.data_ptr(), mlp.N1,
                                       mlp.w2.data_ptr(), mlp.b2.data_ptr(), mlp.N2, out.data_ptr(), out.shape[1],
                                       int(c_off), _stream()), "v3d_sa_fused")
    return out


# =============================================================================================
# SECOND head glue (engine-internal)
# =============================================================================================
def second_head_decode(reg_out, anchors, anchor_idx, n_cls, n_yaw, topk, boxes=None, nms_in=None):
    """reg_out: conv_reg output (B, n_cls*7*n_yaw, ny, nx), any memory format; anchor_idx (B, n_cls, topk) int64.
    Returns boxes (N,7), nms_in (N,5) with the batched_nms_rotated group offsets applied."""
    B, _, ny, nx = reg_out.shape
    N = B * n_cls * topk
    dev = reg_out.device
    if boxes is None:
        boxes = torch.empty((N, 7), dtype=_F32, device=dev)
    if nms_in is None:
        nms_in = torch.empty((N, 5), dtype=_F32, device=dev)
    strides = (ctypes.c_longlong * 4)(*[int(v) for v in reg_out.stride()])
    with torch.cuda.device(dev):
        check(_lib.load().v3d_second_head_decode(reg_out.data_ptr(), strides, anchors.data_ptr(),
                                                 anchor_idx.data_ptr(), B, int(n_cls), int(n_yaw), ny, nx,
                                                 int(topk), boxes.data_ptr(), nms_in.data_ptr(), _stream()),
              "v3d_second_head_decode")
    return boxes, nms_in


def pack_detections(boxes, scores, keep, count, thr, n_cls, topk, counters_ptrs, n_counters, result):
    N = boxes.shape[0]
    with torch.cuda.device(boxes.device):
        check(_lib.load().v3d_pack_detections(boxes.data_ptr(), scores.data_ptr(), keep.data_ptr(), count.data_ptr(),
                                              thr.data_ptr(), N, int(n_cls), int(topk),
                                              counters_ptrs.data_ptr() if counters_ptrs is not None else None,
                                              int(n_counters), result.data_ptr(), _stream()), "v3d_pack_detections")
    return result


def head_cls_logits(fmap, weight, bias, out=None):
    """fmap (B, 128, ny, nx) in channels_last memory; weight (n_out, 128[,1,1]); -> logits (B, n_out, ny*nx)."""
    B, C, ny, nx = fmap.shape
    assert fmap.is_contiguous(memory_format=torch.channels_last) and fmap.dtype == _F32
    n_out = weight.shape[0]
    if out is None:
        out = torch.empty((B, n_out, ny * nx), dtype=_F32, device=fmap.device)
    with torch.cuda.device(fmap.device):
        check(_lib.load().v3d_head_cls_logits(fmap.data_ptr(), B, ny * nx, C, weight.data_ptr(),
                                              bias.data_ptr() if bias is not None else None, n_out, out.data_ptr(),
                                              _stream()), "v3d_head_cls_logits")
    return out


def topk_rows(values, k, out_values=None, out_index=None, workspace=None):
    """values (rows, L) f32 -> (top values (rows,k) descending, indices (rows,k) int64); ties -> lower index.
    `workspace` (uint8, v3d_topk_rows_workspace_bytes) enables the two-stage path for long rows."""
    rows, L = values.shape
    if out_values is None:
        out_values = torch.empty((rows, k), dtype=_F32, device=values.device)
    if out_index is None:
        out_index = torch.empty((rows, k), dtype=torch.int64, device=values.device)
    if workspace is None:
        workspace = torch.empty(_lib.load().v3d_topk_rows_workspace_bytes(rows, int(k)), dtype=torch.uint8,
                                device=values.device)
    with torch.cuda.device(values.device):
        check(_lib.load().v3d_topk_rows(values.data_ptr(), rows, L, int(k), out_values.data_ptr(),
                                        out_index.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()),
              "v3d_topk_rows")
    return out_values, out_index


def head_reg_gather(fmap, w_reg, b_reg, top_logits, anchor_idx, n_cls, n_yaw, topk, deltas, scores):
    B, C, ny, nx = fmap.shape
    with torch.cuda.device(fmap.device):
        check(_lib.load().v3d_head_reg_gather(fmap.data_ptr(), C, w_reg.data_ptr(),
                                              b_reg.data_ptr() if b_reg is not None else None, top_logits.data_ptr(),
                                              anchor_idx.data_ptr(), B, int(n_cls), int(n_yaw), ny, nx, int(topk),
                                              deltas.data_ptr(), scores.data_ptr(), _stream()), "v3d_head_reg_gather")
    return deltas, scores


def second_head_decode_compact(deltas, anchors, anchor_idx, B, n_cls, n_yaw, ny, nx, topk, boxes, nms_in):
    with torch.cuda.device(deltas.device):
        check(_lib.load().v3d_second_head_decode_compact(deltas.data_ptr(), anchors.data_ptr(), anchor_idx.data_ptr(),
                                                         int(B), int(n_cls), int(n_yaw), int(ny), int(nx), int(topk),
                                                         boxes.data_ptr(), nms_in.data_ptr(), _stream()),
              "v3d_second_head_decode_compact")
    return boxes, nms_in
