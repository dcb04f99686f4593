"""TEST INFRASTRUCTURE ONLY: ctypes front-end for the CPU oracle (oracle/v3d_oracle.cpp) and,
when built, for the reference's own sources compiled into oracle/_ref/.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. vision3d_b200 never imports it.
"""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_c_f = ctypes.POINTER(ctypes.c_float)
_c_i = ctypes.POINTER(ctypes.c_int)
_c_l = ctypes.POINTER(ctypes.c_int64)


def build():
    """Compile libv3d_oracle.so (and oracle/_ref when /root/reference is mounted)."""
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libv3d_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_iou_single.restype = ctypes.c_float
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_c_f)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_c_i)


def _i3(v):
    return (ctypes.c_int * 3)(*[int(x) for x in v])


# ---- IoU / NMS ------------------------------------------------------------------------------
def box_iou_rotated(b1, b2, variant=1):
    b1, p1 = _f(b1)
    b2, p2 = _f(b2)
    out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
    lib().orc_box_iou_rotated(p1, b1.shape[0], p2, b2.shape[0], out.ctypes.data_as(_c_f), variant)
    return out


def nms_rotated(dets, scores, thr, variant=1):
    dets, pd = _f(dets)
    scores, ps = _f(scores)
    keep = np.empty(dets.shape[0], np.int64)
    k = lib().orc_nms_rotated(pd, ps, dets.shape[0], ctypes.c_float(thr), variant,
                              keep.ctypes.data_as(_c_l))
    return keep[:k].copy()


# ---- voxelize ---------------------------------------------------------------------------------
def grid_size(voxel_size, bounds):
    """spconv VoxelGenerator: round((hi - lo) / voxel_size) in fp32 -> cells per axis (xyz)."""
    b = np.asarray(bounds, np.float32)
    v = np.asarray(voxel_size, np.float32)
    return np.round((b[3:] - b[:3]) / v).astype(np.int64)


def voxelize(points, voxel_size, bounds, max_pts, max_voxels, cap_policy=0):
    points, pp = _f(points)
    n, c = points.shape
    lo, plo = _f(np.asarray(bounds, np.float32)[:3])
    vs, pvs = _f(voxel_size)
    g = grid_size(voxel_size, bounds)
    voxels = np.zeros((max_voxels, max_pts, c), np.float32)
    coords = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    m = lib().orc_voxelize(pp, n, c, plo, pvs, _i3(g), max_pts, max_voxels, cap_policy,
                           voxels.ctypes.data_as(_c_f), coords.ctypes.data_as(_c_i),
                           num.ctypes.data_as(_c_i))
    return voxels[:m].copy(), coords[:m].copy(), num[:m].copy()


def vfe_mean(voxels, num):
    voxels, pv = _f(voxels)
    num, pn = _i(num)
    m, k, c = voxels.shape
    out = np.empty((m, c), np.float32)
    lib().orc_vfe_mean(pv, pn, m, k, c, out.ctypes.data_as(_c_f))
    return out


# ---- rule book / sparse conv / dense -------------------------------------------------------------
def _t3(v):
    return [int(v)] * 3 if np.isscalar(v) else [int(x) for x in v]


def rulebook_subm(indices, shape, ksize=3, dilation=1):
    indices, pi = _i(indices)
    ks, dl = _t3(ksize), _t3(dilation)
    n = indices.shape[0]
    nbr = np.empty((ks[0] * ks[1] * ks[2], n), np.int32)
    lib().orc_rulebook_subm(pi, n, _i3(shape), _i3(ks), _i3(dl), nbr.ctypes.data_as(_c_i))
    return nbr


def rulebook_conv(indices, shape, ksize, stride, padding=0, dilation=1):
    indices, pi = _i(indices)
    ks, st, pd, dl = _t3(ksize), _t3(stride), _t3(padding), _t3(dilation)
    n = indices.shape[0]
    kv = ks[0] * ks[1] * ks[2]
    cap = max(n * kv, 1)
    out_shape = (ctypes.c_int * 3)()
    out_idx = np.empty((cap, 4), np.int32)
    nbr = np.empty((kv, cap), np.int32)
    m = lib().orc_rulebook_conv(pi, n, _i3(shape), _i3(ks), _i3(st), _i3(pd), _i3(dl), out_shape,
                                out_idx.ctypes.data_as(_c_i), nbr.ctypes.data_as(_c_i), cap)
    assert m >= 0
    return out_idx[:m].copy(), np.ascontiguousarray(nbr[:, :m]), list(out_shape)


def sparse_conv(feat, weight, nbr, scale=None, shift=None, relu=False):
    """weight (k0,k1,k2,Cin,Cout) or (KV,Cin,Cout); nbr (KV, n_out)."""
    feat, pf = _f(feat)
    w, pw = _f(np.asarray(weight, np.float32).reshape(-1, weight.shape[-2], weight.shape[-1]))
    nbr, pn = _i(nbr)
    kv, n_out = nbr.shape
    cin, cout = w.shape[1], w.shape[2]
    out = np.empty((n_out, cout), np.float32)
    if scale is not None:
        scale, psc = _f(scale)
        shift, psh = _f(shift)
    else:
        psc = psh = None
    lib().orc_sparse_conv(pf, pw, pn, n_out, n_out, kv, cin, cout, psc, psh, int(relu),
                          out.ctypes.data_as(_c_f))
    return out


def dense(feat, indices, batch_size, shape):
    feat, pf = _f(feat)
    indices, pi = _i(indices)
    n, c = feat.shape
    out = np.empty((batch_size, c, shape[0], shape[1], shape[2]), np.float32)
    lib().orc_dense(pf, pi, n, c, batch_size, _i3(shape), out.ctypes.data_as(_c_f))
    return out


# ---- point ops ------------------------------------------------------------------------------------
def fps(xyz, m):
    xyz, px = _f(xyz)
    b, n, _ = xyz.shape
    out = np.empty((b, m), np.int32)
    lib().orc_fps(px, b, n, m, out.ctypes.data_as(_c_i))
    return out


def gather(feat, idx):
    feat, pf = _f(feat)
    idx, pi = _i(idx)
    b, c, n = feat.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().orc_gather(pf, pi, b, c, n, m, out.ctypes.data_as(_c_f))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    out = np.empty((b, m, nsample), np.int32)
    lib().orc_ball_query(px, pq, b, n, m, ctypes.c_float(radius), nsample, out.ctypes.data_as(_c_i))
    return out


def group(feat, idx):
    feat, pf = _f(feat)
    idx, pi = _i(idx)
    b, c, n = feat.shape
    _, m, ns = idx.shape
    out = np.empty((b, c, m, ns), np.float32)
    lib().orc_group(pf, pi, b, c, n, m, ns, out.ctypes.data_as(_c_f))
    return out


def query_and_group(xyz, new_xyz, feat, idx):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    idx, pi = _i(idx)
    b, n, _ = xyz.shape
    _, m, ns = idx.shape
    if feat is not None:
        feat, pf = _f(feat)
        c = feat.shape[1]
    else:
        pf, c = None, 0
    out = np.empty((b, 3 + c, m, ns), np.float32)
    lib().orc_query_and_group(px, pq, pf, pi, b, c, n, m, ns, out.ctypes.data_as(_c_f))
    return out


# ---- the reference's own sources, compiled (oracle/_ref) --------------------------------------------
def ref_available(name):
    return os.path.exists(os.path.join(HERE, "_ref", name))


def _ref_shim(name):
    l = ctypes.CDLL(os.path.join(HERE, "_ref", name))
    return l


def ref_shim_iou(b1, b2, nvcc_view=False):
    """Reference header single_box_iou_rotated via oracle/ref_iou_shim.cpp."""
    l = _ref_shim("libref_iou_nvccview.so" if nvcc_view else "libref_iou_host.so")
    b1, p1 = _f(b1)
    b2, p2 = _f(b2)
    out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
    l.ref_box_iou_rotated(p1, b1.shape[0], p2, b2.shape[0], out.ctypes.data_as(_c_f))
    return out


def ref_shim_nms(dets, scores, thr, nvcc_view=False):
    l = _ref_shim("libref_iou_nvccview.so" if nvcc_view else "libref_iou_host.so")
    dets, pd = _f(dets)
    scores, ps = _f(scores)
    keep = np.empty(dets.shape[0], np.int64)
    k = l.ref_nms_rotated(pd, ps, dets.shape[0], ctypes.c_float(thr), keep.ctypes.data_as(_c_l))
    return keep[:k].copy()


def ref_torch_module(cuda=False):
    """The reference's pybind module (box_iou_rotated, nms_rotated) built by build_ref.py."""
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    name = "ref_C_cuda" if cuda else "ref_C_cpu"
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, "_ref", name + ".so"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
