"""Per-op device timings (CUDA events, median of N after warm-up) for the hot-path kernels.
Not the contract bench (that is bench.py); this is the builder's quick look at each op.
    python scripts/microbench.py [--out gpurun_out/microbench.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision3d_b200 import ops, synth  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    res = {}

    def rec(name, fn, **extra):
        med, mn = timeit(fn)
        res[name] = dict(us_median=round(med, 2), us_min=round(mn, 2), **extra)
        print(name, res[name], flush=True)

    # NMS / IoU
    for n in (100, 1600, 2400, 19200):
        boxes, scores, idxs = synth.make_nms_boxes(0, n)
        b = torch.from_numpy(synth.apply_group_offsets(boxes, idxs)).to(dev)
        s = torch.from_numpy(scores).to(dev)
        ws = ops.nms_workspace(n, dev)
        keep = torch.empty(n, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        rec("nms_rotated_N%d" % n, lambda: ops.nms_rotated_padded(b, s, 0.01, ws, keep, cnt), kept=int(cnt.item()))
    boxes, _, _ = synth.make_nms_boxes(1, 512, degrees=True)
    bb = torch.from_numpy(boxes).to(dev)
    rec("box_iou_rotated_512x512", lambda: ops.box_iou_rotated(bb, bb))

    # voxelize T16
    for B in (1, 16, 64):
        clouds = synth.make_batch(0, B)
        pts = torch.from_numpy(np.concatenate(clouds, 0)).to(dev)
        off = torch.from_numpy(np.arange(B + 1, dtype=np.int32) * 16384).to(dev)
        vz = ops.Voxelizer(synth.VOXEL_SIZE, synth.GRID_BOUNDS, synth.MAX_VOXELS, synth.MAX_OCCUPANCY, B,
                           B * 16384, device=dev)
        out = vz.alloc_outputs(4, with_mean=True)
        rec("voxelize_B%d" % B, lambda: vz.run(pts, off, 16384, out))
        M = int(out["voxel_offsets"][-1].item())
        bytes_alg = 16 * B * 16384 + 100 * M
        res["voxelize_B%d" % B].update(voxels=M, alg_bytes=bytes_alg,
                                       gbs=round(bytes_alg / res["voxelize_B%d" % B]["us_median"] / 1e3, 1))
        out_nm = dict(out, mean=None)
        rec("voxelize_nomean_B%d" % B, lambda: vz.run(pts, off, 16384, out_nm))

    # C4: 40k sites, 16 -> 64
    shape = [41, 400, 352]
    for name, idx in (("random", synth.make_active_sites(0, 40000, shape, 1)),
                      ("clustered", synth.make_clustered_sites(0, 40000, shape, 1))):
        n = len(idx)
        ind = torch.from_numpy(idx).to(dev)
        n_rows = torch.tensor([n], dtype=torch.int32, device=dev)
        table = ops.SiteTable(n, dev)
        rec("site_table_build_%s" % name, lambda: table.build(ind, n_rows, shape))
        nbr = torch.empty((27, n), dtype=torch.int32, device=dev)
        rec("rulebook_subm_%s" % name, lambda: ops.rulebook_subm(table, ind, n_rows, shape, 3, 1, nbr))
        pairs = int((nbr >= 0).sum().item())
        feat = torch.randn(n, 16, device=dev)
        w = torch.randn(27, 16, 64, device=dev) / np.sqrt(27 * 16)
        out = torch.empty(n, 64, device=dev)
        rec("sparse_conv_16x64_%s" % name, lambda: ops.sparse_conv(feat, w, nbr, n_rows, n, out=out),
            pairs=pairs, p_over_n=round(pairs / n, 2))
        feat64 = torch.randn(n, 64, device=dev)
        w64 = torch.randn(27, 64, 64, device=dev) / np.sqrt(27 * 64)
        rec("sparse_conv_64x64_%s" % name, lambda: ops.sparse_conv(feat64, w64, nbr, n_rows, n, out=out),
            pairs=pairs)
        wsz = ops.ConvRulebookWorkspace(1, ops.conv_out_shape(shape, [3] * 3, [2] * 3, [1] * 3, [1] * 3), 4 * n, 27, dev)
        oi = torch.empty((4 * n, 4), dtype=torch.int32, device=dev)
        no = torch.zeros(1, dtype=torch.int32, device=dev)
        nb2 = torch.empty((27, 4 * n), dtype=torch.int32, device=dev)
        rec("rulebook_conv_s2_%s" % name,
            lambda: ops.rulebook_conv(table, ind, n_rows, 1, shape, 3, 2, 1, 1, 4 * n, oi, no, nb2, wsz),
            n_out=int(no.item()))

    # dense (B=16 final level)
    B, shape = 16, [2, 200, 176]
    idx = synth.make_active_sites(0, 20000, shape, B)
    ind = torch.from_numpy(idx).to(dev)
    n_rows = torch.tensor([len(idx)], dtype=torch.int32, device=dev)
    feat = torch.randn(len(idx), 64, device=dev)
    dout = torch.empty((B, 64, *shape), device=dev)
    wsd = torch.empty(ops._lib.load().v3d_sparse_to_dense_workspace_bytes(B, ops.i3(shape)), dtype=torch.uint8, device=dev)
    rec("dense_B16", lambda: ops.sparse_to_dense(feat, ind, n_rows, len(idx), B, shape, dout, wsd))
    nbytes = dout.numel() * 4 + len(idx) * (64 * 4 + 16)
    res["dense_B16"].update(alg_bytes=nbytes, gbs=round(nbytes / res["dense_B16"]["us_median"] / 1e3, 1))

    # FPS / ball query (C3)
    xyz = torch.from_numpy(np.stack([c[:, :3] for c in synth.make_batch(0, 8)], 0)).to(dev)
    rec("fps_B8_N16384_m2048", lambda: ops.furthest_point_sample(xyz, 2048))
    kp_idx = ops.furthest_point_sample(xyz, 2048)
    kp = ops.gather_operation(xyz.transpose(1, 2).contiguous(), kp_idx).transpose(1, 2).contiguous()
    for r, ns in ((0.4, 16), (0.8, 32), (4.8, 32)):
        rec("ball_query_r%.1f_ns%d" % (r, ns), lambda: ops.ball_query(r, ns, xyz, kp))
    idx = ops.ball_query(0.8, 32, xyz, kp)
    f = torch.randn(8, 64, 16384, device=dev)
    rec("query_and_group_C64_ns32", lambda: ops.query_and_group(xyz, kp, f, idx))

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
