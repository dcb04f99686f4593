"""BASELINE config C4: SparseConv3d microbench -- 40 000 active voxels in a [41, 400, 352] grid (B = 1), 16 -> 64
channels, rule-book sweep: SubM k=3 on occupancy patterns with P/N ~ {1, 4, 8, 13, 20, 27} pairs per site plus the
strided k3 s2 conv, through the tcgen05 kernel and the exact-fp32 SIMT kernel.
    python scripts/c4_sweep.py > profiles/rXX_c4_sweep.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision3d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPE, N, CIN, COUT = [41, 400, 352], 40000, 16, 64
rng = np.random.default_rng(0)


def sites(target):
    """N sites whose SubM k=3 rule book has ~target pairs per site: a box filled with probability (target-1)/26."""
    if target <= 1:
        flat = rng.choice(SHAPE[0] * SHAPE[1] * SHAPE[2], size=N, replace=False)
    else:
        p = min(1.0, (target - 1) / 26.0)
        vol = int(N / p * 1.02) + 64
        dz = min(SHAPE[0], 32)
        side = int(np.ceil(np.sqrt(vol / dz)))
        zz, yy, xx = np.meshgrid(np.arange(dz), np.arange(side), np.arange(side), indexing="ij")
        cells = np.stack([zz.ravel(), yy.ravel(), xx.ravel()], 1)
        keep = rng.random(len(cells)) < p
        cells = cells[keep]
        cells = cells[rng.permutation(len(cells))[:N]]
        flat = (cells[:, 0] * SHAPE[1] + cells[:, 1]) * SHAPE[2] + cells[:, 2]
    z, rem = np.divmod(flat, SHAPE[1] * SHAPE[2])
    y, x = np.divmod(rem, SHAPE[2])
    idx = np.stack([np.zeros_like(z), z, y, x], 1).astype(np.int32)
    return idx[rng.permutation(len(idx))]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


rows = []
w = torch.from_numpy((rng.normal(size=(27, CIN, COUT)) / np.sqrt(27 * CIN)).astype(np.float32)).to(dev)
pw = ops.PreparedWeights(w)
for target, kind in [(1, "subm"), (4, "subm"), (8, "subm"), (13, "subm"), (20, "subm"), (27, "subm"), (8, "strided")]:
    idx = sites(target)
    n = len(idx)
    ind = torch.from_numpy(idx).to(dev)
    n_rows = torch.tensor([n], dtype=torch.int32, device=dev)
    table = ops.SiteTable(n, dev).build(ind, n_rows, SHAPE)
    feat = torch.randn((n, CIN), device=dev)
    if kind == "subm":
        nbr = ops.rulebook_subm(table, ind, n_rows, SHAPE, 3, 1)
        n_out_dev, n_out, cap = n_rows, n, n
        t_rb = timed(lambda: ops.rulebook_subm(table, ind, n_rows, SHAPE, 3, 1, nbr))
    else:
        cap = 8 * n
        out_idx, n_out_dev, nbr, _ = ops.rulebook_conv(table, ind, n_rows, 1, SHAPE, 3, 2, 1, 1, cap)
        n_out = int(n_out_dev.item())
        ws = ops.ConvRulebookWorkspace(1, [21, 200, 176], cap, 27, dev)
        t_rb = timed(lambda: ops.rulebook_conv(table, ind, n_rows, 1, SHAPE, 3, 2, 1, 1, cap, out_idx, n_out_dev, nbr, ws))
    P = int((nbr[:, :n_out] >= 0).sum().item())
    packed = ops.pack_features(feat, n_rows)
    outp = torch.empty((cap, 2 * COUT), dtype=torch.bfloat16, device=dev)
    out32 = torch.empty((cap, COUT), device=dev)
    t_tc = timed(lambda: ops.sparse_conv(packed, pw, nbr, n_out_dev, cap, out_packed=outp, write_f32=False))
    t_simt = timed(lambda: ops.sparse_conv(feat, w, nbr, n_out_dev, cap, out=out32))
    got = ops.sparse_conv(packed, pw, nbr, n_out_dev, cap)[:n_out]
    ref = ops.sparse_conv(feat, w, nbr, n_out_dev, cap)[:n_out]
    err = float((got - ref).abs().max() / ref.abs().max())
    alg = 4 * (n * CIN + n_out * COUT) + 8 * P + 4 * 27 * CIN * COUT
    stream = P * (4 * CIN + 8 * COUT + 8)
    flops = 2 * P * CIN * COUT
    rows.append(dict(kind=kind, target_pairs_per_site=target, n_in=n, n_out=n_out, pairs=P, pairs_per_out=round(P / n_out, 2),
                     us_tcgen05=round(t_tc, 1), us_simt_fp32=round(t_simt, 1), us_rulebook=round(t_rb, 1),
                     alg_bytes=alg, gbs_alg_tcgen05=round(alg / t_tc / 1e3, 1), ref_style_stream_bytes=stream,
                     tflops_tcgen05=round(flops / t_tc / 1e6, 2), tflops_simt=round(flops / t_simt / 1e6, 2),
                     max_rel_diff_tc_vs_fp32=err))
print(json.dumps({"config": "C4: 40k active voxels, [41,400,352], B=1, 16->64, fp32 features; SubM k3 occupancy sweep + strided k3 s2",
                  "note": "40k rows = 313 tiles of 128 on 148 SMs: ~2 tiles per SM, launch and pipeline fill dominate at this size",
                  "rows": rows}, indent=1))
