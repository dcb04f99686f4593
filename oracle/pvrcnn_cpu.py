"""TEST INFRASTRUCTURE ONLY: the PV-RCNN keypoint stage (BASELINE config 3) on the host CPU, composed op for op the
way the reference stack would run it (detector/model.py:46-74, detector/roi_grid_pool.py:51-72,
detector/sparse_cnn.py:91-146, detector/layers.py:20-50) from the oracle's C++ restatements of the pointnet2 ops
and the model's own torch modules on CPU. Checker of vision3d_b200.pvrcnn.KeypointStage and the C3 CPU baseline.

pad_batch (sparse_cnn.py:118-126) is not drawn: each frame's voxel set is queried as it is. Duplicate rows appended
after the real ones can only take slots the first hit would have filled and carry the features of real hits of the
same ball, so the max-pooled result is the same (checked in tests/test_oracle_cpu.py::test_pad_batch_is_pool_invariant).
"""
import time

import numpy as np
import torch

import oracle
from oracle import second_cpu


def to_global(idx, voxel_size, stride, offset):
    """sparse_cnn.py:91-105: flip (b,z,y,x) -> (x,y,z), * (base_voxel_size * stride) + voxel_offset, fp32."""
    vs = (np.asarray(voxel_size, np.float32) * np.float32(stride)).astype(np.float32)
    return (idx[:, [3, 2, 1]].astype(np.float32) * vs + np.asarray(offset, np.float32)).astype(np.float32)


def _sa_module(pnet, xyz_frames, feat_frames, new_xyz, nsamples, idx_out=None):
    """PointnetSAModuleMSG.forward with per-frame (ragged) sources. xyz_frames[b] (Nb, 3), feat_frames[b] (Nb, C),
    new_xyz (B, M, 3) numpy -> (B, sum Cout, M) torch."""
    B = len(xyz_frames)
    outs = []
    for r, (grouper, mlp) in enumerate(zip(pnet.groupers, pnet.mlps)):
        per_frame = []
        for b in range(B):
            x = np.ascontiguousarray(xyz_frames[b][None])
            q = np.ascontiguousarray(new_xyz[b][None])
            idx = oracle.ball_query(grouper.radius, nsamples[r], x, q)
            if idx_out is not None:
                idx_out.setdefault(r, []).append(idx[0])
            f = np.ascontiguousarray(feat_frames[b].T[None])          # (1, C, Nb)
            g = oracle.query_and_group(x, q, f, idx)                   # (1, 3+C, M, ns)
            with torch.no_grad():
                per_frame.append(mlp(torch.from_numpy(g)).amax(dim=3))
        outs.append(torch.cat(per_frame, 0))
    return torch.cat(outs, 1)


@torch.no_grad()
def keypoint_stage(model, clouds, gridpoints, stages=None, timings=None):
    """model: vision3d_b200.pvrcnn.PVRCNNB200 on CPU, eval. clouds: list of (N, 4); gridpoints (B, n*16, 3) =
    sample_gridpoints(proposals, noise) (roi_grid_pool.py:51-62), generated once and injected into both
    implementations. Returns pooled RoI features (B, n, 256) torch."""
    from vision3d_b200 import pvrcnn
    cfg = model.cfg
    B = len(clouds)
    pts = np.stack(clouds, 0).astype(np.float32)
    t = time.perf_counter()
    kp_idx = oracle.fps(np.ascontiguousarray(pts[..., :3]), cfg.NUM_KEYPOINTS)
    kp = np.stack([pts[b, kp_idx[b], :3] for b in range(B)], 0)
    t_fps = time.perf_counter() - t
    v, c, n = second_cpu.voxelize_batch(cfg, clouds)
    feat = torch.from_numpy(oracle.vfe_mean(v, n))
    bev, levels, level_feats = second_cpu.sparse_middle(model, feat, c, B, return_levels="features")
    t = time.perf_counter()
    lo = np.asarray(cfg.GRID_BOUNDS[:3], np.float32)
    sources = [([pts[b, :, :3] for b in range(B)], [pts[b, :, 3:4] for b in range(B)])]
    for lv in range(4):
        idx = levels[lv][0]
        xyz = to_global(idx, cfg.VOXEL_SIZE, cfg.STRIDES[lv], lo)
        f = level_feats[lv].numpy()
        starts = np.searchsorted(idx[:, 0], np.arange(B + 1))           # compute_pad_amounts, sparse_cnn.py:107-116
        sources.append(([xyz[starts[b]:starts[b + 1]] for b in range(B)], [f[starts[b]:starts[b + 1]] for b in range(B)]))
    sa_idx = {}
    pf = []
    for i, (xs, fs) in enumerate(sources):
        io = {}
        pf.append(_sa_module(model.pnets[i], xs, fs, kp, cfg.SAMPLES_PN, io))
        sa_idx[i] = io
    bevf = pvrcnn.bev_gather(cfg, bev, torch.from_numpy(kp))
    kp_features = torch.cat(pf + [bevf], 1)                                # (B, 512, M)
    t_vsa = time.perf_counter() - t
    t = time.perf_counter()
    grid = np.ascontiguousarray(np.asarray(gridpoints, np.float32).reshape(B, -1, 3))
    roi_idx = {}
    f = _sa_module(model.roi_grid_pool.pnet, [kp[b] for b in range(B)],
                   [kp_features[b].T.contiguous().numpy() for b in range(B)], grid, cfg.SAMPLES_PN, roi_idx)
    m = cfg.GRIDPOOL_NUM_GRIDPOINTS
    nprop = grid.shape[1] // m
    f = f.view(B, -1, nprop, m).permute(0, 2, 1, 3).contiguous().view(B, nprop, -1)
    pooled = model.roi_grid_pool.reduction(f)
    t_roi = time.perf_counter() - t
    if stages is not None:
        stages.update(kp_idx=kp_idx, keypoints=kp, levels=levels, level_feats=level_feats, bev=bev, sa_idx=sa_idx,
                      kp_features=kp_features, gridpoints=grid, roi_idx=roi_idx)
    if timings is not None:
        timings.update(fps_s=t_fps, vsa_s=t_vsa, roi_s=t_roi)
    return pooled
