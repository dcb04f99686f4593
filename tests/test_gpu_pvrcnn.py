"""GPU parity of the PV-RCNN keypoint stage (BASELINE config 3): the production
`pvrcnn.KeypointStage` against the CPU composition of the reference stack (oracle/pvrcnn_cpu.py) on the same clouds
and the same injected grid points. Indices (FPS, every ball query, frame offsets) bit-exact; metric voxel centres
bit-exact; pooled features <= 1e-4 of their scale (north_star: "<= 1e-4 rel for conv/pooling")."""
import numpy as np
import pytest
import torch

import oracle
from oracle import pvrcnn_cpu
from vision3d_b200 import ops, pvrcnn, synth

pytestmark = pytest.mark.gpu


def _inputs(B, n, seed=0, npts=16384):
    clouds = synth.make_batch(seed, B, npts)
    props = pvrcnn.make_proposals(clouds, n, seed)
    noise = pvrcnn.make_grid_noise(B, n, 16, seed)
    grid = pvrcnn.sample_gridpoints(torch.from_numpy(props), noise).reshape(B, -1, 3).numpy()
    return clouds, props, grid


@pytest.mark.parametrize("B,n,fused", [(2, 16, True), (2, 16, False), (8, 100, True)])
def test_keypoint_stage_vs_cpu_composition(cuda, B, n, fused):
    """(8, 100) is config C3 at full size: FPS-2048 on 8 clouds of 16 384 points, 5-source VSA, RoI-grid pool
    with 100 proposals x 16 grid points per frame."""
    cfg = pvrcnn.PVRCNNConfig()
    model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0)
    clouds, props, grid = _inputs(B, n)
    # fused=True: grouping -> shared MLP -> max in ONE tensor-core kernel per scale;
    # fused=False: a10 grouping kernels + the MLP / max in torch (the reference-shaped composition)
    stage = pvrcnn.KeypointStage(model, B, 16384, n, cuda, fused_sa=fused)
    stage.load(clouds, grid)
    pooled = stage.step()
    torch.cuda.synchronize()

    cpu_model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0).eval()
    st = {}
    want = pvrcnn_cpu.keypoint_stage(cpu_model, clouds, grid, stages=st)

    # a7/a8: FPS indices and gathered keypoints, ALL clouds
    assert np.array_equal(stage.kp_idx.cpu().numpy(), st["kp_idx"])
    assert np.array_equal(stage.keypoints.cpu().numpy(), st["keypoints"])
    # a15 + to_global: frame offsets and metric voxel centres of every level
    for lv in range(4):
        idx = st["levels"][lv][0]
        m = len(idx)
        assert int(stage.eng.n_rows[lv].item()) == m
        starts = np.searchsorted(idx[:, 0], np.arange(B + 1))
        assert np.array_equal(stage.level_offsets[lv].cpu().numpy(), starts)
        xyz = pvrcnn_cpu.to_global(idx, cfg.VOXEL_SIZE, cfg.STRIDES[lv], cfg.GRID_BOUNDS[:3])
        assert np.array_equal(stage.level_xyz[lv][:m].cpu().numpy(), xyz)
        f = stage.eng.vox_out["mean"] if lv == 0 else stage.eng.feat[lv][0]
        ref = st["level_feats"][lv].numpy()
        assert np.abs(f[:m].cpu().numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-6), lv
    # a9: every ball query of the 5 set-abstraction sources, both radii, bit-exact
    for i in range(5):
        for r in range(2):
            got = stage.sa_idx[i][r].cpu().numpy()
            ref = np.stack(st["sa_idx"][i][r], 0)
            assert np.array_equal(got, ref), (i, r)
    # a10 + MLP + max, BEV gather: the 512-channel keypoint features
    kf, kr = stage.kp_features.cpu(), st["kp_features"]
    for c0, c1 in [(0, 32), (32, 64), (64, 128), (128, 256), (256, 384), (384, 512)]:
        scale = kr[:, c0:c1].abs().max()
        err = float((kf[:, c0:c1] - kr[:, c0:c1]).abs().max() / scale)
        assert err <= 1e-4, (c0, c1, err)
    # a11: RoI-grid ball queries exact, pooled proposal features
    for r in range(2):
        assert np.array_equal(stage.roi_idx[r].cpu().numpy(), np.stack(st["roi_idx"][r], 0)), r
    err = float((pooled.cpu() - want).abs().max() / want.abs().max())
    print("C3 B=%d n=%d fused_sa=%s: pooled err %.2e" % (B, n, fused, err))
    assert pooled.shape == (B, n, 256) and err <= 1e-4, err


def test_reference_shaped_methods_match_stage(cuda):
    """PVRCNNB200's eager, reference-shaped methods (sample_keypoints / _pointnets / point_feature_extract through the
    compat pointnet2 drop-ins and a dense pad_batch) give the stage's keypoint features: padding with duplicate rows
    does not change the pooled result."""
    B, n = 2, 8
    cfg = pvrcnn.PVRCNNConfig()
    model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0)
    clouds, props, grid = _inputs(B, n, seed=3)
    stage = pvrcnn.KeypointStage(model, B, 16384, n, cuda, fused_sa=False)
    stage.load(clouds, grid)
    stage.step()
    pts = stage.points
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        kp = model.sample_keypoints(pts)
        assert torch.equal(kp, stage.keypoints)
        cnn_features = []
        for lv in range(4):
            offs = stage.level_offsets[lv]
            counts = (offs[1:] - offs[:-1])
            cap = int(counts.max().item())
            f = stage.eng.vox_out["mean"] if lv == 0 else stage.eng.feat[lv][0]
            xyz_d = ops.pad_batch(stage.level_xyz[lv], offs, B, cap, seed=11)
            f_d = ops.pad_batch(f, offs, B, cap, seed=11)
            o = offs.cpu().numpy()
            for b in range(B):   # rows below each frame's count are the frame's own rows (the rest: duplicates)
                cnt = int(o[b + 1] - o[b])
                assert torch.equal(xyz_d[b, :cnt], stage.level_xyz[lv][o[b]:o[b + 1]])
                assert torch.equal(f_d[b, :cnt], f[o[b]:o[b + 1]])
            cnn_features.append((xyz_d, f_d))
        feats = model.point_feature_extract(pts, kp, cnn_features, stage.eng.bev_nhwc)
    ref = stage.kp_features
    assert float((feats - ref).abs().max() / ref.abs().max()) <= 1e-5
    # RoiGridPool.forward, reference-shaped (noise injected), vs the stage
    noise = pvrcnn.make_grid_noise(B, n, 16, 3).to(cuda)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        g_dev = pvrcnn.sample_gridpoints(torch.from_numpy(props).to(cuda), noise).view(B, -1, 3)
        assert float((g_dev.cpu() - torch.from_numpy(grid)).abs().max()) <= 1e-4   # cos/sin: last-bit differences only
        model.to(cuda).eval()
        pooled = model.roi_grid_pool.pnet(kp, ref, stage.gridpoints.contiguous())[1]
    assert pooled.shape == (B, 192, n * 16)


def test_point_op_kernels_new_entry_points(cuda):
    """v3d_fps_keypoints (stride 3 and 4), v3d_ball_query_msg (dense + ragged) == the single-op entry points == oracle;
    v3d_batch_offsets, v3d_to_global, v3d_pad_batch known answers."""
    rng = np.random.default_rng(0)
    clouds = synth.make_batch(7, 3, 8192)
    pts = torch.from_numpy(np.stack(clouds)).to(cuda)
    idx4, kp4 = ops.fps_keypoints(pts, 512)
    idx3, kp3 = ops.fps_keypoints(pts[..., :3].contiguous(), 512)
    want = oracle.fps(np.stack(clouds)[..., :3].copy(), 512)
    assert np.array_equal(idx4.cpu().numpy(), want) and torch.equal(idx3, idx4) and torch.equal(kp3, kp4)
    assert torch.equal(ops.furthest_point_sample(pts[..., :3].contiguous(), 512), idx4)
    assert np.array_equal(kp4.cpu().numpy(), np.stack([c[want[b], :3] for b, c in enumerate(clouds)]))
    # multi-radius == single-radius kernel == oracle (incl. a radius with no hits and one that saturates instantly)
    radii, ns = [0.05, 0.4, 3.0], [16, 32, 8]
    outs = ops.ball_query_msg(radii, ns, pts, kp4)
    for r in range(3):
        single = ops.ball_query(radii[r], ns[r], pts[..., :3].contiguous(), kp4)
        assert torch.equal(outs[r], single)
        assert np.array_equal(outs[r].cpu().numpy(), oracle.ball_query(radii[r], ns[r], np.stack(clouds)[..., :3].copy(),
                                                                       kp4.cpu().numpy()))
    # ragged sources: frames of different length packed back to back
    lens = [5000, 8192, 123]
    rows = torch.cat([pts[b, :lens[b], :3] for b in range(3)]).contiguous()
    offs = torch.tensor([0, 5000, 13192, 13315], dtype=torch.int32, device=cuda)
    outs = ops.ball_query_msg([0.8, 1.6], [16, 32], rows, kp4, offs)
    for b in range(3):
        for r, (rad, n_) in enumerate(zip([0.8, 1.6], [16, 32])):
            ref = oracle.ball_query(rad, n_, clouds[b][None, :lens[b], :3].copy(), kp4[b:b + 1].cpu().numpy())
            assert np.array_equal(outs[r][b].cpu().numpy(), ref[0]), (b, r)
    # culled variant (chunk bounding boxes): bit-identical on ragged sources in random order and in sorted order
    bnd = ops.BallQueryBounds(3, rows.shape[0], cuda).build(rows, offs)
    culled = ops.ball_query_msg([0.8, 1.6], [16, 32], rows, kp4, offs, bounds=bnd)
    assert torch.equal(culled[0], outs[0]) and torch.equal(culled[1], outs[1])
    srt = torch.cat([rows[int(offs[b]):int(offs[b + 1])][torch.argsort(
        (rows[int(offs[b]):int(offs[b + 1]), 2] * 10).floor() * 1e6 + (rows[int(offs[b]):int(offs[b + 1]), 1] * 10).floor() * 1e3
        + rows[int(offs[b]):int(offs[b + 1]), 0])] for b in range(3)]).contiguous()
    plain_s = ops.ball_query_msg([0.3, 2.5], [16, 32], srt, kp4, offs)
    culled_s = ops.ball_query_msg([0.3, 2.5], [16, 32], srt, kp4, offs, bounds=ops.BallQueryBounds(3, 8192, cuda).build(srt, offs))
    assert torch.equal(plain_s[0], culled_s[0]) and torch.equal(plain_s[1], culled_s[1])
    dense_b = ops.BallQueryBounds(3, 8192, cuda).build(pts)
    assert torch.equal(ops.ball_query_msg([0.4], [16], pts, kp4, bounds=dense_b)[0], ops.ball_query_msg([0.4], [16], pts, kp4)[0])
    # selecting variant (x-bucketed copy + nsample smallest original indices): bit-identical, ragged and dense,
    # dense balls (many more hits than nsample), points outside the x range (clamped buckets), nsample < 16
    sel = ops.BallQuerySorted(3, rows.shape[0], rows.shape[0], (0.0, 70.4), cuda).build(rows, offs)
    got = sel.query([0.8, 1.6], [16, 32], kp4)
    assert torch.equal(got[0], outs[0]) and torch.equal(got[1], outs[1])
    got = sel.query([0.3, 2.5, 6.0], [5, 32, 32], kp4)
    want = ops.ball_query_msg([0.3, 2.5, 6.0], [5, 32, 32], rows, kp4, offs)
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    seld = ops.BallQuerySorted(3, 8192, 3 * 8192, (10.0, 30.0), cuda).build(pts)     # stride-4 dense rows, narrow range
    got = seld.query([0.4, 3.0], [16, 32], kp4)
    want = ops.ball_query_msg([0.4, 3.0], [16, 32], pts, kp4)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    got2 = seld.build(pts).query([0.4, 3.0], [16, 32], kp4)                          # workspace re-armed by the scan
    assert torch.equal(got2[0], want[0]) and torch.equal(got2[1], want[1])
    feat = torch.randn((rows.shape[0], 5), device=cuda)
    g = ops.query_and_group_rows(rows, feat, kp4, outs[0], offs)
    for b in range(3):
        sl = slice(int(offs[b]), int(offs[b + 1]))
        ref = oracle.query_and_group(rows[sl][None].cpu().numpy(), kp4[b:b + 1].cpu().numpy(),
                                     feat[sl].t()[None].contiguous().cpu().numpy(), outs[0][b:b + 1].cpu().numpy())
        assert np.array_equal(g[b].cpu().numpy(), ref[0])
    # strided feature column (intensity of x,y,z,i rows)
    g2 = ops.query_and_group_rows(pts, pts[..., 3:], kp4, ops.ball_query_msg([0.8], [16], pts, kp4)[0])
    assert g2.shape == (3, 4, 512, 16)
    i0 = ops.ball_query_msg([0.8], [16], pts, kp4)[0].long()
    assert torch.equal(g2[:, 3], torch.gather(pts[..., 3], 1, i0.view(3, -1)).view(3, 512, 16))
    # batch offsets / to_global / pad_batch
    ind = torch.tensor([[0, 1, 2, 3], [0, 4, 5, 6], [2, 0, 0, 1], [2, 7, 8, 9], [2, 9, 9, 9]], dtype=torch.int32, device=cuda)
    n_rows = torch.tensor([5], dtype=torch.int32, device=cuda)
    o = ops.batch_offsets(ind, n_rows, 4)
    assert o.tolist() == [0, 2, 2, 5, 5]
    xyz = ops.to_global(ind, n_rows, [0.1, 0.2, 0.4], [0.0, -40.0, -3.0])
    want = ind[:, [3, 2, 1]].float().cpu() * torch.tensor([0.1, 0.2, 0.4]) + torch.tensor([0.0, -40.0, -3.0])
    assert torch.equal(xyz.cpu(), want)
    src = torch.arange(10, dtype=torch.float32, device=cuda).view(5, 2)
    d = ops.pad_batch(src, o, 4, 4, seed=5)
    assert d.shape == (4, 4, 2) and torch.equal(d[0, :2], src[:2]) and torch.equal(d[2, :3], src[2:5])
    assert bool((d[1] == 0).all()) and bool((d[3] == 0).all())                     # empty frames -> zeros
    assert all(any(torch.equal(d[0, j], src[k]) for k in range(2)) for j in range(2, 4))
    assert torch.equal(d, ops.pad_batch(src, o, 4, 4, seed=5))                     # same seed, same picks


def test_fused_anchor_matching_vs_expression_path_and_reference_golden(cuda):
    """v3d_match_anchors (IoU + Matcher fused, no M x 70 400 matrix) == the reference's expression
    sequence on the a13 kernel (bit-exact matches / labels / IoU maxima), and the dense targets it yields are the ones
    the reference's own ProposalTargetAssigner produced on its CPU op (golden), up to IoU values that sit within
    rounding of a threshold (the CPU and CUDA builds of the reference header differ in the hull sort)."""
    import os
    from vision3d_b200 import second, targets
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "targets_golden.npz"))
    cfg = second.three_class_config()
    fused = targets.ProposalTargetAssignerB200(cfg, device=cuda, fused=True)
    plain = targets.ProposalTargetAssignerB200(cfg, device=cuda, fused=False)
    for tag in ("a", "b"):
        boxes = torch.from_numpy(gold[tag + "_boxes"]).to(cuda)
        cls = torch.from_numpy(gold[tag + "_class_idx"]).to(cuda)
        i1, i2 = dict(boxes=boxes, class_idx=cls), dict(boxes=boxes, class_idx=cls)
        with torch.no_grad():
            fused(i1)
            plain(i2)
        for k in ("G_cls", "M_cls", "G_reg", "M_reg"):
            assert torch.equal(i1[k], i2[k]), k
        # kernel-level: maxima and first-max indices against the materialised matrix
        for c in range(3):
            m = cls == c
            if not bool(m.any()):
                continue
            a = fused.anchors[c].view(-1, 7)[:, [0, 1, 3, 4, 6]].contiguous()
            q = ops.box_iou_rotated(boxes[m][:, [0, 1, 3, 4, 6]].contiguous(), a)
            mt, lab, vals = ops.match_anchors(boxes[m][:, [0, 1, 3, 4, 6]].contiguous(), a, [0.45, 0.6], [0, -1, 1], True)
            v, _ = q.max(0)
            assert torch.equal(vals, v)
            assert torch.equal(q.gather(0, mt[None])[0], v)                  # index points at a maximum ...
            first = (q == v[None]).float().argmax(0)
            assert torch.equal(mt, first)                                     # ... the first one
        pos = torch.nonzero(i1["G_cls"].reshape(-1) == 1).squeeze(1).cpu().numpy()
        ign = torch.nonzero(~i1["M_cls"].reshape(-1)).squeeze(1).cpu().numpy()
        diff = len(set(pos) ^ set(gold[tag + "_pos"])) + len(set(ign) ^ set(gold[tag + "_ign"]))
        assert diff <= 2, diff                                               # threshold-boundary flips only
        common = np.intersect1d(pos, gold[tag + "_pos"], return_indices=True)
        reg = i1["G_reg"].reshape(-1, 7)[torch.from_numpy(common[0]).to(cuda)].cpu().numpy()
        np.testing.assert_allclose(reg, gold[tag + "_reg_pos"][common[2]], rtol=0, atol=1e-5)


@pytest.mark.parametrize("C,widths,ns", [(1, (8, 16), 16), (4, (8, 16), 32), (32, (32, 32), 16), (64, (64, 64), 32),
                                         (512, (192, 96), 16)])
def test_sa_fused_kernel_vs_torch_expression(cuda, C, widths, ns):
    """v3d_sa_fused (one scale: grouping -> 1x1 conv + ReLU -> 1x1 conv + ReLU -> max) vs QueryAndGroup + torch
    fp32 convolutions on the same ball-query indices; ragged sources; <= 1e-4 of the output scale."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(C)
    B, M = 2, 300                                     # 600 queries -> partial last tile for ns = 16 (9600 rows = 75 tiles)
    lens = [900, 1500]
    offs = torch.tensor([0, 900, 2400], dtype=torch.int32, device=cuda)
    xyz = (torch.rand((2400, 3), generator=g) * 4).to(cuda)
    feat = torch.randn((2400, C), generator=g).to(cuda)
    q = (torch.rand((B, M, 3), generator=g) * 4).to(cuda)
    idx = ops.ball_query_msg([0.7], [ns], xyz, q, offs)[0]
    n1, n2 = widths
    w1 = (torch.randn((n1, 3 + C), generator=g) / (3 + C) ** 0.5).to(cuda)
    b1 = (torch.randn(n1, generator=g) * 0.1).to(cuda)
    w2 = (torch.randn((n2, n1), generator=g) / n1 ** 0.5).to(cuda)
    b2 = (torch.randn(n2, generator=g) * 0.1).to(cuda)
    grouped = ops.query_and_group_rows(xyz, feat, q, idx, offs)                       # (B, 3 + C, M, ns)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h = F.relu(F.conv2d(grouped, w1[:, :, None, None], b1))
        want = F.relu(F.conv2d(h, w2[:, :, None, None], b2)).amax(3)                  # (B, n2, M)
    Cp = -(-C // 8) * 8
    packed = ops.pack_channel_major(feat.t().unsqueeze(0), Cp)                        # (1, C, rows) view -> (rows, 2 Cp)
    mlp = ops.PreparedSaMlp([(w1, b1), (w2, b2)], C)
    out = torch.full((B, n2 + 5, M), -7.0, device=cuda)
    ops.sa_fused(packed, xyz, q, idx, mlp, out, c_off=3, row_offsets=offs)
    torch.cuda.synchronize()
    err = float((out[:, 3:3 + n2] - want).abs().max() / want.abs().max())
    assert err <= 1e-4, err
    assert bool((out[:, :3] == -7.0).all()) and bool((out[:, 3 + n2:] == -7.0).all())   # only its channel slice is written


@pytest.mark.gpu
def test_bev_gather_vs_torch_expression(cuda):
    """v3d_bev_gather == the reference's BEVFeatureGatherer expression (torch ops + F.grid_sample), including keypoints
    outside the map (clamped), the (size - 2) normaliser and the x <-> W swap; partial keypoint tiles; C = 128 and 8."""
    from vision3d_b200 import pvrcnn
    cfg = pvrcnn.PVRCNNConfig()
    g = torch.Generator().manual_seed(3)
    px = (np.asarray(cfg.VOXEL_SIZE[:2], np.float32) * np.float32(cfg.STRIDES[-1])).tolist()
    for (B, C, H, W, M) in [(2, 128, 200, 176, 2048), (3, 8, 37, 21, 70)]:
        fmap = torch.randn((B, C, H, W), generator=g).to(cuda).contiguous(memory_format=torch.channels_last)
        lo = torch.tensor([cfg.GRID_BOUNDS[0] - 3.0, cfg.GRID_BOUNDS[1] - 3.0, -3.0])
        span = torch.tensor([W * px[0] + 6.0, H * px[1] + 6.0, 4.0])
        kp = (torch.rand((B, M, 3), generator=g) * span + lo).to(cuda)
        kp[0, 0, :2] = torch.tensor(cfg.GRID_BOUNDS[:2])                      # exactly on the first pixel
        want = pvrcnn.bev_gather(cfg, fmap, kp)
        got = ops.bev_gather(fmap, kp, cfg.GRID_BOUNDS[:2], px)
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max()), float((got - want).abs().max())
        out = torch.zeros((B, C + 5, M), device=cuda)
        ops.bev_gather(fmap, kp, cfg.GRID_BOUNDS[:2], px, out=out, c_off=5)
        assert torch.equal(out[:, 5:], got) and bool((out[:, :5] == 0).all())
    with pytest.raises(ops.V3DError):
        ops.bev_gather(torch.zeros((1, 8, 4, 4), device=cuda), torch.zeros((1, 4, 3), device=cuda), [0, 0], [1, 1])
