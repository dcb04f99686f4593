"""GPU tests of the end-to-end SECOND path: the production engine (static buffers, CUDA graph) against the
CPU port of the reference stack (oracle/second_cpu.py), the reference-shaped eager path built on the
compat `spconv` drop-in, and the drop-in modules themselves."""
import numpy as np
import pytest
import torch

import oracle
from oracle import second_cpu
from vision3d_b200 import second, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _match(a, b, tol_xy=2e-2, tol_s=2e-3):
    """fraction of detections in a (boxes, bidx, cidx, scores) that have a partner in b"""
    if len(a[0]) == 0:
        return 1.0
    hit = 0
    for i in range(len(a[0])):
        m = (b[1] == a[1][i]) & (b[2] == a[2][i])
        if not m.any():
            continue
        d = np.abs(b[0][m][:, :2] - a[0][i, :2]).max(1)
        s = np.abs(b[3][m] - a[3][i])
        hit += bool(((d < tol_xy) & (s < tol_s)).any())
    return hit / len(a[0])


@pytest.mark.parametrize("cfg_name,B", [("car", 2), ("three", 1)])
def test_engine_vs_cpu_port(cuda, no_tf32, cfg_name, B):
    cfg = second.car_config() if cfg_name == "car" else second.three_class_config()
    model = second.init_for_benchmark(second.SecondB200(cfg), 1)
    clouds = synth.make_batch(10, B, 16384)
    stages = {}
    cpu_model = second.init_for_benchmark(second.SecondB200(cfg), 1).eval()
    want = second_cpu.infer(cpu_model, clouds, second.make_anchors(cfg), nms_variant="oracle", stages=stages)
    for use_graph in (False, True):
        eng = second.SecondEngine(model, B, B * 16384, cuda, use_graph=use_graph).capture()
        got = eng.infer(clouds)
        # level counts and the dense BEV map come only from vision3d_b200 kernels: tight tolerance
        assert int(eng.n_rows[0].item()) == stages["n_voxels"]
        bev = eng.dense_out.view(B, -1, 200, 176).cpu()
        ref = stages["bev"]
        assert torch.equal(bev != 0, ref != 0) or ((bev != 0) ^ (ref != 0)).float().mean() < 1e-4
        assert (bev - ref).abs().max() <= 1e-4 * ref.abs().max()
        # NMS inside the pipeline: the engine's own candidates through the oracle must give its keep list
        k = int(eng.count.item())
        want_keep = oracle.nms_rotated(eng._nms_in.cpu().numpy(), eng._scores.cpu().numpy(), cfg.NMS_THRESH, 1)
        assert np.array_equal(eng.keep[:k].cpu().numpy(), want_keep)
        # candidates: RPN runs in cuDNN vs CPU MKL, so compare the per-(frame, class) score profiles
        gs = eng._scores.view(B, cfg.NUM_CLASSES, -1).sort(-1, descending=True)[0].cpu()
        cs = stages["cand_scores"].view(B, cfg.NUM_CLASSES, -1).sort(-1, descending=True)[0]
        assert (gs - cs).abs().max() < 5e-3, (gs - cs).abs().max()
        # final detections as matched sets (top-k / NMS decisions may flip on near ties across back ends)
        fa, fb = _match(got, want), _match(want, got)
        assert abs(len(got[0]) - len(want[0])) <= max(3, len(want[0]) // 10), (len(got[0]), len(want[0]))
        assert fa >= 0.8 and fb >= 0.8, (fa, fb, len(got[0]), len(want[0]))
        # replay determinism
        r1 = eng.h_result.clone()
        eng.infer(clouds)
        assert torch.equal(r1, eng.h_result)


def test_eager_compat_path_equals_engine(cuda, no_tf32):
    """SecondB200.inference on the compat spconv drop-in, fed exactly like the reference's Preprocessor
    feeds Second (voxels/coords/occupancy from VoxelGenerator.generate), vs the fused engine."""
    from vision3d_b200.compat import spconv
    cfg = second.car_config()
    model = second.init_for_benchmark(second.SecondB200(cfg), 2).to(cuda).eval()
    clouds = synth.make_batch(20, 2, 16384)
    gen = spconv.utils.VoxelGenerator(cfg.VOXEL_SIZE, cfg.GRID_BOUNDS, cfg.MAX_OCCUPANCY, cfg.MAX_VOXELS)
    f, c, o = [], [], []
    for i, p in enumerate(clouds):
        v, cc, n = gen.generate(p)
        vo, co, no = oracle.voxelize(p, cfg.VOXEL_SIZE, cfg.GRID_BOUNDS, cfg.MAX_OCCUPANCY, cfg.MAX_VOXELS)
        assert np.array_equal(v, vo) and np.array_equal(cc, co) and np.array_equal(n, no)
        f.append(v)
        c.append(np.pad(cc, ((0, 0), (1, 0)), constant_values=i))
        o.append(n)
    item = dict(features=torch.from_numpy(np.concatenate(f)).to(cuda),
                coordinates=torch.from_numpy(np.concatenate(c)).to(cuda),
                occupancy=torch.from_numpy(np.concatenate(o)).to(cuda), batch_size=2,
                anchors=second.make_anchors(cfg).to(cuda))
    with torch.no_grad():
        boxes, bidx, cidx, scores = model.inference(item)
    eng = second.SecondEngine(model, 2, 2 * 16384, cuda, use_graph=True).capture()
    got = eng.infer(clouds)
    assert len(got[0]) == len(boxes)
    order = np.lexsort((-got[3], got[1]))
    ref_order = np.lexsort((-scores.cpu().numpy(), bidx.cpu().numpy()))
    np.testing.assert_allclose(got[0][order], boxes.cpu().numpy()[ref_order], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got[3][order], scores.cpu().numpy()[ref_order], rtol=1e-4, atol=1e-5)


def test_spconv_dropin_api(cuda):
    """The surface vision3d uses: nestable/indexable SparseSequential, SubMConv3d with the stray
    positional stride, tuple kernel/stride/padding, .dense(), eval-mode BN folding == unfolded."""
    from vision3d_b200.compat import spconv
    from torch import nn
    shape, B = [9, 30, 28], 2
    idx = synth.make_clustered_sites(1, 500, shape, B)
    feats = torch.randn(len(idx), 4, device=cuda)
    net = spconv.SparseSequential(
        spconv.SparseSequential(spconv.SubMConv3d(4, 16, 3, 3, indice_key="subm0", bias=False),
                                nn.BatchNorm1d(16, eps=1e-3, momentum=0.01), nn.ReLU()),
        spconv.SparseSequential(spconv.SubMConv3d(16, 16, 3, 3, indice_key="subm0", bias=False),
                                nn.BatchNorm1d(16, eps=1e-3, momentum=0.01), nn.ReLU()),
        spconv.SparseSequential(spconv.SparseConv3d(16, 32, (3, 1, 1), (2, 1, 1), padding=[0, 0, 0], bias=False),
                                nn.BatchNorm1d(32, eps=1e-3, momentum=0.01), nn.ReLU()),
    ).to(cuda)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_()
    assert isinstance(net[0][0], spconv.SubMConv3d) and len(net) == 3
    x = spconv.SparseConvTensor(feats, torch.from_numpy(idx).to(cuda), shape, B)
    net.eval()
    with torch.no_grad():
        y = net(x)
        d = y.dense()
        assert d.shape == (B, 32, 4, 30, 28)
        # unfolded reference: run modules one by one (no BN folding)
        z = spconv.SparseConvTensor(feats, torch.from_numpy(idx).to(cuda), shape, B)
        for blk in net:
            z = blk[0](z)
            z.features = blk[2](blk[1](z.features))
        assert torch.allclose(y.features, z.features, rtol=1e-4, atol=1e-5)
    # training mode: autograd through the sparse conv (torch-composed backward)
    net.train()
    feats.requires_grad_(True)
    out = net(spconv.SparseConvTensor(feats, torch.from_numpy(idx).to(cuda), shape, B))
    out.features.sum().backward()
    assert feats.grad is not None and net[0][0].weight.grad is not None


def test_pointnet2_dropin(cuda):
    from vision3d_b200.compat.pointnet2 import pointnet2_modules, pointnet2_utils
    from copy import deepcopy
    xyz = torch.from_numpy(np.stack([c[:4096, :3] for c in synth.make_batch(0, 2)], 0)).to(cuda)
    idx = pointnet2_utils.furthest_point_sample(xyz, 128)
    assert idx.dtype == torch.int32 and idx.shape == (2, 128)
    kp = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    assert torch.equal(kp[0, 5], xyz[0, idx[0, 5].long()])
    mlps = [[4, 8, 16], [4, 8, 16]]
    sa = pointnet2_modules.PointnetSAModuleMSG(npoint=-1, radii=[0.4, 0.8], nsamples=[16, 32], mlps=deepcopy(mlps),
                                               use_xyz=True).to(cuda).eval()
    feats = torch.randn(2, 4, 4096, device=cuda)
    with torch.no_grad():
        new_xyz, out = sa(xyz, feats, kp)
    assert out.shape == (2, 32, 128) and new_xyz is kp
    # grouped tensor equals the oracle's
    g = sa.groupers[0](xyz, kp, feats).cpu().numpy()
    want_idx = oracle.ball_query(0.4, 16, xyz.cpu().numpy(), kp.cpu().numpy())
    want = oracle.query_and_group(xyz.cpu().numpy(), kp.cpu().numpy(), feats.cpu().numpy(), want_idx)
    assert np.array_equal(g, want)


def test_c_ext_and_searchsorted_shims(cuda):
    from vision3d_b200.compat import c_ext, torchsearchsorted
    b = torch.tensor([[0, 0, 2, 2, 0], [1, 1, 2, 2, 0], [0, 0, 2, 2, 45], [10, 10, 2, 2, 0]], dtype=torch.float32,
                     device=cuda)
    iou = c_ext.box_iou_rotated(b[:1], b)
    assert abs(float(iou[0, 1]) - 1 / 7) < 1e-6 and abs(float(iou[0, 2]) - 0.70710678) < 1e-6
    keep = c_ext.nms_rotated(b, torch.tensor([.9, .8, .7, .6], device=cuda), 0.1)
    assert keep.tolist() == [0, 3] and keep.dtype == torch.int64 and keep.is_cuda
    assert c_ext.get_cuda_version().startswith("12.")
    a = torch.tensor([[0, 0, 1, 1, 1, 3]], dtype=torch.int32, device=cuda)
    v = torch.arange(5, dtype=torch.int32, device=cuda)[None]
    assert torchsearchsorted.searchsorted(a, v).tolist() == [[0, 2, 5, 5, 6]]


def test_bev_nhwc_matches_dense_and_engine_modes_agree(cuda, no_tf32):
    """The channels-last BEV writer == dense().view(B, C*D, H, W); engine results do not depend on the RPN
    execution mode (module / BN-folded fused / fused channels_last)."""
    from vision3d_b200 import ops
    rng = np.random.default_rng(0)
    shape, B, C = [2, 200, 176], 2, 64
    idx = synth.make_active_sites(3, 6000, shape, B)
    feat = torch.from_numpy(rng.normal(size=(len(idx), C)).astype(np.float32)).to(cuda)
    ind = torch.from_numpy(idx).to(cuda)
    n_rows = torch.tensor([len(idx)], dtype=torch.int32, device=cuda)
    dense = ops.sparse_to_dense(feat, ind, n_rows, len(idx), B, shape).view(B, C * 2, 200, 176)
    nhwc = ops.sparse_to_bev_nhwc(feat, ind, n_rows, len(idx), B, shape)
    assert nhwc.is_contiguous(memory_format=torch.channels_last) and torch.equal(nhwc, dense)

    cfg = second.car_config()
    model = second.init_for_benchmark(second.SecondB200(cfg), 3)
    clouds = synth.make_batch(30, 2, 16384)
    outs = []
    for mode in ("module", "fused", "fused_nhwc"):
        eng = second.SecondEngine(model, 2, 2 * 16384, cuda, use_graph=True, rpn_mode=mode).capture()
        outs.append(eng.infer(clouds))
    for o in outs[1:]:
        assert abs(len(o[0]) - len(outs[0][0])) <= 2
        assert _match(o, outs[0]) >= 0.9 and _match(outs[0], o) >= 0.9


def test_fused_head_kernels_match_torch_expressions(cuda, no_tf32):
    """v3d_second_head_decode / v3d_pack_detections vs the reference's torch expression sequence."""
    cfg = second.three_class_config()
    model = second.init_for_benchmark(second.SecondB200(cfg), 4)
    clouds = synth.make_batch(40, 2, 16384)
    res = []
    for fused in (False, True):
        eng = second.SecondEngine(model, 2, 2 * 16384, cuda, use_graph=False, fused_head=fused).capture()
        out = eng.infer(clouds)
        res.append((out, eng._boxes.clone(), eng._nms_in.clone(), eng._scores.clone(), eng.h_result.clone()))
    (o0, b0, n0, s0, r0), (o1, b1, n1, s1, r1) = res
    assert torch.equal(s0, s1)
    assert torch.allclose(b0, b1, rtol=1e-5, atol=1e-5)
    assert torch.allclose(n0, n1, rtol=1e-5, atol=1e-3)   # offsets ~1e4: fp32 spacing ~1e-3
    assert len(o0[0]) == len(o1[0])
    n_kept, N = int(r0[-1, 0]), r0.shape[0] - 1
    assert n_kept == int(r1[-1, 0]) and n_kept > 0
    # rows past the kept count are padding (unspecified in the torch path, zero in the fused one)
    np.testing.assert_allclose(r0[:n_kept].numpy(), r1[:n_kept].numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_array_equal(r0[:N, 10].numpy(), r1[:N, 10].numpy())   # valid flags
    np.testing.assert_array_equal(r0[N, :6].numpy(), r1[N, :6].numpy())     # counters row


def test_topk_rows_matches_torch_topk(cuda):
    from vision3d_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    vals = torch.randn((6, 70400), generator=g)
    vals[1] = torch.round(vals[1] * 4) / 4          # heavy ties everywhere, also at the k-th value
    vals[2, :] = 0.5                                # all equal: the k lowest indices win
    vals[3, 100:40000] = -7.25                      # clustered top bytes
    vals[4] = vals[4].abs() * 1e-30                 # tiny positive numbers / denormal range
    vals[5, 5] = float("inf")
    v = vals.to(cuda)
    for k in (1, 100, 256):
        tv, ti = ops.topk_rows(v, k)
        rv, _ = torch.topk(v, k, dim=-1)
        assert torch.equal(tv, rv)                                  # same values, sorted descending
        assert torch.equal(torch.gather(v, 1, ti), tv)              # indices point at those values
        assert all(len(set(r.tolist())) == k for r in ti.cpu())     # no duplicates
        # ties -> lower index first: inside a run of equal values the indices ascend
        same = tv[:, 1:] == tv[:, :-1]
        assert bool(((ti[:, 1:] > ti[:, :-1]) | ~same).all())
    # the selected set under ties is the lowest indices
    tv, ti = ops.topk_rows(v[2:3], 100)
    assert ti.cpu().flatten().tolist() == list(range(100))


def test_native_head_kernels_match_torch_ops(cuda, no_tf32):
    """cls logits / top-k / reg gather / decode on the NHWC map vs conv2d + sigmoid + topk + gather (proposal.py:61-78)."""
    from vision3d_b200 import ops
    cfg = second.three_class_config()
    torch.manual_seed(5)
    head = second.HeadB200(cfg).to(cuda).eval()
    with torch.no_grad():
        head.conv_cls.weight.normal_(std=0.05)
        head.conv_reg.weight.normal_(std=0.05)
        head.conv_reg.bias.normal_(std=0.1)
    B, ny, nx, n_cls, n_yaw, K = 2, 200, 176, cfg.NUM_CLASSES, cfg.NUM_YAW, cfg.TOPK
    fmap = torch.randn((B, 128, ny, nx), device=cuda).contiguous(memory_format=torch.channels_last)
    anchors = second.make_anchors(cfg).to(cuda)
    with torch.no_grad():
        cls_ref = head.conv_cls(fmap).reshape(B, n_cls * n_yaw, -1)
        reg_ref = head.conv_reg(fmap)
    w_cls = head.conv_cls.weight.detach().reshape(n_cls * n_yaw, -1).contiguous()
    logits = ops.head_cls_logits(fmap, w_cls, head.conv_cls.bias.detach())
    assert torch.allclose(logits, cls_ref, rtol=1e-5, atol=1e-5)
    top, a_idx = ops.topk_rows(logits.view(B * n_cls, -1), K)
    N = B * n_cls * K
    deltas, scores = torch.empty((N, 7), device=cuda), torch.empty(N, device=cuda)
    w_reg = head.conv_reg.weight.detach().reshape(n_cls * n_yaw * 7, -1).contiguous()
    ops.head_reg_gather(fmap, w_reg, head.conv_reg.bias.detach(), top, a_idx, n_cls, n_yaw, K, deltas, scores)
    assert torch.equal(scores, torch.sigmoid(top).reshape(-1))      # same expression as torch.sigmoid
    # reference gather of the regression map at the same anchors (reshape order of proposal.py:20-26)
    reg5 = reg_ref.reshape(B, n_cls, 7, n_yaw * ny * nx)
    want = torch.gather(reg5, 3, a_idx.view(B, n_cls, 1, K).expand(-1, -1, 7, -1)).permute(0, 1, 3, 2).reshape(N, 7)
    assert torch.allclose(deltas, want, rtol=1e-5, atol=1e-5)
    boxes, nms_in = torch.empty((N, 7), device=cuda), torch.empty((N, 5), device=cuda)
    ops.second_head_decode_compact(deltas, anchors, a_idx, B, n_cls, n_yaw, ny, nx, K, boxes, nms_in)
    b2, n2 = ops.second_head_decode(reg_ref, anchors, a_idx.view(B, n_cls, K).contiguous(), n_cls, n_yaw, K)
    assert torch.allclose(boxes, b2, rtol=1e-5, atol=1e-5) and torch.allclose(nms_in, n2, rtol=1e-5, atol=1e-3)


def test_engine_native_head_vs_torch_head(cuda, no_tf32):
    cfg = second.car_config()
    model = second.init_for_benchmark(second.SecondB200(cfg), 2)
    clouds = synth.make_batch(60, 2, 16384)
    res = []
    for fused in (False, True):
        eng = second.SecondEngine(model, 2, 2 * 16384, cuda, use_graph=fused, fused_head=fused,
                                  rpn_mode="fused_nhwc").capture()
        out = eng.infer(clouds)
        res.append((out, eng._scores.clone().view(2, -1), eng._boxes.clone().view(2, -1, 7)))
    (o0, s0, b0), (o1, s1, b1) = res
    assert torch.allclose(s0, s1, rtol=0, atol=2e-6)               # candidate scores, sorted per frame
    close = torch.isclose(b0, b1, rtol=1e-4, atol=1e-4).all(-1)    # same candidates (order may differ on near ties)
    assert close.float().mean().item() > 0.97
    for f in range(2):
        assert abs(len(o0[f]) - len(o1[f])) <= 1
