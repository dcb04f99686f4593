"""Parity of the BENCHMARKED configurations, in the mode they are measured.

bench.py measures workload "t16" (SECOND car-only, 16 frames per GPU) and "c5" (3-class, global batch 64 -> 8
frames per GPU on 8 GPUs), both built by `second.make_bench_engine` with `second.BENCH_ENGINE_FLAGS` (CUDA graph,
tcgen05 sparse conv, channels_last cuDNN RPN with TF32 as torch defaults it, native head, grouped NMS). These
tests build the SAME engine through the SAME factory and compare stage by stage with the CPU port of the
reference stack (oracle/second_cpu.py, reference detector/second.py:20-35, proposal.py:47-80):

  voxel rows / coords, active sites and indices of every level ......... exact
  dense BEV map (14 sparse layers, bf16x3 tensor-core path) ............. <= 1e-4 of the map's scale
  RPN (cuDNN, TF32 allowed -- the only non-fp32 arithmetic on the path) .. <= RPN_TF32_TOL of the map's scale
                                                                           against the fp32 CPU RPN on the
                                                                           engine's own BEV (stated tolerance)
  head: the engine's own RPN output through the CPU head (1x1 convs, sigmoid, top-k, gather, decode):
        same anchors selected (boundary ties aside), scores <= 2e-6, boxes <= 1e-4
  NMS : the engine's own candidates through the oracle NMS -> identical keep list, identical final tuple
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import second_cpu
from vision3d_b200 import second, synth

pytestmark = pytest.mark.gpu

RPN_TF32_TOL = 1e-2   # of max|fmap|; cuDNN TF32 (10-bit mantissa operands, fp32 accumulate) over 7 conv layers


@pytest.mark.parametrize("workload,frames", [("t16", 16), ("c5", 8)])
def test_benchmarked_engine_stage_by_stage(cuda, workload, frames):
    eng, model, cfg = second.make_bench_engine(workload, frames, cuda)
    assert eng.graph is not None and eng.rpn_mode == "fused_nhwc" and eng.fused_head and eng.grouped_nms
    assert eng.overlap_rulebooks   # rule-book chain on the second stream, as benchmarked
    B, n_cls, K = frames, cfg.NUM_CLASSES, cfg.TOPK
    clouds = synth.make_batch(0, B, second.PTS_PER_FRAME)   # bench.py's first batch of rank 0
    got = eng.infer(clouds)

    # ---- CPU port, sparse half
    cpu_model = second.init_for_benchmark(second.SecondB200(cfg), 0).eval()
    v, c, n = second_cpu.voxelize_batch(cfg, clouds)
    m = len(c)
    assert int(eng.n_rows[0].item()) == m
    assert np.array_equal(eng.vox_out["coords"][:m].cpu().numpy(), c)
    assert np.array_equal(eng.vox_out["num_points"][:m].cpu().numpy(), n)
    assert np.array_equal(eng.vox_out["voxels"][:m].cpu().numpy(), v)
    feat = torch.from_numpy(oracle.vfe_mean(v, n))
    assert np.array_equal(eng.vox_out["mean"][:m].cpu().numpy(), feat.numpy())
    with torch.no_grad():
        bev_ref, levels, _ = second_cpu.sparse_middle(cpu_model, feat, c, B, return_levels=True)
    for lv in range(1, 5):
        idx_ref = levels[lv][0]
        assert int(eng.n_rows[lv].item()) == len(idx_ref), lv
        assert np.array_equal(eng.indices[lv][:len(idx_ref)].cpu().numpy(), idx_ref), lv
    bev = eng.bev_nhwc.cpu().contiguous()          # logical (B, 128, 200, 176)
    scale = bev_ref.abs().max()
    assert torch.equal(bev != 0, bev_ref != 0) or ((bev != 0) ^ (bev_ref != 0)).float().mean() < 1e-5
    err_bev = float((bev - bev_ref).abs().max() / scale)
    assert err_bev <= 1e-4, err_bev

    # ---- RPN: cuDNN with the TF32 setting bench.py runs under, vs fp32 on the CPU, SAME input (engine's BEV)
    with torch.no_grad():
        fmap_ref = cpu_model.rpn(bev)
    fmap = eng._fmap.cpu().contiguous()
    err_rpn = float((fmap - fmap_ref).abs().max() / fmap_ref.abs().max())
    print("workload %s: bev err %.2e, rpn err %.2e (cudnn.allow_tf32=%s)" % (
        workload, err_bev, err_rpn, torch.backends.cudnn.allow_tf32))
    assert err_rpn <= RPN_TF32_TOL, err_rpn

    # ---- head: the ENGINE's RPN output through the CPU head (proposal.py:61-78)
    anchors = second.make_anchors(cfg)
    with torch.no_grad():
        boxes_ref, scores_ref = cpu_model.head.candidates(fmap, anchors)            # (B, n_cls, K, 7), (B, n_cls, K)
        cls_map, _ = cpu_model.head(fmap)
        _, a_ref = cls_map.sigmoid().reshape(B, n_cls, -1).topk(K, -1)
    s_gpu = eng._scores.view(B, n_cls, K).cpu()
    b_gpu = eng._boxes.view(B, n_cls, K, 7).cpu()
    a_gpu = eng._a_idx.view(B, n_cls, K).cpu()
    assert (s_gpu - scores_ref).abs().max() <= 2e-6      # both sorted descending
    n_swapped = 0
    for b in range(B):
        for k in range(n_cls):
            ga, ra = a_gpu[b, k].tolist(), a_ref[b, k].tolist()
            common = set(ga) & set(ra)
            # anchors present on one side only must sit at the top-k boundary (score ties within rounding)
            kth = float(scores_ref[b, k, -1])
            for a, sc in list(zip(ga, s_gpu[b, k].tolist())) + list(zip(ra, scores_ref[b, k].tolist())):
                if a not in common:
                    n_swapped += 1
                    assert abs(sc - kth) <= 2e-6, (b, k, a, sc, kth)
            gi = {a: i for i, a in enumerate(ga)}
            ri = {a: i for i, a in enumerate(ra)}
            sel = sorted(common)
            bg = b_gpu[b, k][[gi[a] for a in sel]]
            br = boxes_ref[b, k][[ri[a] for a in sel]]
            assert torch.allclose(bg, br, rtol=1e-4, atol=1e-4), (b, k, (bg - br).abs().max())
    assert n_swapped <= 2 * B * n_cls   # boundary ties are rare

    # ---- NMS + final tuple: the engine's own candidates through the oracle, exactly
    kcount = int(eng.count.item())
    nms_in, scores = eng._nms_in.cpu().numpy(), eng._scores.cpu().numpy()
    want_keep = oracle.nms_rotated(nms_in, scores, cfg.NMS_THRESH, 1)
    assert np.array_equal(eng.keep[:kcount].cpu().numpy(), want_keep)
    boxes_all = eng._boxes.cpu().numpy()
    c_idx = np.tile(np.repeat(np.arange(n_cls), K), B)
    b_idx = np.repeat(np.arange(B), n_cls * K)
    thr = np.array([a["score_thresh"] for a in cfg.ANCHORS], np.float32)
    keep = want_keep[scores[want_keep] > thr[c_idx[want_keep]]]
    assert np.array_equal(got[0], boxes_all[keep]) and np.array_equal(got[3], scores[keep])
    assert np.array_equal(got[1], b_idx[keep]) and np.array_equal(got[2], c_idx[keep])
    assert len(got[0]) > 0

    # ---- the oracle's own offsets trick on the engine's boxes reproduces the NMS input (group separation)
    g = torch.from_numpy(c_idx + n_cls * b_idx)
    want_in = second.group_offsets(torch.from_numpy(boxes_all[:, [0, 1, 3, 4, 6]]), g).numpy()
    np.testing.assert_allclose(nms_in, want_in, rtol=0, atol=2e-3)   # offsets ~1e4: fp32 spacing ~1e-3


def test_level_capacity_overflow_raises_and_stays_in_bounds(cuda):
    """a level that outgrows its row capacity must (a) raise from finalize(), and (b) never index
    past the rule / feature buffers on the way there (ranks >= capacity read as 'no neighbour')."""
    from vision3d_b200 import ops
    cfg = second.car_config()
    model = second.init_for_benchmark(second.SecondB200(cfg), 0)
    clouds = synth.make_batch(0, 1, 16384)
    eng = second.SecondEngine(model, 1, 16384, cuda, level_caps=[None, 3000, 40000, 24000, 24000],
                              use_graph=False).capture()
    with pytest.raises(ops.V3DError, match="capacity"):
        eng.infer(clouds)
    torch.cuda.synchronize()

    # kernel level, with canaries: level 1 of a real cloud (~27k sites) into capacity 4096
    v, c, n = oracle.voxelize(clouds[0], cfg.VOXEL_SIZE, cfg.GRID_BOUNDS, 5, 20000)
    idx = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    shape = [41, 1600, 1408]
    ind = torch.from_numpy(idx).to(cuda)
    n_rows = torch.tensor([len(idx)], dtype=torch.int32, device=cuda)
    table = ops.SiteTable(len(idx), cuda).build(ind, n_rows, shape)
    cap, stride, CANARY = 4096, 4096 + 512, -7
    nbr = torch.full((27, stride), CANARY, dtype=torch.int32, device=cuda)
    out_idx = torch.full((cap + 512, 4), CANARY, dtype=torch.int32, device=cuda)
    n_out = torch.zeros(1, dtype=torch.int32, device=cuda)
    ws = ops.ConvRulebookWorkspace(1, [21, 800, 704], cap, 27, cuda)
    ops.rulebook_conv(table, ind, n_rows, 1, shape, 3, 2, 1, 1, cap, out_idx, n_out, nbr, ws)
    assert int(n_out.item()) > cap                                   # overflow is reported un-clamped
    assert bool((nbr[:, cap:] == CANARY).all()) and bool((out_idx[cap:] == CANARY).all())
    # SubM rule book of the overflowed level: neighbour ranks >= cap must read as -1, mirrors stay inside
    nbr2 = torch.full((27, stride), CANARY, dtype=torch.int32, device=cuda)
    nbr2[14:, :cap] = 0
    ops.rulebook_subm(ws, out_idx[:cap], n_out, [21, 800, 704], 3, 1, nbr2[:, :stride], capacity=cap)
    got = nbr2[:, :cap]
    assert int(got.max().item()) < cap and int(got.min().item()) >= -1
    assert bool((nbr2[:14, cap:] == CANARY).all())                   # lower half: nothing written past cap
    # and the first `cap` rows agree with the oracle restricted to those rows
    oi, _, _ = oracle.rulebook_conv(idx, shape, 3, 2, 1)
    want = oracle.rulebook_subm(oi, [21, 800, 704], 3)[:, :cap]
    want = np.where(want >= cap, -1, want)
    assert np.array_equal(got.cpu().numpy(), want)


def test_error_paths(cuda):
    """Bad arguments are refused with an exception before any kernel runs (reference: AT_ASSERTM -> RuntimeError)."""
    from vision3d_b200 import _lib, ops
    lib = _lib.load()
    d = torch.rand((128, 5), device=cuda)
    s = torch.rand(128, device=cuda)
    keep = torch.empty(128, dtype=torch.int64, device=cuda)
    cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
    small = torch.empty(64, dtype=torch.uint8, device=cuda)
    with pytest.raises(ops.V3DError, match="workspace"):
        ops.nms_rotated_padded(d, s, 0.1, workspace=small, keep=keep, count=cnt)
    with pytest.raises(ops.V3DError, match="workspace"):
        ops.nms_rotated_padded(d, s, 0.1, workspace=small, keep=keep, count=cnt, group_size=64)
    # documented limit of the device-side greedy scan (the reference's host loop takes any N): N <= 65536
    assert lib.v3d_nms_rotated(1, 1, 65537, 0.1, 1, 1, 1, 1 << 40, 0) == -1   # V3D_ERR_INVALID_ARGUMENT, no launch
    with pytest.raises(ops.V3DError):
        ops.nms_rotated_padded(d, s[:100], 0.1)
    with pytest.raises(ops.V3DError, match="CUDA tensor"):
        ops.box_iou_rotated(torch.rand(3, 5), torch.rand(3, 5))
    with pytest.raises(ops.V3DError, match="last dimension"):
        ops.box_iou_rotated(torch.rand((3, 4), device=cuda), torch.rand((3, 5), device=cuda))
    # conv rule-book workspace too small
    ind = torch.zeros((16, 4), dtype=torch.int32, device=cuda)
    n_rows = torch.tensor([16], dtype=torch.int32, device=cuda)
    table = ops.SiteTable(16, cuda).build(ind, n_rows, [8, 8, 8])
    ws = ops.ConvRulebookWorkspace(1, [4, 4, 4], 64, 27, cuda)
    ws.buf = ws.buf[:128]
    with pytest.raises(ops.V3DError, match="workspace"):
        ops.rulebook_conv(table, ind, n_rows, 1, [8, 8, 8], 3, 2, 1, 1, 64, workspace=ws)
    # sparse conv: unsupported tensor-core shape is refused by the C-ABI, not silently rerouted
    assert lib.v3d_sparse_conv_prepared_bytes(27, 24, 64) == 0
    # points capacity of the engine
    eng = second.SecondEngine(second.SecondB200(second.car_config()), 1, 1000, cuda, use_graph=False)
    with pytest.raises(AssertionError, match="capacity"):
        eng.load_host([synth.make_cloud(0, 2000)])


def test_iou_many_rows_chunked_launch(cuda):
    """M beyond one launch's 65535 row tiles (the reference transposes instead, box_iou_rotated_cuda.cu:84-95):
    rows are cut over several launches; spot rows from every chunk are bit-exact vs the oracle."""
    from vision3d_b200 import ops
    rng = np.random.default_rng(5)
    M, N = 65535 * 16 + 40, 3
    b1 = np.stack([rng.uniform(0, 4, M), rng.uniform(0, 4, M), rng.uniform(0.5, 3, M), rng.uniform(0.5, 3, M),
                   rng.uniform(-90, 90, M)], 1).astype(np.float32)
    b2 = b1[:N].copy()
    got = ops.box_iou_rotated(torch.from_numpy(b1).to(cuda), torch.from_numpy(b2).to(cuda)).cpu().numpy()
    sel = np.r_[0:64, M // 2:M // 2 + 64, 65535 * 16 - 8:M]
    want = oracle.box_iou_rotated(b1[sel], b2, 1)
    assert np.array_equal(got[sel].view(np.uint32), want.view(np.uint32))


def test_rulebook_chain_on_second_stream_is_bit_identical(cuda):
    """The captured step with the site-table / rule-book chain on its own stream (a parallel graph branch; the
    convolutions wait per level) returns exactly what the single-stream step returns, replay after replay, on
    alternating inputs (a missing dependency would show up as stale rule tables of the previous batch)."""
    frames = 4
    e1, _, _ = second.make_bench_engine("t16", frames, cuda, overlap_rulebooks=True)
    e0, _, _ = second.make_bench_engine("t16", frames, cuda, overlap_rulebooks=False)
    assert e1.graph is not None and e0.graph is not None
    for seed in (0, 1, 0, 2, 1):
        clouds = synth.make_batch(seed, frames, second.PTS_PER_FRAME)
        a, b = e1.infer(clouds), e0.infer(clouds)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert torch.equal(e1.result, e0.result)
        for lv in range(5):
            n = int(e0.n_rows[lv].item())
            assert int(e1.n_rows[lv].item()) == n and torch.equal(e1.indices[lv][:n], e0.indices[lv][:n])
