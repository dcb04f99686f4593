"""CPU checks of the PV-RCNN keypoint-stage mirror (rows a11 / a15 / f3): the pure-torch glue of vision3d_b200.pvrcnn
and the oracle's restatements against golden vectors generated from the reference's OWN Python
(tests/golden/make_pvrcnn_golden.py), the pool-invariance argument that lets the product skip pad_batch, and the
oracle composition of the whole stage at a small size."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import pvrcnn_cpu
from vision3d_b200 import pvrcnn, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "pvrcnn_golden.npz"))


def test_sample_gridpoints_matches_reference_golden(gold):
    got = pvrcnn.sample_gridpoints(torch.from_numpy(gold["grid_boxes"]), torch.from_numpy(gold["grid_noise"]))
    assert np.array_equal(got.numpy(), gold["grid_points"])          # same torch expressions: bit-identical


def test_bev_gather_matches_reference_golden(gold):
    cfg = pvrcnn.PVRCNNConfig()
    got = pvrcnn.bev_gather(cfg, torch.from_numpy(gold["bev_map"]), torch.from_numpy(gold["bev_kp"]))
    assert np.array_equal(got.numpy(), gold["bev_out"])


def test_to_global_and_pad_amounts_match_reference_golden(gold):
    cfg = pvrcnn.PVRCNNConfig()
    for s in cfg.STRIDES:
        xyz = pvrcnn_cpu.to_global(gold["glob_idx"], cfg.VOXEL_SIZE, s, cfg.GRID_BOUNDS[:3])
        assert np.array_equal(xyz, gold["glob_xyz_s%d" % s]), s
    starts = np.searchsorted(gold["pad_batch_index"], np.arange(5))
    cnt = starts[1:] - starts[:-1]
    assert np.array_equal(cnt, gold["pad_count"]) and np.array_equal(cnt.max() - cnt, gold["pad_pad"])


def test_mlp_mirror_matches_reference_golden(gold):
    mlp = pvrcnn.MLPB200([48, 32, 16])
    keys = [str(k) for k in gold["mlp_keys"]]
    assert keys == list(mlp.state_dict().keys())                     # reference state_dicts load by name
    mlp.load_state_dict({k: torch.from_numpy(gold["mlp_" + k.replace(".", "_")]) for k in keys})
    with torch.no_grad():
        y = mlp(torch.from_numpy(gold["mlp_x"]).clone())
    assert np.array_equal(y.numpy(), gold["mlp_y"])


def test_pad_batch_is_pool_invariant():
    """pad_batch (sparse_cnn.py:118-126) appends random duplicates of a frame's own rows. With xyz and features padded
    by the SAME picks, ball_query -> group -> max over samples is unchanged for ANY picks -- which is why the product
    queries the ragged levels in place (ops.ball_query_msg with row_offsets) instead of drawing a padding."""
    rng = np.random.default_rng(0)
    n, C, M = 300, 5, 40
    xyz = rng.uniform(0, 4, (1, n, 3)).astype(np.float32)
    feat = rng.normal(size=(1, C, n)).astype(np.float32)
    q = rng.uniform(0, 4, (1, M, 3)).astype(np.float32)
    for radius, ns in [(0.5, 16), (1.5, 32), (0.05, 16)]:
        idx = oracle.ball_query(radius, ns, xyz, q)
        base = oracle.query_and_group(xyz, q, feat, idx).max(-1)
        for trial in range(3):
            pick = rng.integers(0, n, 150)
            xyz_p = np.concatenate([xyz, xyz[:, pick]], 1)
            feat_p = np.concatenate([feat, feat[:, :, pick]], 2)
            idx_p = oracle.ball_query(radius, ns, xyz_p, q)
            pooled = oracle.query_and_group(xyz_p, q, feat_p, idx_p).max(-1)
            hit = (idx_p != 0).any(-1) | (np.linalg.norm(xyz[0, 0] - q[0], axis=-1) < radius)
            assert np.array_equal(pooled[..., hit[0]], base[..., hit[0]])
            assert np.array_equal(pooled, base)


def test_oracle_keypoint_stage_small():
    cfg = pvrcnn.PVRCNNConfig()
    cfg.NUM_KEYPOINTS = 128
    model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0).eval()
    # parameter names of the reference PV_RCNN tree (detector/model.py:24-32, roi_grid_pool.py:23-24)
    names = set(model.state_dict().keys())
    for k in ["pnets.0.mlps.0.layer0.conv.weight", "pnets.4.mlps.1.layer1.bn.bn.running_var",
              "roi_grid_pool.pnet.mlps.0.layer0.conv.weight", "roi_grid_pool.reduction.linear_1.weight",
              "cnn.blocks.0.0.0.weight", "proposal_layer.conv_cls.bias"]:
        assert k in names, k
    assert tuple(model.state_dict()["roi_grid_pool.pnet.mlps.0.layer0.conv.weight"].shape) == (192, 515, 1, 1)
    clouds = synth.make_batch(0, 2, 4096)
    props = pvrcnn.make_proposals(clouds, 4, 0)
    grid = pvrcnn.sample_gridpoints(torch.from_numpy(props), pvrcnn.make_grid_noise(2, 4, 16, 0)).reshape(2, -1, 3).numpy()
    st = {}
    out = pvrcnn_cpu.keypoint_stage(model, clouds, grid, stages=st)
    assert out.shape == (2, 4, 256) and st["kp_features"].shape == (2, 512, 128)
    assert st["kp_idx"].shape == (2, 128) and (st["kp_idx"][:, 0] == 0).all()
    assert len(set(st["kp_idx"][0].tolist())) == 128 and float(out.abs().max()) > 0
    # keypoints are cloud points; every RoI ball-query index addresses a keypoint
    assert np.array_equal(st["keypoints"][1], clouds[1][st["kp_idx"][1], :3])
    assert all(int(np.stack(v).max()) < 128 for v in st["roi_idx"].values())


# ---------------------------------------------------------------------------------------------------------------
# training-side IoU consumer: ProposalTargetAssigner mirror pinned to the reference's own run
# ---------------------------------------------------------------------------------------------------------------
def _dense_targets(item):
    G_cls, M_cls, G_reg = item["G_cls"], item["M_cls"], item["G_reg"]
    pos = torch.nonzero(G_cls.reshape(-1) == 1).squeeze(1)
    ign = torch.nonzero(~M_cls.reshape(-1)).squeeze(1)
    return pos.numpy(), ign.numpy(), G_reg.reshape(-1, 7)[pos].numpy(), int(item["M_reg"].sum())


def test_target_assigner_mirror_matches_reference_golden():
    """targets.ProposalTargetAssignerB200 (reference expression path, IoU = the oracle's restatement of the reference
    CPU op) reproduces the reference's own ProposalTargetAssigner.forward: positives, ignored anchors, matched boxes and
    encoded regression targets (tests/golden/make_targets_golden.py)."""
    from vision3d_b200 import second, targets
    gold = np.load(os.path.join(ROOT, "tests", "golden", "targets_golden.npz"))
    cfg = second.three_class_config()

    def iou_cpu(a, b):
        return torch.from_numpy(oracle.box_iou_rotated(a.numpy(), b.numpy(), 0))   # variant 0 = reference CPU build

    assigner = targets.ProposalTargetAssignerB200(cfg, fused=False, iou_fn=iou_cpu)
    for tag in ("a", "b"):
        item = dict(boxes=torch.from_numpy(gold[tag + "_boxes"]), class_idx=torch.from_numpy(gold[tag + "_class_idx"]))
        with torch.no_grad():
            assigner(item)
        assert tuple(item["G_cls"].shape) == tuple(gold[tag + "_shape"])
        pos, ign, reg, n_reg = _dense_targets(item)
        assert np.array_equal(pos, gold[tag + "_pos"]) and np.array_equal(ign, gold[tag + "_ign"])
        assert n_reg == int(gold[tag + "_n_reg_mask"])
        np.testing.assert_allclose(reg, gold[tag + "_reg_pos"], rtol=0, atol=1e-6)
        box_idx, _ = assigner.match_all_classes(item["boxes"], item["class_idx"])
        assert np.array_equal(box_idx.reshape(-1)[torch.from_numpy(pos)].numpy(), gold[tag + "_pos_box"])


def test_encode_decode_roundtrip_and_matcher_strata():
    from vision3d_b200 import second, targets
    g = torch.Generator().manual_seed(0)
    anchors = second.make_anchors(second.three_class_config()).view(-1, 7)[:64]
    boxes = anchors + torch.randn((64, 7), generator=g) * 0.1
    boxes[:, 6] = anchors[:, 6] + torch.rand(64, generator=g) * 3.0       # encode wraps yaw differences into [0, pi)
    back = second.decode_boxes(targets.encode_boxes(boxes, anchors), anchors)
    assert torch.allclose(back, boxes, atol=1e-5)
    m = targets.Matcher([0.45, 0.6], [0, -1, 1])
    q = torch.tensor([[0.1, 0.5, 0.7, 0.45, 0.6], [0.2, 0.1, 0.0, 0.0, 0.0]])
    matches, labels = m(q)
    assert matches.tolist() == [1, 0, 0, 0, 0] and labels.tolist() == [0, -1, 1, -1, 1]
    e, l = m(torch.zeros((0, 3)))
    assert e.tolist() == [0, 0, 0] and l.tolist() == [0, 0, 0]
