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
