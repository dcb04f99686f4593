"""GPU parity tests (-m gpu): every kernel, called through the C-ABI (vision3d_b200.ops -> ctypes ->
libv3d_b200.so), against the CPU oracle on the same seeded inputs and against the committed golden
vectors generated from the reference's own compiled sources.

Bars: bit-exact for IoU values (vs oracle variant 1 = reference header as nvcc sees it, no FMA),
NMS keep indices, voxelize outputs, rule books, FPS / ball-query / grouping indices and copies;
<= 1e-4 relative (to the tensor's max magnitude) for the sparse convolution."""
import numpy as np
import pytest
import torch

import oracle
from vision3d_b200 import synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


CASES = ["known", "spread", "cluster", "degenerate", "offset"]


# ------------------------------------------------------------------------------------ IoU / NMS
@pytest.mark.parametrize("case", CASES)
def test_iou_golden_bit_exact(cuda, golden, case):
    from vision3d_b200 import ops
    b = golden[case + "_boxes"]
    got = ops.box_iou_rotated(_t(b, cuda), _t(b, cuda)).cpu().numpy()
    assert np.array_equal(_bits(got), _bits(golden[case + "_iou_nvcc"]))


@pytest.mark.parametrize("m,n", [(1, 1), (3, 130), (17, 1000), (300, 257), (10, 7040)])
def test_iou_random_vs_oracle_bit_exact(cuda, m, n):
    from vision3d_b200 import ops
    rng = np.random.default_rng(m * 1000 + n)

    def mk(k):
        return np.stack([rng.uniform(0, 20, k), rng.uniform(-10, 10, k), rng.uniform(0.2, 3, k),
                         rng.uniform(0.2, 5, k), rng.uniform(-200, 200, k)], 1).astype(np.float32)
    b1, b2 = mk(m), mk(n)
    got = ops.box_iou_rotated(_t(b1, cuda), _t(b2, cuda)).cpu().numpy()
    want = oracle.box_iou_rotated(b1, b2, 1)
    assert m * n < 1000 or (want > 0).sum() > 0
    assert np.array_equal(_bits(got), _bits(want))


def test_iou_empty(cuda):
    from vision3d_b200 import ops
    out = ops.box_iou_rotated(torch.zeros((0, 5), device=cuda), torch.zeros((4, 5), device=cuda))
    assert out.shape == (0, 4)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("thr", [0.01, 0.1, 0.5])
def test_nms_golden_exact(cuda, golden, case, thr):
    from vision3d_b200 import ops
    b, s = golden[case + "_boxes"], golden[case + "_scores"]
    keep = ops.nms_rotated(_t(b, cuda), _t(s, cuda), thr).cpu().numpy()
    assert keep.dtype == np.int64
    assert np.array_equal(keep, golden["keep_nvcc_%s_%03d" % (case, int(thr * 100))])


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 100, 129, 1600, 2400])
def test_nms_grouped_vs_oracle_exact(cuda, n):
    """KITTI-shape boxes, groups of 100 with the wrapper's fp32 offsets, thr 0.01 (proposal.py:54)."""
    from vision3d_b200 import ops
    boxes, scores, idxs = synth.make_nms_boxes(n, n)
    b = synth.apply_group_offsets(boxes, idxs)
    keep = ops.nms_rotated(_t(b, cuda), _t(scores, cuda), 0.01).cpu().numpy()
    want = oracle.nms_rotated(b, scores, 0.01, 1)
    assert np.array_equal(keep, want)
    # descending score order, as the reference returns (nms_rotated_cuda.cu:131-133)
    assert np.all(np.diff(scores[keep]) <= 0)


def test_nms_dense_cluster_and_ties(cuda):
    from vision3d_b200 import ops
    rng = np.random.default_rng(9)
    n = 700
    b = np.stack([rng.uniform(0, 12, n), rng.uniform(-6, 6, n), np.full(n, 1.6), np.full(n, 3.9),
                  rng.uniform(0, 180, n)], 1).astype(np.float32)
    s = np.round(rng.random(n), 2).astype(np.float32)  # many exact score ties -> stable index order
    for thr in (0.01, 0.3, 0.7):
        keep = ops.nms_rotated(_t(b, cuda), _t(s, cuda), thr).cpu().numpy()
        assert np.array_equal(keep, oracle.nms_rotated(b, s, thr, 1))


def test_nms_full_size_group_decomposition(cuda):
    """Config-5 size (B=64, 3 classes: N=19200, 192 groups). Size-independent property: groups do not
    interact, so NMS(all) == per-group oracle NMS merged by descending score; idempotence."""
    from vision3d_b200 import ops
    n = 19200
    boxes, scores, idxs = synth.make_nms_boxes(5, n)
    b = synth.apply_group_offsets(boxes, idxs)
    keep = ops.nms_rotated(_t(b, cuda), _t(scores, cuda), 0.01).cpu().numpy()
    parts = []
    for g in range(n // 100):
        sl = slice(g * 100, (g + 1) * 100)
        parts.append(oracle.nms_rotated(b[sl], scores[sl], 0.01, 1) + g * 100)
    want = np.concatenate(parts)
    want = want[np.argsort(-scores[want], kind="stable")]
    assert np.array_equal(keep, want)
    keep2 = ops.nms_rotated(_t(b[keep], cuda), _t(scores[keep], cuda), 0.01).cpu().numpy()
    assert np.array_equal(keep2, np.arange(len(keep)))


def test_nms_empty_and_padded(cuda):
    from vision3d_b200 import ops
    assert ops.nms_rotated(torch.zeros((0, 5), device=cuda), torch.zeros((0,), device=cuda), 0.1).numel() == 0
    boxes, scores, idxs = synth.make_nms_boxes(1, 300)
    keep, count = ops.nms_rotated_padded(_t(boxes, cuda), _t(scores, cuda), 0.01)
    k = int(count.item())
    assert np.array_equal(keep[:k].cpu().numpy(), oracle.nms_rotated(boxes, scores, 0.01, 1))


@pytest.mark.skipif(not oracle.ref_available("ref_C_cuda.so"), reason="reference CUDA build not present")
def test_against_reference_cuda_kernels(cuda):
    """The reference's own SIMT kernels recompiled for sm_100a (oracle/_ref/ref_C_cuda.so) as on-GPU
    comparator: identical keep; IoU within 1e-5 (the reference build contracts FMAs, ours must not)."""
    from vision3d_b200 import ops
    ref = oracle.ref_torch_module(cuda=True)
    boxes, scores, idxs = synth.make_nms_boxes(3, 1600)
    b = _t(synth.apply_group_offsets(boxes, idxs), cuda)
    s = _t(scores, cuda)
    assert torch.equal(ops.nms_rotated(b, s, 0.01), ref.nms_rotated(b, s, 0.01))
    d = _t(boxes[:400], cuda)
    d[:, :2] *= 0.1
    a, r = ops.box_iou_rotated(d, d), ref.box_iou_rotated(d, d)
    assert (r > 0).sum() > 1000
    assert torch.allclose(a, r, atol=1e-5, rtol=0)


# ------------------------------------------------------------------------------------ voxelize
def _run_voxelize(cuda, clouds, max_voxels, policy, C=4, with_mean=True):
    from vision3d_b200 import ops
    total = sum(len(c) for c in clouds)
    B = len(clouds)
    vz = ops.Voxelizer(synth.VOXEL_SIZE, synth.GRID_BOUNDS, max_voxels, synth.MAX_OCCUPANCY, B, max(total, 1),
                       device=cuda, cap_policy=policy)
    out = vz.alloc_outputs(C, with_mean)
    for key in ("voxels", "coords", "num_points"):
        out[key].fill_(-7)  # poison: padding must be written by the kernel
    off = np.zeros(B + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in clouds])
    pts = np.concatenate(clouds, 0) if total else np.zeros((0, C), np.float32)
    for _ in range(2):  # second call reuses the epoch-tagged table without clearing
        vz.run(_t(pts, cuda), _t(off, cuda), max(len(c) for c in clouds), out)
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}


def _check_voxelize(got, clouds, max_voxels, policy):
    vo = got["voxel_offsets"]
    assert vo[0] == 0
    for b, cloud in enumerate(clouds):
        v, c, n = oracle.voxelize(cloud, synth.VOXEL_SIZE, synth.GRID_BOUNDS, synth.MAX_OCCUPANCY, max_voxels,
                                  policy)
        lo, hi = vo[b], vo[b + 1]
        assert hi - lo == len(v), (b, hi - lo, len(v))
        assert np.array_equal(got["coords"][lo:hi, 0], np.full(len(v), b))
        assert np.array_equal(got["coords"][lo:hi, 1:], c)
        assert np.array_equal(got["num_points"][lo:hi], n)
        assert np.array_equal(_bits(got["voxels"][lo:hi]), _bits(v))
        if got["mean"] is not None and len(v):
            np.testing.assert_allclose(got["mean"][lo:hi], oracle.vfe_mean(v, n), rtol=1e-6, atol=1e-7)


def test_voxelize_config1_single_cloud(cuda):
    clouds = [synth.make_cloud(0)]
    _check_voxelize(_run_voxelize(cuda, clouds, synth.MAX_VOXELS, 0), clouds, synth.MAX_VOXELS, 0)


def test_voxelize_ragged_batch_with_empty_and_outliers(cuda):
    clouds = [synth.make_cloud(1, 5000), np.zeros((0, 4), np.float32), synth.make_cloud(2, 16384),
              synth.make_cloud(3, 1025), synth.make_cloud(4, 1)]
    clouds[0][::7, 0] = -3.0  # outside the range -> dropped
    clouds[2][:4000] = clouds[2][4000:8000]  # heavy duplication: voxels with > 5 points
    for policy in (0, 1):
        _check_voxelize(_run_voxelize(cuda, clouds, synth.MAX_VOXELS, policy), clouds, synth.MAX_VOXELS, policy)


@pytest.mark.parametrize("policy", [0, 1])
def test_voxelize_max_voxels_cap(cuda, policy):
    clouds = [synth.make_cloud(5, 8000), synth.make_cloud(6, 3000)]
    _check_voxelize(_run_voxelize(cuda, clouds, 2500, policy), clouds, 2500, policy)


def test_voxelize_five_channel_points(cuda):
    rng = np.random.default_rng(0)
    c = synth.make_cloud(7, 3000)
    clouds = [np.concatenate([c, rng.random((3000, 1)).astype(np.float32)], 1)]
    got = _run_voxelize(cuda, clouds, synth.MAX_VOXELS, 0, C=5)
    _check_voxelize(got, clouds, synth.MAX_VOXELS, 0)


def test_voxelize_t16_batch_properties(cuda):
    """Target-line size (B=16 x 16384 pts): per-frame oracle parity plus conservation properties."""
    clouds = synth.make_batch(0, 16)
    got = _run_voxelize(cuda, clouds, synth.MAX_VOXELS, 0)
    _check_voxelize(got, clouds, synth.MAX_VOXELS, 0)
    rows = got["voxel_offsets"][-1]
    n = got["num_points"][:rows]
    assert n.min() >= 1 and n.max() <= synth.MAX_OCCUPANCY
    flat = got["coords"][:rows].astype(np.int64)
    key = ((flat[:, 0] * 41 + flat[:, 1]) * 1600 + flat[:, 2]) * 1408 + flat[:, 3]
    assert len(np.unique(key)) == rows  # one voxel per occupied cell


# ------------------------------------------------------------------------------------ rule book / conv / dense
def _site_setup(cuda, idx, shape, cap_extra=37):
    from vision3d_b200 import ops
    n = len(idx)
    cap = n + cap_extra
    ind = torch.zeros((cap, 4), dtype=torch.int32, device=cuda)
    ind[:n] = _t(idx, cuda)
    n_rows = torch.tensor([n], dtype=torch.int32, device=cuda)
    table = ops.SiteTable(cap, cuda).build(ind, n_rows, shape)
    return ind, n_rows, table, cap


@pytest.mark.parametrize("shape,n,B", [([9, 14, 12], 150, 2), ([41, 400, 352], 20000, 1)])
def test_rulebook_subm_exact(cuda, shape, n, B):
    from vision3d_b200 import ops
    idx = synth.make_clustered_sites(1, n, shape, B)
    ind, n_rows, table, cap = _site_setup(cuda, idx, shape)
    for _ in range(2):  # rebuild -> epoch bump
        table.build(ind, n_rows, shape)
        nbr = ops.rulebook_subm(table, ind, n_rows, shape, 3, 1)
    want = oracle.rulebook_subm(idx, shape, 3)
    assert (want >= 0).sum() > len(idx)
    assert np.array_equal(nbr[:, :len(idx)].cpu().numpy(), want)


@pytest.mark.parametrize("ks,st,pd", [(3, 2, 1), (3, 2, [0, 1, 1]), ([3, 1, 1], [2, 1, 1], 0)])
def test_rulebook_conv_exact(cuda, ks, st, pd):
    from vision3d_b200 import ops
    shape, B = [11, 40, 36], 3
    idx = synth.make_clustered_sites(2, 900, shape, B)
    ind, n_rows, table, cap = _site_setup(cuda, idx, shape)
    want_idx, want_nbr, want_shape = oracle.rulebook_conv(idx, shape, ks, st, pd)
    out_cap = len(want_idx) + 11
    for _ in range(2):
        out_idx, n_out, nbr, oshape = ops.rulebook_conv(table, ind, n_rows, B, shape, ks, st, pd, 1, out_cap)
    assert oshape == want_shape
    assert int(n_out.item()) == len(want_idx)
    assert np.array_equal(out_idx[:len(want_idx)].cpu().numpy(), want_idx)
    assert np.array_equal(nbr[:, :len(want_idx)].cpu().numpy(), want_nbr)


def test_rulebook_conv_kitti_level0(cuda):
    """Real level-0 geometry [41,1600,1408] -> [21,800,704] from a voxelized synthetic cloud."""
    from vision3d_b200 import ops
    clouds = synth.make_batch(0, 2)
    rows = []
    for b, c in enumerate(clouds):
        _, coords, _ = oracle.voxelize(c, synth.VOXEL_SIZE, synth.GRID_BOUNDS, 5, 20000)
        rows.append(np.concatenate([np.full((len(coords), 1), b, np.int32), coords], 1))
    idx = np.concatenate(rows, 0)
    shape = [41, 1600, 1408]
    ind, n_rows, table, cap = _site_setup(cuda, idx, shape)
    want_idx, want_nbr, _ = oracle.rulebook_conv(idx, shape, 3, 2, 1)
    out_idx, n_out, nbr, oshape = ops.rulebook_conv(table, ind, n_rows, 2, shape, 3, 2, 1, 1, len(want_idx) + 100)
    assert oshape == [21, 800, 704] and int(n_out.item()) == len(want_idx)
    assert np.array_equal(out_idx[:len(want_idx)].cpu().numpy(), want_idx)
    assert np.array_equal(nbr[:, :len(want_idx)].cpu().numpy(), want_nbr)


def _conv_case(cuda, cin, cout, subm, bn, seed=0):
    from vision3d_b200 import ops
    rng = np.random.default_rng(seed)
    shape, B = [9, 30, 28], 2
    idx = synth.make_clustered_sites(seed + 3, 700, shape, B)
    ind, n_rows, table, cap = _site_setup(cuda, idx, shape)
    feat = rng.normal(size=(len(idx), cin)).astype(np.float32)
    w = (rng.normal(size=(27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    fd = torch.zeros((cap, cin), device=cuda)
    fd[:len(idx)] = _t(feat, cuda)
    if subm:
        nbr = ops.rulebook_subm(table, ind, n_rows, shape, 3, 1)
        n_out, out_cap, want_nbr = n_rows, cap, oracle.rulebook_subm(idx, shape, 3)
    else:
        oi, want_nbr, _ = oracle.rulebook_conv(idx, shape, 3, 2, 1)
        out_cap = len(oi) + 5
        _, n_out, nbr, _ = ops.rulebook_conv(table, ind, n_rows, B, shape, 3, 2, 1, 1, out_cap)
    scale = shift = None
    if bn:
        scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.normal(size=cout).astype(np.float32)
    out = ops.sparse_conv(fd, _t(w, cuda), nbr, n_out, out_cap, _t(scale, cuda) if bn else None,
                          _t(shift, cuda) if bn else None, relu=bn)
    want = oracle.sparse_conv(feat, w, want_nbr, scale, shift, bn)
    got = out[:len(want)].cpu().numpy()
    tol = 1e-4 * np.abs(want).max()
    assert np.abs(got - want).max() <= tol, (np.abs(got - want).max(), tol)


@pytest.mark.parametrize("cin,cout", [(4, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (16, 64)])
def test_sparse_conv_subm_vs_oracle(cuda, cin, cout):
    _conv_case(cuda, cin, cout, subm=True, bn=True)
    _conv_case(cuda, cin, cout, subm=True, bn=False, seed=1)


@pytest.mark.parametrize("cin,cout", [(16, 32), (32, 64), (64, 64)])
def test_sparse_conv_strided_vs_oracle(cuda, cin, cout):
    _conv_case(cuda, cin, cout, subm=False, bn=True)


def test_sparse_conv_row_permutation_equivariance(cuda):
    """Property: permuting the voxel rows permutes the SubM output rows."""
    from vision3d_b200 import ops
    rng = np.random.default_rng(3)
    shape = [9, 30, 28]
    idx = synth.make_clustered_sites(8, 600, shape, 1)
    feat = rng.normal(size=(600, 16)).astype(np.float32)
    w = _t(rng.normal(size=(27, 16, 32)).astype(np.float32), cuda)
    outs = []
    perm = rng.permutation(600)
    for order in (np.arange(600), perm):
        ind, n_rows, table, cap = _site_setup(cuda, idx[order], shape, cap_extra=0)
        nbr = ops.rulebook_subm(table, ind, n_rows, shape, 3, 1)
        outs.append(ops.sparse_conv(_t(feat[order], cuda), w, nbr, n_rows, cap).cpu().numpy())
    np.testing.assert_allclose(outs[1], outs[0][perm], rtol=0, atol=1e-5 * np.abs(outs[0]).max())


def test_dense_exact(cuda):
    from vision3d_b200 import ops
    rng = np.random.default_rng(0)
    for shape, B, n, C in [([2, 200, 176], 3, 9000, 64), ([5, 13, 7], 2, 100, 16), ([3, 9, 67], 1, 50, 5)]:
        idx = synth.make_active_sites(1, n, shape, B)
        feat = rng.normal(size=(len(idx), C)).astype(np.float32)
        ind, n_rows, _, cap = _site_setup(cuda, idx, shape)
        fd = torch.zeros((cap, C), device=cuda)
        fd[:len(idx)] = _t(feat, cuda)
        out = ops.sparse_to_dense(fd, ind, n_rows, cap, B, shape)
        assert out.shape == (B, C, *shape)
        assert np.array_equal(_bits(out.cpu().numpy()), _bits(oracle.dense(feat, idx, B, shape)))
        # the reference then views (B, C*D, H, W) (sparse_cnn.py:131-132): must be a free view
        assert out.view(B, C * shape[0], shape[1], shape[2]).is_contiguous()


# ------------------------------------------------------------------------------------ point ops
@pytest.mark.parametrize("B,N,m", [(2, 4096, 256), (3, 1000, 64), (1, 16384, 2048), (2, 37, 37)])
def test_fps_exact(cuda, B, N, m):
    from vision3d_b200 import ops
    rng = np.random.default_rng(N)
    xyz = rng.uniform(-30, 30, size=(B, N, 3)).astype(np.float32)
    if N > 100:
        xyz[:, N - 50:] = xyz[:, :50]  # duplicated points: exact distance ties (batch padding)
    got = ops.furthest_point_sample(_t(xyz, cuda), m).cpu().numpy()
    assert got.dtype == np.int32
    assert np.array_equal(got, oracle.fps(xyz, m))


def test_fps_config3_batch8_synthetic_clouds(cuda):
    from vision3d_b200 import ops
    xyz = np.stack([c[:, :3] for c in synth.make_batch(0, 8)], 0)
    got = ops.furthest_point_sample(_t(xyz, cuda), 2048).cpu().numpy()
    want = oracle.fps(xyz[:2], 2048)
    assert np.array_equal(got[:2], want)
    for b in range(8):  # property: a sample never repeats while distinct points remain
        assert len(np.unique(got[b])) == 2048


def test_ball_query_group_gather_exact(cuda):
    from vision3d_b200 import ops
    rng = np.random.default_rng(1)
    B, N, M, C = 2, 5000, 300, 7
    xyz = rng.uniform(-10, 10, size=(B, N, 3)).astype(np.float32)
    q = xyz[:, rng.choice(N, M, replace=False)] + rng.normal(0, 0.05, size=(B, M, 3)).astype(np.float32)
    q[0, 0] = 500.0
    feat = rng.normal(size=(B, C, N)).astype(np.float32)
    for radius, ns in [(0.4, 16), (0.8, 32), (2.4, 16)]:
        idx = ops.ball_query(radius, ns, _t(xyz, cuda), _t(q, cuda))
        want = oracle.ball_query(radius, ns, xyz, q)
        assert np.array_equal(idx.cpu().numpy(), want)
        g = ops.grouping_operation(_t(feat, cuda), idx).cpu().numpy()
        assert np.array_equal(_bits(g), _bits(oracle.group(feat, want)))
        qg = ops.query_and_group(_t(xyz, cuda), _t(q, cuda), _t(feat, cuda), idx).cpu().numpy()
        assert np.array_equal(_bits(qg), _bits(oracle.query_and_group(xyz, q, feat, want)))
    fi = oracle.fps(xyz, 64)
    got = ops.gather_operation(_t(feat, cuda), _t(fi, cuda)).cpu().numpy()
    assert np.array_equal(_bits(got), _bits(oracle.gather(feat, fi)))


# ------------------------------------------------------------------------------------ tcgen05 sparse conv
def _conv_tc_case(cuda, cin, cout, subm, n_sites=3000, shape=(9, 60, 56), B=2, seed=0, bn=True):
    """tcgen05 3xTF32 kernel vs the fp64-accumulating oracle (<= 1e-4 rel of max, expected ~1e-6) and vs
    the exact-fp32 SIMT kernel."""
    from vision3d_b200 import ops
    rng = np.random.default_rng(seed)
    shape = list(shape)
    idx = synth.make_clustered_sites(seed + 11, n_sites, shape, B)
    ind, n_rows, table, cap = _site_setup(cuda, idx, shape)
    feat = rng.normal(size=(len(idx), cin)).astype(np.float32)
    w = (rng.normal(size=(27, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    fd = torch.zeros((cap, cin), device=cuda)
    fd[:len(idx)] = _t(feat, cuda)
    if subm:
        nbr = ops.rulebook_subm(table, ind, n_rows, shape, 3, 1)
        n_out, out_cap, want_nbr = n_rows, cap, oracle.rulebook_subm(idx, shape, 3)
    else:
        oi, want_nbr, _ = oracle.rulebook_conv(idx, shape, 3, 2, 1)
        out_cap = len(oi) + 5
        _, n_out, nbr, _ = ops.rulebook_conv(table, ind, n_rows, B, shape, 3, 2, 1, 1, out_cap)
    scale = shift = None
    if bn:
        scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.normal(size=cout).astype(np.float32)
    sc, sh = (_t(scale, cuda), _t(shift, cuda)) if bn else (None, None)
    pw = ops.PreparedWeights(_t(w, cuda))
    assert pw.buf is not None
    got = ops.sparse_conv(fd, pw, nbr, n_out, out_cap, sc, sh, relu=bn)
    simt = ops.sparse_conv(fd, _t(w, cuda), nbr, n_out, out_cap, sc, sh, relu=bn)
    want = oracle.sparse_conv(feat, w, want_nbr, scale, shift, bn)
    n = len(want)
    err = np.abs(got[:n].cpu().numpy() - want).max() / np.abs(want).max()
    err_simt = (got[:n] - simt[:n]).abs().max().item() / np.abs(want).max()
    assert err <= 1e-4, err          # the contract (BASELINE north_star)
    assert err_simt <= 3e-5, err_simt  # bf16x3: ~5e-6 expected
    # packed output = the bf16 (h1 | h2) split of the fp32 output, bit for bit what v3d_feature_pack produces
    outp = torch.empty((out_cap, 2 * cout), dtype=torch.bfloat16, device=cuda)
    got2 = ops.sparse_conv(fd, pw, nbr, n_out, out_cap, sc, sh, relu=bn, out_packed=outp)
    assert torch.equal(got2[:n], got[:n])
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda)
    assert torch.equal(ops.pack_features(got2.contiguous(), n_dev)[:n], outp[:n])
    rel = (ops.unpack_features(outp[:n]) - got[:n]).abs().max().item() / max(got[:n].abs().max().item(), 1e-30)
    assert rel <= 2.0 ** -16, rel
    only_packed = torch.zeros_like(outp)
    ops.sparse_conv(ops.pack_features(fd, torch.tensor([fd.shape[0]], dtype=torch.int32, device=cuda)), pw, nbr,
                    n_out, out_cap, sc, sh, relu=bn, out_packed=only_packed, write_f32=False)
    assert torch.equal(only_packed[:n], outp[:n])
    return err


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (16, 64)])
def test_sparse_conv_tc_subm(cuda, cin, cout):
    err = _conv_tc_case(cuda, cin, cout, subm=True)
    assert err < 3e-5, err  # bf16x3 split: ~5e-6 of the output scale


@pytest.mark.parametrize("cin,cout", [(16, 32), (32, 64), (64, 64)])
def test_sparse_conv_tc_strided(cuda, cin, cout):
    _conv_tc_case(cuda, cin, cout, subm=False, bn=False)


def test_sparse_conv_tc_many_tiles_and_tail(cuda):
    """More tiles than SMs (persistent loop, TMEM double buffering) and a ragged last tile."""
    _conv_tc_case(cuda, 64, 64, subm=True, n_sites=148 * 128 + 77, shape=(11, 120, 110), B=2, seed=5)
    _conv_tc_case(cuda, 32, 32, subm=True, n_sites=130, shape=(5, 20, 20), B=1, seed=6)


def test_prepared_weights_fall_back_to_exact_path_for_cin4(cuda):
    from vision3d_b200 import ops
    pw = ops.PreparedWeights(torch.randn(27, 4, 16, device=cuda))
    assert pw.buf is None  # Cin = 4 stays on the exact-fp32 SIMT kernel


def test_rank_indexed_rulebooks_match_hash_and_oracle(cuda):
    """Levels produced by a strided conv are indexed by that conv's bitmap/prefix workspace: SubM and the next
    strided conv built from it must equal the oracle (and therefore the hash-table path)."""
    from vision3d_b200 import ops
    shape, B = [11, 120, 96], 3
    idx0 = synth.make_clustered_sites(21, 5000, shape, B)
    ind, n_rows, table, cap = _site_setup(cuda, idx0, shape)
    # level 0 -> 1 (hash-indexed input)
    o1, nbr1_want, shape1 = oracle.rulebook_conv(idx0, shape, 3, 2, 1)
    cap1 = len(o1) + 100
    ws1 = ops.ConvRulebookWorkspace(B, shape1, cap1, 27, cuda)
    out1, n1, nbr1, sh1 = ops.rulebook_conv(table, ind, n_rows, B, shape, 3, 2, 1, 1, cap1, workspace=ws1)
    assert sh1 == shape1 and int(n1.item()) == len(o1)
    assert np.array_equal(out1[:len(o1)].cpu().numpy(), o1)
    # SubM on level 1 through the rank index
    nbr_s = ops.rulebook_subm(ws1, out1, n1, shape1, 3, 1, capacity=cap1)
    assert np.array_equal(nbr_s[:, :len(o1)].cpu().numpy(), oracle.rulebook_subm(o1, shape1, 3))
    # level 1 -> 2 with the rank index as input index
    o2, nbr2_want, shape2 = oracle.rulebook_conv(o1, shape1, 3, 2, [0, 1, 1])
    cap2 = len(o2) + 50
    out2, n2, nbr2, sh2 = ops.rulebook_conv(ws1, out1, n1, B, shape1, 3, 2, [0, 1, 1], 1, cap2)
    assert sh2 == shape2 and int(n2.item()) == len(o2)
    assert np.array_equal(out2[:len(o2)].cpu().numpy(), o2)
    assert np.array_equal(nbr2[:, :len(o2)].cpu().numpy(), nbr2_want)


def test_voxelize_large_frame_uses_global_hash_path(cuda):
    """Frames above 32768 points take the multi-kernel global-hash path (the cluster/DSMEM kernel covers
    frames up to 32768 points); both must give the oracle's result."""
    big = np.concatenate([synth.make_cloud(40, 16384), synth.make_cloud(41, 16384), synth.make_cloud(42, 8000)], 0)
    clouds = [big, synth.make_cloud(43, 2000)]
    _check_voxelize(_run_voxelize(cuda, clouds, 60000, 0), clouds, 60000, 0)
    mid = [np.concatenate([synth.make_cloud(44, 16384), synth.make_cloud(45, 9000)], 0)]  # PPT = 4 cluster path
    _check_voxelize(_run_voxelize(cuda, mid, 60000, 1), mid, 60000, 1)


def test_voxelize_cluster_dsmem_variant_subprocess(cuda):
    """The opt-in single-kernel cluster/DSMEM voxelizer (V3D_VOXELIZE_CLUSTER=1, read once per process) must
    give the same bit-exact result as the default path; run it in a child process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import torch\n"
        "from tests.test_gpu_parity import _run_voxelize, _check_voxelize\n"
        "from vision3d_b200 import synth\n"
        "dev = torch.device('cuda:0')\n"
        "clouds = [synth.make_cloud(1, 5000), np.zeros((0, 4), np.float32), synth.make_cloud(2, 16384), synth.make_cloud(3, 1025)]\n"
        "clouds[2][:4000] = clouds[2][4000:8000]\n"
        "for pol, mv in ((0, 20000), (1, 20000), (0, 2500), (1, 2500)):\n"
        "    _check_voxelize(_run_voxelize(dev, clouds, mv, pol), clouds, mv, pol)\n"
        "print('cluster-ok')\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, V3D_VOXELIZE_CLUSTER="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert "cluster-ok" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_grouped_nms_equals_global_nms_on_offset_groups(cuda):
    """v3d_nms_rotated_grouped == v3d_nms_rotated == oracle on groups separated by the batched_nms offsets."""
    from vision3d_b200 import ops
    rng = np.random.default_rng(17)
    for gs, ng in ((100, 16), (128, 3), (37, 5), (1, 4)):
        n = gs * ng
        ctr = rng.uniform(0, 40, (n, 2))
        wl = rng.uniform(1.0, 6.0, (n, 2))
        ang = rng.uniform(-180, 180, (n, 1))
        boxes = np.concatenate([ctr, wl, ang], 1).astype(np.float32)
        scores = rng.random(n).astype(np.float32)
        scores[rng.integers(0, n, n // 5)] = 0.5                      # ties -> lower index first
        grp = np.repeat(np.arange(ng), gs).astype(np.float32)
        span = np.float32(boxes[:, :2].max() + boxes[:, 2:4].max() / 2 - (boxes[:, :2].min() - boxes[:, 2:4].min() / 2) + 1)
        off = boxes.copy()
        off[:, :2] += (grp * span)[:, None]                           # ops/iou_nms.py:121-132
        d, s = _t(off, cuda), _t(scores, cuda)
        k0, c0 = ops.nms_rotated_padded(d, s, 0.01)
        k1, c1 = ops.nms_rotated_padded(d, s, 0.01, group_size=gs)
        n0, n1 = int(c0.item()), int(c1.item())
        assert n0 == n1 and torch.equal(k0[:n0], k1[:n1])
        assert np.array_equal(k1[:n1].cpu().numpy(), oracle.nms_rotated(off, scores, 0.01, 1))


# ------------------------------------------------------------------------------------ sparse conv backward (8f-1)
@pytest.mark.parametrize("subm,cin,cout", [(True, 8, 16), (False, 16, 32), (True, 64, 64), (False, 4, 16)])
def test_sparse_conv_backward_vs_dense_conv3d_autograd(cuda, subm, cin, cout):
    """Hand-written backward (dX: forward kernel on the inverted rule table with W^T; dW: v3d_sparse_conv_bwd_weight)
    through the compat spconv autograd Function, against torch autograd of a float64 dense conv3d on the densified
    input, gradients read back at the active sites. <= 1e-4 of the gradient scale (fp32 kernels: ~1e-6)."""
    import torch.nn.functional as F
    from vision3d_b200.compat import spconv
    shape, B = [7, 14, 12], 2
    idx = synth.make_clustered_sites(5, 260, shape, B)
    ii = torch.from_numpy(idx.astype(np.int64))
    torch.manual_seed(1)
    feats = torch.randn(len(idx), cin, device=cuda, requires_grad=True)
    if subm:
        conv = spconv.SubMConv3d(cin, cout, 3, indice_key="k", bias=False).to(cuda)
        kw = dict(padding=1)
    else:
        conv = spconv.SparseConv3d(cin, cout, 3, 2, padding=[0, 1, 1], bias=False).to(cuda)
        kw = dict(stride=2, padding=(0, 1, 1))
    conv.train()
    y = conv(spconv.SparseConvTensor(feats, torch.from_numpy(idx).to(cuda), shape, B))
    R = torch.randn_like(y.features)
    (y.features * R).sum().backward()
    # float64 dense reference on the CPU
    dense_in = torch.zeros((B, cin, *shape), dtype=torch.float64)
    dense_in[ii[:, 0], :, ii[:, 1], ii[:, 2], ii[:, 3]] = feats.detach().double().cpu()
    dense_in.requires_grad_(True)
    w64 = conv.weight.detach().double().cpu().requires_grad_(True)           # (k0, k1, k2, Cin, Cout)
    out = F.conv3d(dense_in, w64.permute(4, 3, 0, 1, 2), **kw)
    oi = y.indices.long().cpu()
    got_fwd = y.features.detach().double().cpu()
    ref_fwd = out[oi[:, 0], :, oi[:, 1], oi[:, 2], oi[:, 3]]
    assert float((got_fwd - ref_fwd).abs().max() / ref_fwd.abs().max()) <= 1e-5
    (ref_fwd * R.double().cpu()).sum().backward()
    g_in = dense_in.grad[ii[:, 0], :, ii[:, 1], ii[:, 2], ii[:, 3]]
    err_x = float((feats.grad.double().cpu() - g_in).abs().max() / g_in.abs().max())
    err_w = float((conv.weight.grad.double().cpu() - w64.grad).abs().max() / w64.grad.abs().max())
    assert err_x <= 1e-4 and err_w <= 1e-4, (err_x, err_w)


def test_rulebook_invert_is_the_inverse(cuda):
    from vision3d_b200 import ops
    shape, B = [9, 30, 28], 2
    idx = synth.make_clustered_sites(2, 400, shape, B)
    ind, n_rows, table, cap = _site_setup(cuda, idx, shape)
    oi, want_nbr, _ = oracle.rulebook_conv(idx, shape, 3, 2, 1)
    _, n_out, nbr, _ = ops.rulebook_conv(table, ind, n_rows, B, shape, 3, 2, 1, 1, len(oi) + 3)
    inv = ops.rulebook_invert(nbr, n_out, len(oi) + 3, len(idx)).cpu().numpy()
    want = np.full((27, len(idx)), -1, np.int32)
    for k in range(27):
        o = np.nonzero(want_nbr[k] >= 0)[0]
        want[k, want_nbr[k, o]] = o
    assert np.array_equal(inv, want)
