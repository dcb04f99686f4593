"""Host-side mirror of the reference's PV-RCNN keypoint stage (BASELINE config 3: "FPS-2048 + ball_query +
RoI-grid pool, batch 8"), on the vision3d_b200 kernels.

The reference wires only `PV_RCNN.proposal` (detector/model.py:76-82); `point_feature_extract` (:68-74) and
`RoiGridPool.forward` (detector/roi_grid_pool.py:64-72) are defined but never called and `PV_RCNN.forward`
raises (:84-85) -- SURVEY 3.2. This module is the driver SURVEY 3.2 asks for:

`PVRCNNB200`   nn.Module with the reference's module tree / parameter names (pnets.*, roi_grid_pool.pnet,
               roi_grid_pool.reduction.linear_*, cnn.blocks.*, proposal_layer.conv_cls/conv_reg), built on the compat
               `pointnet2` / `spconv` drop-ins exactly the way the reference builds it, plus the reference-shaped
               eager methods (sample_keypoints, _pointnets, point_feature_extract, RoiGridPool.forward).
`KeypointStage` the production path bench.py measures: raw points -> fused FPS+gather (128-bit point loads) ->
               voxelize/VFE/sparse backbone with every level kept (SecondEngine plan) -> device-side frame offsets +
               to_global (no .cpu() round trip, no pad_batch: the ball query reads the ragged levels in place) ->
               5 x [multi-radius warp-per-query ball query -> grouping -> shared MLP -> max] -> bilinear BEV gather ->
               RoI-grid pool on injected grid points -> reduction MLP.

RNG: `sample_gridpoints` draws `torch.rand` (roi_grid_pool.py:59) and `pad_batch` draws `torch.randint`
(sparse_cnn.py:33-37,123); neither stream is reproducible across devices, so grid points are an INPUT here
(`sample_gridpoints(boxes, noise)` takes the uniform noise) and no padding is drawn at all (see ops.ball_query_msg).
"""
import math
from copy import deepcopy
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops, second
from .compat.pointnet2.pointnet2_modules import PointnetSAModuleMSG


@dataclass
class PVRCNNConfig(second.SecondConfig):
    """core/config.py:4-80 (defaults: 3 classes)."""
    NUM_KEYPOINTS: int = 2048
    SAMPLES_PN: List[int] = field(default_factory=lambda: [16, 32])
    PSA_RADII: List[List[float]] = field(default_factory=lambda: [[0.4, 0.8], [0.4, 0.8], [0.8, 1.2], [1.2, 2.4],
                                                                  [2.4, 4.8]])
    PSA_MLPS: List[List[List[int]]] = field(default_factory=lambda: [
        [[1, 8, 16], [1, 8, 16]], [[4, 8, 16], [4, 8, 16]], [[32, 32, 32], [32, 32, 32]],
        [[64, 64, 64], [64, 64, 64]], [[64, 64, 64], [64, 64, 64]]])
    GRIDPOOL_NUM_GRIDPOINTS: int = 16
    GRIDPOOL_RADII_PN: List[float] = field(default_factory=lambda: [0.8, 1.6])
    GRIDPOOL_MLPS_PN: List[List[int]] = field(default_factory=lambda: [[512, 192, 96], [512, 192, 96]])
    GRIDPOOL_MLPS_REDUCTION: List[int] = field(default_factory=lambda: [16 * 192, 256, 256])


class MLPB200(nn.Sequential):
    """detector/layers.py:53-75 with the defaults RoiGridPool uses (bias=False, bn=False, relu=True)."""

    def __init__(self, channels):
        super().__init__()
        for i in range(len(channels) - 1):
            self.add_module("linear_%d" % i, nn.Linear(channels[i], channels[i + 1], bias=False))
            nn.init.normal_(self[-1].weight, std=0.01)
            self.add_module("relu_%d" % i, nn.ReLU(inplace=True))


def rotate_z(points, theta):
    """roi_grid_pool.py:35-49. points (b, n, m, 3), theta (b, n)."""
    b, n, m, _ = points.shape
    theta = theta.unsqueeze(-1).expand(-1, -1, m)
    xy, z = torch.split(points, [2, 1], dim=-1)
    c, s = torch.cos(theta), torch.sin(theta)
    R = torch.stack((c, -s, s, c), dim=-1).view(b, n, m, 2, 2)
    xy = torch.matmul(R, xy.unsqueeze(-1))
    return torch.cat((xy.squeeze(-1), z), dim=-1)


def sample_gridpoints(boxes, noise):
    """roi_grid_pool.py:51-62 with the uniform noise `torch.rand((b, n, m, 3))` injected. -> (b, n, m, 3)."""
    g = boxes[:, :, None, 3:6] * (noise - 0.5)
    return boxes[:, :, None, 0:3] + rotate_z(g, boxes[..., -1])


class RoiGridPoolB200(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.pnet = PointnetSAModuleMSG(npoint=-1, radii=cfg.GRIDPOOL_RADII_PN, nsamples=cfg.SAMPLES_PN,
                                        mlps=deepcopy(cfg.GRIDPOOL_MLPS_PN), use_xyz=True)
        self.reduction = MLPB200(cfg.GRIDPOOL_MLPS_REDUCTION)

    def forward(self, proposals, keypoint_xyz, keypoint_features, noise):
        """roi_grid_pool.py:64-72; `noise` replaces the torch.rand draw inside sample_gridpoints."""
        b, n, _ = proposals.shape
        m = self.cfg.GRIDPOOL_NUM_GRIDPOINTS
        gridpoints = sample_gridpoints(proposals, noise).view(b, -1, 3)
        features = self.pnet(keypoint_xyz, keypoint_features, gridpoints)[1]
        features = features.view(b, -1, n, m).permute(0, 2, 1, 3).contiguous().view(b, n, -1)
        return self.reduction(features)


def bev_gather(cfg, feature_map, keypoint_xyz):
    """BEVFeatureGatherer.forward (detector/layers.py:20-50): bilinear F.grid_sample of the BEV map at the
    keypoints' fractional pixel indices (x <-> W swap included). feature_map (B, C, H, W) -> (B, C, M)."""
    _, _, H, W = feature_map.shape
    pixel_offset = feature_map.new_tensor(cfg.GRID_BOUNDS[:2])
    base_pixel = feature_map.new_tensor(cfg.VOXEL_SIZE[:2])
    ind = keypoint_xyz[:, None, :, :2] - pixel_offset
    ind = ind / (base_pixel * cfg.STRIDES[-1])
    dims = ind.new_tensor([W - 1, H - 1])
    ind = torch.min(torch.clamp(ind, min=0), dims)
    ind = (2 * (ind / (dims - 1)) - 1).flip(3)
    return F.grid_sample(feature_map, ind, align_corners=True).squeeze(2)


class PVRCNNB200(nn.Module):
    """Module tree of detector/model.py:16-33 (refinement_layer left out: unimplemented upstream, refinement.py:32-33)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        pn = []
        for i, mlps in enumerate(cfg.PSA_MLPS):  # build_pointnets, model.py:35-44 (deepcopy: the module mutates it)
            pn.append(PointnetSAModuleMSG(npoint=-1, radii=cfg.PSA_RADII[i], nsamples=cfg.SAMPLES_PN,
                                          mlps=deepcopy(mlps), use_xyz=True))
        self.pnets = nn.Sequential(*pn)
        self.roi_grid_pool = RoiGridPoolB200(cfg)
        self.cnn = second.MiddleB200(cfg)
        self.proposal_layer = second.HeadB200(cfg)

    @property
    def head(self):  # SecondEngine's name for the proposal layer
        return self.proposal_layer

    # ---- reference-shaped eager methods (the compat drop-ins underneath are vision3d_b200 kernels)
    def sample_keypoints(self, points):
        """model.py:46-56."""
        from .compat.pointnet2.pointnet2_utils import furthest_point_sample, gather_operation
        points = points[..., :3].contiguous()
        indices = furthest_point_sample(points, self.cfg.NUM_KEYPOINTS)
        keypoints = gather_operation(points.transpose(1, 2).contiguous(), indices)
        return keypoints.transpose(1, 2).contiguous()

    def _pointnets(self, cnn_out, keypoint_xyz):
        """model.py:58-66. cnn_out: [(xyz (B,N,3), features (B,N,C))] dense per source."""
        out = []
        for (voxel_xyz, voxel_features), pnet in zip(cnn_out, self.pnets):
            out.append(pnet(voxel_xyz.contiguous(), voxel_features.transpose(1, 2).contiguous(), keypoint_xyz)[1])
        return out

    def point_feature_extract(self, points, keypoints, cnn_features, bev_map):
        """model.py:68-74 -> (B, 384 + 128, M)."""
        points_split = torch.split(points, [3, 1], dim=-1)
        pf = self._pointnets([points_split] + cnn_features, keypoints)
        return torch.cat(pf + [bev_gather(self.cfg, bev_map, keypoints)], dim=1)


def init_for_benchmark(model, seed=0):
    """O(1) activations everywhere (see second.init_for_benchmark): sparse backbone He-init at the effective fan-in,
    shared MLPs kaiming (their default), BN statistics randomised so that the eval-mode fold is exercised."""
    second.init_for_benchmark_backbone(model, seed)
    g = torch.Generator().manual_seed(seed + 17)
    with torch.no_grad():
        for m in list(model.pnets.modules()) + list(model.roi_grid_pool.pnet.modules()):
            if isinstance(m, nn.Conv2d):   # the modules' own kaiming draw comes from the global RNG: redraw it here so
                fan = m.in_channels       # that two models built with the same seed are identical
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan))
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        for m in model.roi_grid_pool.reduction.modules():
            if isinstance(m, nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / m.in_features))
    return model


def make_proposals(clouds, n, seed=0, wlh=(1.6, 3.9, 1.56)):
    """n boxes per frame centred on random cloud points, wlh = car anchor, yaw ~ U(-pi, pi)."""
    rng = np.random.default_rng(seed)
    out = np.empty((len(clouds), n, 7), np.float32)
    for b, c in enumerate(clouds):
        pick = rng.integers(0, len(c), n)
        out[b, :, :3] = c[pick, :3]
        out[b, :, 3:6] = wlh
        out[b, :, 6] = rng.uniform(-np.pi, np.pi, n)
    return out


def make_grid_noise(B, n, m, seed=0):
    """The uniform draw of sample_gridpoints, generated ONCE on the host (torch.manual_seed(seed)) and injected
    into both implementations."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((B, n, m, 3), generator=g)


def _fold_shared_mlp(mlp):
    """SharedMLP (Conv2d 1x1 no bias -> BatchNorm2d(eval) -> ReLU) -> [(W (Cout, Cin), b (Cout))] with BN folded."""
    layers = []
    for blk in mlp:
        conv, bn = blk.conv, blk.bn.bn
        inv = torch.rsqrt(bn.running_var + bn.eps)
        s = (bn.weight * inv).detach()
        w = (conv.weight.detach().reshape(conv.out_channels, -1) * s[:, None]).float().contiguous()
        b = (bn.bias - bn.running_mean * s).detach().float().contiguous()
        layers.append((w, b))
    return layers


class KeypointStage:
    """The C3 production path (see module docstring). All buffers static, no host synchronisation inside `step`.

    step inputs : self.points (B, N, 4) raw clouds, self.gridpoints (B, n*16, 3) = sample_gridpoints(proposals, noise)
                  evaluated ONCE on the host and injected (cos/sin differ in the last bit across back
                  ends, which would move grid points across ball boundaries)
    step outputs: self.keypoints (B, 2048, 3), self.kp_idx (B, 2048) int32, self.kp_features (B, 512, 2048),
                  self.pooled (B, n, 256); per-source ball-query indices in self.sa_idx for the parity tests."""

    def __init__(self, model: PVRCNNB200, batch_size, points_per_frame, n_proposals, device, level_caps=None,
                 fused_sa=True, overlap_rulebooks=True):
        cfg = model.cfg
        self.cfg, self.B, self.N, self.n = cfg, int(batch_size), int(points_per_frame), int(n_proposals)
        self.dev = dev = torch.device(device)
        self.model = model.to(dev).eval()
        B, M = self.B, cfg.NUM_KEYPOINTS
        # sparse backbone: the SecondEngine plan up to the dense BEV, every level's fp32 rows kept
        self.eng = second.SecondEngine(model, B, B * self.N, dev, level_caps=level_caps, use_graph=False,
                                       rpn_mode="none", keep_level_features=True,
                                       overlap_rulebooks=overlap_rulebooks)
        self.points = torch.zeros((B, self.N, 4), dtype=torch.float32, device=dev)
        self.eng.points = self.points.view(B * self.N, 4)      # the voxelizer reads the same buffer
        self.eng.frame_off.copy_(torch.arange(B + 1, dtype=torch.int32) * self.N)
        self.gridpoints = torch.zeros((B, self.n * cfg.GRIDPOOL_NUM_GRIDPOINTS, 3), dtype=torch.float32, device=dev)
        self.kp_idx = torch.zeros((B, M), dtype=torch.int32, device=dev)
        self.keypoints = torch.zeros((B, M, 3), dtype=torch.float32, device=dev)
        # per sparse level: frame offsets + metric voxel centres (to_global), ragged
        lo = np.asarray(cfg.GRID_BOUNDS[:3], np.float32)
        vs = np.asarray(cfg.VOXEL_SIZE, np.float32)
        self.level_vs = [(vs * np.float32(s)).astype(np.float32) for s in cfg.STRIDES]
        self.level_off = lo
        self.level_offsets = [self.eng.vox_out["voxel_offsets"]] + [
            torch.zeros(B + 1, dtype=torch.int32, device=dev) for _ in range(3)]
        self.level_xyz = [torch.zeros((self.eng.caps[lv], 3), dtype=torch.float32, device=dev) for lv in range(4)]
        self.sa_mlps = [[_fold_shared_mlp(m) for m in p.mlps] for p in self.model.pnets]
        self.roi_mlps = [_fold_shared_mlp(m) for m in self.model.roi_grid_pool.pnet.mlps]
        self.sa_idx = [[torch.zeros((B, M, ns), dtype=torch.int32, device=dev) for ns in cfg.SAMPLES_PN]
                       for _ in range(5)]
        Mg = self.n * cfg.GRIDPOOL_NUM_GRIDPOINTS
        self.roi_idx = [torch.zeros((B, Mg, ns), dtype=torch.int32, device=dev) for ns in cfg.SAMPLES_PN]
        # levels 1-3 are in ascending (b,z,y,x) order: chunk bounding boxes make their ball queries ~20x cheaper
        self.sa_bounds = [None, None] + [ops.BallQueryBounds(B, self.eng.caps[lv], dev) for lv in (1, 2, 3)]
        # raw points (shuffled) and level-0 voxels (first-appearance order): x-bucketed copy + index selection
        xr = (cfg.GRID_BOUNDS[0], cfg.GRID_BOUNDS[3])
        self.sa_sorted = [ops.BallQuerySorted(B, self.N, B * self.N, xr, dev),
                          ops.BallQuerySorted(B, self.eng.caps[0], self.eng.caps[0], xr, dev)]
        self.kp_features = torch.zeros((B, 512, M), dtype=torch.float32, device=dev)
        # fused set abstraction: grouping -> shared MLP -> max in one tensor-core kernel per scale
        self.fused_sa = bool(fused_sa)
        if self.fused_sa:
            src_c = [1, 4, 32, 64, 64]
            src_cp = [8, 16, 32, 64, 64]   # channels of the packed source rows (x0: the engine's 16-channel pack)
            self.sa_prepared = [[ops.PreparedSaMlp(l, src_c[i], src_cp[i]) for l in self.sa_mlps[i]] for i in range(5)]
            self.roi_prepared = [ops.PreparedSaMlp(l, 512) for l in self.roi_mlps]
            self.pts_packed = torch.zeros((B * self.N, 16), dtype=torch.bfloat16, device=dev)
            self.kp_packed = torch.zeros((B * M, 1024), dtype=torch.bfloat16, device=dev)
            self.roi_feat = torch.zeros((B, 192, Mg), dtype=torch.float32, device=dev)
        self.pooled = None
        self.timings = {}
        self._build_plan()

    # level sources: (xyz rows, feature rows, row_offsets or None, stride of xyz rows)
    def _source(self, i):
        e = self.eng
        if i == 0:   # raw points: xyz = points[..., :3], feature = intensity (model.py:69-70); dense batch
            return self.points, self.points[..., 3:], None
        lv = i - 1
        feat = e.vox_out["mean"] if lv == 0 else e.feat[lv][0]
        return self.level_xyz[lv], feat, self.level_offsets[lv]

    @staticmethod
    def _mlp_max(grouped, layers):
        """shared MLP (BN folded) + ReLU per layer, then max over the samples. grouped (B, C, M, ns) -> (B, Cout, M).
        fp32 GEMMs (the pooling contract is 1e-4: TF32 is switched off for these convolutions)."""
        x = grouped
        for w, b in layers:
            x = F.relu(F.conv2d(x, w[:, :, None, None], b), inplace=True)
        return x.amax(dim=3)

    # every set-abstraction module is three plan ops (so that bench.py can time a9 / a10 / the MLP separately):
    # all ball queries of the module, then per radius the grouping and the shared MLP + max
    def _sa_query(self, i):
        xyz, _, offs = self._source(i)
        radii = [g.radius for g in self.model.pnets[i].groupers]
        if i < 2:
            self.sa_sorted[i].build(xyz, offs).query(radii, self.cfg.SAMPLES_PN, self.keypoints, out=self.sa_idx[i])
            return
        bounds = self.sa_bounds[i].build(xyz, offs) if self.sa_bounds[i] is not None else None
        ops.ball_query_msg(radii, self.cfg.SAMPLES_PN, xyz, self.keypoints, offs, out=self.sa_idx[i], bounds=bounds)

    def _sa_group(self, i, r):
        xyz, feat, offs = self._source(i)
        self._grouped = ops.query_and_group_rows(xyz, feat, self.keypoints, self.sa_idx[i][r], offs)

    def _sa_mlp(self, i, r):
        o = self._mlp_max(self._grouped, self.sa_mlps[i][r])
        c0 = self._sa_c0[i] + sum(m[-1][0].shape[0] for m in self.sa_mlps[i][:r])
        self.kp_features[:, c0:c0 + o.shape[1]] = o

    def _packed_source(self, i):
        e = self.eng
        if i == 0:
            return self.pts_packed
        return e.featp[0][2] if i == 1 else e.featp[i - 1][0]

    def _sa_fused(self, i, r):
        xyz, _, offs = self._source(i)
        ops.sa_fused(self._packed_source(i), xyz, self.keypoints, self.sa_idx[i][r], self.sa_prepared[i][r],
                     self.kp_features, c_off=self._sa_c0[i] + sum(m.N2 for m in self.sa_prepared[i][:r]), row_offsets=offs)

    def _pack_points(self):   # intensity column of the raw (B, N, 4) rows -> 8-channel packed rows
        ops.pack_channel_major(self.points[..., 3].unsqueeze(1), 8, out=self.pts_packed)

    def _roi_pack(self):
        ops.pack_channel_major(self.kp_features, 512, out=self.kp_packed)

    def _roi_fused(self, r):
        ops.sa_fused(self.kp_packed, self.keypoints, self.gridpoints, self.roi_idx[r], self.roi_prepared[r],
                     self.roi_feat, c_off=96 * r)

    def _roi_reduce_fused(self):
        B, n, m = self.B, self.n, self.cfg.GRIDPOOL_NUM_GRIDPOINTS
        f = self.roi_feat.view(B, -1, n, m).permute(0, 2, 1, 3).contiguous().view(B, n, -1)   # roi_grid_pool.py:69-70
        self.pooled = self.model.roi_grid_pool.reduction(f)

    def _roi_query(self):
        pn = self.model.roi_grid_pool.pnet
        ops.ball_query_msg([g.radius for g in pn.groupers], self.cfg.SAMPLES_PN, self.keypoints, self.gridpoints, None,
                           out=self.roi_idx)

    def _roi_group(self, r):
        self._grouped = ops.query_and_group(self.keypoints, self.gridpoints, self.kp_features, self.roi_idx[r])

    def _roi_mlp(self, r):
        self._roi_out[r] = self._mlp_max(self._grouped, self.roi_mlps[r])       # (B, 96, n*16)

    def _roi_reduce(self):
        B, n, m = self.B, self.n, self.cfg.GRIDPOOL_NUM_GRIDPOINTS
        f = torch.cat(self._roi_out, 1)                                          # (B, 192, n*16)
        f = f.view(B, -1, n, m).permute(0, 2, 1, 3).contiguous().view(B, n, -1)  # roi_grid_pool.py:69-70
        self.pooled = self.model.roi_grid_pool.reduction(f)

    def _levels(self):
        e = self.eng
        for lv in range(4):
            if lv > 0:
                ops.batch_offsets(e.indices[lv], e.n_rows[lv], self.B, out=self.level_offsets[lv])
            ops.to_global(e.indices[lv], e.n_rows[lv], self.level_vs[lv], self.level_off, out=self.level_xyz[lv])

    def _bev(self):
        ops.bev_gather(self.eng.bev_nhwc, self.keypoints, self.cfg.GRID_BOUNDS[:2], self._bev_pixel, out=self.kp_features,
                       c_off=384)

    def _build_plan(self):
        chans = [sum(m[-1][0].shape[0] for m in sa) for sa in self.sa_mlps]      # 32, 32, 64, 128, 128
        self._sa_c0 = [int(sum(chans[:i])) for i in range(5)]
        assert sum(chans) == 384
        plan = [("fps+gather", lambda: ops.fps_keypoints(self.points, self.cfg.NUM_KEYPOINTS, self.kp_idx, self.keypoints))]
        for name, _, fn in self.eng.plan[:self.eng.n_backbone_ops]:
            plan.append(("backbone/" + name, fn))
        plan.append(("offsets+to_global", self._levels))
        if self.fused_sa:
            plan.append(("pack_points", self._pack_points))
        for i in range(5):
            plan.append(("sa%d/ball_query" % i, (lambda i=i: self._sa_query(i))))
            for r in range(len(self.cfg.SAMPLES_PN)):
                if self.fused_sa:
                    plan.append(("sa%d/fused_group+mlp+max_r%d" % (i, r), (lambda i=i, r=r: self._sa_fused(i, r))))
                else:
                    plan.append(("sa%d/group_r%d" % (i, r), (lambda i=i, r=r: self._sa_group(i, r))))
                    plan.append(("sa%d/mlp+max_r%d(torch)" % (i, r), (lambda i=i, r=r: self._sa_mlp(i, r))))
        self._bev_pixel = (np.asarray(self.cfg.VOXEL_SIZE[:2], np.float32) * np.float32(self.cfg.STRIDES[-1])).tolist()
        plan.append(("bev_gather", self._bev))
        plan.append(("roi/ball_query", self._roi_query))
        self._roi_out = [None] * len(self.cfg.SAMPLES_PN)
        if self.fused_sa:
            plan.append(("roi/pack_keypoint_features", self._roi_pack))
            for r in range(len(self.cfg.SAMPLES_PN)):
                plan.append(("roi/fused_group+mlp+max_r%d" % r, (lambda r=r: self._roi_fused(r))))
            plan.append(("roi/reduction_mlp(torch)", self._roi_reduce_fused))
        else:
            for r in range(len(self.cfg.SAMPLES_PN)):
                plan.append(("roi/group_r%d" % r, (lambda r=r: self._roi_group(r))))
                plan.append(("roi/mlp+max_r%d(torch)" % r, (lambda r=r: self._roi_mlp(r))))
            plan.append(("roi/reduction_mlp(torch)", self._roi_reduce))
        self.plan = plan

    def load(self, clouds, gridpoints):
        pts = np.stack(clouds, 0).astype(np.float32)
        assert pts.shape == (self.B, self.N, 4)
        self.points.copy_(torch.from_numpy(pts))
        self.gridpoints.copy_(torch.as_tensor(gridpoints).reshape(self.gridpoints.shape))

    def step(self):
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            if self.eng.overlap_rulebooks:   # rule-book chain of the backbone on the engine's second stream
                self.eng.run_overlapped([(name[9:] if name.startswith("backbone/") else name, fn) for name, fn in self.plan])
                return self.pooled
            for _, fn in self.plan:
                fn()
        return self.pooled

    def profile(self, iters=3):
        """[(op, us)] with CUDA events on the current stream, after one full step."""
        out = []
        self.step()
        torch.cuda.synchronize(self.dev)
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            for name, fn in self.plan:
                fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    fn()
                b.record()
                b.synchronize()
                out.append((name, a.elapsed_time(b) * 1e3 / iters))
        return out
