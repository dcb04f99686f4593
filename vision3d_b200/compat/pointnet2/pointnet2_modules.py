"""pointnet2.pointnet2_modules drop-in: PointnetSAModuleMSG as vision3d constructs and calls it
(detector/model.py:39-43,64; detector/roi_grid_pool.py:28-32,68): grouping by the vision3d_b200
kernels, the shared MLP (1x1 Conv2d + BatchNorm2d + ReLU) and the max over samples in torch."""
from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from . import pointnet2_utils


class SharedMLP(nn.Sequential):
    """pytorch_utils.SharedMLP: Conv2d 1x1 (bias only without BN) -> BatchNorm2d -> ReLU per layer."""

    def __init__(self, args: List[int], bn: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            block = nn.Sequential()
            conv = nn.Conv2d(args[i], args[i + 1], kernel_size=(1, 1), stride=(1, 1), bias=not bn)
            nn.init.kaiming_normal_(conv.weight)
            if conv.bias is not None:
                nn.init.constant_(conv.bias, 0)
            block.add_module("conv", conv)
            if bn:
                norm = nn.BatchNorm2d(args[i + 1])
                nn.init.constant_(norm.weight, 1.0)
                nn.init.constant_(norm.bias, 0)
                block.add_module("bn", nn.Sequential(OrderedNamed("bn", norm)))
            block.add_module("activation", nn.ReLU(inplace=True))
            self.add_module("layer{}".format(i), block)


def OrderedNamed(name, module):
    from collections import OrderedDict
    return OrderedDict([(name, module)])


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = "max_pool"

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        """xyz (B, N, 3), features (B, C, N), new_xyz (B, M, 3) -> (new_xyz, (B, sum Cout, M))."""
        new_features_list = []
        if new_xyz is None:
            if self.npoint is not None:
                xyz_flipped = xyz.transpose(1, 2).contiguous()
                idx = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
                new_xyz = pointnet2_utils.gather_operation(xyz_flipped, idx).transpose(1, 2).contiguous()
        for i in range(len(self.groupers)):
            new_features = self.groupers[i](xyz, new_xyz, features)       # (B, C, M, ns)
            new_features = self.mlps[i](new_features)                      # (B, Cout, M, ns)
            if self.pool_method == "max_pool":
                new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            elif self.pool_method == "avg_pool":
                new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            else:
                raise NotImplementedError
            new_features_list.append(new_features.squeeze(-1))            # (B, Cout, M)
        return new_xyz, torch.cat(new_features_list, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping set abstraction. NOTE: like upstream, adds 3 to mlps[i][0] IN PLACE when
    use_xyz -- vision3d deep-copies its channel lists because of this (detector/model.py:36,42)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, pool_method="max_pool", instance_norm=False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for i in range(len(radii)):
            radius, nsample = radii[i], nsamples[i]
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            mlp_spec = mlps[i]
            if use_xyz:
                mlp_spec[0] += 3
            self.mlps.append(SharedMLP(mlp_spec, bn=bn))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method="max_pool", instance_norm=False):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm)
