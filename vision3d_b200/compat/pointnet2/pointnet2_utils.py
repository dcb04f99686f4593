"""pointnet2.pointnet2_utils drop-in: the torch.autograd.Function aliases vision3d calls
(furthest_point_sample, gather_operation at detector/model.py:53-54) plus what
PointnetSAModuleMSG needs (ball_query, grouping_operation, QueryAndGroup, GroupAll).
Forward passes are vision3d_b200 sm_100a kernels; index outputs are non-differentiable; the two
copy ops get a torch scatter-add backward (training side)."""
import torch
from torch import nn
from torch.autograd import Function

from ... import ops


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B, N, 3) contiguous f32 -> (B, npoint) int32, start index 0."""
        out = ops.furthest_point_sample(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B, C, N), idx (B, m) int32 -> (B, C, m)."""
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return ops.gather_operation(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, m = grad_out.shape
        g = grad_out.new_zeros((B, C, ctx.n))
        g.scatter_add_(2, idx.long().unsqueeze(1).expand(-1, C, -1), grad_out.contiguous())
        return g, None


gather_operation = GatherOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        """xyz (B, N, 3), new_xyz (B, M, 3) -> (B, M, nsample) int32."""
        out = ops.ball_query(radius, nsample, xyz, new_xyz)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B, C, N), idx (B, M, ns) int32 -> (B, C, M, ns)."""
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return ops.grouping_operation(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, M, ns = grad_out.shape
        g = grad_out.new_zeros((B, C, ctx.n))
        g.scatter_add_(2, idx.long().view(B, 1, M * ns).expand(-1, C, -1), grad_out.reshape(B, C, M * ns))
        return g, None


grouping_operation = GroupingOperation.apply


class _QueryGroupFused(Function):
    """[xyz[idx] - new_xyz ; features[idx]] in one kernel (no 3 separate tensors + cat)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = xyz.shape[1]
        ctx.has_feat = features is not None
        return ops.query_and_group(xyz, new_xyz, features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        """Upstream QueryAndGroup is differentiable w.r.t. xyz and new_xyz too (grouping_operation(xyz) minus
        new_xyz): channels 0..2 scatter-add into xyz[idx] and sum (negated) over the samples into new_xyz."""
        (idx,) = ctx.saved_tensors
        B, CT, M, ns = grad_out.shape
        flat = idx.long().view(B, 1, M * ns)
        g_xyz = g_new = g_feat = None
        if ctx.needs_input_grad[0]:
            g = grad_out.new_zeros((B, 3, ctx.n))
            g.scatter_add_(2, flat.expand(-1, 3, -1), grad_out[:, :3].reshape(B, 3, M * ns))
            g_xyz = g.transpose(1, 2).contiguous()
        if ctx.needs_input_grad[1]:
            g_new = -grad_out[:, :3].sum(-1).transpose(1, 2).contiguous()
        if ctx.has_feat and ctx.needs_input_grad[2]:
            gf = grad_out[:, 3:].reshape(B, CT - 3, M * ns)
            g_feat = grad_out.new_zeros((B, CT - 3, ctx.n))
            g_feat.scatter_add_(2, flat.expand(-1, CT - 3, -1), gf)
        return g_xyz, g_new, g_feat, None


class QueryAndGroup(nn.Module):
    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        """xyz (B, N, 3), new_xyz (B, M, 3), features (B, C, N) -> (B, 3 + C, M, nsample)."""
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if self.use_xyz:
            return _QueryGroupFused.apply(xyz, new_xyz, features, idx)
        assert features is not None, "Cannot have not features and not use xyz as a feature!"
        return grouping_operation(features, idx)


class GroupAll(nn.Module):
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                return torch.cat([grouped_xyz, grouped_features], dim=1)
            return grouped_features
        return grouped_xyz
