"""Drop-in for the subset of `spconv` (v1.x API) that vision3d imports
(vision3d/detector/second.py:3,42; detector/sparse_cnn.py:12,16-30,136,153-186;
core/preprocess.py:7,18): SparseConvTensor, SparseSequential, SubMConv3d, SparseConv3d and
spconv.utils.VoxelGenerator. Every computation is a vision3d_b200 sm_100a kernel (no CPU path).

Parameter names/shapes match spconv v1.x so reference checkpoints load: conv weight
(k0, k1, k2, Cin, Cout), optional bias (Cout,).
"""
import math
from collections import OrderedDict

import torch
from torch import nn

from ... import ops
from . import utils  # noqa: F401  (spconv.utils.VoxelGenerator)

__all__ = ["SparseConvTensor", "SparseModule", "SparseSequential", "SparseConvolution", "SubMConv3d",
           "SparseConv3d", "utils"]


class SparseConvTensor(object):
    """features (N, C) f32, indices (N, 4) int32 [b, z, y, x] (second.py:42-44)."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices if indices.dtype == torch.int32 else indices.int()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return int(torch.tensor(self.spatial_shape).prod())

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def _n_rows(self):
        n = getattr(self, "_n_rows_dev", None)
        if n is None or int(self._n_rows_host) != self.features.shape[0]:
            self._n_rows_host = self.features.shape[0]
            self._n_rows_dev = torch.tensor([self._n_rows_host], dtype=torch.int32, device=self.features.device)
        return self._n_rows_dev

    def dense(self, channels_first=True):
        """(B, C, D, H, W) zero-filled (sparse_cnn.py:130)."""
        n = self.features.shape[0]
        out = ops.sparse_to_dense(self.features.contiguous(), self.indices.contiguous(), self._n_rows(), n,
                                  self.batch_size, self.spatial_shape)
        if not channels_first:
            return out.permute(0, 2, 3, 4, 1).contiguous()
        return out

    @property
    def sparity(self):  # (sic) upstream spelling
        return self.indices.shape[0] / (self.spatial_size * self.batch_size)


class SparseModule(nn.Module):
    """Marker base: modules that consume and return a SparseConvTensor."""
    pass


def _is_sparse(module):
    return isinstance(module, SparseModule)


class SparseSequential(SparseModule):
    """nn.Sequential that threads a SparseConvTensor through sparse modules and applies plain
    nn.Modules (BatchNorm1d, ReLU) to `.features`. Nestable and indexable (`self.blocks[0]`,
    sparse_cnn.py:139). In eval mode a [conv, BatchNorm1d, ReLU] run is folded into the conv kernel's
    epilogue (scale/shift/relu), which is what makes a layer a single launch."""

    def __init__(self, *args, **kwargs):
        super(SparseSequential, self).__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        k = 0
        while k < len(mods):
            module = mods[k]
            if isinstance(module, SparseConvolution) and not self.training and k + 1 < len(mods) and \
                    isinstance(mods[k + 1], nn.BatchNorm1d) and mods[k + 1].track_running_stats:
                bn = mods[k + 1]
                relu = k + 2 < len(mods) and isinstance(mods[k + 2], nn.ReLU)
                input = module(input, fold_bn=bn, fold_relu=relu)
                k += 3 if relu else 2
                continue
            if _is_sparse(module):
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input.features = module(input.features)
            else:
                input = module(input)
            k += 1
        return input


class _SparseConvFunction(torch.autograd.Function):
    """SubMConvFunction / SparseConvFunction of spconv v1.x (torch.autograd.Function). Forward is the
    fused sm_100a kernel; backward (training side) is composed from torch index ops."""

    @staticmethod
    def forward(ctx, features, weight, nbr, n_out_dev, n_out, scale, shift, relu, subm=False):
        prepared = isinstance(weight, ops.PreparedWeights)
        out = ops.sparse_conv(features.contiguous(), weight if prepared else weight.contiguous(), nbr, n_out_dev,
                              max(n_out, 1), scale, shift, relu)
        if not prepared:
            ctx.save_for_backward(features, weight, nbr)
        ctx.n_out, ctx.n_out_dev, ctx.subm = n_out, n_out_dev, bool(subm)
        ctx.fused = scale is not None or relu or prepared
        return out[:n_out]

    @staticmethod
    def backward(ctx, grad_out):
        """Hand-written kernels: dX = forward kernel on the inverted rule table with W^T, dW =
        v3d_sparse_conv_bwd_weight (ops.sparse_conv_backward)."""
        if ctx.fused:
            raise RuntimeError("backward through a BN/ReLU-folded sparse conv is not defined; call .train()")
        features, weight, nbr = ctx.saved_tensors
        g_feat, g_w = ops.sparse_conv_backward(features, weight, nbr, ctx.n_out_dev, ctx.n_out, grad_out.contiguous(),
                                               ctx.subm, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g_feat, (g_w.reshape(weight.shape) if g_w is not None else None), None, None, None, None, None, None, None


def _triple(v, ndim=3):
    if isinstance(v, (list, tuple)):
        assert len(v) == ndim
        return [int(x) for x in v]
    return [int(v)] * ndim


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 fused_bn=False):
        super(SparseConvolution, self).__init__()
        assert groups == 1 and ndim == 3 and not transposed and not inverse
        self.ndim = ndim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.subm = subm
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.Tensor(*self.kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def _rulebook(self, input):
        dev = input.features.device
        n = input.features.shape[0]
        n_dev = input._n_rows()
        ind = input.indices.contiguous()
        cached = input.find_indice_pair(self.indice_key)
        level = input.indice_dict.get(("__table__", tuple(input.spatial_shape), n))
        if self.subm:
            if cached is not None:
                return cached
            if level is None:
                level = ops.SiteTable(max(n, 1), dev).build(ind, n_dev, input.spatial_shape)
                input.indice_dict[("__table__", tuple(input.spatial_shape), n)] = level
            nbr = ops.rulebook_subm(level, ind, n_dev, input.spatial_shape, self.kernel_size, self.dilation)
            rb = (nbr, ind, n_dev, n, list(input.spatial_shape))
            if self.indice_key is not None:
                input.indice_dict[self.indice_key] = rb
            return rb
        if level is None:
            level = ops.SiteTable(max(n, 1), dev).build(ind, n_dev, input.spatial_shape)
            input.indice_dict[("__table__", tuple(input.spatial_shape), n)] = level
        out_shape = ops.conv_out_shape(input.spatial_shape, self.kernel_size, self.stride, self.padding, self.dilation)
        cells = input.batch_size * out_shape[0] * out_shape[1] * out_shape[2]
        reach = 1
        for d in range(3):  # outputs one input can reach along a dimension
            reach *= min(self.kernel_size[d], (self.kernel_size[d] - 1) * self.dilation[d] // self.stride[d] + 1)
        cap = int(max(1, min(cells, n * reach)))
        out_idx, n_out_dev, nbr, _ = ops.rulebook_conv(level, ind, n_dev, input.batch_size, input.spatial_shape,
                                                       self.kernel_size, self.stride, self.padding, self.dilation, cap)
        n_out = int(n_out_dev.item())  # the reference API exposes exact-size tensors: one sync here
        if n_out > cap:
            raise ops.V3DError("strided sparse conv produced %d sites > capacity %d" % (n_out, cap))
        return (nbr, out_idx[:n_out], n_out_dev, n_out, out_shape)

    def forward(self, input, fold_bn=None, fold_relu=False):
        assert isinstance(input, SparseConvTensor)
        nbr, out_idx, n_out_dev, n_out, out_shape = self._rulebook(input)
        scale = shift = None
        if fold_bn is not None:
            inv = torch.rsqrt(fold_bn.running_var + fold_bn.eps)
            gamma = fold_bn.weight if fold_bn.affine else torch.ones_like(inv)
            beta = fold_bn.bias if fold_bn.affine else torch.zeros_like(inv)
            scale = (gamma * inv).float().contiguous()
            shift = (beta - fold_bn.running_mean * gamma * inv).float()
            if self.bias is not None:
                shift = shift + self.bias * scale
            shift = shift.contiguous()
        elif self.bias is not None:
            scale = torch.ones_like(self.bias)
            shift = self.bias
        weight = self.weight
        if not (torch.is_grad_enabled() and (self.weight.requires_grad or input.features.requires_grad)):
            # inference: tcgen05 path with the weight image prepared once per weight version
            ver = (self.weight._version, self.weight.data_ptr())
            if getattr(self, "_prepared_ver", None) != ver:
                self._prepared, self._prepared_ver = ops.PreparedWeights(self.weight), ver
            weight = self._prepared
        feats = _SparseConvFunction.apply(input.features, weight, nbr, n_out_dev, n_out,
                                          scale.detach() if scale is not None else None,
                                          shift.detach() if shift is not None else None,
                                          bool(fold_relu and fold_bn is not None), self.subm)
        out = SparseConvTensor(feats, out_idx, out_shape, input.batch_size)
        out.indice_dict = input.indice_dict if self.subm else {}
        if self.subm:
            out._n_rows_host, out._n_rows_dev = n_out, n_out_dev
        out.grid = input.grid
        return out


class SubMConv3d(SparseConvolution):
    """vision3d passes a stray positional `3` that lands in the stride slot
    (detector/sparse_cnn.py:15-17,154-172): accepted and ignored, submanifold stride is 1."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None):
        super(SubMConv3d, self).__init__(3, in_channels, out_channels, kernel_size, 1, padding, dilation, groups,
                                         bias, True, indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None):
        super(SparseConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation,
                                           groups, bias, indice_key=indice_key)
