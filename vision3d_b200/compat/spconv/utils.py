"""spconv.utils.VoxelGenerator drop-in (constructed at vision3d/core/preprocess.py:17-24, called :30).
numpy in, numpy out like upstream, but the hashing/scatter runs on the GPU (vision3d_b200 kernels)."""
import numpy as np
import torch

from ... import ops


class VoxelGenerator:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, cap_policy=0):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = int(max_num_points)
        self._max_voxels = int(max_voxels)
        self._grid_size = np.round(grid_size).astype(np.int64)
        self._cap_policy = cap_policy
        self._impl = None  # created lazily: constructing a Preprocessor must not need a device

    def generate(self, points, max_voxels=None):
        """points (N, C) float32 numpy -> voxels (M, K, C) f32 zero padded, coordinates (M, 3) int32
        z,y,x, num_points_per_voxel (M,) int32."""
        points = np.ascontiguousarray(points, dtype=np.float32)
        n, c = points.shape
        mv = int(max_voxels or self._max_voxels)
        impl = self._impl
        if impl is None or impl[0].P < n or impl[0].max_voxels != mv or impl[1]["voxels"].shape[2] != c:
            vz = ops.Voxelizer(self._voxel_size, self._point_cloud_range, mv, self._max_num_points, 1,
                               max(n, 1 << 14), device="cuda", cap_policy=self._cap_policy)
            impl = self._impl = (vz, vz.alloc_outputs(c, with_mean=False))
        vz, out = impl
        dev = vz.device
        off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        vz.run(torch.from_numpy(points).to(dev), off, n, out)
        m = int(out["voxel_offsets"][1].item())
        return (out["voxels"][:m].cpu().numpy(), out["coords"][:m, 1:].cpu().numpy(),
                out["num_points"][:m].cpu().numpy())

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size
