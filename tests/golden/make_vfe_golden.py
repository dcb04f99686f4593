"""Golden vectors for VoxelFeatureExtractor.forward (detector/layers.py:10-17) from the reference's own class:
    python tests/golden/make_vfe_golden.py     # needs /root/reference; writes tests/golden/vfe_golden.npz"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_layers", "/root/reference/vision3d/detector/layers.py")
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
rng = np.random.default_rng(3)
N = 512
occ = rng.integers(1, 6, N).astype(np.int32)
v = (rng.normal(size=(N, 5, 4)) * 30).astype(np.float32)
for i in range(5):
    v[occ <= i, i] = 0           # zero padding past the occupancy, as the voxel generator leaves it
mean = m.VoxelFeatureExtractor()(torch.from_numpy(v), torch.from_numpy(occ)).numpy()
np.savez_compressed(os.path.join(HERE, "vfe_golden.npz"), voxels=v, occupancy=occ, mean=mean)
print(mean.shape)
