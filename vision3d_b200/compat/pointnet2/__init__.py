"""Drop-in for the subset of sshaoshuai/Pointnet2.PyTorch that vision3d imports
(vision3d/detector/model.py:6-7; detector/roi_grid_pool.py:5)."""
from . import pointnet2_modules, pointnet2_utils  # noqa: F401
