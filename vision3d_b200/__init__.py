"""vision3d_b200: B200-native (sm_100a) kernels for the per-frame LiDAR hot path of jhultman/vision3d."""
from ._lib import V3DError  # noqa: F401
