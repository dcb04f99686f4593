"""torchsearchsorted.searchsorted shim (vision3d/detector/sparse_cnn.py:11,112): per-row binary search.
No kernel of ours is needed -- torch.searchsorted is an exact functional stand-in."""
import torch


def searchsorted(a, v, out=None, side="left"):
    res = torch.searchsorted(a.contiguous(), v.contiguous(), right=(side == "right"))
    if out is not None:
        out.copy_(res)
        return out
    return res
