"""Provider of the `vision3d._C` extension module surface (reference vision3d/ops/csrc/vision.cpp:60-65,
built by reference setup.py:51-59): box_iou_rotated, nms_rotated, get_compiler_version,
get_cuda_version. vision3d/ops/iou_nms.py:8-9,85 is the caller and is used unchanged."""
from .. import _lib, ops


def box_iou_rotated(boxes1, boxes2):
    """(M,5),(N,5) -> (M,N) f32 on the inputs' device (box_iou_rotated.h:20-32). CUDA only."""
    return ops.box_iou_rotated(boxes1, boxes2)


def nms_rotated(dets, scores, iou_threshold):
    """(N,5),(N,),float -> (K,) int64 on the inputs' device, descending score (nms_rotated.h:22-36)."""
    return ops.nms_rotated(dets, scores, iou_threshold)


def get_cuda_version():
    # vision.cpp:21-32 formatting of CUDART_VERSION
    v = _lib.load().v3d_cudart_version()
    s = "%d.%d" % (v // 1000, v // 10 % 100)
    if v % 10:
        s += ".%d" % (v % 10)
    return s


def get_compiler_version():
    return "nvcc sm_100a (vision3d_b200)"
