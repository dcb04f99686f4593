"""ctypes binding of the C-ABI in include/v3d_b200.h (vision3d_b200/libv3d_b200.so).

There is no fallback of any kind: if the shared library is missing or a call fails, an exception is
raised. Build with `python -m vision3d_b200.build` (or __graft_entry__.build()).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libv3d_b200.so")

c_int, c_float, c_size_t, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
P = c_void_p  # every device pointer and the stream travel as integers

# name -> (restype, argtypes); mirrors include/v3d_b200.h declaration by declaration
SIGNATURES = {
    "v3d_abi_version": (c_int, []),
    "v3d_status_string": (ctypes.c_char_p, [c_int]),
    "v3d_last_cuda_error": (ctypes.c_char_p, []),
    "v3d_cudart_version": (c_int, []),
    "v3d_check_device": (c_int, []),
    "v3d_sparse_conv_tc_variant": (c_int, []),
    "v3d_box_iou_rotated": (c_int, [P, c_int, P, c_int, P, P]),
    "v3d_nms_rotated_workspace_bytes": (c_size_t, [c_int]),
    "v3d_nms_rotated": (c_int, [P, P, c_int, c_float, P, P, P, c_size_t, P]),
    "v3d_nms_rotated_grouped": (c_int, [P, P, c_int, c_int, c_float, P, P, P, c_size_t, P]),
    "v3d_match_anchors": (c_int, [P, c_int, P, c_int, c_int, P, P, P, P, P, P, P]),
    "v3d_voxelize_workspace_bytes": (c_size_t, [c_int, c_int]),
    "v3d_voxelize_workspace_init": (c_int, [P, c_size_t, c_int, c_int, P]),
    "v3d_voxelize_batch": (c_int, [P, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, c_int,
                                   P, P, P, P, P, P, c_size_t, c_int, P]),
    "v3d_site_table_bytes": (c_size_t, [c_int]),
    "v3d_site_table_init": (c_int, [P, c_size_t, c_int, P]),
    "v3d_site_table_build": (c_int, [P, P, P, c_int, P, P]),
    "v3d_rulebook_subm": (c_int, [P, P, P, c_int, P, P, P, P, c_int, P]),
    "v3d_rulebook_conv_workspace_bytes": (c_size_t, [c_int, P, c_int, c_int]),
    "v3d_conv_out_shape": (None, [P, P, P, P, P, P]),
    "v3d_rulebook_conv": (c_int, [P, P, P, c_int, c_int, P, P, P, P, P, P, P, c_int, P, c_int, P,
                                  c_size_t, P]),
    "v3d_rulebook_subm_ranked": (c_int, [P, c_int, c_int, P, P, c_int, P, P, P, P, c_int, P]),
    "v3d_rulebook_conv_ranked": (c_int, [P, c_int, P, P, c_int, c_int, P, P, P, P, P, P, P, c_int, P, c_int, P,
                                         c_size_t, P]),
    "v3d_sparse_conv_fwd": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P]),
    "v3d_rulebook_invert": (c_int, [P, c_int, P, c_int, c_int, P, c_int, P]),
    "v3d_sparse_conv_bwd_weight": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, P]),
    "v3d_sparse_conv_prepared_bytes": (c_size_t, [c_int, c_int, c_int]),
    "v3d_sparse_conv_prepare": (c_int, [P, c_int, c_int, c_int, P, c_size_t, P]),
    "v3d_feature_pack": (c_int, [P, P, c_int, c_int, c_int, P, P]),
    "v3d_sparse_conv_fwd_tc": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P]),
    "v3d_sparse_to_dense_workspace_bytes": (c_size_t, [c_int, P]),
    "v3d_sparse_to_dense": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "v3d_sparse_to_dense_nhwc": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "v3d_second_head_decode": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "v3d_head_cls_logits": (c_int, [P, c_int, c_int, c_int, P, P, c_int, P, P]),
    "v3d_topk_rows_workspace_bytes": (c_size_t, [c_int, c_int]),
    "v3d_topk_rows": (c_int, [P, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "v3d_head_reg_gather": (c_int, [P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "v3d_second_head_decode_compact": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "v3d_pack_detections": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P, c_int, P, P]),
    "v3d_fps_workspace_bytes": (c_size_t, [c_int, c_int]),
    "v3d_fps": (c_int, [P, c_int, c_int, c_int, P, P, c_size_t, P]),
    "v3d_gather": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "v3d_ball_query": (c_int, [P, P, c_int, c_int, c_int, c_float, c_int, P, P]),
    "v3d_group": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P]),
    "v3d_query_and_group": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P]),
    "v3d_fps_keypoints": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P]),
    "v3d_ball_query_msg": (c_int, [P, c_int, P, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "v3d_ball_query_bounds_bytes": (c_size_t, [c_int, c_int]),
    "v3d_ball_query_bounds": (c_int, [P, c_int, P, c_int, c_int, c_int, P, P]),
    "v3d_ball_query_msg_culled": (c_int, [P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "v3d_ball_query_sort_workspace_bytes": (c_size_t, [c_int]),
    "v3d_ball_query_sort_x": (c_int, [P, c_int, P, c_int, c_int, c_int, c_float, c_float, P, P, P]),
    "v3d_ball_query_msg_select": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "v3d_bev_gather": (c_int, [P, c_int, c_int, c_int, c_int, P, c_int, c_float, c_float, c_float, c_float, P, c_int,
                               c_int, P]),
    "v3d_query_and_group_rows": (c_int, [P, c_int, P, c_int, c_int, c_int, P, P, P, c_int, c_int, c_int, P, P]),
    "v3d_sa_mlp_prepared_bytes": (c_size_t, [c_int, c_int]),
    "v3d_sa_mlp_prepare": (c_int, [P, c_int, c_int, P, c_size_t, P]),
    "v3d_sa_fused": (c_int, [P, c_int, P, c_int, P, c_int, P, P, c_int, c_int, c_int, P, P, c_int, P, P, c_int, P,
                             c_int, c_int, P]),
    "v3d_pack_channel_major": (c_int, [P, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int,
                                       c_int, P, P]),
    "v3d_batch_offsets": (c_int, [P, P, c_int, c_int, P, P]),
    "v3d_to_global": (c_int, [P, P, c_int, P, P, P, P]),
    "v3d_pad_batch": (c_int, [P, c_int, P, c_int, c_int, ctypes.c_ulonglong, P, P]),
}

_lib = None


class V3DError(RuntimeError):
    """Raised for any non-zero status from the C-ABI (the reference raises RuntimeError from
    AT_ASSERTM / AT_ERROR: box_iou_rotated_cuda.cu:69-70, box_iou_rotated.h:28)."""


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise V3DError(
                "vision3d_b200: %s is missing. There is no CPU or PyTorch fallback; build it with "
                "`python -m vision3d_b200.build`." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.v3d_status_string(status).decode()
        if status == -3:
            msg += ": " + lib.v3d_last_cuda_error().decode()
        raise V3DError("%s failed: %s" % (what, msg))


def i3(v):
    """Host int[3] argument."""
    return (ctypes.c_int * 3)(*[int(x) for x in v])


def f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])
