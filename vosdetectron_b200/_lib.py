"""ctypes binding of libvosd_b200.so (the C ABI declared in include/vosd_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
returns a negative status this module raises.  The product never computes on
the CPU and never imports anything under oracle/.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VOSD_B200_LIB: A/B tuning builds (tools/ab_build.sh); the product is always the in-tree library
LIB_PATH = os.environ.get("VOSD_B200_LIB") or os.path.join(HERE, "libvosd_b200.so")

MAX_LEVELS = 8
MAX_ANCHORS = 16
MAX_TOPK = 16384

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
vp = ctypes.c_void_p


class RpnLevel(ctypes.Structure):
    """struct vosd_rpn_level"""
    _fields_ = [("scores", vp), ("deltas", vp),
                ("height", ctypes.c_int), ("width", ctypes.c_int), ("num_anchors", ctypes.c_int),
                ("feat_stride", ctypes.c_double),
                ("anchors", ctypes.c_double * (4 * MAX_ANCHORS))]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "vosd_version": (ctypes.c_char_p, []),
    "vosd_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "vosd_launch_count": (ctypes.c_ulonglong, []),
    "vosd_set_device": (ctypes.c_int, [ctypes.c_int]),
    "vosd_debug_force_generic": (ctypes.c_int, [ctypes.c_int]),
    "vosd_roialign_fwd": (ctypes.c_int, [vp, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "vosd_roialign_bwd": (ctypes.c_int, [vp, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                         ctypes.c_int, vp]),
    "vosd_roialign_ml_fwd": (ctypes.c_int, [ctypes.POINTER(vp), c_int_p, c_int_p, c_float_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            vp, vp, vp, vp, vp]),
    "vosd_roialign_fwd_workspace_bytes": (ctypes.c_size_t, [c_int_p, c_int_p] + [ctypes.c_int] * 6),
    "vosd_roialign_ml_fwd_ws": (ctypes.c_int, [ctypes.POINTER(vp), c_int_p, c_int_p, c_float_p, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
    "vosd_roialign_ml_fwd_nhwc": (ctypes.c_int, [ctypes.POINTER(vp), c_int_p, c_int_p, c_float_p, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, vp, vp, vp, vp, vp]),
    "vosd_roialign_ml_bwd": (ctypes.c_int, [vp, ctypes.POINTER(vp), c_int_p, c_int_p, c_float_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, vp, vp, vp, ctypes.c_int, vp]),
    "vosd_roialign_ml_bwd_nhwc": (ctypes.c_int, [vp, ctypes.POINTER(vp), c_int_p, c_int_p, c_float_p, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, vp, vp, vp, ctypes.c_int, vp]),
    "vosd_proposals_capacity": (ctypes.c_int, [ctypes.POINTER(RpnLevel), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "vosd_generate_proposals_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(RpnLevel), ctypes.c_int,
                                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "vosd_generate_proposals": (ctypes.c_int, [ctypes.POINTER(RpnLevel), ctypes.c_int, ctypes.c_int, vp,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                               vp, vp, vp, vp, ctypes.c_size_t, vp]),
    "vosd_decode_anchors": (ctypes.c_int, [ctypes.POINTER(RpnLevel), ctypes.c_int, vp, vp, vp]),
    "vosd_any_nan": (ctypes.c_int, [vp, ctypes.c_size_t, vp, vp]),
    "vosd_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "vosd_nms": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_float, vp, vp, vp, ctypes.c_size_t, vp]),
    "vosd_collect_distribute_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "vosd_collect_distribute": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                               ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
    "vosd_distribute": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       vp, vp, vp, vp, vp]),
    "vosd_bbox_transform": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_float,
                                           ctypes.c_float, vp, vp]),
    "vosd_box_results_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "vosd_box_results": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_size_t, vp]),
    "vosd_pack_mask_bits": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_longlong, vp, vp]),
    "vosd_paste_rle_smem_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "vosd_paste_rle": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_float, vp, ctypes.c_longlong, vp, ctypes.c_longlong, vp, vp, vp, vp, vp,
                                      vp, vp]),
    "vosd_paste_masks_packed": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_float, vp, vp, vp]),
    "vosd_flow_align_fwd": (ctypes.c_int, [ctypes.c_int] * 4 + [vp, vp, vp, vp]),
    "vosd_flow_align_bwd": (ctypes.c_int, [ctypes.c_int] * 4 + [vp, vp, vp, vp, vp, ctypes.c_int, vp]),
    "vosd_flow_align_ml_fwd": (ctypes.c_int, [ctypes.c_int] * 3 + [c_int_p, c_int_p, ctypes.POINTER(vp),
                                              ctypes.POINTER(vp), ctypes.POINTER(vp), vp]),
    "vosd_flow_align_ml_bwd": (ctypes.c_int, [ctypes.c_int] * 3 + [c_int_p, c_int_p] + [ctypes.POINTER(vp)] * 5
                               + [ctypes.c_int, vp]),
    "vosd_debug_flow_align_fast": (ctypes.c_int, [ctypes.c_int]),
    "vosd_mask_iou_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "vosd_mask_iou_nms": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_longlong, vp, ctypes.c_double, vp, vp, vp,
                                         ctypes.c_size_t, vp]),
    "vosd_rle_to_bits": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_longlong, vp, ctypes.c_int, vp]),
    "vosd_bbox_overlaps": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_int, vp, vp, vp, vp]),
    "vosd_bbox_targets": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, vp, vp, vp, vp]),
    "vosd_sample_rois": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_float, vp, vp, vp, vp]),
    "vosd_paste_masks": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_float, vp, vp, vp]),
}

_lib = None


class VosdError(RuntimeError):
    def __init__(self, fn, status, msg):
        super().__init__("%s failed with status %d (%s)" % (fn, status, msg))
        self.status = status


def load():
    """Load (once) and return the ctypes handle; raises if the .so is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s not found: build it with `python -m vosdetectron_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(fn_name, status):
    if status != 0:
        raise VosdError(fn_name, status, load().vosd_status_string(status).decode())


def call(fn_name, *args):
    """Call an int-returning entry point and raise on a negative status."""
    check(fn_name, getattr(load(), fn_name)(*args))


def launch_count():
    return int(load().vosd_launch_count())
