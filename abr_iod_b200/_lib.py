"""ctypes binding of ``libabr_b200.so`` (the C ABI declared in ``include/abr_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing the import of
any op fails with instructions to build it, and every op raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ABR_B200_LIB: another build of the same library (A/B measurements of kernel variants with tools/*; nothing else)
LIB_PATH = os.environ.get("ABR_B200_LIB") or os.path.join(_HERE, "lib", "libabr_b200.so")

ABR_F32, ABR_BF16 = 0, 1
ABR_NCHW, ABR_NHWC, ABR_NCHW_MAPS_NHWC_POOLED = 0, 1, 2
ABR_MAX_LEVELS = 8

_vp, _int, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


class PasteImage(ctypes.Structure):
    """``abr_paste_image_t``"""
    _fields_ = [("offset", ctypes.c_int64), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("first_op", ctypes.c_int32), ("n_ops", ctypes.c_int32)]


class PasteOp(ctypes.Structure):
    """``abr_paste_op_t``"""
    _fields_ = [("kind", ctypes.c_int32), ("y0", ctypes.c_int32), ("x0", ctypes.c_int32), ("y1", ctypes.c_int32),
                ("x1", ctypes.c_int32), ("src_width", ctypes.c_int32), ("sy0", ctypes.c_int32), ("sx0", ctypes.c_int32),
                ("src_offset", ctypes.c_int64), ("lam", ctypes.c_double), ("fill", ctypes.c_int32),
                ("pad_", ctypes.c_int32)]


class ResizeJob(ctypes.Structure):
    """``abr_resize_job_t``"""
    _fields_ = [("src_offset", ctypes.c_int64), ("dst_offset", ctypes.c_int64), ("tmp_offset", ctypes.c_int64),
                ("src_h", ctypes.c_int32), ("src_w", ctypes.c_int32), ("dst_h", ctypes.c_int32), ("dst_w", ctypes.c_int32),
                ("x_taps", ctypes.c_int32), ("x_ksize", ctypes.c_int32), ("y_taps", ctypes.c_int32), ("y_ksize", ctypes.c_int32)]


PASTE_FILL, PASTE_COPY, PASTE_BLEND = 0, 1, 2

# name -> (restype, argtypes); mirrors include/abr_b200.h one to one (tests/test_abi.py checks the header)
SIGNATURES = {
    "abr_version": (_int, []),
    "abr_last_error": (ctypes.c_char_p, []),
    "abr_launch_count": (ctypes.c_uint64, []),
    "abr_set_option": (_int, [ctypes.c_char_p, _int]),
    "abr_stage_timing_begin": (_int, [_int]),
    "abr_stage_timing_begin_every": (_int, [_int, _int]),
    "abr_stage_timing_end": (_int, [_vp, _int]),
    "abr_roi_align_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "abr_roi_align_workspace_bytes_nchw": (_sz, [_int, _int, _int, _int, _int, _int, ctypes.c_longlong, _int]),
    "abr_roi_align_workspace_bytes_layout": (_sz, [_int, _int, _int, _int, _int, _int, ctypes.c_longlong, _int, _int]),
    "abr_roi_align_forward": (_int, [_vp, _vp, _vp] + [_int] * 7 + [_f, _int, _int, _int, _vp, _sz, _int, _vp]),
    "abr_roi_align_backward": (_int, [_vp, _vp, _vp] + [_int] * 7 + [_f, _int, _int, _int, _int, _vp, _sz, _int, _vp]),
    "abr_roi_align_multilevel_forward": (_int, [_vp, _vp, _vp, _vp, _int, _vp, _vp, _vp] + [_int] * 8 + [_vp, _sz, _int, _vp]),
    "abr_roi_align_multilevel_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int] + [_int] * 9 + [_vp, _sz, _int, _vp]),
    "abr_fpn_map_levels": (_int, [_vp, _vp, _int, _f, _f, _f, _f, _f, _vp]),
    "abr_roi_pool_forward": (_int, [_vp, _vp, _vp, _vp] + [_int] * 7 + [_f, _int, _int, _vp]),
    "abr_roi_pool_backward": (_int, [_vp, _vp, _vp, _vp] + [_int] * 7 + [_int, _int, _int, _vp]),
    "abr_nms_workspace_bytes": (_sz, [_vp, _int]),
    "abr_nms_batched": (_int, [_vp, _vp, _vp, _int, _f, _int, _int, _vp, _int, _vp, _vp, _sz, _vp]),
    "abr_rpn_proposals_workspace_bytes": (_sz, [_int] * 6),
    "abr_rpn_proposals": (_int, [_vp, _vp, _vp, ctypes.c_longlong, _vp] + [_int] * 7 + [_f, _int, _f, _vp, _f,
                                 _vp, _vp, _vp, _vp, _int, _vp, _sz, _vp]),
    "abr_box_postprocess_workspace_bytes": (_sz, [_vp, _int, _int]),
    "abr_box_postprocess": (_int, [_vp, _vp, _int, _int, _vp, _vp, _vp, _int, _int, _f, _f, _int, _int, _vp, _f,
                                   _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _int, _vp, _sz, _vp]),
    "abr_ard_workspace_bytes": (_sz, [_int, _int, _int]),
    "abr_ard_forward_backward": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _f, _f, _int, _int, _vp, _sz, _vp]),
    "abr_roi_ard_fused_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "abr_roi_ard_fused": (_int, [_vp] * 7 + [_int] * 7 + [_f, _int, _f, _f, _int, _int, _int, _vp, _sz, _int, _vp]),
    "abr_match_proposals": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "abr_box_iou": (_int, [_vp, _int, _vp, _int, _vp, _vp]),
    "abr_logit_loss_workspace_bytes": (_sz, [_int]),
    "abr_roi_distillation_id": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "abr_fastrcnn_loss": (_int, [_vp, _vp, _int, _vp, _vp, _int, _int, _int, _int, _f, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "abr_channel_mean": (_int, [_vp, _int, _int, _int, _int, _int, _vp, _vp]),
    "abr_prototype_distances": (_int, [_vp, _int, _int, _vp, _vp, _vp]),
    "abr_prototype_herding": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _sz, _vp]),
    "abr_sample_fg_bg": (_int, [_vp, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "abr_scale_if_needed": (_int, [_vp, _sz, _vp, _f, _int, _vp]),
    "abr_paste_batch": (_int, [_vp, _vp, _int, _vp, _int, _vp, _int, _vp]),
    "abr_resize_bicubic_batch": (_int, [_vp, _vp, _vp, _int, _vp, _int, _vp]),
}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "abr_iod_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C abr_iod_b200/csrc`); there is no CPU or PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name, None)
            if fn is None and os.environ.get("ABR_B200_LIB"):
                continue  # an older build under A/B test may lack the newest entry points
            if fn is None:
                raise ImportError("abr_iod_b200: %s does not export %s -- rebuild it" % (LIB_PATH, name))
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    """The reference surfaces native failures as RuntimeError (AT_ASSERTM / THCudaCheck); so do we."""
    if rc != 0:
        raise RuntimeError("abr_b200 error %d: %s" % (rc, lib().abr_last_error().decode()))


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: abr_iod_b200 has no CPU path (got device %s)" % (what, t.device))


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return ABR_F32
    if t.dtype == torch.bfloat16:
        return ABR_BF16
    raise RuntimeError("unsupported dtype %s (float32 and bfloat16 are native; float16 is upcast by the wrappers)" % t.dtype)


def as_compute_dtype(t: torch.Tensor) -> torch.Tensor:
    """``amp.float_function`` of the reference (layers/roi_align.py:58, nms.py:8) runs these ops in fp32;
    bfloat16 additionally has a native path here, everything else is upcast."""
    return t if t.dtype in (torch.float32, torch.bfloat16) else t.float()


def is_channels_last(t: torch.Tensor) -> bool:
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


NCHW_STAGING = True  # run contiguous-NCHW ROIAlign calls through the channels-last kernels (costs scratch memory)
# Contiguous feature maps, but channels-last RoI features (same logical [R,C,P,P] tensor, channels_last strides): saves the
# two passes over the pooled tensor that a contiguous result costs (1.6x on the RoI path of a contiguous model).  Off by
# default because a consumer that calls ``.view(R, -1)`` on the pooled tensor (the FPN MLP head) needs ``.reshape``.
POOLED_CHANNELS_LAST = False


def roi_align_layout(x: torch.Tensor) -> int:
    """Layout code of a ROIAlign call on feature map ``x``."""
    if is_channels_last(x):
        return ABR_NHWC
    return ABR_NCHW_MAPS_NHWC_POOLED if POOLED_CHANNELS_LAST else ABR_NCHW


def roi_align_workspace(R, PH, PW, max_h, device, channels_last=True, nchw_staging=None, layout=None):
    """Scratch for ROIAlign (caller-owned, per call): the per-RoI plans of the channels-last kernels and, for a
    contiguous-NCHW call (``nchw_staging=(B, C, sum_hw, dtype_code)``), room for channels-last copies of the maps and
    of the pooled tensor.  ``layout`` (a code of ``roi_align_layout``) overrides ``channels_last``.
    Returns (tensor|None, bytes)."""
    if R == 0:
        return None, 0
    if layout is not None:
        channels_last = layout == ABR_NHWC
    if layout == ABR_NCHW_MAPS_NHWC_POOLED:
        B, C, sum_hw, code = nchw_staging
        n = int(lib().abr_roi_align_workspace_bytes_layout(R, PH, PW, max_h, B, C, sum_hw, code, layout))
    elif channels_last:
        n = int(lib().abr_roi_align_workspace_bytes(R, PH, PW, max_h))
    elif nchw_staging is not None and NCHW_STAGING:
        B, C, sum_hw, code = nchw_staging
        n = int(lib().abr_roi_align_workspace_bytes_nchw(R, PH, PW, max_h, B, C, sum_hw, code))
    else:
        return None, 0
    return torch.empty((n,), dtype=torch.uint8, device=device), n


def plan_only(ws, R, PH, PW, max_h):
    """What an autograd context should keep of a ROIAlign workspace: only the per-RoI plans at its start (a few MB), not the
    channels-last staging copies behind them (a B*C*H*W map copy plus an R*C*P*P pooled copy for a contiguous-NCHW call)."""
    if ws is None:
        return None
    n = int(lib().abr_roi_align_workspace_bytes(R, PH, PW, max_h))
    return ws if ws.numel() <= n else ws[:n].clone()


def workspace_with_plan(plan, R, PH, PW, max_h, device, layout, nchw_staging):
    """A full-size workspace for the backward of a call whose forward kept ``plan_only``: the staging room is allocated
    afresh and the plans are copied to its start.  Returns (tensor, bytes)."""
    ws, n = roi_align_workspace(R, PH, PW, max_h, device, layout=layout, nchw_staging=nchw_staging)
    if ws is None or n <= plan.numel():
        return plan, plan.numel()
    ws[: plan.numel()].copy_(plan)
    return ws, n


def set_option(key: str, value: int) -> None:
    """``abr_set_option``: kernel-family switches for measurements and tests ("roi_v2", "fwd_tma", ...)."""
    check(lib().abr_set_option(key.encode(), int(value)))


FUSED_STAGES = ("plan", "pool_teacher_student", "ard_coefficients", "backward")


def stage_timing_begin(max_calls: int, every: int = 1) -> None:
    """Record stage events for every ``every``-th ``abr_roi_ard_fused`` call from now on (at most ``max_calls`` calls)."""
    check(lib().abr_stage_timing_begin_every(int(max_calls), int(every)))


def stage_timing_end():
    """(calls recorded, {stage: average ms per call}) of the ``abr_roi_ard_fused`` calls since ``stage_timing_begin``."""
    buf = (ctypes.c_float * len(FUSED_STAGES))()
    n = int(lib().abr_stage_timing_end(buf, len(FUSED_STAGES)))
    if n < 0:
        raise RuntimeError("abr_b200 error: %s" % lib().abr_last_error().decode())
    return n, {k: float(buf[i]) for i, k in enumerate(FUSED_STAGES)}


def launch_count() -> int:
    return int(lib().abr_launch_count())
