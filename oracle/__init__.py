"""CPU oracle for the ABR_IOD RoI hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package, and only as the checker.  ``abr_iod_b200`` never imports it.

* :mod:`oracle` (this file) -- ctypes front end of ``abr_oracle.c`` (plain C restatement
  of the reference's ROIAlign / ROIPool / NMS / ARD / paste arithmetic) and of
  ``_ref/libabr_ref_cpu.so`` (the reference's own ``csrc/cpu/*.cpp`` compiled in place).
* :mod:`oracle.ard_torch` -- PyTorch restatement of ``distillation/distillation.py:86-130``.
* :mod:`oracle.paste` -- numpy restatement of ``data/datasets/voc_abr.py:512-858``.
* :mod:`oracle.pooler` -- restatement of ``modeling/poolers.py`` over the C oracle.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libabr_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libabr_ref_cpu.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f64p = ctypes.POINTER(ctypes.c_double)
_int = ctypes.c_int


def build(ref: bool = False) -> None:
    """Compile the C restatement (and, with ``ref=True`` and /root/reference present, the
    reference's CPU ops into ``oracle/_ref``)."""
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir(os.environ.get("ABR_REFERENCE", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "abr_oracle.c")
        ):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_roi_align_fwd.argtypes = [_f32p, _f32p, _f32p] + [_int] * 7 + [ctypes.c_float, _int]
        L.orc_roi_align_bwd.argtypes = [_f32p, _f32p, _f32p] + [_int] * 7 + [ctypes.c_float, _int]
        L.orc_roi_pool_fwd.argtypes = [_f32p, _f32p, _f32p, _i32p] + [_int] * 7 + [ctypes.c_float]
        L.orc_roi_pool_bwd.argtypes = [_f32p, _i32p, _f32p, _f32p] + [_int] * 7
        L.orc_nms.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_float, _int, _i64p]
        L.orc_nms.restype = ctypes.c_int64
        L.orc_ard.argtypes = [_f32p, _f32p, _int, _int, _int, ctypes.c_double, _f64p, _f32p]
        L.orc_paste_mixup.argtypes = [_u8p, _int, _int, _u8p, _int, _int] + [_int] * 6 + [ctypes.c_double]
        L.orc_paste_copy.argtypes = [_u8p, _int, _int, _u8p, _int, _int] + [_int] * 6
        _lib = L
    return _lib


_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref_lib() -> ctypes.CDLL:
    """The reference's compiled CPU ops (``oracle/_ref``); needs libtorch, so torch is imported first."""
    global _ref
    if _ref is None:
        import torch  # noqa: F401  (loads libtorch/libc10 so the shim's dependencies resolve)

        L = ctypes.CDLL(_REF_PATH)
        L.ref_roi_align_forward_cpu.argtypes = [_f32p, _f32p, _f32p] + [_int] * 7 + [ctypes.c_float, _int]
        L.ref_nms_cpu.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_float, _i64p]
        L.ref_nms_cpu.restype = ctypes.c_int64
        _ref = L
    return _ref


# ---------------------------------------------------------------- the reference's CUDA kernels (same-box comparator)
_REF_CUDA_PATH = os.path.join(_HERE, "_ref", "libabr_ref_cuda.so")
_ref_cuda = None


def ref_cuda_available() -> bool:
    return os.path.exists(_REF_CUDA_PATH)


def ref_cuda_lib() -> ctypes.CDLL:
    """``oracle/_ref/libabr_ref_cuda.so``: the reference's own csrc/cuda/{ROIAlign_cuda,ROIPool_cuda,nms}.cu compiled in
    place for sm_100a (oracle/ref_cuda_shim.cu).  All arguments are DEVICE pointers of contiguous fp32 NCHW tensors."""
    global _ref_cuda
    if _ref_cuda is None:
        import torch  # noqa: F401

        L = ctypes.CDLL(_REF_CUDA_PATH)
        vp, f = ctypes.c_void_p, ctypes.c_float
        L.ref_cuda_roi_align_forward.argtypes = [vp, vp, vp] + [_int] * 7 + [f, _int]
        L.ref_cuda_roi_align_backward.argtypes = [vp, vp, vp] + [_int] * 7 + [f, _int]
        L.ref_cuda_roi_align_forward_nocopy.argtypes = [vp, vp] + [_int] * 7 + [f, _int]
        L.ref_cuda_roi_align_backward_nocopy.argtypes = [vp, vp] + [_int] * 7 + [f, _int]
        L.ref_cuda_roi_pool_forward.argtypes = [vp, vp, vp, vp] + [_int] * 7 + [f]
        L.ref_cuda_roi_pool_backward.argtypes = [vp, vp, vp, vp, vp] + [_int] * 7 + [f]
        L.ref_cuda_nms.argtypes = [vp, ctypes.c_int64, f, vp]
        L.ref_cuda_nms.restype = ctypes.c_int64
        _ref_cuda = L
    return _ref_cuda


def ref_cuda_roi_align_forward(x, rois, scale, ph, pw, ratio):
    """The reference's ROIAlign_forward_cuda on torch CUDA tensors (contiguous NCHW fp32)."""
    import torch

    x, rois = x.contiguous().float(), rois.contiguous().float()
    B, C, H, W = x.shape
    out = torch.empty((rois.shape[0], C, ph, pw), device=x.device)
    ref_cuda_lib().ref_cuda_roi_align_forward(x.data_ptr(), rois.data_ptr(), out.data_ptr(), B, C, H, W, rois.shape[0], ph, pw,
                                             float(scale), int(ratio))
    return out


def ref_cuda_roi_align_backward(grad, rois, scale, ph, pw, B, C, H, W, ratio):
    import torch

    grad, rois = grad.contiguous().float(), rois.contiguous().float()
    gin = torch.empty((B, C, H, W), device=grad.device)
    ref_cuda_lib().ref_cuda_roi_align_backward(grad.data_ptr(), rois.data_ptr(), gin.data_ptr(), B, C, H, W, rois.shape[0], ph, pw,
                                              float(scale), int(ratio))
    return gin


def ref_cuda_roi_pool_forward(x, rois, scale, ph, pw):
    import torch

    x, rois = x.contiguous().float(), rois.contiguous().float()
    B, C, H, W = x.shape
    out = torch.empty((rois.shape[0], C, ph, pw), device=x.device)
    arg = torch.empty((rois.shape[0], C, ph, pw), device=x.device, dtype=torch.int32)
    ref_cuda_lib().ref_cuda_roi_pool_forward(x.data_ptr(), rois.data_ptr(), out.data_ptr(), arg.data_ptr(), B, C, H, W,
                                            rois.shape[0], ph, pw, float(scale))
    return out, arg


def ref_cuda_roi_pool_backward(grad, x, rois, argmax, scale, ph, pw):
    import torch

    grad, x, rois = grad.contiguous().float(), x.contiguous().float(), rois.contiguous().float()
    B, C, H, W = x.shape
    gin = torch.empty((B, C, H, W), device=x.device)
    ref_cuda_lib().ref_cuda_roi_pool_backward(grad.data_ptr(), x.data_ptr(), rois.data_ptr(), argmax.contiguous().data_ptr(),
                                             gin.data_ptr(), B, C, H, W, rois.shape[0], ph, pw, float(scale))
    return gin


def ref_cuda_nms(boxes, scores, thr):
    """The reference's nms_cuda ('>' rule) on torch CUDA tensors; returns ascending original indices (device, int64)."""
    import torch

    b = torch.cat((boxes.float(), scores.float().unsqueeze(1)), 1).contiguous()  # csrc/nms.h:19-20
    n = b.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=b.device)
    k = ref_cuda_lib().ref_cuda_nms(b.data_ptr(), n, float(thr), keep.data_ptr())
    return keep[:k]


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- ROIAlign
def roi_align_forward(inp, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, use_ref=False):
    """``_C.roi_align_forward`` (csrc/ROIAlign.h:11-25) on NCHW fp32 numpy arrays."""
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    B, C, H, W = inp.shape
    R = rois.shape[0]
    out = np.empty((R, C, pooled_h, pooled_w), np.float32)
    if out.size == 0:
        return out
    fn = ref_lib().ref_roi_align_forward_cpu if use_ref else lib().orc_roi_align_fwd
    fn(_p(inp, _f32p), _p(rois, _f32p), _p(out, _f32p), B, C, H, W, R, pooled_h, pooled_w,
       float(spatial_scale), int(sampling_ratio))
    return out


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, B, C, H, W, sampling_ratio):
    """``_C.roi_align_backward`` (csrc/ROIAlign.h:27-46)."""
    grad, rois = _f32(grad), _f32(rois).reshape(-1, 5)
    R = rois.shape[0]
    gin = np.zeros((B, C, H, W), np.float32)
    if grad.size == 0:
        return gin
    lib().orc_roi_align_bwd(_p(grad, _f32p), _p(rois, _f32p), _p(gin, _f32p), B, C, H, W, R, pooled_h, pooled_w,
                            float(spatial_scale), int(sampling_ratio))
    return gin


# ---------------------------------------------------------------- ROIPool
def roi_pool_forward(inp, rois, spatial_scale, pooled_h, pooled_w):
    """``_C.roi_pool_forward`` (csrc/ROIPool.h:9-25): returns (output, int32 argmax)."""
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    B, C, H, W = inp.shape
    R = rois.shape[0]
    out = np.empty((R, C, pooled_h, pooled_w), np.float32)
    arg = np.zeros((R, C, pooled_h, pooled_w), np.int32)
    if out.size:
        lib().orc_roi_pool_fwd(_p(inp, _f32p), _p(rois, _f32p), _p(out, _f32p), _p(arg, _i32p), B, C, H, W, R,
                               pooled_h, pooled_w, float(spatial_scale))
    return out, arg


def roi_pool_backward(grad, argmax, rois, B, C, H, W):
    """``_C.roi_pool_backward`` (csrc/ROIPool.h:27-48)."""
    grad, rois = _f32(grad), _f32(rois).reshape(-1, 5)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    R, _, PH, PW = grad.shape
    gin = np.zeros((B, C, H, W), np.float32)
    if grad.size:
        lib().orc_roi_pool_bwd(_p(grad, _f32p), _p(argmax, _i32p), _p(rois, _f32p), _p(gin, _f32p), B, C, H, W, R,
                               PH, PW)
    return gin


# ---------------------------------------------------------------- NMS
def nms(boxes, scores, thresh, flavour="cuda", use_ref=False):
    """``_C.nms`` (csrc/nms.h:10-28).  ``flavour='cuda'`` suppresses on IoU > thr
    (csrc/cuda/nms.cu:60), ``'cpu'`` on IoU >= thr (csrc/cpu/nms_cpu.cpp:60).
    Returns int64 original indices, ascending."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), np.int64)
    if n == 0:
        return keep[:0]
    if use_ref:
        assert flavour == "cpu"
        k = ref_lib().ref_nms_cpu(_p(boxes, _f32p), _p(scores, _f32p), n, float(thresh), _p(keep, _i64p))
    else:
        k = lib().orc_nms(_p(boxes, _f32p), _p(scores, _f32p), n, float(thresh), int(flavour == "cpu"),
                          _p(keep, _i64p))
    return keep[:k].copy()


# ---------------------------------------------------------------- ARD
def ard(f_old, f_new, gamma=1.0, want_grad=True):
    """ARD loss of distillation/distillation.py:86-100 with (teacher, student) argument order.
    Returns (loss, loss_afd, loss_pad, dF_new or None); double accumulation."""
    f_old, f_new = _f32(f_old), _f32(f_new)
    N, C = f_old.shape[:2]
    HW = int(np.prod(f_old.shape[2:]))
    loss3 = np.zeros(3, np.float64)
    g = np.empty_like(f_new) if want_grad else None
    lib().orc_ard(_p(f_old, _f32p), _p(f_new, _f32p), N, C, HW, float(gamma), _p(loss3, _f64p),
                  _p(g, _f32p) if want_grad else None)
    return float(loss3[0]), float(loss3[1]), float(loss3[2]), g


# ---------------------------------------------------------------- paste pixels
def paste_mixup(img, src, y0, x0, y1, x1, sy0, sx0, lam):
    """In-place ``img[y0:y1,x0:x1] = lam*img[...] + (1-lam)*src[sy0:, sx0:]`` with numpy's
    float64->uint8 truncation (voc_abr.py:659-678)."""
    assert img.dtype == np.uint8 and src.dtype == np.uint8 and img.flags.c_contiguous and src.flags.c_contiguous
    lib().orc_paste_mixup(_p(img, _u8p), img.shape[0], img.shape[1], _p(src, _u8p), src.shape[0], src.shape[1],
                          y0, x0, y1, x1, sy0, sx0, float(lam))


def paste_copy(img, src, y0, x0, y1, x1, sy0, sx0):
    """In-place ``img[y0:y1,x0:x1] = src[sy0:sy0+.., sx0:sx0+..]`` (voc_abr.py:763)."""
    assert img.dtype == np.uint8 and src.dtype == np.uint8 and img.flags.c_contiguous and src.flags.c_contiguous
    lib().orc_paste_copy(_p(img, _u8p), img.shape[0], img.shape[1], _p(src, _u8p), src.shape[0], src.shape[1],
                         y0, x0, y1, x1, sy0, sx0)
