"""Attentive RoI Distillation with the reference's entry point
(distillation/distillation.py:86-130: ``calculate_attentive_roi_feature_distillation``), as ONE fused
forward+backward kernel (``abr_ard_forward_backward`` of libabr_b200).

Argument roles follow the CALL SITE, tools/train_incremental.py:115:
``calculate_attentive_roi_feature_distillation(roi_align_features_source, roi_align_features_target, gamma)``
-- the first argument is the old model's (teacher's) pooled features, the second the student's; the attention of the
FIRST argument weights the feature term.  Gradients flow to the second argument only (the reference computes the
teacher under ``torch.no_grad()``, train_incremental.py:83-85); asking for a gradient w.r.t. the first raises.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib


def _ard_launch(f_first, f_second, gamma, want_grad, grad_scale=1.0):
    _lib.require_cuda(f_first, "f_map_s")
    _lib.require_cuda(f_second, "f_map_t")
    if f_first.shape != f_second.shape or f_first.dim() != 4:
        raise RuntimeError("ARD expects two [N,C,H,W] tensors of the same shape, got %s and %s"
                           % (tuple(f_first.shape), tuple(f_second.shape)))
    if f_first.dtype != f_second.dtype:
        f_first = f_first.to(f_second.dtype)
    nhwc = _lib.is_channels_last(f_second)
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    fo = f_first.detach().contiguous(memory_format=fmt)
    fn = f_second.detach().contiguous(memory_format=fmt)
    N, C, H, W = fn.shape
    loss3 = torch.empty((3,), dtype=torch.float32, device=fn.device)
    grad = torch.empty_like(fn, memory_format=fmt) if want_grad else None
    L = _lib.lib()
    ws_bytes = int(L.abr_ard_workspace_bytes(N, C, H * W))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=fn.device)
    with torch.cuda.device(fn.device):
        _lib.check(L.abr_ard_forward_backward(
            fo.data_ptr(), fn.data_ptr(), grad.data_ptr() if want_grad else None, loss3.data_ptr(), N, C, H * W,
            float(gamma), float(grad_scale), _lib.dtype_code(fn), _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW,
            ws.data_ptr(), ws_bytes, _lib.stream_ptr(fn.device)))
    return loss3, grad


class _AttentiveRoIDistillation(Function):
    @staticmethod
    def forward(ctx, f_first, f_second, gamma):
        want_grad = f_second.requires_grad
        loss3, grad = _ard_launch(f_first, f_second, gamma, want_grad)
        ctx.first_needs_grad = f_first.requires_grad
        ctx.grad = grad  # dL/df_second for an upstream gradient of 1
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss, _grad_parts):
        if ctx.first_needs_grad:
            raise RuntimeError("ARD: gradient w.r.t. the first argument (old model features) is not implemented; "
                               "the reference computes them under torch.no_grad() (train_incremental.py:83-85)")
        g = ctx.grad
        if g is None:
            return None, None, None
        scale = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().abr_scale_if_needed(g.data_ptr(), g.numel(), scale.data_ptr(), 1.0,
                                                      _lib.dtype_code(g), _lib.stream_ptr(g.device)))
        ctx.grad = None
        return None, g, None


def attentive_roi_distillation_terms(f_map_s, f_map_t, gamma=1.0):
    """Returns the device tensor ``[loss, loss_afd, loss_pad]`` (no autograd)."""
    loss3, _ = _ard_launch(_lib.as_compute_dtype(f_map_s), _lib.as_compute_dtype(f_map_t), gamma, False)
    return loss3


def calculate_attentive_roi_feature_distillation(f_map_s, f_map_t, gamma=1.0):
    """Drop-in for distillation/distillation.py:86-100.  Returns the scalar ``loss_afd + gamma * loss_pad``."""
    loss, _ = _AttentiveRoIDistillation.apply(_lib.as_compute_dtype(f_map_s), _lib.as_compute_dtype(f_map_t), gamma)
    return loss


# --------------------------------------------------------------------------------------------------------------------
# Inclusive distillation of the box head's outputs (distillation/distillation.py:164-241 of the reference)
def _scale_in_place(t, upstream):
    scale = upstream.detach().to(torch.float32).reshape(1).contiguous()
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().abr_scale_if_needed(t.data_ptr(), t.numel(), scale.data_ptr(), 1.0, _lib.dtype_code(t),
                                                  _lib.stream_ptr(t.device)))
    return t


class _RoIDistillationID(Function):
    @staticmethod
    def forward(ctx, soften_scores, soften_bboxes, target_scores, target_bboxes):
        _lib.require_cuda(target_scores, "target_scores")
        if soften_scores.requires_grad or soften_bboxes.requires_grad:
            raise RuntimeError("roi distillation: the soften (old model) results carry no gradient in the reference "
                               "(train_incremental.py:83-85); detach them")
        R, Co = soften_scores.shape
        Ct = target_scores.shape[1]
        if target_scores.shape[0] != R or tuple(soften_bboxes.shape) != (R, Co, 4) or tuple(target_bboxes.shape) != (R, Ct, 4):
            raise RuntimeError("roi distillation: scores [R,C] and bboxes [R,C,4] of both models must agree on R, got %s %s %s %s"
                               % (tuple(soften_scores.shape), tuple(soften_bboxes.shape), tuple(target_scores.shape), tuple(target_bboxes.shape)))
        f32 = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731
        ss, sb, ts, tb = f32(soften_scores), f32(soften_bboxes), f32(target_scores), f32(target_bboxes)
        dev = ts.device
        loss3 = torch.empty((3,), dtype=torch.float32, device=dev)
        gs = torch.empty_like(ts) if target_scores.requires_grad else None
        gb = torch.empty_like(tb) if target_bboxes.requires_grad else None
        L = _lib.lib()
        ws_bytes = int(L.abr_logit_loss_workspace_bytes(R))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.abr_roi_distillation_id(ss.data_ptr(), sb.data_ptr(), ts.data_ptr(), tb.data_ptr(), R, Co, Ct, 1.0,
                                                 gs.data_ptr() if gs is not None else None,
                                                 gb.data_ptr() if gb is not None else None, loss3.data_ptr(), ws.data_ptr(),
                                                 ws_bytes, _lib.stream_ptr(dev)))
        ctx.grads = (gs, gb)
        ctx.dtypes = (target_scores.dtype, target_bboxes.dtype)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss, _grad_parts):
        gs, gb = ctx.grads
        ctx.grads = (None, None)
        if gs is not None:
            gs = _scale_in_place(gs, grad_loss).to(ctx.dtypes[0])
        if gb is not None:
            gb = _scale_in_place(gb, grad_loss).to(ctx.dtypes[1])
        return None, None, gs, gb


def roi_distillation_id_terms(soften_results, target_results):
    """Device tensor ``[class term + box term, class term, box term]`` of the 'id' distillation, with autograd into the
    target (student) results through the first element."""
    soften_scores, soften_bboxes = soften_results
    target_scores, target_bboxes = target_results
    return _RoIDistillationID.apply(soften_scores, soften_bboxes, target_scores, target_bboxes)


def calculate_roi_distillation_losses(soften_results, target_results, dist="l2", soften_proposal=None):
    """Drop-in for distillation/distillation.py:220-241 with ``dist='id'`` (inclusive distillation: unbiased cross-entropy
    + L2 box term), the setting the ABR runs use (scripts/run_SI.sh).  One fused forward+backward kernel.
    The legacy ``dist='l2'`` branch (mean-normalised logits, Faster-ILOD) is outside this path."""
    if dist != "id":
        raise NotImplementedError("calculate_roi_distillation_losses: only dist='id' is part of the accelerated path")
    loss, _ = roi_distillation_id_terms(soften_results, target_results)
    return loss
