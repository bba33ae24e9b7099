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
        ctx.wanted_grad = want_grad
        ctx.grad = grad  # dL/df_second for an upstream gradient of 1
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss, _grad_parts):
        if ctx.first_needs_grad:
            raise RuntimeError("ARD: gradient w.r.t. the first argument (old model features) is not implemented; "
                               "the reference computes them under torch.no_grad() (train_incremental.py:83-85)")
        if not ctx.wanted_grad:
            return None, None, None
        g = ctx.grad
        if g is None:
            # the gradient buffer (as large as the pooled tensor) is scaled in place and handed to autograd once
            raise RuntimeError("ARD: backward called a second time over the same graph; the fused kernel's gradient buffer "
                               "has been released (call the loss again instead of retain_graph=True)")
        scale = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().abr_scale_if_needed(g.data_ptr(), g.numel(), scale.data_ptr(), 1.0,
                                                      _lib.dtype_code(g), _lib.stream_ptr(g.device)))
        ctx.grad = None
        return None, g, None


# --------------------------------------------------------------------------------------------------------------------
# Fused ARD step: teacher pooling + student pooling + ARD loss + its backward through the student's ROIAlign
class _PooledAttentiveRoIDistillation(Function):
    """``abr_roi_ard_fused``.  Outputs: (teacher RoI features, student RoI features, loss, [loss, afd, pad]).  The
    gradient of the loss w.r.t. the student's feature map is computed by the same call (for an upstream gradient of 1)
    and scaled in ``backward``; a gradient arriving at the student's RoI features (the box head's) takes the ordinary
    ROIAlign backward and is added."""

    @staticmethod
    def forward(ctx, teacher_map, student_map, rois, output_size, spatial_scale, sampling_ratio, gamma):
        from ..layers.roi_align import _prep_rois

        ph, pw = output_size
        t = teacher_map.detach()
        s = student_map.detach()
        rois = _prep_rois(rois, s.device)
        B, C, H, W = s.shape
        R = rois.size(0)
        dev = s.device
        f_old = torch.empty((R, C, ph, pw), dtype=s.dtype, device=dev, memory_format=torch.channels_last)
        f_new = torch.empty_like(f_old, memory_format=torch.channels_last)
        loss3 = torch.empty((3,), dtype=torch.float32, device=dev)
        want_grad = student_map.requires_grad
        gmap = torch.empty_like(s, memory_format=torch.channels_last) if want_grad else None
        L = _lib.lib()
        ws_bytes = int(L.abr_roi_ard_fused_workspace_bytes(R, C, ph, pw))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.abr_roi_ard_fused(
                t.data_ptr(), s.data_ptr(), rois.data_ptr(), f_old.data_ptr(), f_new.data_ptr(),
                gmap.data_ptr() if want_grad else None, loss3.data_ptr(), B, C, H, W, R, ph, pw, float(spatial_scale),
                int(sampling_ratio), float(gamma), 1.0, _lib.ABR_F32, _lib.ABR_NHWC, 1, ws.data_ptr(), ws_bytes, 0,
                _lib.stream_ptr(dev)))
        ctx.gmap = gmap
        ctx.teacher_needs_grad = teacher_map.requires_grad
        ctx.geometry = (rois, ph, pw, float(spatial_scale), int(sampling_ratio), (B, C, H, W))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(f_old, loss3)
        return f_old, f_new, loss3[0].clone(), loss3

    @staticmethod
    @once_differentiable
    def backward(ctx, _g_old, g_new, g_loss, _g_parts):
        from ..layers.roi_align import roi_align_backward

        if ctx.teacher_needs_grad:
            raise RuntimeError("fused ARD: gradient w.r.t. the old model's feature map is not implemented; the reference "
                               "computes the teacher under torch.no_grad() (train_incremental.py:83-85)")
        if ctx.gmap is None:
            return (None,) * 7
        rois, ph, pw, scale, ratio, (B, C, H, W) = ctx.geometry
        grad = None
        if g_loss is not None:
            grad = ctx.gmap * g_loss.to(ctx.gmap.dtype)  # out of place: a second backward over the graph stays correct
        if g_new is not None:
            extra = roi_align_backward(g_new, rois, scale, ph, pw, B, C, H, W, ratio, layout=_lib.ABR_NHWC)
            grad = extra if grad is None else grad + extra
        return None, grad, None, None, None, None, None


def pooled_attentive_roi_distillation(teacher_map, student_map, rois, output_size, spatial_scale, sampling_ratio, gamma=1.0):
    """The RoI part of the distillation step of tools/train_incremental.py:84-115 in one call: ROIAlign of the old model's
    and of the student's feature map over the SAME ``rois`` ([R,5]), the ARD loss between the two pooled tensors and --
    through autograd -- its backward into the student's map.  Returns ``(teacher_roi_features, student_roi_features,
    loss)``; both feature tensors are ``[R,C,PH,PW]`` in channels-last storage and can feed the box head as usual.
    fp32 runs the fused kernels (``abr_roi_ard_fused``); other dtypes take the separate ops."""
    from torch.nn.modules.utils import _pair

    from ..layers.roi_align import roi_align

    _lib.require_cuda(teacher_map, "teacher feature map")
    _lib.require_cuda(student_map, "student feature map")
    output_size = _pair(output_size)
    if teacher_map.shape != student_map.shape:
        raise RuntimeError("fused ARD: the two feature maps must have the same shape, got %s and %s"
                           % (tuple(teacher_map.shape), tuple(student_map.shape)))
    if (student_map.dtype != torch.float32 or teacher_map.dtype != torch.float32 or max(output_size) > 16
            or rois.size(0) == 0):
        f_old = roi_align(teacher_map.detach(), rois, output_size, spatial_scale, sampling_ratio)
        f_new = roi_align(student_map, rois, output_size, spatial_scale, sampling_ratio)
        return f_old, f_new, calculate_attentive_roi_feature_distillation(f_old, f_new, gamma)
    t = teacher_map.contiguous(memory_format=torch.channels_last)
    s = student_map.contiguous(memory_format=torch.channels_last)
    f_old, f_new, loss, _ = _PooledAttentiveRoIDistillation.apply(t, s, rois, output_size, spatial_scale, sampling_ratio, gamma)
    return f_old, f_new, loss


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
        # out of place ([R, C]-sized tensors): a second backward over the same graph stays correct
        gs, gb = ctx.grads
        if gs is not None:
            gs = (gs * grad_loss.to(gs.dtype)).to(ctx.dtypes[0])
        if gb is not None:
            gb = (gb * grad_loss.to(gb.dtype)).to(ctx.dtypes[1])
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


# --------------------------------------------------------------------------------------------------------------------
# The reference's helper names (distillation/distillation.py:103-130), kept so that code importing them keeps working.
# They return the same quantities, computed by the fused kernel (loss terms) or one reduction kernel (the attention map).
def activation_at(f_map, temp=2):
    """distillation.py:121-130: ``H*W*softmax_{hw}(mean_c |f|^temp)`` as ``[N,H,W]``; only ``temp == 2`` is on the path."""
    if temp != 2:
        raise NotImplementedError("activation_at: only temp=2 (the value calculate_attentive_roi_feature_distillation uses)")
    _lib.require_cuda(f_map, "f_map")
    f = _lib.as_compute_dtype(f_map).detach()
    nhwc = _lib.is_channels_last(f)
    f = f.contiguous(memory_format=torch.channels_last if nhwc else torch.contiguous_format)
    N, C, H, W = f.shape
    sq = _channel_mean_of_squares(f, nhwc)
    return (H * W * torch.softmax(sq, dim=1)).view(N, H, W)


def _channel_mean_of_squares(f, nhwc):
    """mean_c f^2 per position through abr_channel_mean on the squared tensor."""
    N, C, H, W = f.shape
    sq_in = (f.float() * f.float()).contiguous(memory_format=torch.channels_last if nhwc else torch.contiguous_format)
    out = torch.empty((N, H * W), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(_lib.lib().abr_channel_mean(sq_in.data_ptr(), N, C, H * W, _lib.ABR_F32,
                                               _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW, out.data_ptr(), _lib.stream_ptr(f.device)))
    return out


def afd_loss(f_map_s, f_map_t, S_t=None):
    """distillation.py:103-111 at the call site's argument roles: mean of ``A_first * (f_first - f_second)^2``.  ``S_t`` is
    accepted for signature compatibility; the fused kernel recomputes the attention of the first argument."""
    return attentive_roi_distillation_terms(f_map_s, f_map_t, 1.0)[1]


def pad_loss(S_s, S_t):
    """distillation.py:114-118: mean absolute difference of two attention maps (plain reduction)."""
    return (S_s - S_t).abs().mean()
