"""``ROIAlign`` / ``roi_align`` with the reference's signatures (maskrcnn_benchmark/layers/roi_align.py:12-70),
backed by ``abr_roi_align_forward/backward`` of libabr_b200 (sm_100a).

Layout: a channels-last input (``x.contiguous(memory_format=torch.channels_last)``) takes the vectorised NHWC
kernels and yields a channels-last output; anything else is treated as the reference's contiguous NCHW.
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _lib


def _prep_rois(rois, device):
    if rois.dim() != 2 or rois.size(1) != 5:
        raise RuntimeError("rois must be [R,5] (batch_index, x1, y1, x2, y2), got %s" % (tuple(rois.shape),))
    _lib.require_cuda(rois, "rois")
    return rois.detach().to(dtype=torch.float32).contiguous()


def roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, return_plan=False, plan=None):
    """``_C.roi_align_forward`` (csrc/ROIAlign.h:11-25).  With ``return_plan`` also returns the workspace holding the
    per-RoI plans, which ``roi_align_backward(..., plan=...)`` -- or another forward over the SAME RoIs, output size,
    sampling ratio, scale and map shape (the teacher / student pair of the distillation step), via ``plan=`` -- can reuse.  The result is channels-last
    for a channels-last ``input`` (and for a contiguous one when ``_lib.POOLED_CHANNELS_LAST`` is set)."""
    _lib.require_cuda(input, "input")
    rois = _prep_rois(rois, input.device)
    layout = _lib.roi_align_layout(input)
    x = input if layout == _lib.ABR_NHWC else input.contiguous()
    B, C, H, W = x.shape
    R = rois.size(0)
    out = torch.empty((R, C, pooled_h, pooled_w), dtype=x.dtype, device=x.device,
                      memory_format=torch.contiguous_format if layout == _lib.ABR_NCHW else torch.channels_last)
    if out.numel() == 0:
        return (out, None) if return_plan else out
    with torch.cuda.device(x.device):
        if plan is not None:
            ws, ws_bytes = plan, plan.numel()
        else:
            ws, ws_bytes = _lib.roi_align_workspace(R, pooled_h, pooled_w, H, x.device, layout=layout,
                                                    nchw_staging=(B, C, H * W, _lib.dtype_code(x)))
        _lib.check(_lib.lib().abr_roi_align_forward(
            x.data_ptr(), rois.data_ptr(), out.data_ptr(), B, C, H, W, R, pooled_h, pooled_w,
            float(spatial_scale), int(sampling_ratio), _lib.dtype_code(x), layout,
            ws.data_ptr() if ws is not None else None, ws_bytes, int(plan is not None), _lib.stream_ptr(x.device)))
    return (out, ws) if return_plan else out


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, batch_size, channels, height, width,
                       sampling_ratio, channels_last=None, plan=None, layout=None):
    """``_C.roi_align_backward`` (csrc/ROIAlign.h:27-46).  ``channels_last=None`` follows ``grad``'s layout; ``layout``
    (a code of ``_lib.roi_align_layout``) overrides it; ``plan`` is the workspace a forward over the same RoIs returned
    (skips re-planning)."""
    _lib.require_cuda(grad, "grad")
    rois = _prep_rois(rois, grad.device)
    if layout is None:
        nhwc = _lib.is_channels_last(grad) if channels_last is None else bool(channels_last)
        layout = _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW
    pooled_fmt = torch.contiguous_format if layout == _lib.ABR_NCHW else torch.channels_last
    map_fmt = torch.channels_last if layout == _lib.ABR_NHWC else torch.contiguous_format
    g = grad.contiguous(memory_format=pooled_fmt)
    gin = torch.empty((batch_size, channels, height, width), dtype=g.dtype, device=g.device, memory_format=map_fmt)
    if gin.numel() == 0:
        return gin
    with torch.cuda.device(g.device):
        has_plan = int(plan is not None)
        if has_plan:
            # (a contiguous-NCHW call needs its staging room again: the context only kept the plans)
            ws, ws_bytes = _lib.workspace_with_plan(plan, rois.size(0), pooled_h, pooled_w, height, g.device, layout,
                                                    (batch_size, channels, height * width, _lib.dtype_code(g)))
        else:
            ws, ws_bytes = _lib.roi_align_workspace(rois.size(0), pooled_h, pooled_w, height, g.device, layout=layout,
                                                    nchw_staging=(batch_size, channels, height * width, _lib.dtype_code(g)))
        _lib.check(_lib.lib().abr_roi_align_backward(
            g.data_ptr(), rois.data_ptr(), gin.data_ptr(), batch_size, channels, height, width, rois.size(0),
            pooled_h, pooled_w, float(spatial_scale), int(sampling_ratio), _lib.dtype_code(g), layout, 1,
            ws.data_ptr() if ws is not None else None, ws_bytes, has_plan, _lib.stream_ptr(g.device)))
    return gin


class _ROIAlign(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        ctx.save_for_backward(roi)
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.sampling_ratio = sampling_ratio
        ctx.input_shape = input.size()
        ctx.layout = _lib.roi_align_layout(input)
        out, ws = roi_align_forward(input, roi, spatial_scale, ctx.output_size[0], ctx.output_size[1], sampling_ratio,
                                    return_plan=True)
        ctx.plan = _lib.plan_only(ws, roi.size(0), ctx.output_size[0], ctx.output_size[1], input.size(2))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        bs, ch, h, w = ctx.input_shape
        grad_input = roi_align_backward(grad_output, rois, ctx.spatial_scale, ctx.output_size[0], ctx.output_size[1],
                                        bs, ch, h, w, ctx.sampling_ratio, layout=ctx.layout, plan=ctx.plan)
        return grad_input, None, None, None, None


def roi_align(input, rois, output_size, spatial_scale, sampling_ratio):
    return _ROIAlign.apply(_lib.as_compute_dtype(input), rois, output_size, spatial_scale, sampling_ratio)


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super(ROIAlign, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio)
