"""``ROIPool`` / ``roi_pool`` with the reference's signatures (maskrcnn_benchmark/layers/roi_pool.py:12-65),
backed by ``abr_roi_pool_forward/backward`` of libabr_b200."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _lib
from .roi_align import _prep_rois


def roi_pool_forward(input, rois, spatial_scale, pooled_h, pooled_w):
    """``_C.roi_pool_forward`` (csrc/ROIPool.h:9-25): returns (output, int32 argmax)."""
    _lib.require_cuda(input, "input")
    rois = _prep_rois(rois, input.device)
    nhwc = _lib.is_channels_last(input)
    x = input if nhwc else input.contiguous()
    B, C, H, W = x.shape
    R = rois.size(0)
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    out = torch.empty((R, C, pooled_h, pooled_w), dtype=x.dtype, device=x.device, memory_format=fmt)
    argmax = torch.empty((R, C, pooled_h, pooled_w), dtype=torch.int32, device=x.device, memory_format=fmt)
    if out.numel() == 0:
        return out, argmax
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().abr_roi_pool_forward(
            x.data_ptr(), rois.data_ptr(), out.data_ptr(), argmax.data_ptr(), B, C, H, W, R, pooled_h, pooled_w,
            float(spatial_scale), _lib.dtype_code(x), _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW,
            _lib.stream_ptr(x.device)))
    return out, argmax


def roi_pool_backward(grad, input, rois, argmax, spatial_scale, pooled_h, pooled_w, batch_size, channels, height, width):
    """``_C.roi_pool_backward`` (csrc/ROIPool.h:27-48); ``input`` and ``spatial_scale`` are unused, as in the reference."""
    _lib.require_cuda(grad, "grad")
    rois = _prep_rois(rois, grad.device)
    nhwc = _lib.is_channels_last(argmax)
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    g = grad.contiguous(memory_format=fmt)
    gin = torch.empty((batch_size, channels, height, width), dtype=g.dtype, device=g.device, memory_format=fmt)
    if gin.numel() == 0:
        return gin
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().abr_roi_pool_backward(
            g.data_ptr(), argmax.data_ptr(), rois.data_ptr(), gin.data_ptr(), batch_size, channels, height, width,
            rois.size(0), pooled_h, pooled_w, _lib.dtype_code(g), _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW, 1,
            _lib.stream_ptr(g.device)))
    return gin


class _ROIPool(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale):
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.input_shape = input.size()
        output, argmax = roi_pool_forward(input, roi, spatial_scale, ctx.output_size[0], ctx.output_size[1])
        ctx.save_for_backward(roi, argmax)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, argmax = ctx.saved_tensors
        bs, ch, h, w = ctx.input_shape
        grad_input = roi_pool_backward(grad_output, None, rois, argmax, ctx.spatial_scale, ctx.output_size[0],
                                       ctx.output_size[1], bs, ch, h, w)
        return grad_input, None, None, None


def roi_pool(input, rois, output_size, spatial_scale):
    return _ROIPool.apply(_lib.as_compute_dtype(input), rois, output_size, spatial_scale)


class ROIPool(nn.Module):
    def __init__(self, output_size, spatial_scale):
        super(ROIPool, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois):
        return roi_pool(input, rois, self.output_size, self.spatial_scale)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s)" % (self.__class__.__name__, self.output_size, self.spatial_scale)
