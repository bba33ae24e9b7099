"""``LevelMapper`` / ``Pooler`` / ``make_pooler`` with the reference's interface (modeling/poolers.py:11-117).

The single-level path is one ROIAlign launch.  The multi-level (FPN) path replaces the reference's per-level
``nonzero`` (host sync) + gather + launch + indexed scatter loop (:93-105) by two launches and no sync:
``abr_fpn_map_levels`` computes the level of every RoI on the device and ``abr_roi_align_multilevel_forward`` pools
every RoI from its own level straight into its output row.
"""
import ctypes

import torch
from torch import nn
from torch.nn.modules.utils import _pair
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib
from ..layers import ROIAlign
from ..layers.roi_align import _prep_rois


def cat(tensors, dim=0):
    """modeling/utils.py:9-16"""
    assert isinstance(tensors, (list, tuple))
    return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)


def _fpn_levels(rois, k_min, k_max, s0, lvl0, eps):
    levels = torch.empty((rois.size(0),), dtype=torch.int32, device=rois.device)
    if levels.numel():
        with torch.cuda.device(rois.device):
            _lib.check(_lib.lib().abr_fpn_map_levels(rois.data_ptr(), levels.data_ptr(), rois.size(0), float(k_min),
                                                     float(k_max), float(s0), float(lvl0), float(eps),
                                                     _lib.stream_ptr(rois.device)))
    return levels


class LevelMapper(object):
    """Eqn.(1) of the FPN paper (modeling/poolers.py:11-42).  Returns int64 levels relative to ``k_min``."""

    def __init__(self, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
        self.k_min = k_min
        self.k_max = k_max
        self.s0 = canonical_scale
        self.lvl0 = canonical_level
        self.eps = eps

    def levels_of_rois(self, rois):
        return _fpn_levels(rois, self.k_min, self.k_max, self.s0, self.lvl0, self.eps)

    def __call__(self, boxlists):
        boxes = cat([b.convert("xyxy").bbox for b in boxlists])
        _lib.require_cuda(boxes, "boxes")
        rois = torch.cat([boxes.new_zeros((boxes.size(0), 1)), boxes.float()], dim=1).contiguous()
        return self.levels_of_rois(rois).to(torch.int64)


def _level_arrays(feats, scales):
    L = len(feats)
    ptrs = (ctypes.c_void_p * L)(*[f.data_ptr() for f in feats])
    hs = (ctypes.c_int * L)(*[int(f.shape[2]) for f in feats])
    ws = (ctypes.c_int * L)(*[int(f.shape[3]) for f in feats])
    sc = (ctypes.c_float * L)(*[float(s) for s in scales])
    return ptrs, hs, ws, sc


class _MultiLevelROIAlign(Function):
    @staticmethod
    def forward(ctx, rois, levels, output_size, scales, sampling_ratio, *feats):
        layout = _lib.roi_align_layout(feats[0])
        fmt = torch.channels_last if layout == _lib.ABR_NHWC else torch.contiguous_format
        feats = [f.contiguous(memory_format=fmt) for f in feats]
        B, C = feats[0].shape[:2]
        R = rois.size(0)
        PH, PW = output_size
        out = torch.empty((R, C, PH, PW), dtype=feats[0].dtype, device=feats[0].device,
                          memory_format=torch.contiguous_format if layout == _lib.ABR_NCHW else torch.channels_last)
        ctx.save_for_backward(rois, levels)
        ctx.meta = (output_size, tuple(scales), sampling_ratio, [tuple(f.shape) for f in feats], layout)
        ctx.plan = None
        if out.numel():
            ptrs, hs, ws, sc = _level_arrays(feats, scales)
            with torch.cuda.device(out.device):
                wk, wk_bytes = _lib.roi_align_workspace(
                    R, PH, PW, max(f.shape[2] for f in feats), out.device, layout=layout,
                    nchw_staging=(B, C, sum(f.shape[2] * f.shape[3] for f in feats), _lib.dtype_code(out)))
                _lib.check(_lib.lib().abr_roi_align_multilevel_forward(
                    ptrs, hs, ws, sc, len(feats), rois.data_ptr(), levels.data_ptr(), out.data_ptr(), B, C, R, PH, PW,
                    int(sampling_ratio), _lib.dtype_code(out), layout,
                    wk.data_ptr() if wk is not None else None, wk_bytes, 0, _lib.stream_ptr(out.device)))
                ctx.plan = _lib.plan_only(wk, R, PH, PW, max(f.shape[2] for f in feats))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, levels = ctx.saved_tensors
        (PH, PW), scales, sampling_ratio, shapes, layout = ctx.meta
        fmt = torch.channels_last if layout == _lib.ABR_NHWC else torch.contiguous_format
        g = grad_output.contiguous(memory_format=torch.contiguous_format if layout == _lib.ABR_NCHW else torch.channels_last)
        grads = [torch.empty(s, dtype=g.dtype, device=g.device, memory_format=fmt) for s in shapes]
        B, C = shapes[0][:2]
        ptrs, hs, ws, sc = _level_arrays(grads, scales)
        with torch.cuda.device(g.device):
            has_plan = int(ctx.plan is not None)
            if has_plan:
                wk, wk_bytes = _lib.workspace_with_plan(
                    ctx.plan, rois.size(0), PH, PW, max(s[2] for s in shapes), g.device, layout,
                    (B, C, sum(s[2] * s[3] for s in shapes), _lib.dtype_code(g)))
            else:
                wk, wk_bytes = _lib.roi_align_workspace(
                    rois.size(0), PH, PW, max(s[2] for s in shapes), g.device, layout=layout,
                    nchw_staging=(B, C, sum(s[2] * s[3] for s in shapes), _lib.dtype_code(g)))
            _lib.check(_lib.lib().abr_roi_align_multilevel_backward(
                g.data_ptr(), rois.data_ptr(), levels.data_ptr(), ptrs, hs, ws, sc, len(grads), B, C, rois.size(0),
                PH, PW, int(sampling_ratio), _lib.dtype_code(g), layout, 1,
                wk.data_ptr() if wk is not None else None, wk_bytes, has_plan, _lib.stream_ptr(g.device)))
        return (None, None, None, None, None) + tuple(grads)


class Pooler(nn.Module):
    """Pooler for detection with or without FPN (modeling/poolers.py:45-105): ``forward(x, boxes)`` with
    ``x`` a list of per-level feature maps and ``boxes`` a list of BoxLists -> ``[R,C,P,P]`` in RoI order."""

    def __init__(self, output_size, scales, sampling_ratio):
        super(Pooler, self).__init__()
        self.poolers = nn.ModuleList(
            [ROIAlign(output_size, spatial_scale=scale, sampling_ratio=sampling_ratio) for scale in scales])
        self.output_size = output_size
        self.scales = tuple(scales)
        self.sampling_ratio = sampling_ratio
        lvl_min = -torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item()
        lvl_max = -torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item()
        self.map_levels = LevelMapper(lvl_min, lvl_max)

    def convert_to_roi_format(self, boxes):
        concat_boxes = cat([b.bbox for b in boxes], dim=0)
        device, dtype = concat_boxes.device, concat_boxes.dtype
        ids = cat([torch.full((len(b), 1), i, dtype=dtype, device=device) for i, b in enumerate(boxes)], dim=0)
        return torch.cat([ids, concat_boxes], dim=1)

    def forward(self, x, boxes):
        rois = self.convert_to_roi_format(boxes)
        if len(self.poolers) == 1:
            return self.poolers[0](x[0], rois)
        if len(x) > _lib.ABR_MAX_LEVELS:
            raise RuntimeError("at most %d feature levels are supported" % _lib.ABR_MAX_LEVELS)
        dtype = x[0].dtype
        feats = [_lib.as_compute_dtype(f) for f in x]
        rois32 = _prep_rois(rois, feats[0].device)
        levels = self.map_levels.levels_of_rois(rois32)
        out = _MultiLevelROIAlign.apply(rois32, levels, _pair(self.output_size), self.scales, self.sampling_ratio, *feats)
        return out.to(dtype)


def make_pooler(cfg, head_name):
    """modeling/poolers.py:108-117"""
    resolution = cfg.MODEL[head_name].POOLER_RESOLUTION
    scales = cfg.MODEL[head_name].POOLER_SCALES
    sampling_ratio = cfg.MODEL[head_name].POOLER_SAMPLING_RATIO
    return Pooler(output_size=(resolution, resolution), scales=scales, sampling_ratio=sampling_ratio)
