"""Stand-in for the reference's pybind11 module ``maskrcnn_benchmark._C`` (csrc/vision.cpp:9-24), restricted to the
hot path: the same five functions with the same positional signatures, backed by libabr_b200.  With
``abr_iod_b200.compat.install()`` the reference's own ``layers/roi_align.py``, ``roi_pool.py`` and ``nms.py`` import
this module as ``maskrcnn_benchmark._C`` and run unchanged."""
from .layers.nms import nms as _nms
from .layers.roi_align import roi_align_backward as _ra_bwd
from .layers.roi_align import roi_align_forward as _ra_fwd
from .layers.roi_pool import roi_pool_backward as _rp_bwd
from .layers.roi_pool import roi_pool_forward as _rp_fwd
from . import _lib


def nms(dets, scores, threshold):
    """csrc/nms.h:10-28"""
    return _nms(dets, scores, threshold)


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """csrc/ROIAlign.h:11-25"""
    return _ra_fwd(_lib.as_compute_dtype(input), rois, spatial_scale, pooled_height, pooled_width, sampling_ratio)


def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
                       sampling_ratio):
    """csrc/ROIAlign.h:27-46"""
    return _ra_bwd(_lib.as_compute_dtype(grad), rois, spatial_scale, pooled_height, pooled_width, batch_size, channels,
                   height, width, sampling_ratio)


def roi_pool_forward(input, rois, spatial_scale, pooled_height, pooled_width):
    """csrc/ROIPool.h:9-25"""
    return _rp_fwd(_lib.as_compute_dtype(input), rois, spatial_scale, pooled_height, pooled_width)


def roi_pool_backward(grad, input, rois, argmax, spatial_scale, pooled_height, pooled_width, batch_size, channels,
                      height, width):
    """csrc/ROIPool.h:27-48"""
    return _rp_bwd(_lib.as_compute_dtype(grad), input, rois, argmax, spatial_scale, pooled_height, pooled_width,
                   batch_size, channels, height, width)


def _not_on_hot_path(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("maskrcnn_benchmark._C.%s is outside the RoI hot path that abr_iod_b200 replaces" % name)
    fn.__name__ = name
    return fn


for _name in ("sigmoid_focalloss_forward", "sigmoid_focalloss_backward", "deform_conv_forward",
              "deform_conv_backward_input", "deform_conv_backward_parameters", "modulated_deform_conv_forward",
              "modulated_deform_conv_backward", "deform_psroi_pooling_forward", "deform_psroi_pooling_backward"):
    globals()[_name] = _not_on_hot_path(_name)
