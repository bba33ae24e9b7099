"""Mirror of ``maskrcnn_benchmark.layers`` for the ops on the hot path (layers/__init__.py:10-14)."""
from .nms import nms, nms_batched
from .roi_align import ROIAlign, roi_align
from .roi_pool import ROIPool, roi_pool

__all__ = ["nms", "nms_batched", "roi_align", "ROIAlign", "roi_pool", "ROIPool"]
