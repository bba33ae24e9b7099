"""Prototype Box Selection scoring on the device: the two numeric steps of the reference's offline
tool -- the channel-mean descriptor of every ground-truth box (tools/prototype_box_selection.py:96-101) and the
nearest-to-class-mean ranking of ``Mem.mean_feature_sampling`` (tools/extract_memory.py:111-161) -- backed by
``abr_channel_mean`` / ``abr_prototype_distances`` of libabr_b200.  The image cropping and file writing around them
(``creat_and_save_box_image``) are replaced by the packed store of ``abr_iod_b200.data.prototype_store``.  The other two
selection rules of the reference's ``Mem`` are here as well: ``herding_ranking`` (``herding_feature_sampling``,
extract_memory.py:163-211, one kernel) and ``random_ranking`` (``rnd_sampling``, :83-109: a ``random.shuffle``, host glue)."""
import random

import torch

from .. import _lib


def roi_descriptors(roi_align_features):
    """``torch.mean(roi_align_features.cpu(), dim=1)`` without the device-to-host copy of the pooled tensor:
    [R,C,P,P] (contiguous or channels-last, fp32/bf16) -> [R,P,P] fp32 on the device."""
    _lib.require_cuda(roi_align_features, "roi_align_features")
    x = _lib.as_compute_dtype(roi_align_features.detach())
    nhwc = _lib.is_channels_last(x)
    x = x.contiguous(memory_format=torch.channels_last if nhwc else torch.contiguous_format)
    R, C, H, W = x.shape
    out = torch.empty((R, H, W), dtype=torch.float32, device=x.device)
    if R:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().abr_channel_mean(x.data_ptr(), R, C, H * W, _lib.dtype_code(x),
                                                   _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW, out.data_ptr(), _lib.stream_ptr(x.device)))
    return out


def mean_feature_ranking(features, num_bbox_per_cls):
    """Ranking of one class, extract_memory.py:115-147.  ``features``: [n,P,P] (or [n,F]) descriptors of the class's boxes.
    Like the reference, a class with fewer than ``num_bbox_per_cls`` boxes is first topped up with copies of its first
    boxes (:116-120).  Returns (indices into the topped-up list, nearest first, cut to num_bbox_per_cls; distances of all
    topped-up entries, float64; the source box of every topped-up entry).  Equal distances (the copies) keep list order."""
    f = torch.as_tensor(features)
    _lib.require_cuda(f, "features")
    f = f.detach().to(torch.float32).reshape(f.shape[0], -1)
    n = f.shape[0]
    if n == 0:
        raise RuntimeError("mean_feature_ranking: a class without boxes cannot be ranked")
    source = torch.arange(n, device=f.device)
    if n < num_bbox_per_cls:
        deficit = num_bbox_per_cls - n
        f = torch.cat([f, f[:deficit]], 0)
        source = torch.cat([source, source[:deficit]], 0)
    f = f.contiguous()
    mean = torch.empty((f.shape[1],), dtype=torch.float64, device=f.device)
    dist = torch.empty((f.shape[0],), dtype=torch.float64, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(_lib.lib().abr_prototype_distances(f.data_ptr(), f.shape[0], f.shape[1], mean.data_ptr(), dist.data_ptr(),
                                                      _lib.stream_ptr(f.device)))
    order = torch.sort(dist, stable=True)[1][:num_bbox_per_cls]
    return order, dist, source


def _top_up(n, num_bbox_per_cls):
    """extract_memory.py:116-120 / :169-173 / :93-94: a class with fewer boxes is topped up with copies of its first ones."""
    source = list(range(n))
    if n < num_bbox_per_cls:
        source.extend(source[: num_bbox_per_cls - n])
    return source


def herding_ranking(features, num_bbox_per_cls):
    """Selection of one class by ``Mem.herding_feature_sampling`` (extract_memory.py:163-211): greedily, the box whose
    inclusion brings the mean of the chosen descriptors closest to the normalised class mean.  ``features``: [n,P,P] or
    [n,F].  Returns (indices into the topped-up list in selection order, the source box of every topped-up entry).
    The reference's loop runs over all boxes and keeps the first ``num_bbox_per_cls`` picks; the kernel stops there."""
    f = torch.as_tensor(features)
    _lib.require_cuda(f, "features")
    f = f.detach().to(torch.float32).reshape(f.shape[0], -1)
    if f.shape[0] == 0:
        raise RuntimeError("herding_ranking: a class without boxes cannot be ranked")
    source = torch.tensor(_top_up(f.shape[0], num_bbox_per_cls), device=f.device)
    f = f[source].contiguous()
    n, F = f.shape
    k = min(n, num_bbox_per_cls)
    mean = torch.empty((F,), dtype=torch.float64, device=f.device)
    dist = torch.empty((n,), dtype=torch.float64, device=f.device)
    selected = torch.empty((k,), dtype=torch.int64, device=f.device)
    ws_bytes = F * 8 + n
    ws = torch.empty((ws_bytes + 8,), dtype=torch.uint8, device=f.device)
    with torch.cuda.device(f.device):
        st = _lib.stream_ptr(f.device)
        _lib.check(_lib.lib().abr_prototype_distances(f.data_ptr(), n, F, mean.data_ptr(), dist.data_ptr(), st))  # the class mean
        _lib.check(_lib.lib().abr_prototype_herding(f.data_ptr(), n, F, k, mean.data_ptr(), selected.data_ptr(), ws.data_ptr(),
                                                    ws_bytes, st))
    return selected, source


def random_ranking(n, num_bbox_per_cls):
    """``Mem.rnd_sampling`` (extract_memory.py:83-109) for one class of ``n`` boxes: ``random.shuffle`` of the class's list
    (the same draw from Python's ``random``), top-up with the first shuffled entries, the first ``num_bbox_per_cls`` kept.
    Returns the chosen source boxes in order.  Host-side index glue, no kernel."""
    order = list(range(n))
    random.shuffle(order)
    if n < num_bbox_per_cls:
        order.extend(order[: num_bbox_per_cls - n])
    return order[:num_bbox_per_cls]
