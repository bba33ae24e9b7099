"""numpy restatement of modeling/poolers.py (LevelMapper + Pooler) and of
structures/boxlist_ops.py:boxlist_nms over the C oracle -- TEST INFRASTRUCTURE ONLY."""
import numpy as np

from . import nms, roi_align_forward


def map_levels(boxes_xyxy, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
    """modeling/poolers.py:31-42 in fp32, area with the +1 convention (structures/bounding_box.py:227-231)."""
    b = np.asarray(boxes_xyxy, np.float32)
    one = np.float32(1)
    area = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    s = np.sqrt(area)
    lvl = np.floor(np.float32(canonical_level) + np.log2(s / np.float32(canonical_scale) + np.float32(eps)))
    lvl = np.clip(lvl, np.float32(k_min), np.float32(k_max))
    return lvl.astype(np.int64) - int(k_min)


def to_roi_format(boxes_per_image):
    """modeling/poolers.py:73-78: [R,5] fp32 rows (image index, x1, y1, x2, y2)."""
    rows = [np.concatenate([np.full((len(b), 1), i, np.float32), np.asarray(b, np.float32)], 1)
            for i, b in enumerate(boxes_per_image)]
    return np.concatenate(rows, 0)


def pooler(feats, boxes_per_image, output_size, scales, sampling_ratio):
    """modeling/poolers.py:80-105."""
    rois = to_roi_format(boxes_per_image)
    P = output_size
    if len(scales) == 1:
        return roi_align_forward(feats[0], rois, scales[0], P, P, sampling_ratio)
    k_min = -np.log2(np.float32(scales[0]))
    k_max = -np.log2(np.float32(scales[-1]))
    levels = map_levels(rois[:, 1:], k_min, k_max)
    out = np.zeros((len(rois), feats[0].shape[1], P, P), np.float32)
    for lvl, (f, s) in enumerate(zip(feats, scales)):
        idx = np.nonzero(levels == lvl)[0]
        out[idx] = roi_align_forward(f, rois[idx], s, P, P, sampling_ratio)
    return out


def boxlist_nms(boxes_xyxy, scores, thresh, max_proposals=-1, flavour="cuda"):
    """structures/boxlist_ops.py:9-31 on xyxy boxes: returns the kept indices (ascending, truncated)."""
    if thresh <= 0:
        return np.arange(len(boxes_xyxy))
    keep = nms(boxes_xyxy, scores, thresh, flavour)
    return keep[:max_proposals] if max_proposals > 0 else keep
