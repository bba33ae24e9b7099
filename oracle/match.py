"""numpy restatement of the box head's proposal <-> ground-truth matching -- TEST INFRASTRUCTURE ONLY.

Follows, in fp32 and in the reference's operation order:
  * boxlist_iou                                  structures/boxlist_ops.py:53-88
  * Matcher.__call__ (no low-quality matches)    modeling/matcher.py:52-81
  * FastRCNNLossComputation.prepare_targets      modeling/roi_heads/box_head/loss.py:57-84
  * BoxCoder.encode                              modeling/box_coder.py:22-50  (torch.log on CPU: the third-party arithmetic
                                                 the reference itself calls)
Pinned by tests/golden/match.npz, produced by running the reference's FastRCNNLossComputation.prepare_targets here.
"""
import numpy as np
import torch

F = np.float32


def area(b):
    return (b[:, 2] - b[:, 0] + F(1)) * (b[:, 3] - b[:, 1] + F(1))


def box_iou(boxes1, boxes2):
    b1, b2 = np.asarray(boxes1, F).reshape(-1, 4), np.asarray(boxes2, F).reshape(-1, 4)
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.maximum(rb - lt + F(1), F(0))
    inter = wh[:, :, 0] * wh[:, :, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (area(b1)[:, None] + area(b2)[None, :] - inter)


def match(quality, high, low):
    """quality [G, n] -> matched index per column (first maximum), -1 below low, -2 in [low, high)."""
    vals = quality.max(axis=0)
    idx = quality.argmax(axis=0).astype(np.int64)
    out = idx.copy()
    out[vals < F(low)] = -1
    out[(vals >= F(low)) & (vals < F(high))] = -2
    return out


def encode(reference_boxes, proposals, weights):
    r, p = np.asarray(reference_boxes, F), np.asarray(proposals, F)
    one, half = F(1), F(0.5)
    ex_w, ex_h = p[:, 2] - p[:, 0] + one, p[:, 3] - p[:, 1] + one
    ex_cx, ex_cy = p[:, 0] + half * ex_w, p[:, 1] + half * ex_h
    gt_w, gt_h = r[:, 2] - r[:, 0] + one, r[:, 3] - r[:, 1] + one
    gt_cx, gt_cy = r[:, 0] + half * gt_w, r[:, 1] + half * gt_h
    wx, wy, ww, wh = (F(w) for w in weights)
    log = lambda x: torch.from_numpy(np.ascontiguousarray(x, F)).log().numpy()  # noqa: E731
    return np.stack([wx * (gt_cx - ex_cx) / ex_w, wy * (gt_cy - ex_cy) / ex_h, ww * log(gt_w / ex_w), wh * log(gt_h / ex_h)], 1)


def prepare_targets(proposals, gt_boxes, gt_labels, high, low, weights):
    """One image: (matched_idxs [n], labels [n], regression_targets [n,4])."""
    m = match(box_iou(gt_boxes, proposals), high, low)
    clamped = np.maximum(m, 0)
    labels = np.asarray(gt_labels, np.int64)[clamped].copy()
    labels[m == -1] = 0
    labels[m == -2] = -1
    return m, labels, encode(np.asarray(gt_boxes, F)[clamped], proposals, weights)
