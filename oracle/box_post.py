"""numpy restatement of the box-head PostProcessor -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows modeling/roi_heads/box_head/inference.py of the reference:
  * PostProcessor.forward          :42-83   softmax, BoxCoder.decode of every class' deltas, per-image split, clip
  * PostProcessor.prepare_boxlist  :85-103
  * PostProcessor.filter_results   :105-151 score threshold, per-class NMS (class 0 = background kept apart), cat of the
                                            foreground classes, detections_per_img cut by kthvalue (ties survive)
softmax / exp are torch's CPU kernels (the third-party arithmetic the reference itself calls).  Pinned by
tests/golden/box_post.npz, produced by running the reference's PostProcessor here.
"""
import numpy as np
import torch

from . import nms
from .rpn import BBOX_XFORM_CLIP, clip_to_image, decode

F = np.float32


def softmax(logits):
    return torch.softmax(torch.from_numpy(np.ascontiguousarray(logits, F)), -1).numpy()


def decode_all_classes(box_regression, boxes, weights, clip=BBOX_XFORM_CLIP):
    """box_coder.decode on [R, 4K] codes: the same anchor box for every class column (box_coder.py:67-93)."""
    R = box_regression.shape[0]
    K = box_regression.shape[1] // 4
    out = np.empty((R, 4 * K), F)
    for j in range(K):
        out[:, 4 * j:4 * j + 4] = decode(box_regression[:, 4 * j:4 * j + 4], boxes, weights, clip)
    return out


def filter_results(boxes, scores, num_classes, score_thresh, nms_thresh, detections_per_img, flavour="cuda"):
    """inference.py:105-151 for one image.  boxes [n, 4C] (clipped), scores [n, C].
    Returns (boxes [d,4], scores [d], labels [d]) and the background triple."""
    per_class = []
    for j in range(num_classes):
        inds = np.nonzero(scores[:, j] > F(score_thresh))[0]
        s_j = scores[inds, j]
        b_j = boxes[inds, 4 * j:4 * j + 4]
        if nms_thresh > 0:  # structures/boxlist_ops.py:22-23
            keep = nms(b_j, s_j, nms_thresh, flavour)
            b_j, s_j, inds = b_j[keep], s_j[keep], inds[keep]
        per_class.append((b_j, s_j, np.full((len(s_j),), j, np.int64), inds))
    background = per_class[0]
    fg = per_class[1:]
    b = np.concatenate([x[0] for x in fg], 0) if fg else np.zeros((0, 4), F)
    s = np.concatenate([x[1] for x in fg], 0) if fg else np.zeros((0,), F)
    lab = np.concatenate([x[2] for x in fg], 0) if fg else np.zeros((0,), np.int64)
    rows = np.concatenate([x[3] for x in fg], 0) if fg else np.zeros((0,), np.int64)
    d = len(s)
    if d > detections_per_img > 0:
        kth = np.sort(s, kind="stable")[d - detections_per_img]  # torch.kthvalue(k = d - det + 1), 1-based
        keep = np.nonzero(s >= kth)[0]
        b, s, lab, rows = b[keep], s[keep], lab[keep], rows[keep]
    return (b, s, lab, rows), background


def box_postprocess(class_logits, box_regression, proposals, image_sizes, score_thresh=0.05, nms_thresh=0.5,
                    detections_per_img=100, weights=(10.0, 10.0, 5.0, 5.0), clip=BBOX_XFORM_CLIP,
                    cls_agnostic_bbox_reg=False, flavour="cuda"):
    """PostProcessor.forward.  proposals: list of [n_i,4] xyxy per image.  Returns (results, backgrounds): per image
    (boxes, scores, labels, proposal row) for the foreground classes and for class 0."""
    class_logits, box_regression = np.asarray(class_logits, F), np.asarray(box_regression, F)
    prob = softmax(class_logits)
    C = prob.shape[1]
    concat = np.concatenate([np.asarray(p, F).reshape(-1, 4) for p in proposals], 0)
    reg = box_regression[:, -4:] if cls_agnostic_bbox_reg else box_regression
    dec = decode_all_classes(reg.reshape(len(concat), -1), concat, weights, clip)
    if cls_agnostic_bbox_reg:
        dec = np.tile(dec, (1, C))
    results, backgrounds = [], []
    start = 0
    for p, size in zip(proposals, image_sizes):
        n = len(p)
        boxes = clip_to_image(dec[start:start + n].reshape(-1, 4), size).reshape(n, 4 * C)
        res, bg = filter_results(boxes, prob[start:start + n], C, score_thresh, nms_thresh, detections_per_img, flavour)
        results.append(res)
        backgrounds.append(bg)
        start += n
    return results, backgrounds
