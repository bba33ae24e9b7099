"""numpy restatement of the RPN proposal path around NMS -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows, in fp32 and in the reference's operation order:
  * RPNPostProcessor.forward_for_single_feature_map   modeling/rpn/inference.py:76-118
  * permute_and_flatten                                modeling/rpn/utils.py:10-14
  * BoxCoder.decode                                    modeling/box_coder.py:52-95
  * BoxList.clip_to_image(remove_empty=False)          structures/bounding_box.py:214-225
  * remove_small_boxes                                 structures/boxlist_ops.py:34-48
  * boxlist_nms                                        structures/boxlist_ops.py:9-31   (via the C oracle's orc_nms)

Tie contract (the reference's torch.topk leaves the order of equal scores unspecified): candidates are ranked by
(logit descending, anchor index ascending).  sigmoid is monotonic, so with distinct sigmoid values this is exactly
the reference's order.  Pinned by tests/golden/rpn.npz, produced by running the reference's RPNPostProcessor here.
"""
import math

import numpy as np
import torch

from . import nms

F = np.float32
BBOX_XFORM_CLIP = math.log(1000.0 / 16)  # modeling/box_coder.py:20


def permute_and_flatten(layer, N, A, C, H, W):
    """modeling/rpn/utils.py:10-14: [N, A*C, H, W] -> [N, H*W*A, C]."""
    return np.ascontiguousarray(layer.reshape(N, A, C, H, W).transpose(0, 3, 4, 1, 2)).reshape(N, -1, C)


def decode(rel_codes, boxes, weights=(1.0, 1.0, 1.0, 1.0), clip=BBOX_XFORM_CLIP):
    """modeling/box_coder.py:52-95 for [n,4] codes, every operation rounded to fp32 like the tensor ops."""
    rel_codes, boxes = np.asarray(rel_codes, F), np.asarray(boxes, F)
    one, half = F(1), F(0.5)
    widths = boxes[:, 2] - boxes[:, 0] + one
    heights = boxes[:, 3] - boxes[:, 1] + one
    ctr_x = boxes[:, 0] + half * widths
    ctr_y = boxes[:, 1] + half * heights
    wx, wy, ww, wh = (F(w) for w in weights)
    dx, dy = rel_codes[:, 0] / wx, rel_codes[:, 1] / wy
    dw = np.minimum(rel_codes[:, 2] / ww, F(clip))
    dh = np.minimum(rel_codes[:, 3] / wh, F(clip))
    pred_ctr_x = dx * widths + ctr_x
    pred_ctr_y = dy * heights + ctr_y
    pred_w = _exp(dw) * widths
    pred_h = _exp(dh) * heights
    out = np.empty_like(rel_codes)
    out[:, 0] = pred_ctr_x - half * pred_w
    out[:, 1] = pred_ctr_y - half * pred_h
    out[:, 2] = pred_ctr_x + half * pred_w - one
    out[:, 3] = pred_ctr_y + half * pred_h - one
    return out


def clip_to_image(boxes, size):
    """structures/bounding_box.py:214-219; size = (width, height)."""
    w, h = size
    b = boxes.copy()
    b[:, 0] = np.clip(b[:, 0], F(0), F(w - 1))
    b[:, 1] = np.clip(b[:, 1], F(0), F(h - 1))
    b[:, 2] = np.clip(b[:, 2], F(0), F(w - 1))
    b[:, 3] = np.clip(b[:, 3], F(0), F(h - 1))
    return b


def small_box_mask(boxes, min_size):
    """structures/boxlist_ops.py:43-47 (xywh widths carry the +1: structures/bounding_box.py:86-90)."""
    ws = boxes[:, 2] - boxes[:, 0] + F(1)
    hs = boxes[:, 3] - boxes[:, 1] + F(1)
    return (ws >= F(min_size)) & (hs >= F(min_size))


def _exp(x):
    """torch.exp on CPU: the third-party arithmetic the reference itself calls (box_coder.py:79-80)."""
    return torch.from_numpy(np.ascontiguousarray(x, F)).exp().numpy()


def sigmoid(x):
    """torch.sigmoid on CPU (inference.py:89)."""
    return torch.from_numpy(np.ascontiguousarray(x, F)).sigmoid().numpy()


def select_topk(logits, k):
    """indices of the k largest logits, ranked (logit descending, index ascending)."""
    order = np.argsort(-np.asarray(logits, F), kind="stable")
    return order[:k]


def candidates(objectness, box_regression, anchors, image_sizes, pre_nms_top_n, min_size, weights=(1.0, 1.0, 1.0, 1.0),
               clip=BBOX_XFORM_CLIP):
    """Everything before the NMS (inference.py:88-110 + the two filters of :113-115): per image the decoded, clipped,
    size-filtered boxes in rank order, their scores and the anchor index each came from."""
    objectness, box_regression = np.asarray(objectness, F), np.asarray(box_regression, F)
    N, A, H, W = objectness.shape
    logits = permute_and_flatten(objectness, N, A, 1, H, W).reshape(N, -1)
    reg = permute_and_flatten(box_regression, N, A, 4, H, W)
    k = min(int(pre_nms_top_n), A * H * W)
    scores_all = sigmoid(logits)  # on the whole [N, H*W*A] tensor like inference.py:89 (same vector/tail split in torch)
    out = []
    for n in range(N):
        idx = select_topk(logits[n], k)
        anc = np.asarray(anchors[n] if len(anchors) == N else anchors[0], F).reshape(-1, 4)[idx]
        boxes = clip_to_image(decode(reg[n][idx], anc, weights, clip), image_sizes[n])
        ok = small_box_mask(boxes, min_size)
        out.append((boxes[ok], scores_all[n][idx][ok], idx[ok]))
    return out


def rpn_proposals(objectness, box_regression, anchors, image_sizes, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size,
                  weights=(1.0, 1.0, 1.0, 1.0), clip=BBOX_XFORM_CLIP, flavour="cuda"):
    """forward_for_single_feature_map: list over images of (proposals [m,4], objectness [m], anchor index [m])."""
    out = []
    for boxes, scores, idx in candidates(objectness, box_regression, anchors, image_sizes, pre_nms_top_n, min_size,
                                         weights, clip):
        if nms_thresh > 0:  # structures/boxlist_ops.py:22-23: a non-positive threshold returns the list untouched
            keep = nms(boxes, scores, nms_thresh, flavour)
            if post_nms_top_n > 0:
                keep = keep[:post_nms_top_n]
            boxes, scores, idx = boxes[keep], scores[keep], idx[keep]
        out.append((boxes, scores, idx))
    return out
