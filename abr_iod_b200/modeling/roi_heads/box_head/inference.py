"""``PostProcessor`` with the reference's interface (modeling/roi_heads/box_head/inference.py:12-174), backed by
``abr_box_postprocess`` of libabr_b200: softmax, per-class box decode, clip, score threshold, the NMS of every
(image, class) pair as ONE batched call and the detections_per_img cut run on the device with one host
synchronisation for the batch (the reference: ~21 per image)."""
import ctypes

import torch
from torch import nn

from .... import _lib
from ....structures.bounding_box import make_boxlist
from ...box_coder import BoxCoder


def box_postprocess(class_logits, box_regression, proposals, boxes_per_image, image_sizes, score_thresh=0.05, nms_thresh=0.5,
                    detections_per_img=100, weights=(10.0, 10.0, 5.0, 5.0), bbox_xform_clip=None,
                    cls_agnostic_bbox_reg=False, cpu_tie_rule=False, det_stride=None):
    """Device part of ``PostProcessor.forward`` for the whole batch.

    Arguments:
        class_logits (Tensor[R,C]), box_regression (Tensor[R,4C] or any width when class-agnostic: the last 4 columns)
        proposals (Tensor[R,4]): the images' reference boxes back to back; boxes_per_image (list[int])
    Returns a dict of padded device tensors: boxes [N,S,4], scores [N,S], labels [N,S] (int64), rows [N,S] (proposal
    index inside its image), n [N] (int32) and the class-0 results bg_boxes [N,B,4], bg_scores [N,B], bg_n [N].
    """
    _lib.require_cuda(class_logits, "class_logits")
    dev = class_logits.device
    logits = class_logits.detach().to(torch.float32).contiguous()
    reg = box_regression.detach().to(torch.float32).contiguous()
    R, C = logits.shape
    reg = reg.reshape(R, -1)
    props = proposals.detach().to(device=dev, dtype=torch.float32).contiguous()
    N = len(boxes_per_image)
    if sum(boxes_per_image) != R or props.shape != (R, 4) or len(image_sizes) != N:
        raise RuntimeError("box_postprocess: %d logits rows, %s proposals, boxes_per_image=%s" % (R, tuple(props.shape), list(boxes_per_image)))
    if reg.shape[1] < (4 if cls_agnostic_bbox_reg else 4 * C):
        raise RuntimeError("box_regression should have %d columns, got %d" % (4 * C, reg.shape[1]))
    if bbox_xform_clip is None:
        bbox_xform_clip = BoxCoder((1, 1, 1, 1)).bbox_xform_clip
    max_n = max(list(boxes_per_image) + [1])
    full = max(1, (C - 1) * max_n)
    if det_stride is None:
        det_stride = min(full, detections_per_img + 64) if detections_per_img > 0 else full
    counts = (ctypes.c_int * N)(*[int(n) for n in boxes_per_image])
    sizes = (ctypes.c_int * (2 * N))(*[int(v) for s in image_sizes for v in s])
    wts = (ctypes.c_float * 4)(*[float(w) for w in weights])
    L = _lib.lib()
    ws_bytes = int(L.abr_box_postprocess_workspace_bytes(counts, N, C))
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    while True:
        out = {
            "boxes": torch.empty((N, det_stride, 4), dtype=torch.float32, device=dev),
            "scores": torch.empty((N, det_stride), dtype=torch.float32, device=dev),
            "labels": torch.empty((N, det_stride), dtype=torch.int64, device=dev),
            "rows": torch.empty((N, det_stride), dtype=torch.int32, device=dev),
            "n": torch.empty((N,), dtype=torch.int32, device=dev),
            "bg_boxes": torch.empty((N, max_n, 4), dtype=torch.float32, device=dev),
            "bg_scores": torch.empty((N, max_n), dtype=torch.float32, device=dev),
            "bg_n": torch.empty((N,), dtype=torch.int32, device=dev),
        }
        with torch.cuda.device(dev):
            _lib.check(L.abr_box_postprocess(
                logits.data_ptr(), reg.data_ptr(), reg.shape[1], int(bool(cls_agnostic_bbox_reg)), props.data_ptr(), counts,
                sizes, N, C, float(score_thresh), float(nms_thresh), int(bool(cpu_tie_rule)), int(detections_per_img), wts,
                float(bbox_xform_clip), out["boxes"].data_ptr(), out["scores"].data_ptr(), out["labels"].data_ptr(),
                out["rows"].data_ptr(), out["n"].data_ptr(), det_stride, out["bg_boxes"].data_ptr(),
                out["bg_scores"].data_ptr(), out["bg_n"].data_ptr(), max_n, ws.data_ptr(), ws_bytes, _lib.stream_ptr(dev)))
        both = torch.stack((out["n"], out["bg_n"])).tolist()  # the one host synchronisation (and D2H copy) of the batch
        out["n_host"], out["bg_n_host"] = both
        worst = max(out["n_host"] + [0])
        if worst <= det_stride:
            return out
        det_stride = worst  # equal scores tied at the cut: run again with room for all of them


class PostProcessor(nn.Module):
    """From a set of classification scores, box regression and proposals, computes the post-processed boxes, and applies
    NMS to obtain the final results (same constructor and ``forward`` as the reference's class)."""

    def __init__(self, score_thresh=0.05, nms=0.5, detections_per_img=100, box_coder=None, cls_agnostic_bbox_reg=False):
        super(PostProcessor, self).__init__()
        self.score_thresh = score_thresh
        self.nms = nms
        self.detections_per_img = detections_per_img
        if box_coder is None:
            box_coder = BoxCoder(weights=(10.0, 10.0, 5.0, 5.0))
        self.box_coder = box_coder
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg

    def forward(self, x, boxes):
        """
        Arguments:
            x (tuple[tensor, tensor]): class logits and box regression of the box head
            boxes (list[BoxList]): the reference boxes, one BoxList per image
        Returns:
            results (list[BoxList]) with fields ``scores`` and ``labels``, and the last image's background-class
            BoxList (what the reference's loop leaves in ``results_background``)
        """
        class_logits, box_regression = x
        image_shapes = [box.size for box in boxes]
        boxes_per_image = [len(box) for box in boxes]
        concat_boxes = torch.cat([a.bbox for a in boxes], dim=0)
        out = box_postprocess(class_logits, box_regression, concat_boxes, boxes_per_image, image_shapes, self.score_thresh,
                              self.nms, self.detections_per_img, self.box_coder.weights, self.box_coder.bbox_xform_clip,
                              self.cls_agnostic_bbox_reg)
        bg_counts = out["bg_n_host"]
        results = []
        for i, size in enumerate(image_shapes):
            n = out["n_host"][i]
            boxlist = make_boxlist(out["boxes"][i, :n], size, mode="xyxy")
            boxlist.add_field("scores", out["scores"][i, :n])
            boxlist.add_field("labels", out["labels"][i, :n])
            results.append(boxlist)
        results_background = None
        if image_shapes:
            i = len(image_shapes) - 1
            results_background = make_boxlist(out["bg_boxes"][i, : bg_counts[i]], image_shapes[i], mode="xyxy")
            results_background.add_field("scores", out["bg_scores"][i, : bg_counts[i]])
            results_background.add_field("labels", torch.zeros((bg_counts[i],), dtype=torch.int64, device=class_logits.device))
        return results, results_background


def make_roi_box_post_processor(cfg):
    """inference.py:154-174: reads the same config keys."""
    heads = cfg.MODEL.ROI_HEADS
    return PostProcessor(heads.SCORE_THRESH, heads.NMS, heads.DETECTIONS_PER_IMG, BoxCoder(weights=heads.BBOX_REG_WEIGHTS),
                         cfg.MODEL.CLS_AGNOSTIC_BBOX_REG)
