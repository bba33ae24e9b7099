"""``RPNPostProcessor`` with the reference's interface (modeling/rpn/inference.py:15-196), backed by
``abr_rpn_proposals`` of libabr_b200: sigmoid, top-k, box decode, clip, small-box filter and NMS of the WHOLE batch run
on the device with one host synchronisation (the per-image proposal counts) instead of the reference's per-image loop."""
import ctypes

import torch

from ... import _lib
from ...structures.bounding_box import BoxList, make_boxlist
from ...structures.boxlist_ops import cat_boxlist
from ..box_coder import BoxCoder


def rpn_proposals(objectness, box_regression, anchors, image_sizes, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size,
                  weights=(1.0, 1.0, 1.0, 1.0), bbox_xform_clip=None, cpu_tie_rule=False, return_anchor_index=False):
    """Device part of ``forward_for_single_feature_map`` (inference.py:76-118) for the whole batch.

    Arguments:
        objectness (Tensor[N,A,H,W]), box_regression (Tensor[N,4A,H,W]): RPN head outputs, contiguous or channels-last
        anchors (Tensor[M,4] shared by all images, or Tensor[N,M,4]), M = A*H*W in (h, w, a) order
        image_sizes (list[(width, height)])
    Returns:
        proposals (Tensor[N,S,4]), scores (Tensor[N,S]), n_out (IntTensor[N], device) [, anchor_index (IntTensor[N,S])]
        where S = the padded per-image capacity; image i owns the first n_out[i] rows.
    """
    _lib.require_cuda(objectness, "objectness")
    N, A, H, W = objectness.shape
    M = A * H * W
    if box_regression.shape != (N, 4 * A, H, W):
        raise RuntimeError("box_regression should be [N, 4A, H, W] = %s, got %s" % ((N, 4 * A, H, W), tuple(box_regression.shape)))
    nhwc = _lib.is_channels_last(objectness)
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    obj = objectness.detach().to(torch.float32).contiguous(memory_format=fmt)
    reg = box_regression.detach().to(torch.float32).contiguous(memory_format=fmt)
    anchors = anchors.detach().to(device=obj.device, dtype=torch.float32).contiguous()
    if anchors.dim() == 2:
        stride = 0
        if anchors.shape != (M, 4):
            raise RuntimeError("anchors should be [%d,4], got %s" % (M, tuple(anchors.shape)))
    else:
        stride = M * 4
        if anchors.shape != (N, M, 4):
            raise RuntimeError("anchors should be [%d,%d,4], got %s" % (N, M, tuple(anchors.shape)))
    if len(image_sizes) != N:
        raise RuntimeError("need one (width, height) per image")
    if bbox_xform_clip is None:
        bbox_xform_clip = BoxCoder((1, 1, 1, 1)).bbox_xform_clip
    k = min(int(pre_nms_top_n), M)
    cap = min(k, int(post_nms_top_n)) if (post_nms_top_n > 0 and nms_thresh > 0) else k
    dev = obj.device
    proposals = torch.empty((N, cap, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((N, cap), dtype=torch.float32, device=dev)
    n_out = torch.empty((N,), dtype=torch.int32, device=dev)
    anchor_index = torch.empty((N, cap), dtype=torch.int32, device=dev) if return_anchor_index else None
    sizes = (ctypes.c_int * (2 * N))(*[int(v) for s in image_sizes for v in s])
    wts = (ctypes.c_float * 4)(*[float(w) for w in weights])
    L = _lib.lib()
    ws_bytes = int(L.abr_rpn_proposals_workspace_bytes(N, A, H, W, int(pre_nms_top_n), int(post_nms_top_n)))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.abr_rpn_proposals(
            obj.data_ptr(), reg.data_ptr(), anchors.data_ptr(), stride, sizes, N, A, H, W,
            _lib.ABR_NHWC if nhwc else _lib.ABR_NCHW, int(pre_nms_top_n), int(post_nms_top_n), float(nms_thresh),
            int(bool(cpu_tie_rule)), float(min_size), wts, float(bbox_xform_clip), proposals.data_ptr(), scores.data_ptr(),
            anchor_index.data_ptr() if anchor_index is not None else None, n_out.data_ptr(), cap, ws.data_ptr(), ws_bytes,
            _lib.stream_ptr(dev)))
    if return_anchor_index:
        return proposals, scores, n_out, anchor_index
    return proposals, scores, n_out


class RPNPostProcessor(torch.nn.Module):
    """Performs post-processing on the outputs of the RPN boxes, before feeding the proposals to the heads
    (same constructor and methods as the reference's class)."""

    def __init__(self, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size, box_coder=None, fpn_post_nms_top_n=None,
                 fpn_post_nms_per_batch=True):
        super(RPNPostProcessor, self).__init__()
        self.pre_nms_top_n = pre_nms_top_n
        self.post_nms_top_n = post_nms_top_n
        self.nms_thresh = nms_thresh
        self.min_size = min_size
        if box_coder is None:
            box_coder = BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        self.box_coder = box_coder
        if fpn_post_nms_top_n is None:
            fpn_post_nms_top_n = post_nms_top_n
        self.fpn_post_nms_top_n = fpn_post_nms_top_n
        self.fpn_post_nms_per_batch = fpn_post_nms_per_batch

    def add_gt_proposals(self, proposals, targets):
        """inference.py:52-74: ground-truth boxes join the proposals with objectness 1."""
        device = proposals[0].bbox.device
        out = []
        for proposal, target in zip(proposals, targets):
            gt = BoxList(target.bbox.to(device), target.size, target.mode).convert(proposal.mode)
            gt.add_field("objectness", torch.ones(len(gt), device=device))
            out.append(cat_boxlist((proposal, gt)))
        return out

    def forward_for_single_feature_map(self, anchors, objectness, box_regression):
        """
        Arguments:
            anchors: list[BoxList]
            objectness: tensor of size N, A, H, W
            box_regression: tensor of size N, A * 4, H, W
        """
        boxes = [a.bbox for a in anchors]
        if all(b.data_ptr() == boxes[0].data_ptr() and b.shape == boxes[0].shape for b in boxes):
            anchor_t = boxes[0]
        else:
            anchor_t = torch.stack(boxes, 0)
        image_sizes = [a.size for a in anchors]
        proposals, scores, n_out = rpn_proposals(
            objectness, box_regression, anchor_t, image_sizes, self.pre_nms_top_n, self.post_nms_top_n, self.nms_thresh,
            self.min_size, self.box_coder.weights, self.box_coder.bbox_xform_clip)
        counts = n_out.tolist()  # the only host synchronisation of the batch
        result = []
        for i, size in enumerate(image_sizes):
            boxlist = make_boxlist(proposals[i, : counts[i]], size, mode="xyxy")
            boxlist.add_field("objectness", scores[i, : counts[i]])
            result.append(boxlist)
        return result

    def forward(self, anchors, objectness, box_regression, targets=None):
        """anchors: per image, per level BoxLists; objectness / box_regression: one tensor per level.
        Returns one BoxList of proposals per image (decoded, clipped, NMS-filtered; with the ground truth appended while
        training, inference.py:120-147)."""
        levels = len(objectness)
        per_level_anchors = list(zip(*anchors))  # level -> the images' anchor BoxLists
        per_level = [self.forward_for_single_feature_map(per_level_anchors[lvl], objectness[lvl], box_regression[lvl])
                     for lvl in range(levels)]
        proposals = [cat_boxlist([per_level[lvl][img] for lvl in range(levels)]) for img in range(len(per_level[0]))]
        if levels > 1:
            proposals = self.select_over_all_levels(proposals)
        if self.training and targets is not None:
            proposals = self.add_gt_proposals(proposals, targets)
        return proposals

    def select_over_all_levels(self, boxlists):
        """FPN only (inference.py:149-178; a handful of small tensors per batch, kept as tensor glue).  Training with
        ``fpn_post_nms_per_batch``: the best ``fpn_post_nms_top_n`` proposals of the WHOLE batch survive, each image
        keeping its own order; otherwise every image keeps its best ``fpn_post_nms_top_n``, best first."""
        budget = self.fpn_post_nms_top_n
        if self.training and self.fpn_post_nms_per_batch:
            counts = [len(b) for b in boxlists]
            scores = torch.cat([b.get_field("objectness") for b in boxlists], dim=0)
            winners = torch.topk(scores, min(budget, scores.numel()), dim=0, sorted=True)[1]
            keep = torch.zeros(scores.shape, dtype=torch.bool, device=scores.device)
            keep[winners] = True
            return [b[m] for b, m in zip(boxlists, keep.split(counts))]
        out = []
        for b in boxlists:
            scores = b.get_field("objectness")
            out.append(b[torch.topk(scores, min(budget, scores.numel()), dim=0, sorted=True)[1]])
        return out


def make_rpn_postprocessor(config, rpn_box_coder, is_train):
    """inference.py:181-204: reads the same config keys."""
    rpn = config.MODEL.RPN
    return RPNPostProcessor(
        pre_nms_top_n=rpn.PRE_NMS_TOP_N_TRAIN if is_train else rpn.PRE_NMS_TOP_N_TEST,
        post_nms_top_n=rpn.POST_NMS_TOP_N_TRAIN if is_train else rpn.POST_NMS_TOP_N_TEST,
        nms_thresh=rpn.NMS_THRESH,
        min_size=rpn.MIN_SIZE,
        box_coder=rpn_box_coder,
        fpn_post_nms_top_n=rpn.FPN_POST_NMS_TOP_N_TRAIN if is_train else rpn.FPN_POST_NMS_TOP_N_TEST,
        fpn_post_nms_per_batch=rpn.FPN_POST_NMS_PER_BATCH,
    )
