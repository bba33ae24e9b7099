"""``FastRCNNLossComputation.__call__`` with the reference's interface (modeling/roi_heads/box_head/loss.py:15-184):
the classification loss (inclusive, ``dist_type='id'``, or plain cross-entropy) and the smooth-L1 box loss, forward and
backward in ONE kernel (``abr_fastrcnn_loss`` of libabr_b200) instead of ~25 + ~40 tiny tensor kernels."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

import ctypes

from .... import _lib
from ....structures.bounding_box import BoxList


def match_proposals(proposals, targets, matcher, box_coder):
    """Device matching of a batch (``abr_match_proposals``).  proposals / targets: list[BoxList] (targets carry a
    ``labels`` field).  Returns three lists of per-image tensors: labels [n] int64, regression_targets [n,4], matched_idxs [n]."""
    if getattr(matcher, "allow_low_quality_matches", False):
        raise NotImplementedError("match_proposals: allow_low_quality_matches (the RPN loss setting) is outside this path")
    n_images = len(proposals)
    if n_images != len(targets) or n_images == 0:
        raise RuntimeError("match_proposals: need one target BoxList per proposal BoxList")
    dev = proposals[0].bbox.device
    _lib.require_cuda(proposals[0].bbox, "proposals")
    for p, t in zip(proposals, targets):
        if p.size != t.size:
            raise RuntimeError("boxlists should have same image size, got {}, {}".format(t, p))  # boxlist_ops.py:67-69
        if len(t) == 0:
            raise ValueError("No ground-truth boxes available for one of the images during training")  # matcher.py:55-58
        if len(p) == 0:
            raise ValueError("No proposal boxes available for one of the images during training")  # matcher.py:59-62
    counts = [len(p) for p in proposals]
    gcounts = [len(t) for t in targets]
    cat = lambda ts: ts[0] if len(ts) == 1 else torch.cat(ts, 0)  # noqa: E731
    boxes = cat([p.convert("xyxy").bbox for p in proposals]).detach().to(torch.float32).contiguous()
    gt = cat([t.convert("xyxy").bbox.to(dev) for t in targets]).detach().to(torch.float32).contiguous()
    gl = cat([t.get_field("labels").to(dev) for t in targets]).to(torch.int64).contiguous()
    R = boxes.shape[0]
    matched = torch.empty((R,), dtype=torch.int64, device=dev)
    labels = torch.empty((R,), dtype=torch.int64, device=dev)
    reg = torch.empty((R, 4), dtype=torch.float32, device=dev)
    wts = (ctypes.c_float * 4)(*[float(w) for w in box_coder.weights])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().abr_match_proposals(
            boxes.data_ptr(), (ctypes.c_int * n_images)(*counts), gt.data_ptr(), gl.data_ptr(), (ctypes.c_int * n_images)(*gcounts),
            n_images, float(matcher.high_threshold), float(matcher.low_threshold), wts, matched.data_ptr(), labels.data_ptr(),
            reg.data_ptr(), _lib.stream_ptr(dev)))
    return list(labels.split(counts)), list(reg.split(counts)), list(matched.split(counts))


class _FastRCNNLoss(Function):
    @staticmethod
    def forward(ctx, class_logits, box_regression, labels, regression_targets, n_old, cls_agnostic, beta):
        _lib.require_cuda(class_logits, "class_logits")
        R, C = class_logits.shape
        logits = class_logits.detach().to(torch.float32).contiguous()
        reg = box_regression.detach().to(torch.float32).contiguous().reshape(R, -1)
        lab = labels.detach().to(device=logits.device, dtype=torch.int64).contiguous()
        tgt = regression_targets.detach().to(device=logits.device, dtype=torch.float32).contiguous()
        if lab.shape != (R,) or tgt.shape != (R, 4):
            raise RuntimeError("fastrcnn loss: labels [R] and regression_targets [R,4] expected, got %s and %s"
                               % (tuple(lab.shape), tuple(tgt.shape)))
        dev = logits.device
        loss2 = torch.empty((2,), dtype=torch.float32, device=dev)
        gl = torch.empty_like(logits) if class_logits.requires_grad else None
        gr = torch.empty_like(reg) if box_regression.requires_grad else None
        L = _lib.lib()
        ws_bytes = int(L.abr_logit_loss_workspace_bytes(R))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.abr_fastrcnn_loss(logits.data_ptr(), reg.data_ptr(), reg.shape[1], lab.data_ptr(), tgt.data_ptr(), R, C,
                                           int(n_old), int(bool(cls_agnostic)), float(beta), 1.0, 1.0,
                                           gl.data_ptr() if gl is not None else None, gr.data_ptr() if gr is not None else None,
                                           loss2.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(dev)))
        ctx.grads = (gl, gr)
        ctx.meta = (class_logits.dtype, box_regression.dtype, box_regression.shape)
        return loss2[0].clone(), loss2[1].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_cls, grad_box):
        # out of place: the stored gradients (for upstream gradients of 1) survive, so a second backward over the same
        # graph (retain_graph=True) is correct; the tensors are [R, C]-sized
        gl, gr = ctx.grads
        if gl is not None:
            gl = (gl * grad_cls.to(gl.dtype)).to(ctx.meta[0])
        if gr is not None:
            gr = (gr * grad_box.to(gr.dtype)).to(ctx.meta[1]).reshape(ctx.meta[2])
        return gl, gr, None, None, None, None, None


def fastrcnn_loss(class_logits, box_regression, labels, regression_targets, n_old=-1, cls_agnostic_bbox_reg=False, beta=1.0):
    """(classification_loss, box_loss) of loss.py:122-184 for already concatenated tensors.  ``n_old >= 0`` selects the
    inclusive classification loss with that many old classes, ``n_old < 0`` plain cross-entropy."""
    return _FastRCNNLoss.apply(class_logits, box_regression, labels, regression_targets, n_old, cls_agnostic_bbox_reg, beta)


class FastRCNNLossComputation(object):
    """Computes the loss for Faster R-CNN (same constructor, ``subsample`` and ``__call__`` as the reference's class).
    ``prepare_targets`` is ONE device launch for the whole batch (``abr_match_proposals``: IoU, Matcher, labels, box
    encoding); the fg/bg sampling keeps the reference's ``torch.randperm`` stream by default and runs as one launch for the
    batch with a sampler built with ``device_sampling=True`` (see the sampler's docstring)."""

    def __init__(self, proposal_matcher, fg_bg_sampler, box_coder, cls_agnostic_bbox_reg=False, dist_type=None, old_classes=[]):
        self.proposal_matcher = proposal_matcher
        self.fg_bg_sampler = fg_bg_sampler
        self.box_coder = box_coder
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.dist_type = dist_type
        self.n_old_cl = len(old_classes)

    def prepare_targets(self, proposals, targets):
        """loss.py:57-84 for the batch: per image the int64 labels (0 background, -1 ignored) and the regression targets.
        Also returns the matched ground-truth indices (``matched_idxs`` of loss.py:43-55)."""
        return match_proposals(proposals, targets, self.proposal_matcher, self.box_coder)

    def match_targets_to_proposals(self, proposal, target):
        """loss.py:43-55 for one image: the matched targets with ``labels`` and ``matched_idxs`` fields."""
        _, _, matched = match_proposals([proposal], [target], self.proposal_matcher, self.box_coder)
        idx = matched[0].clamp(min=0)
        out = BoxList(target.bbox[idx], target.size, target.mode)
        out.add_field("labels", target.get_field("labels")[idx])
        out.add_field("matched_idxs", matched[0])
        return out

    def subsample(self, proposals, targets):
        """loss.py:86-120: positive/negative sampling; returns the sampled proposals with ``labels`` and
        ``regression_targets`` fields and remembers them for ``__call__``."""
        labels, regression_targets, _ = self.prepare_targets(proposals, targets)
        proposals = list(proposals)
        for lab, tgt, per_image in zip(labels, regression_targets, proposals):
            per_image.add_field("labels", lab)
            per_image.add_field("regression_targets", tgt)
        if getattr(self.fg_bg_sampler, "device_sampling", False):
            # one launch for the batch and ONE host read (the per-image counts) instead of three nonzero() syncs per image
            pos, neg, counts = self.fg_bg_sampler.sample_on_device(labels)
            kept = counts.sum(1).tolist()
            for i, (p_mask, n_mask) in enumerate(zip(pos, neg)):
                idx = torch.nonzero_static(p_mask | n_mask, size=kept[i]).squeeze(1)  # size known: no synchronisation
                proposals[i] = proposals[i][idx]
        else:
            sampled_pos_inds, sampled_neg_inds = self.fg_bg_sampler(labels)
            for i, (pos, neg) in enumerate(zip(sampled_pos_inds, sampled_neg_inds)):
                proposals[i] = proposals[i][torch.nonzero(pos | neg).squeeze(1)]
        self._proposals = proposals
        return proposals

    def __call__(self, class_logits, box_regression):
        """
        Arguments:
            class_logits (list[Tensor]), box_regression (list[Tensor])
        Returns:
            classification_loss (Tensor), box_loss (Tensor)
        """
        class_logits = class_logits[0] if len(class_logits) == 1 else torch.cat(list(class_logits), dim=0)
        box_regression = box_regression[0] if len(box_regression) == 1 else torch.cat(list(box_regression), dim=0)
        if not hasattr(self, "_proposals"):
            raise RuntimeError("subsample needs to be called before")
        proposals = self._proposals
        labels = torch.cat([p.get_field("labels") for p in proposals], dim=0)
        regression_targets = torch.cat([p.get_field("regression_targets") for p in proposals], dim=0)
        return fastrcnn_loss(class_logits, box_regression, labels, regression_targets,
                             self.n_old_cl if self.dist_type == "id" else -1, self.cls_agnostic_bbox_reg, 1.0)
