"""``FastRCNNLossComputation.__call__`` with the reference's interface (modeling/roi_heads/box_head/loss.py:15-184):
the classification loss (inclusive, ``dist_type='id'``, or plain cross-entropy) and the smooth-L1 box loss, forward and
backward in ONE kernel (``abr_fastrcnn_loss`` of libabr_b200) instead of ~25 + ~40 tiny tensor kernels."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .... import _lib
from ....distillation.distillation import _scale_in_place


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
        gl, gr = ctx.grads
        ctx.grads = (None, None)
        if gl is not None:
            gl = _scale_in_place(gl, grad_cls).to(ctx.meta[0])
        if gr is not None:
            gr = _scale_in_place(gr, grad_box).to(ctx.meta[1]).reshape(ctx.meta[2])
        return gl, gr, None, None, None, None, None


def fastrcnn_loss(class_logits, box_regression, labels, regression_targets, n_old=-1, cls_agnostic_bbox_reg=False, beta=1.0):
    """(classification_loss, box_loss) of loss.py:122-184 for already concatenated tensors.  ``n_old >= 0`` selects the
    inclusive classification loss with that many old classes, ``n_old < 0`` plain cross-entropy."""
    return _FastRCNNLoss.apply(class_logits, box_regression, labels, regression_targets, n_old, cls_agnostic_bbox_reg, beta)


class FastRCNNLossComputation(object):
    """Computes the loss for Faster R-CNN (same constructor and ``__call__`` as the reference's class).  ``subsample`` --
    matching + random fg/bg sampling, SURVEY 8f rank 2 -- is not part of this library: set ``_proposals`` (BoxLists
    with ``labels`` and ``regression_targets`` fields), e.g. from the reference's own ``subsample``."""

    def __init__(self, proposal_matcher, fg_bg_sampler, box_coder, cls_agnostic_bbox_reg=False, dist_type=None, old_classes=[]):
        self.proposal_matcher = proposal_matcher
        self.fg_bg_sampler = fg_bg_sampler
        self.box_coder = box_coder
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.dist_type = dist_type
        self.n_old_cl = len(old_classes)

    def subsample(self, proposals, targets):
        raise NotImplementedError("FastRCNNLossComputation.subsample is outside the accelerated path (SURVEY 8f rank 2)")

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
