"""PyTorch (CPU, autograd) restatement of the two logit-level losses of the incremental step -- TEST INFRASTRUCTURE ONLY.

  * roi_distillation_id  : calculate_roi_distillation_losses(dist='id') (distillation/distillation.py:164-241):
                           unbiased cross-entropy between teacher and student class logits + L2 on the old classes' boxes
  * fastrcnn_loss        : FastRCNNLossComputation.__call__ (modeling/roi_heads/box_head/loss.py:122-184):
                           inclusive classification loss (dist_type 'id') or plain cross-entropy, + smooth-L1 box loss
Written against the formulas, not the reference's line structure; pinned by tests/golden/logit_losses.npz, produced by
running the reference's own functions here (fp32 and fp64, with autograd gradients).
"""
import torch
import torch.nn.functional as F


def roi_distillation_id(soften_scores, soften_bboxes, target_scores, target_bboxes):
    """Teacher ("soften") logits [R,Co] / boxes [R,Co,4]; student ("target") logits [R,Ct] / boxes [R,Ct,4], Ct > Co.
    Returns (total, class term, box term)."""
    R, Co = soften_scores.shape
    Ct = target_scores.shape[1]
    den = torch.logsumexp(target_scores, dim=1)
    # the student's background = its own background plus every class the teacher never saw
    bkg_cols = torch.cat([target_scores[:, :1], target_scores[:, Co:]], dim=1)
    log_p_bkg = torch.logsumexp(bkg_cols, dim=1) - den
    log_p_old = target_scores[:, 1:Co] - den[:, None]
    teacher = torch.softmax(soften_scores, dim=1)
    per_row = (teacher[:, 0] * log_p_bkg + (teacher[:, 1:] * log_p_old).sum(dim=1)) / Co
    cls_term = -per_row.mean()
    diff = target_bboxes[:, 1:Co, :] - soften_bboxes[:, 1:, :]
    box_term = (diff * diff).sum(dim=2).mean(dim=1).mean(dim=0)
    return cls_term + box_term, cls_term, box_term


def smooth_l1_sum(x, y, beta):
    """layers/smooth_l1_loss.py:6-18 with size_average=False."""
    n = (x - y).abs()
    return torch.where(n < beta, 0.5 * n * n / beta, n - 0.5 * beta).sum()


def fastrcnn_loss(class_logits, box_regression, labels, regression_targets, n_old=-1, cls_agnostic_bbox_reg=False, beta=1.0):
    """n_old >= 0: inclusive classification loss (loss.py:151-159): the background log-probability is the log of the
    summed probability of background + the n_old old classes, old-class columns score 0; n_old < 0: F.cross_entropy.
    Returns (classification_loss, box_loss)."""
    if n_old >= 0:
        den = torch.logsumexp(class_logits, dim=1)
        outputs = torch.zeros_like(class_logits)
        outputs[:, 0] = torch.logsumexp(class_logits[:, : n_old + 1], dim=1) - den
        outputs[:, n_old + 1:] = class_logits[:, n_old + 1:] - den[:, None]
        cls = F.nll_loss(outputs, labels)
    else:
        cls = F.cross_entropy(class_logits, labels)
    pos = torch.nonzero(labels > 0).squeeze(1)
    lab = labels[pos]
    cols = (torch.tensor([4, 5, 6, 7]) if cls_agnostic_bbox_reg else 4 * lab[:, None] + torch.arange(4))
    box = smooth_l1_sum(box_regression[pos[:, None], cols], regression_targets[pos], beta) / labels.numel()
    return cls, box
