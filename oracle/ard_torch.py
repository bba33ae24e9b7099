"""PyTorch restatement of the reference's Attentive RoI Distillation loss -- TEST INFRASTRUCTURE ONLY.

Follows distillation/distillation.py:86-130 op for op (abs -> pow -> channel mean -> softmax over
H*W scaled by H*W; L1 between the two attention maps; MSE between sqrt(attention)-weighted features)
so that it costs what the reference costs on a CPU and rounds the way the reference rounds in fp32.
Used (a) as the fp32 cross-check of oracle.ard and (b) as the "port" CPU baseline of bench.py.
Argument order is the CALL SITE's (tools/train_incremental.py:115): first = old model (teacher),
second = new model (student); the attention of the FIRST argument weights the feature term.
"""
import torch
import torch.nn.functional as F


def attention_map(f_map: torch.Tensor) -> torch.Tensor:
    """distillation.py:121-130 (``activation_at`` with temp=2, no temperature division)."""
    n, _, h, w = f_map.shape
    energy = f_map.abs().pow(2).mean(dim=1, keepdim=True)
    return (h * w * F.softmax(energy.view(n, -1), dim=1)).view(n, h, w)


def ard_loss(f_first: torch.Tensor, f_second: torch.Tensor, gamma: float = 1.0):
    """distillation.py:86-118.  Returns (loss, loss_afd, loss_pad)."""
    att_first = attention_map(f_first)      # the reference names this S_attention_t
    att_second = attention_map(f_second)    # ... and this S_attention_s
    loss_pad = F.l1_loss(att_second, att_first, reduction="mean")
    root = torch.sqrt(att_first.unsqueeze(1))
    loss_afd = F.mse_loss(f_first * root, f_second * root, reduction="mean")
    return loss_afd + gamma * loss_pad, loss_afd, loss_pad


def ard_fwd_bwd(f_old: torch.Tensor, f_new: torch.Tensor, gamma: float = 1.0):
    """One forward+backward the way train_incremental.py:83-85,115,145 runs it: the teacher tensor
    carries no grad, the student does.  Returns (loss, dL/dF_new)."""
    f_new = f_new.detach().requires_grad_(True)
    loss, _, _ = ard_loss(f_old.detach(), f_new, gamma)
    (grad,) = torch.autograd.grad(loss, f_new)
    return loss.detach(), grad
