"""LabelSmoothingLoss with the reference's interface (LabelSmoothingLoss.py:7-15), computed by sst_ce_sumexp_loss."""
import torch
from torch import nn

from . import lib as L

PAD = 42


class _CeSumExpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_bsc, target, eps, pad):
        B, S, C = logits_bsc.shape
        x = logits_bsc.contiguous().float()
        rows = B * S
        grad = torch.empty(rows, C, dtype=torch.float32, device=x.device)
        ws = torch.empty(2 * rows, dtype=torch.float32, device=x.device)
        out = torch.zeros(1, dtype=torch.float32, device=x.device)
        tgt = target.contiguous().view(-1)
        n_valid = int((tgt != pad).sum())
        L.ce_sumexp_loss(L.F32, L.F32, rows, S, C, x, C, tgt, pad, eps, n_valid, 1.0, ws, grad, C, out)
        ctx.save_for_backward(grad)
        ctx.shape = (B, S, C)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None, None, None


class LabelSmoothingLoss(nn.Module):
    def __init__(self, epsilon=0.1, num_classes=2):
        super().__init__()
        self.epsilon = epsilon
        self.num_classes = num_classes

    def forward(self, input, target):
        """input: (B, C, S) as the reference passes it (recognition_model.py:102)."""
        return _CeSumExpFn.apply(input.permute(0, 2, 1), target, self.epsilon, PAD)
