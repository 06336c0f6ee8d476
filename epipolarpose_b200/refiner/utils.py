"""Refiner loop helpers -- host-side mirror of the reference refiner/utils.py (AverageMeter :4-15,
lr_decay :18-22, step_decay :24-28, save_ckpt :30-36), plus clip_grad_norm_ on the device."""
import os

import torch

from epipolarpose_b200 import ops as _ops

_backend = [_ops]


class AverageMeter:
    """Running value / weighted mean (the four public attributes of the reference's meter)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.sum, self.count, self.avg = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.count = self.count + n
        self.sum = self.sum + n * val
        self.avg = self.sum / self.count


def _exponential_lr(optimizer, step, lr, decay_step, gamma):
    """lr * gamma^(step / decay_step), written into every parameter group; returns the new rate."""
    new_lr = lr * pow(gamma, step / decay_step)
    for group in optimizer.param_groups:
        group["lr"] = new_lr
    return new_lr


# the reference keeps two names for the same schedule (refiner/utils.py:18-28)
lr_decay = _exponential_lr
step_decay = _exponential_lr


def save_ckpt(state, ckpt_path, is_best=True):
    """best.pth.tar / last.pth.tar under ckpt_path (refiner/utils.py:30-36)."""
    torch.save(state, os.path.join(ckpt_path, ("best" if is_best else "last") + ".pth.tar"))


def clip_grad_norm_(parameters, max_norm):
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm) with norm_type 2 (refiner/main.py:57)
    on the device: one sum-of-squares kernel per gradient into a float64 scalar, one scale kernel
    per gradient -- no host synchronisation.  Returns the total norm as a 0-dim device tensor."""
    ops = _backend[0]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros(())
    total = torch.zeros(1, device=grads[0].device, dtype=torch.float64)
    for g in grads:
        if not g.is_contiguous():
            raise ValueError("gradients must be contiguous")
        ops.sumsq(g, g.numel(), total)
    for g in grads:
        ops.clip_scale(g, g.numel(), total, max_norm)
    return total.sqrt().reshape(())
