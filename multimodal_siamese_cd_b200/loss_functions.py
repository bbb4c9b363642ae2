"""Drop-in for the reference's `utils/loss_functions.py`: `get_criterion(loss_type, negative_weight, positive_weight)`
returns a callable `(logits, target) -> 0-d tensor` (utils/loss_functions.py:6-33).

`'PowerJaccardLoss'` — the loss every reference config selects (configs/base.yaml:17,44) — runs in the sm_100a
kernels (b200cd_pj_fwd / _loss / _bwd) behind the `b200cd::power_jaccard` / `b200cd::power_jaccard_backward` custom
ops (torch.library, fake implementations, register_autograd).
Gradients flow to the logits and, when it requires grad, to the target (the MMCR consistency term passes
sigmoid(logits_stream2) as target, train_semisupervised.py:75-105). The other names map to the same torch
compositions the reference uses; they are not on the hot path.

Data-parallel semantics: the reference evaluates the loss on the gathered global batch (nn.DataParallel, SURVEY §0
finding 2). With one process per GPU call `set_data_parallel_group(group)`: the three partial sums are all-reduced
(SUM) between the forward and backward kernels so every rank sees the global ratio.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import ops

_DP_GROUP = None
_DP_ENABLED = False


def set_data_parallel_group(group="default") -> None:
    """Make PowerJaccardLoss a global-batch loss across the ranks of `group` (None disables)."""
    global _DP_GROUP, _DP_ENABLED
    if group is None:
        _DP_GROUP, _DP_ENABLED = None, False
    else:
        _DP_GROUP = None if group == "default" else group
        _DP_ENABLED = True


def _allreduce_sums(sums: torch.Tensor) -> None:
    if _DP_ENABLED:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(_DP_GROUP) > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=_DP_GROUP)


def _nblk(numel: int) -> int:
    return max(1, min(296, numel // 4096))


# torch.library custom ops (namespace b200cd::): forward and backward of the loss as two ops with fake
# (meta) implementations, wired together with register_autograd. `torch.ops.b200cd.power_jaccard(logits, target)`
# returns (loss, sums); gradients flow to the logits and, when it requires grad, to the target.
@torch.library.custom_op("b200cd::power_jaccard", mutates_args=(), device_types="cuda")
def _pj_forward(logits: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    zc = logits.detach().float().contiguous()
    tc = target.detach().float().contiguous()
    n = zc.numel()
    sums = torch.zeros(3, device=logits.device, dtype=torch.float64)
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    with torch.cuda.device(logits.device):
        if n > 0:
            nblk = _nblk(n)
            ws = torch.empty(nblk * 3, device=logits.device, dtype=torch.float64)
            ops.pj_fwd(zc.view(1, -1), tc.view(1, -1), False, None, 0, nblk, ws, sums)
        _allreduce_sums(sums)   # unconditional for every term, empty row sets included: ranks cannot diverge
        ops.pj_loss(sums, loss)
    return loss, sums


@_pj_forward.register_fake
def _pj_forward_fake(logits, target):
    return logits.new_empty((), dtype=torch.float32), logits.new_empty((3,), dtype=torch.float64)


@torch.library.custom_op("b200cd::power_jaccard_backward", mutates_args=(), device_types="cuda")
def _pj_backward(logits: torch.Tensor, target: torch.Tensor, sums: torch.Tensor, grad: torch.Tensor,
                 need_target_grad: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    zc = logits.detach().float().contiguous()
    tc = target.detach().float().contiguous()
    dz = torch.empty_like(zc)
    dt = torch.empty_like(tc) if need_target_grad else torch.empty(0, device=zc.device)
    if zc.numel() > 0:
        with torch.cuda.device(zc.device):
            ops.pj_bwd(zc.view(1, -1), tc.view(1, -1), False, None, 0, sums, grad.float().contiguous(), 1.0, False,
                       dz.view(1, -1), dt.view(1, -1) if need_target_grad else None)
    return dz, dt


@_pj_backward.register_fake
def _pj_backward_fake(logits, target, sums, grad, need_target_grad):
    return (logits.new_empty(logits.shape, dtype=torch.float32),
            target.new_empty(target.shape if need_target_grad else (0,), dtype=torch.float32))


def _pj_setup(ctx, inputs, output):
    logits, target = inputs
    ctx.save_for_backward(logits, target, output[1])


def _pj_autograd(ctx, g_loss, g_sums):
    logits, target, sums = ctx.saved_tensors
    need_z, need_t = ctx.needs_input_grad
    dz, dt = torch.ops.b200cd.power_jaccard_backward(logits, target, sums, g_loss, bool(need_t))
    return (dz.view(logits.shape).to(logits.dtype) if need_z else None,
            dt.view(target.shape).to(target.dtype) if need_t else None)


torch.library.register_autograd("b200cd::power_jaccard", _pj_autograd, setup_context=_pj_setup)


def power_jaccard_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:  # noqa: A002 (reference arg name)
    """1 - I / (sum p^2 + sum t^2 - I + 1e-6), p = sigmoid(input), I = sum p*t, over the whole (global) batch."""
    if not input.is_cuda:
        raise RuntimeError("PowerJaccardLoss (b200cd) runs on CUDA tensors only; there is no CPU fallback")
    if input.numel() != target.numel():
        raise ValueError(f"power_jaccard_loss: logits {tuple(input.shape)} and target {tuple(target.shape)} differ in size")
    if input.numel() % 4 != 0:
        raise NotImplementedError("power_jaccard_loss kernel needs a multiple of 4 elements")
    return torch.ops.b200cd.power_jaccard(input, target)[0]


# ---- non-hot-path names kept for API parity (same formulas as utils/loss_functions.py:36-197) ----------------
def _flat_prob(logit, target):
    return torch.sigmoid(logit).flatten(), target.flatten()


def soft_dice_loss(y_logit, y_true):
    p, t = _flat_prob(y_logit, y_true)
    eps = 1e-6
    return 1 - (2.0 * (p * t).sum() + eps) / (p.sum() + t.sum() + eps)


soft_dice_squared_sum_loss = soft_dice_loss  # identical in the reference (utils/loss_functions.py:47-56, "TODO: fix")


def jaccard_like_loss(input, target):  # noqa: A002
    p, t = _flat_prob(input, target)
    inter = (p * t).sum()
    return 1 - 2.0 * inter / ((p ** 2 + t ** 2).sum() - inter + 1e-6)


def dice_like_loss(input, target):  # noqa: A002
    p, t = _flat_prob(input, target)
    return 1 - 2.0 * (p * t).sum() / ((p ** 2 + t ** 2).sum() + 1e-6)


def iou_loss(y_logit, y_true):
    p, t = _flat_prob(y_logit, y_true)
    inter = (p * t).sum()
    return 1 - inter / ((p + t).sum() - inter + 1e-6)


def soft_dice_loss_balanced(input, target):  # noqa: A002
    p, t = _flat_prob(input, target)
    eps = 1e-6
    dice_pos = (2.0 * (p * t).sum()) / (p.sum() + t.sum() + eps)   # no eps in the numerator (utils/loss_functions.py:192)
    np_, nt = 1 - p, 1 - t
    dice_neg = (2.0 * (np_ * nt).sum()) / (np_.sum() + nt.sum() + eps)
    return 1 - dice_pos - dice_neg


def get_criterion(loss_type, negative_weight: float = 1, positive_weight: float = 1):
    if loss_type == "PowerJaccardLoss":
        return power_jaccard_loss
    if loss_type == "BCEWithLogitsLoss":
        return nn.BCEWithLogitsLoss()
    if loss_type == "CrossEntropyLoss":
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return nn.CrossEntropyLoss(weight=torch.tensor([negative_weight, positive_weight]).float().to(device))
    if loss_type in ("MeanSquareErrorLoss", "L2"):
        return nn.MSELoss()
    table = {
        "SoftDiceLoss": soft_dice_loss,
        "SoftDiceSquaredSumLoss": soft_dice_squared_sum_loss,
        "SoftDiceBalancedLoss": soft_dice_loss_balanced,
        "IoULoss": iou_loss,
        "DiceLikeLoss": dice_like_loss,
    }
    if loss_type in table:
        return table[loss_type]
    raise Exception(f"unknown loss {loss_type}")
