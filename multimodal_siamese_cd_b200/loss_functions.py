"""Drop-in for the reference's `utils/loss_functions.py`: `get_criterion(loss_type, negative_weight, positive_weight)`
returns a callable `(logits, target) -> 0-d tensor` (utils/loss_functions.py:6-33).

`'PowerJaccardLoss'` — the loss every reference config selects (configs/base.yaml:17,44) — runs in the sm_100a
kernels (b200cd_pj_fwd / _loss / _bwd) behind a torch autograd node and the `b200cd::power_jaccard` custom op.
Gradients flow to the logits and, when it requires grad, to the target (the MMCR consistency term passes
sigmoid(logits_stream2) as target, train_semisupervised.py:75-105). The other names map to the same torch
compositions the reference uses; they are not on the hot path.

Data-parallel semantics: the reference evaluates the loss on the gathered global batch (nn.DataParallel, SURVEY §0
finding 2). With one process per GPU call `set_data_parallel_group(group)`: the three partial sums are all-reduced
(SUM) between the forward and backward kernels so every rank sees the global ratio.
"""
from __future__ import annotations

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


class _PowerJaccard(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z: torch.Tensor, t: torch.Tensor):
        if not z.is_cuda:
            raise RuntimeError("PowerJaccardLoss (b200cd) runs on CUDA tensors only; there is no CPU fallback")
        zc = z.detach().float().contiguous()
        tc = t.detach().float().contiguous()
        if zc.numel() != tc.numel():
            raise ValueError(f"power_jaccard_loss: logits {tuple(z.shape)} and target {tuple(t.shape)} differ in size")
        n = zc.numel()
        if n % 4 != 0:
            raise NotImplementedError("power_jaccard_loss kernel needs a multiple of 4 elements")
        sums = torch.zeros(3, device=z.device, dtype=torch.float64)
        loss = torch.empty((), device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            if n > 0:
                nblk = _nblk(n)
                ws = torch.empty(nblk * 3, device=z.device, dtype=torch.float64)
                ops.pj_fwd(zc.view(1, -1), tc.view(1, -1), False, None, 0, nblk, ws, sums)
            _allreduce_sums(sums)
            ops.pj_loss(sums, loss)
        ctx.save_for_backward(zc, tc, sums)
        ctx.shapes = (z.shape, t.shape)
        return loss

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        zc, tc, sums = ctx.saved_tensors
        need_z, need_t = ctx.needs_input_grad
        dz = torch.empty_like(zc)
        dt = torch.empty_like(tc) if need_t else None
        if zc.numel() > 0:
            with torch.cuda.device(zc.device):
                ops.pj_bwd(zc.view(1, -1), tc.view(1, -1), False, None, 0, sums, g.float().contiguous(), 1.0, False,
                           dz.view(1, -1), None if dt is None else dt.view(1, -1))
        dz = dz.view(ctx.shapes[0]) if need_z else None
        dt = dt.view(ctx.shapes[1]) if need_t else None
        return dz, dt


def power_jaccard_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:  # noqa: A002 (reference arg name)
    """1 - I / (sum p^2 + sum t^2 - I + 1e-6), p = sigmoid(input), I = sum p*t, over the whole (global) batch."""
    return _PowerJaccard.apply(input, target)


# torch.library registration: `torch.ops.b200cd.power_jaccard(logits, target)` (forward only; the autograd path is
# the Function above, which is what get_criterion returns).
try:
    _lib_def = torch.library.Library("b200cd", "DEF")
    _lib_def.define("power_jaccard(Tensor logits, Tensor target) -> Tensor")

    def _pj_cuda(logits, target):
        return _PowerJaccard.apply(logits.detach(), target.detach())

    def _pj_meta(logits, target):
        return logits.new_empty(())

    _lib_def.impl("power_jaccard", _pj_cuda, "CUDA")
    _lib_def.impl("power_jaccard", _pj_meta, "Meta")
except Exception:  # noqa: BLE001  (re-import in the same interpreter)
    pass


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
    dice_pos = (2.0 * (p * t).sum() + eps) / (p.sum() + t.sum() + eps)
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
