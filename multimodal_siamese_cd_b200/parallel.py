"""One-process-per-GPU data parallelism with the semantics of the reference's `nn.DataParallel(model)`
(utils/networks.py:27), minus its per-step parameter broadcast / scatter / gather:

  * every rank holds a persistent replica and processes its own contiguous chunk of the batch;
  * BatchNorm statistics stay per replica (the reference has no SyncBN);
  * the power-Jaccard loss is the ratio over the GLOBAL batch: its three partial sums are all-reduced (SUM) between
    the loss forward and backward kernels (loss_functions._allreduce_sums / TrainStep.run);
  * parameter gradients are all-reduced with SUM (not mean): DataParallel reduce-adds replica gradients of a loss that
    is already global.

`enable_data_parallel()` switches the drop-in modules (unchanged training scripts launched with torchrun) to this
behaviour; the fused `TrainStep` picks the default process group up by itself and additionally overlaps the gradient
all-reduce with the rest of backward in buckets.
"""
from __future__ import annotations

import torch

from . import loss_functions

_STATE = {"enabled": False, "group": None}


def enable_data_parallel(group=None) -> None:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("enable_data_parallel: torch.distributed is not initialised")
    _STATE["enabled"] = dist.get_world_size(group) > 1
    _STATE["group"] = group
    loss_functions.set_data_parallel_group("default" if group is None else group)


def disable_data_parallel() -> None:
    _STATE["enabled"] = False
    _STATE["group"] = None
    loss_functions.set_data_parallel_group(None)


def is_enabled() -> bool:
    return _STATE["enabled"]


def allreduce_gradients(flat: torch.Tensor) -> None:
    """SUM all-reduce of a flat gradient buffer across the data-parallel group (no-op when disabled)."""
    if _STATE["enabled"]:
        import torch.distributed as dist
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_STATE["group"])


def shard_rows(batch_size: int, rank: int, world: int) -> slice:
    """Rows of the global batch owned by `rank`: contiguous chunks of ceil(B / world) rows, as
    torch.nn.parallel.scatter does for nn.DataParallel."""
    chunk = -(-batch_size // world)
    lo = min(batch_size, rank * chunk)
    return slice(lo, min(batch_size, lo + chunk))
