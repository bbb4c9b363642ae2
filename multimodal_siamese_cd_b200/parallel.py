"""One-process-per-GPU data parallelism with the semantics of the reference's `nn.DataParallel(model)`
(utils/networks.py:27), minus its per-step parameter broadcast / scatter / gather:

  * every rank holds a persistent replica and processes its own contiguous chunk of the batch;
  * BatchNorm statistics stay per replica (the reference has no SyncBN);
  * the power-Jaccard loss is the ratio over the GLOBAL batch: its three partial sums are all-reduced (SUM) between
    the loss forward and backward kernels (loss_functions._allreduce_sums / TrainStep.run);
  * parameter gradients are all-reduced with SUM (not mean): DataParallel reduce-adds replica gradients of a loss that
    is already global.

Two ways to feed the ranks:

  * `enable_data_parallel()` ("sharded" inputs): every rank passes its OWN rows (`shard_rows`) to the network and the
    loss; logits stay local and the loss all-reduces its three partial sums. What the fused `TrainStep`, bench.py and a
    training script with a DistributedSampler-style loader use.
  * `enable_data_parallel(mode="gather")` (replicated inputs): what the launcher selects for the reference's UNCHANGED
    scripts under torchrun, whose DataLoader hands every rank the same full batch. The network shards the batch by rank
    inside `forward` (DataParallel's scatter), all-gathers the logits (DataParallel's gather) and returns full-batch
    logits on every rank, so the scripts' loss code — including the boolean row indexing of train_semisupervised.py —
    runs as written on the global batch; backward takes this rank's rows of the logit gradient. Evaluation
    (`net.eval()`) runs the whole batch on every rank.

In both modes the gradient all-reduce is bucketed and overlapped with the rest of backward (StepEngine.backward_dp).
"""
from __future__ import annotations

import torch

from . import loss_functions

_STATE = {"enabled": False, "group": None, "mode": "sharded"}

# Objects that hold CUDA graphs with captured NCCL collectives (step.TrainStep). ncclCommDestroy — torch's
# destroy_process_group as well as b200cd_comm_destroy — waits until every graph that references the communicator has
# been destroyed, so the graphs are dropped in disable_data_parallel() before any communicator goes away.
_GRAPH_HOLDERS = __import__("weakref").WeakSet()


def register_graph_holder(obj) -> None:
    """`obj.release_comm_graphs()` is called by disable_data_parallel()."""
    _GRAPH_HOLDERS.add(obj)


def release_comm_graphs() -> None:
    holders = list(_GRAPH_HOLDERS)
    if holders and torch.cuda.is_available():
        torch.cuda.synchronize()
    for h in holders:
        h.release_comm_graphs()


def enable_data_parallel(group=None, mode: str = "sharded") -> None:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("enable_data_parallel: torch.distributed is not initialised")
    if mode not in ("sharded", "gather"):
        raise ValueError(f"enable_data_parallel: mode must be 'sharded' or 'gather' (got {mode!r})")
    _STATE["enabled"] = dist.get_world_size(group) > 1
    _STATE["group"] = group
    _STATE["mode"] = mode
    # gathered logits: the loss already sees the global batch on every rank, nothing to all-reduce there
    loss_functions.set_data_parallel_group(("default" if group is None else group) if mode == "sharded" else None)


def _nccl_path() -> bytes:
    import glob
    import os
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
    return cands[0].encode() if cands else b""


def enable_native_comm() -> bool:
    """Create the LIBRARY-owned NCCL communicator (b200cd_comm_init, include/b200cd.h) over the ranks of the enabled
    data-parallel group: rank 0 draws the NCCL unique id and the 128 bytes travel through the existing
    torch.distributed group once. Afterwards gradient buckets and loss sums are all-reduced by libb200cd itself
    (b200cd_allreduce_bucket / _f64) on the caller's streams — capturable into the step's CUDA graph — and
    torch.distributed is only the bootstrap. Returns False (and changes nothing) when the group is not NCCL-backed."""
    import torch.distributed as dist

    from . import _lib
    if not _STATE["enabled"] or dist.get_backend(_STATE["group"]) != "nccl":
        return False
    if _STATE.get("native"):
        return True
    lib = _lib.load()
    _lib.check(lib.b200cd_comm_load(_nccl_path()))
    rank, world = rank_world()
    dev = torch.device("cuda", torch.cuda.current_device())
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        import ctypes as C
        buf = (C.c_char * 128)()
        _lib.check(lib.b200cd_comm_unique_id(buf))
        idt.copy_(torch.tensor(list(bytes(buf)), dtype=torch.uint8))
    src = dist.get_global_rank(_STATE["group"], 0) if _STATE["group"] is not None else 0
    dist.broadcast(idt, src=src, group=_STATE["group"])
    host = bytes(idt.cpu().numpy().tobytes())
    _lib.check(lib.b200cd_comm_init(host, rank, world))
    _STATE["native"] = True
    return True


def native_comm() -> bool:
    return bool(_STATE.get("native"))


def allreduce_sum_(t: torch.Tensor) -> None:
    """In-place SUM all-reduce of a contiguous fp32 / fp64 CUDA tensor on the current stream: the library's own NCCL
    communicator when it exists, else torch.distributed."""
    if _STATE.get("native"):
        from . import _lib
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64)
        fn = _lib.load().b200cd_allreduce_bucket if t.dtype == torch.float32 else _lib.load().b200cd_allreduce_f64
        _lib.check(fn(t.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream))
    else:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_STATE["group"])


def disable_data_parallel() -> None:
    """Call before torch.distributed.destroy_process_group(): graphs with captured collectives are released first."""
    release_comm_graphs()
    if _STATE.get("native"):
        from . import _lib
        _lib.load().b200cd_comm_destroy()
        _STATE["native"] = False
    _STATE["enabled"] = False
    _STATE["group"] = None
    _STATE["mode"] = "sharded"
    loss_functions.set_data_parallel_group(None)


def is_enabled() -> bool:
    return _STATE["enabled"]


def group():
    return _STATE["group"]


def gather_mode() -> bool:
    return _STATE["enabled"] and _STATE["mode"] == "gather"


def rank_world() -> tuple[int, int]:
    import torch.distributed as dist
    return dist.get_rank(_STATE["group"]), dist.get_world_size(_STATE["group"])


class GatherRows(torch.autograd.Function):
    """nn.DataParallel's gather for one-process-per-GPU replicas: every rank contributes its rows of an output and
    receives the full batch; backward keeps this rank's rows of the incoming gradient (the loss is computed on the
    global batch on every rank, so that slice IS d(global loss) / d(local logits))."""

    @staticmethod
    def forward(ctx, local: torch.Tensor, batch_size: int):
        import torch.distributed as dist
        rank, world = rank_world()
        rows = shard_rows(batch_size, rank, world)
        chunk = -(-batch_size // world)
        assert local.shape[0] == rows.stop - rows.start, (local.shape, rows)
        pad = local.new_zeros((chunk,) + tuple(local.shape[1:]))
        pad[:local.shape[0]].copy_(local)
        full = local.new_empty((world * chunk,) + tuple(local.shape[1:]))
        dist.all_gather(list(full.split(chunk, 0)), pad, group=_STATE["group"])   # views of `full`: filled in place
        ctx.rows = rows
        # ranks own contiguous chunks of `chunk` rows and only the last non-empty chunk can be short
        return full[:batch_size].contiguous() if world * chunk != batch_size else full

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        return g[ctx.rows].contiguous(), None


def allreduce_gradients(flat: torch.Tensor) -> None:
    """SUM all-reduce of a flat gradient buffer across the data-parallel group (no-op when disabled)."""
    if _STATE["enabled"]:
        allreduce_sum_(flat)


def shard_rows(batch_size: int, rank: int, world: int) -> slice:
    """Rows of the global batch owned by `rank`: contiguous chunks of ceil(B / world) rows, as
    torch.nn.parallel.scatter does for nn.DataParallel."""
    chunk = -(-batch_size // world)
    lo = min(batch_size, rank * chunk)
    return slice(lo, min(batch_size, lo + chunk))
