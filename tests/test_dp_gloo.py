"""CPU, gloo, world_size 2: the data-parallel exchange of the B200 path (loss partial sums all-reduced with SUM between
loss forward and backward; gradients all-reduced with SUM; per-replica BatchNorm) reproduces nn.DataParallel's
semantics, and the product helpers behind it (parallel.shard_rows, loss_functions._allreduce_sums,
parallel.allreduce_gradients) work on a real process group."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_siamese_cd_b200 import loss_functions, parallel
from oracle import dp_oracle as D
from oracle import unet_oracle as O

WORLD = 2
MTYPE, CIN, TOPO, B, HW = "siameseunet", 4, (64, 128), 4, 32


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, port: int, q) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        parallel.enable_data_parallel()
        assert parallel.is_enabled()
        # product helper: global partial sums
        sums = torch.tensor([1.0 + rank, 2.0, 3.0 * rank], dtype=torch.float64)
        loss_functions._allreduce_sums(sums)
        assert sums.tolist() == [3.0, 4.0, 3.0]
        sd = O.clone_state(O.reference_state_dict(MTYPE, in_channels=CIN, topology=TOPO, seed=7))
        batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
        rows = parallel.shard_rows(B, rank, WORLD)
        res = D.dp_rank_step(MTYPE, sd, D.shard(batch, rows), loss_functions._allreduce_sums,
                             parallel.allreduce_gradients)
        q.put((rank, res["loss"].item(), res["logits"], {k: v for k, v in res["grads"].items() if v is not None}))
        dist.barrier()
    finally:
        parallel.disable_data_parallel()
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_exchange_reproduces_dataparallel_semantics():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(WORLD):
        rank, loss, logits, grads = q.get(timeout=240)
        got[rank] = (loss, logits, grads)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process emulation of nn.DataParallel on the same global batch
    torch.set_num_threads(4)
    sd = O.clone_state(O.reference_state_dict(MTYPE, in_channels=CIN, topology=TOPO, seed=7))
    batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
    ref = D.dp_emulation_step(MTYPE, sd, batch, WORLD, lambda r: parallel.shard_rows(B, r, WORLD))
    for rank in range(WORLD):
        loss, logits, grads = got[rank]
        assert abs(loss - ref["loss"].item()) < 1e-6                      # every rank sees the GLOBAL loss
        rows = parallel.shard_rows(B, rank, WORLD)
        assert (logits - ref["logits"][rows]).abs().max().item() < 1e-5   # per-replica BatchNorm statistics
        for k, g in grads.items():
            r = ref["grads"][k]
            if k.endswith((".conv.0.bias", ".conv.3.bias")):
                continue                                                  # analytically zero (noise)
            assert (g - r).norm().item() <= 2e-4 * r.norm().item() + 1e-8, k   # SUM of replica gradients
    # both ranks end with identical gradients
    for k in got[0][2]:
        assert torch.equal(got[0][2][k], got[1][2][k]), k


def test_dp_differs_from_full_batch_batchnorm():
    """Guard against 'fixing' the semantics: DataParallel replicas normalise with their OWN chunk statistics, so the
    result is NOT the single-replica full-batch forward."""
    sd = O.clone_state(O.reference_state_dict(MTYPE, in_channels=CIN, topology=TOPO, seed=7))
    batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
    full = O.forward(MTYPE, O.clone_state(sd), batch["x_t1"], batch["x_t2"], train=True).detach()
    dp = D.dp_emulation_step(MTYPE, sd, batch, WORLD, lambda r: parallel.shard_rows(B, r, WORLD))["logits"]
    assert (full - dp).abs().max().item() > 1e-3
