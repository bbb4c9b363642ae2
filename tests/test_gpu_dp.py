"""GPU (needs >= 2 devices, skipped otherwise): one process per GPU over NCCL. The fused TrainStep and the drop-in
modules with parallel.enable_data_parallel() must reproduce nn.DataParallel's semantics (oracle.dp_oracle): global
power-Jaccard loss, SUM of replica gradients, per-replica BatchNorm statistics."""
from __future__ import annotations

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

MTYPE, CIN, TOPO, B, HW, WORLD = "siameseunet", 4, (64, 128), 4, 32, 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, port: int, path: str, q, backend: str = "nccl", precision: str = "fast") -> None:
    import torch.distributed as dist

    from multimodal_siamese_cd_b200 import loss_functions, networks, parallel
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    from oracle import unet_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    # nccl: one GPU per rank. gloo: every rank on cuda:0 (CUDA tensors staged through the host by gloo), so that the
    # data-parallel code paths are exercised on a ONE-GPU box too
    didx = rank if backend == "nccl" else 0
    torch.cuda.set_device(didx)
    dev = torch.device("cuda", didx)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        native = path.endswith("_native")
        path = path.replace("_native", "")
        parallel.enable_data_parallel(mode="gather" if path == "gather" else "sharded")
        if native:      # the library's own NCCL communicator (b200cd_comm_init / b200cd_allreduce_bucket)
            assert parallel.enable_native_comm() and parallel.native_comm()
        cfg = synthetic_cfg(MTYPE, in_channels=CIN, topology=TOPO)
        torch.manual_seed(7)
        net = networks.create_network(cfg).to(dev).train()
        net.module.set_precision(precision)
        batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
        rows = parallel.shard_rows(B, rank, WORLD)
        x1, x2, y = (batch[k][rows].to(dev) for k in ("x_t1", "x_t2", "y_change"))
        if path == "gather":
            # replicated inputs, as an unchanged reference script under torchrun sees them: the full batch on every rank;
            # the network shards by rank inside forward and returns the gathered full-batch logits
            out = net(batch["x_t1"].to(dev), batch["x_t2"].to(dev))
            assert out.shape[0] == B
            loss = loss_functions.get_criterion("PowerJaccardLoss")(out, batch["y_change"].to(dev))
            loss.backward()
            logits = out.detach().cpu()[rows]
            grads = {n: p.grad.detach().cpu() for n, p in net.module.named_parameters() if p.grad is not None}
        elif path == "fused":
            ts = TrainStep(net.module, rows.stop - rows.start, HW, HW, kind="supervised", device=dev)
            sd0 = {k: v.clone() for k, v in net.state_dict().items()}
            for _ in range(4):          # the same step four times: eager, eager, then graph capture and replay
                net.load_state_dict(sd0)
                loss = ts(x1, x2, y_change=y)
            logits = ts.eng.output_tensors()[0].detach().cpu()
            g = ts.eng.grads
            grads = {n: g.views[n].detach().cpu() for n, _ in g.params if n not in g.skip}
        else:
            out = net(x1, x2)
            loss = loss_functions.get_criterion("PowerJaccardLoss")(out, y)
            loss.backward()
            logits = out.detach().cpu()
            grads = {n: p.grad.detach().cpu() for n, p in net.module.named_parameters() if p.grad is not None}
        torch.cuda.synchronize()
        # numpy arrays are pickled by value: torch tensors travel as shared-memory handles that die with this process
        q.put((rank, loss.item(), logits.numpy(), {k: v.numpy() for k, v in grads.items()}))
        dist.barrier()
    finally:
        parallel.disable_data_parallel()
        dist.destroy_process_group()


def _run_world(path: str, backend: str, precision: str):
    import torch.multiprocessing as mp

    from multimodal_siamese_cd_b200 import parallel
    from oracle import dp_oracle as D
    from oracle import unet_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, path, q, backend, precision)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(WORLD):
        rank, loss, logits, grads = q.get(timeout=300)
        got[rank] = (loss, torch.from_numpy(logits), {k: torch.from_numpy(v) for k, v in grads.items()})
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = O.clone_state(O.reference_state_dict(MTYPE, in_channels=CIN, topology=TOPO, seed=7))
    batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
    ref = D.dp_emulation_step(MTYPE, sd, batch, WORLD, lambda r: parallel.shard_rows(B, r, WORLD))
    for rank in range(WORLD):
        loss, logits, grads = got[rank]
        rows = parallel.shard_rows(B, rank, WORLD)
        prec = precision == "precise"
        assert abs(loss - ref["loss"].item()) <= (1e-5 if prec else 1e-4), (loss, ref["loss"].item())   # the GLOBAL loss on every rank
        rl = ((logits - ref["logits"][rows]).norm() / ref["logits"][rows].norm()).item()
        assert rl <= (1e-4 if prec else 3e-2), rl                  # split-bf16 / single-bf16 storage vs exact fp32
        num = den = 0.0
        worst = 0.0
        for k, r in ref["grads"].items():
            if r is None or k.endswith((".conv.0.bias", ".conv.3.bias")):
                continue
            d, nn_ = (grads[k].double() - r.double()).norm().item(), r.double().norm().item()
            num, den = num + d * d, den + nn_ * nn_
            worst = max(worst, d / max(nn_, 1e-30))
        # SUM-reduced gradients, globally and per parameter: the same envelopes as the single-GPU step tests
        # (tests/test_gpu_e2e.py: fast = bf16 noise floor; precise = the split-storage format floor)
        assert (num / den) ** 0.5 <= (3e-2 if prec else 0.25), (num / den) ** 0.5
        assert worst <= (8e-2 if prec else 0.8), worst
    for k in got[0][2]:
        assert torch.equal(got[0][2][k], got[1][2][k]), k                                    # identical on all ranks


@pytest.mark.parametrize("path", ["fused", "dropin", "gather", "fused_native", "gather_native"])
def test_two_gpu_data_parallel_matches_dataparallel_semantics(path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 CUDA devices")
    _run_world(path, "nccl", "fast")


@pytest.mark.parametrize("path,precision", [("fused", "fast"), ("dropin", "fast"), ("gather", "fast"), ("fused", "precise"),
                                            ("gather", "precise")])
def test_one_gpu_two_ranks_data_parallel(path, precision):
    """The same on a ONE-GPU box: two ranks share cuda:0 and exchange over gloo (CUDA tensors) — the loss-sum exchange,
    the bucketed gradient all-reduce of StepEngine.backward_dp and the scatter / gather of the drop-in 'gather' mode."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run_world(path, "gloo", precision)
