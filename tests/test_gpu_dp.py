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


def _worker(rank: int, port: int, path: str, q) -> None:
    import torch.distributed as dist

    from multimodal_siamese_cd_b200 import loss_functions, networks, parallel
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    from oracle import unet_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        parallel.enable_data_parallel()
        cfg = synthetic_cfg(MTYPE, in_channels=CIN, topology=TOPO)
        torch.manual_seed(7)
        net = networks.create_network(cfg).to(dev).train()
        batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
        rows = parallel.shard_rows(B, rank, WORLD)
        x1, x2, y = (batch[k][rows].to(dev) for k in ("x_t1", "x_t2", "y_change"))
        if path == "fused":
            ts = TrainStep(net.module, rows.stop - rows.start, HW, HW, kind="supervised", device=dev)
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
        q.put((rank, loss.item(), logits, grads))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("path", ["fused", "dropin"])
def test_two_gpu_data_parallel_matches_dataparallel_semantics(path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp

    from multimodal_siamese_cd_b200 import parallel
    from oracle import dp_oracle as D
    from oracle import unet_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, path, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(WORLD):
        rank, loss, logits, grads = q.get(timeout=300)
        got[rank] = (loss, logits, grads)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = O.clone_state(O.reference_state_dict(MTYPE, in_channels=CIN, topology=TOPO, seed=7))
    batch = O.synthetic_batch(B, CIN, HW, HW, seed=7)
    ref = D.dp_emulation_step(MTYPE, sd, batch, WORLD, lambda r: parallel.shard_rows(B, r, WORLD))
    for rank in range(WORLD):
        loss, logits, grads = got[rank]
        rows = parallel.shard_rows(B, rank, WORLD)
        assert abs(loss - ref["loss"].item()) <= 1e-4, (loss, ref["loss"].item())          # the GLOBAL loss on every rank
        rl = ((logits - ref["logits"][rows]).norm() / ref["logits"][rows].norm()).item()
        assert rl <= 3e-2, rl                                                                # bf16 storage vs exact fp32
        num = den = 0.0
        for k, r in ref["grads"].items():
            if r is None or k.endswith((".conv.0.bias", ".conv.3.bias")):
                continue
            num += (grads[k].double() - r.double()).norm().item() ** 2
            den += r.double().norm().item() ** 2
        assert (num / den) ** 0.5 <= 0.35, (num / den) ** 0.5                               # SUM-reduced, bf16 noise floor
    for k in got[0][2]:
        assert torch.equal(got[0][2][k], got[1][2][k]), k                                    # identical on all ranks
