"""Device-side timeline of one data-parallel training step (one process per GPU, launched with torch.distributed.run):
where the two exchanges of the step sit relative to the compute stream, and how much of the gradient all-reduce is
exposed (runs after the last backward kernel has finished).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/dp_timeline.py [config] > gpurun_out/dp_timeline_nN.json

The step runs eagerly (no graphs) behind a spin kernel, so the host is a whole step ahead of the device and the CUDA
events — recorded on the stream each piece runs on — measure device time only. Every rank reports its own events;
rank 0 prints one JSON object with, per label, the max over ranks of the offset from the step's first kernel, plus
  exposed_ms   = end of the last all-reduce - end of the last backward compute segment (what the overlap did not hide)
  comm_busy_ms = sum of the all-reduce intervals on the communication stream
  graph_ms     = the same step as it is timed by bench.py (one CUDA graph per step), for scale
and the single-GPU step time of the same box when run with N = 1.
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch
    import torch.distributed as dist

    import bench
    from multimodal_siamese_cd_b200 import networks, parallel
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    cfgname = sys.argv[1] if len(sys.argv) > 1 else "dualstream"
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    out = bench._claim_stdout()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        parallel.enable_data_parallel()
        if os.environ.get("B200CD_NATIVE_COMM", "1") != "0":
            parallel.enable_native_comm()
    mtype, cin, B, kind, alpha, _gf, _yaml = bench.CONFIGS[cfgname]
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    ts = TrainStep(net.module, B, 256, 256, kind=kind, alpha=alpha, device=dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    xc = 6 if mtype in bench.TWO_STREAM else cin
    ts.eng.x_t1.copy_(torch.rand(B, xc, 256, 256, device=dev, generator=g))
    ts.eng.x_t2.copy_(torch.rand(B, xc, 256, 256, device=dev, generator=g))
    for t in ts.targets.values():
        t.copy_((torch.rand(t.shape, device=dev, generator=g) > 0.9).float())
    if kind == "mmcr":
        ts.rowmask.copy_(torch.tensor([i % 3 != 2 for i in range(B)], dtype=torch.uint8))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the step as bench.py times it
    for _ in range(8):
        ts.run()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        ts.run()
    e1.record()
    sync()
    graph_ms = e0.elapsed_time(e1) / 30

    eng = ts.eng

    def ev(stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream or torch.cuda.current_stream())
        return e

    best = None
    for _ in range(3):          # three eager passes; the last one is reported
        sync()
        torch.cuda._sleep(int(8e7))
        tl = [("step_start", ev())]
        eng._run_fwd_eager()
        ts._loss_fwd()
        tl.append(("loss_sums_ready", ev()))
        if world > 1:
            ts._allreduce_sums()
        tl.append(("loss_sums_reduced", ev()))
        ts._loss_bwd()
        tl.append(("backward_start", ev()))
        if world > 1:
            eng.backward_dp(ts.dp, ts.grad_buckets, inner_graphs=False, timeline=tl)
        else:
            eng._run_bwd_eager()
        tl.append(("step_end", ev()))
        torch.cuda.synchronize()
        t0 = tl[0][1]
        best = [(name, t0.elapsed_time(e)) for name, e in tl]
    names = [n for n, _ in best]
    vals = torch.tensor([v for _, v in best], device=dev, dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        per_rank = [[round(x, 4) for x in v.tolist()] for v in allv]
        mx = torch.stack(allv).max(0).values.tolist()
        d = dict(zip(names, mx))
        segs = [k for k in names if k.startswith("seg")]
        ars = sorted({k.split("_")[0] for k in names if k.startswith("ar")})
        res = {"config": cfgname, "n_gpus": world, "batch_per_gpu": B, "graph_ms_per_step": round(graph_ms, 4),
               "eager_step_ms": round(d["step_end"], 4), "labels": names,
               "ms_from_step_start_max_over_ranks": [round(x, 4) for x in mx], "per_rank": per_rank,
               "native_comm": parallel.native_comm(), "grad_buckets": ts.grad_buckets,
               "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
        if world > 1 and segs and ars:
            last_seg = d[segs[-1]]
            starts = {a: next(v for k, v in d.items() if k.startswith(a + "_start")) for a in ars}
            ends = {a: d[a + "_end"] for a in ars}
            res["loss_sum_exchange_ms"] = round(d["loss_sums_reduced"] - d["loss_sums_ready"], 4)
            res["comm_busy_ms"] = round(sum(ends[a] - starts[a] for a in ars), 4)
            res["exposed_ms"] = round(max(ends.values()) - last_seg, 4)
            res["allreduce_ms"] = {a: round(ends[a] - starts[a], 4) for a in ars}
        print(json.dumps(res), file=out, flush=True)
    if world > 1:
        del ts
        parallel.disable_data_parallel()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
