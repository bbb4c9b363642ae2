#!/usr/bin/env python
"""Benchmark of the training-step hot path (fwd + loss + bwd) on synthetic 256x256 S1+S2 patch pairs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config dualstream|siamese|dtsiamese|dtsiamese_ssl|mmcr]
  python bench.py --impl reference ...   # the reference algorithm's CPU path (oracle port) on the host cores

One JSON line on stdout (rank 0). `value` = whole-job patch-pairs/s with inputs resident in HBM (fused TrainStep,
CUDA-graph replay); `e2e` = the same metric through the reference-facing drop-in modules
(net(x_t1, x_t2) -> criterion -> loss.backward()) with pinned HOST inputs copied every step and loss.item() read back.
For N > 1 launch with torch.distributed.run (one process per GPU, NCCL); per-GPU batch is fixed (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "train patch-pairs/s (fwd+bwd, 256x256 S1+S2)"
UNIT = "patch-pairs/s"

# BASELINE.json configs -> (model type, in_channels, per-GPU batch, step kind, alpha, GF per pair fwd+bwd [BASELINE.md §3])
CONFIGS = {
    "siamese": ("siameseunet", 4, 8, "supervised", 0.5, 279.47, "baseline_siamese.yaml"),
    "dualstream": ("dualstreamunet", 6, 16, "supervised", 0.5, 384.38, "baseline_dualstream.yaml TRAINER.BATCH_SIZE 16"),
    "dtsiamese": ("dtsiameseunet", 6, 8, "dualtask", 0.5, 488.70, "dtsiamese.yaml MODEL.IN_CHANNELS 6"),
    "dtsiamese_ssl": ("dtsiameseunet", 6, 8, "mmcr", 0.1, 488.70, "dtsiamese_ssl.yaml MODEL.IN_CHANNELS 6"),
    "mmcr": ("whatevernet", 6, 64, "mmcr", 0.5, 558.38, "siamese_mmcr_alpha0500_16batch.yaml TRAINER.BATCH_SIZE 64/GPU"),
}
H = W = 256


def ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/): mean over the captured launches. None when no capture is committed."""
    import csv
    for name in ("r01_ncu_full_fprop_pair_kernel.csv",):
        f = ROOT / "profiles" / name
        if not f.exists():
            continue
        rows = list(csv.reader(open(f)))
        hdr = rows[0]
        try:
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[2:]:
            if len(r) > max(ir, iw) and kernel_substr in r[ik]:
                try:
                    vals.append(float(r[ir]) * unit.get(rows[1][ir], 1.0) + float(r[iw]) * unit.get(rows[1][iw], 1.0))
                except ValueError:
                    pass
        if vals:
            return {"bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals), "source": f"profiles/{name}"}
    return None


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tensor": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"], "src": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in Path(self.f.name).read_text().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for nm, v in zip(names, r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        sm_sorted = sorted(sm)
        busy = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
def cpu_step_rate(cfgname: str, sample_pairs: int, steps: int, warmup: int, threads=None) -> dict:
    """The reference algorithm on the host cores: oracle port (pinned against the reference by tests/golden) running
    zero_grad -> forward -> loss -> backward in fp32 on `sample_pairs` patch pairs per step."""
    import torch

    from oracle import unet_oracle as O
    mtype, cin, _, kind, alpha, _, _ = CONFIGS[cfgname]
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd0 = O.reference_state_dict(mtype, in_channels=cin, seed=7)
    batch = O.synthetic_batch(max(sample_pairs, 3 if kind == "mmcr" else 1), 6 if cin == 6 else cin, H, W, seed=7)
    n = batch["x_t1"].shape[0]
    times = []
    for i in range(warmup + steps):
        sd = O.clone_state(sd0)
        t0 = time.perf_counter()
        O.train_step(mtype, sd, batch, kind=kind, alpha=alpha, q=False)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": n / med, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} patch pairs per step, {steps} timed steps after {warmup} warm-up, fp32 torch CPU kernels, "
                      f"median step {med:.2f} s", "sec_per_step": med, "pairs": n}


def run_reference_arm(args, out) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mtype, cin, B, kind, alpha, gf, yaml = CONFIGS[args.config]
    steps, warm = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    r = cpu_step_rate(args.config, sample_pairs=8, steps=steps, warmup=warm)   # bounded sample: 8 pairs per step
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{yaml}: {mtype}, fwd+loss+bwd ({kind}), 256x256, S1+S2; CPU sample of {r['pairs']} "
                               f"pairs per step (GPU arm: {B} pairs per GPU per step)"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


# --------------------------------------------------------------------------------------------------------------
def profile_step(ts, steps: int = 2) -> dict:
    """Eager (no graph) passes with CUDA events around every launch: per-kernel-family time, algorithmic FLOPs/bytes."""
    import torch

    from multimodal_siamese_cd_b200 import ops
    eng = ts.eng
    eng.branch_streams = eng.wgrad_side = False   # one stream: every launch is timed on its own
    fam = {}
    l0 = ops.LAUNCHES
    for it in range(steps + 1):
        ops.PROFILE = [] if it > 0 else None
        eng._run_fwd_eager()
        ts._loss_fwd()
        ts._loss_bwd()
        eng._run_bwd_eager()
        torch.cuda.synchronize()
        if ops.PROFILE:
            if it == steps and os.environ.get("B200CD_DUMP_CALLS"):
                with open(os.environ["B200CD_DUMP_CALLS"], "w") as f:
                    for name, fl, by, e0, e1, tag in ops.PROFILE:
                        ms_ = e0.elapsed_time(e1)
                        f.write(f"{name:14s} {ms_ * 1e3:9.1f} us  {fl / ms_ / 1e9 if fl else 0:8.1f} TF  "
                                f"{by / ms_ / 1e6:8.1f} GB/s  {tag}\n")
            for name, fl, by, e0, e1, _tag in ops.PROFILE:
                d = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "calls": 0})
                d["ms"] += e0.elapsed_time(e1)
                d["flops"] += fl
                d["bytes"] += by
                d["calls"] += 1
    ops.PROFILE = None
    fam["_launches_per_step"] = (ops.LAUNCHES - l0) // (steps + 1)
    for d in fam.values():
        if not isinstance(d, dict):
            continue
        for k in ("ms", "flops", "bytes"):
            d[k] /= steps
        d["calls"] //= steps
    return fam


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else a library prints there (NCCL's version banner, ...) goes to
    stderr. Returns a file object on the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main() -> None:
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="dualstream", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args, out)
        return

    import torch
    import torch.distributed as dist

    from multimodal_siamese_cd_b200 import loss_functions, networks, ops, parallel
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        parallel.enable_data_parallel()
    W_ = max(3, args.warmup)
    K = max(1, args.steps)

    mtype, cin, B, kind, alpha, gf_pair, yaml = CONFIGS[args.config]
    if args.batch:
        B = args.batch
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin

    # synthetic batch, rank r uses seed 7 + r; generated on the host, pinned
    g = torch.Generator().manual_seed(7 + rank)
    host = {
        "x_t1": torch.rand(B, xc, H, W, generator=g).pin_memory(),
        "x_t2": torch.rand(B, xc, H, W, generator=g).pin_memory(),
        "y_change": (torch.rand(B, 1, H, W, generator=g) > 0.9).float().pin_memory(),
        "y_sem_t1": (torch.rand(B, 1, H, W, generator=g) > 0.8).float().pin_memory(),
        "y_sem_t2": (torch.rand(B, 1, H, W, generator=g) > 0.8).float().pin_memory(),
    }
    is_labeled = torch.tensor([i % 3 != 2 for i in range(B)])

    ts = TrainStep(net.module, B, H, W, kind=kind, alpha=alpha, device=dev)
    tg = {k: host[k] for k in ts.targets}
    ts.set_inputs(host["x_t1"], host["x_t2"], is_labeled=is_labeled if kind == "mmcr" else None, **tg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(W_):
        ts.run()
    l0 = ops.LAUNCHES
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = ts.run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    clocks = sampler.stop() if sampler else None
    ops.device_status(local)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    loss_val = float(loss.item())

    # ---- end to end through the reference-facing modules ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        crit = loss_functions.get_criterion("PowerJaccardLoss")

        from multimodal_siamese_cd_b200.data import DevicePrefetcher, LossReader

        def e2e_step(b, reader):
            # the body of the reference loop (train_supervised.py:63-79 and its dual-task / semi-supervised variants)
            for p in net.parameters():
                p.grad = None
            outs = net(b["x_t1"], b["x_t2"])
            if kind == "supervised":
                loss = crit(outs, b["y_change"])
            elif kind == "dualtask":
                c, s1, s2 = outs
                loss = (crit(c, b["y_change"]) + (crit(s1, b["y_sem_t1"]) + crit(s2, b["y_sem_t2"])) / 2) / 2
            else:
                f, s1, s2 = outs
                y = b["y_change"]
                lab = b["is_labeled"]
                p2 = torch.sigmoid(s2)
                loss = alpha * (crit(f[lab,], y[lab,]) + crit(s1[lab,], y[lab,]) + crit(s2[lab,], y[lab,])) / 3 + \
                    (1 - alpha) * crit(s1[~lab,], p2[~lab,])
            loss.backward()
            return reader.push(loss)  # device -> host read of every step's loss (train_supervised.py:79), one step late

        keys = {"supervised": ["y_change"], "dualtask": ["y_change", "y_sem_t1", "y_sem_t2"], "mmcr": ["y_change"]}[kind]
        host_batch = {"x_t1": host["x_t1"], "x_t2": host["x_t2"], "is_labeled": is_labeled, **{k: host[k] for k in keys}}
        if os.environ.get("B200CD_DEBUG_E2E_RESIDENT") == "1":   # measurement aid: what the PCIe staging costs (invalid e2e)
            host_batch = {k: (v.to(dev) if torch.is_tensor(v) and k != "is_labeled" else v) for k, v in host_batch.items()}

        def host_batches(n):           # the pinned host batch, staged host -> device again for every step
            for _ in range(n):
                yield host_batch

        # one prefetcher / reader for warm-up and timed region: staging buffers, pinned slots, engine buffers and CUDA
        # graphs all exist before the clock starts (at least 10 warm-up steps: a cold caching allocator was seen to
        # stall one early step by ~0.4 s, which a 30-step measurement does not average out)
        reader = LossReader(dev)
        pf = DevicePrefetcher(host_batches(max(W_, 10)), dev)
        for b in pf:
            e2e_step(b, reader)
        reader.drain()
        barrier()
        pf.batches = host_batches(K)
        pf.bytes_staged = 0
        reader.bytes_read = 0
        e2e_losses = []
        e0.record()
        dbg_t = [time.perf_counter()] if os.environ.get("B200CD_DEBUG_E2E_TIMES") == "1" else None
        for b in pf:
            v = e2e_step(b, reader)
            if v is not None:
                e2e_losses.append(v)
            if dbg_t is not None:
                dbg_t.append(time.perf_counter())
        if dbg_t:
            sys.stderr.write("e2e host ms per step: %s\n" % [round((b_ - a_) * 1e3, 1) for a_, b_ in zip(dbg_t, dbg_t[1:])])
        e2e_losses += reader.drain()
        if dbg_t:
            sys.stderr.write("e2e drain done after %.1f ms\n" % ((time.perf_counter() - dbg_t[-1]) * 1e3))
        e1.record()
        barrier()
        assert len(e2e_losses) == K and all(x == x for x in e2e_losses), "every step's loss must have been read back"
        ems = e0.elapsed_time(e1) / K
        t = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = t.item()
        h2d = pf.bytes_staged // K      # counted from the tensors copied host -> device in the timed region
        e2e = {"value": B * world / ems * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": reader.bytes_read // K, "ms_per_step": ems, "last_loss": e2e_losses[-1],
               "api": "for batch in data.DevicePrefetcher(loader, device): net = networks.create_network(cfg); "
                      "loss = get_criterion('PowerJaccardLoss')(net(x_t1, x_t2), y); loss.backward(); "
                      "data.LossReader.push(loss)  # pinned-host inputs staged one step ahead on a side stream, every "
                      "step's loss copied to pinned host memory and read one step later"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + per-family breakdown (rank 0, eager with events) ----------------------
    pk = peaks()
    fam = profile_step(ts)
    gpu_launches = fam.pop("_launches_per_step") * K   # our own kernels inside the timed region (graph-replayed)
    total_ms = sum(d["ms"] for d in fam.values())
    tensor_fams = ("fprop3x3", "dgrad3x3_bnbwd", "wgrad", "gemm1tap", "convT_dgrad", "convT_dgrad_bnbwd")
    dom = max(fam, key=lambda k: fam[k]["ms"])
    d = fam[dom]
    if dom in tensor_fams:
        ach = d["flops"] / d["ms"] / 1e9
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"]}
    else:
        ach = d["bytes"] / d["ms"] / 1e6
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    tr = ncu_traffic("fprop_pair_kernel") if dom == "fprop3x3" else None
    roof.update({"kernel": dom, "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr,
                 "share_of_step": d["ms"] / total_ms, "peak_source": pk["src"] +
                 (" bf16_tflops_sustained" if roof["bound"] == "tensor" else " hbm_gbs"),
                 "launches_per_step": d["calls"], "ms_per_step": d["ms"]})
    breakdown = {}
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        b = {"ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4), "calls": v["calls"]}
        if k in tensor_fams:
            b["tflops"] = round(v["flops"] / v["ms"] / 1e9, 1)
        else:
            b["gbs"] = round(v["bytes"] / v["ms"] / 1e6, 1)
        breakdown[k] = b

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # bounded sample of the same workload: 8 pairs per step, 1 warm-up + 4 timed steps (~10-15 s of host work)
        r = cpu_step_rate(args.config, sample_pairs=8, steps=4, warmup=1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    value = B * world / ms * 1e3
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"{yaml}: {mtype}, fwd + {kind} power-Jaccard loss + bwd, 256x256, S1 2-band + S2 4-band, "
                        f"{B} patch pairs per GPU",
            "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": f"no flush: one step touches {ts.eng.mem_bytes / 2**30:.1f} GiB of activations/gradients per GPU (>> 126 MB L2)",
            "precision": "bf16 storage and MMA operands, fp32 accumulation / BatchNorm / loss / parameter gradients",
        },
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": gpu_launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernel_breakdown": breakdown,
        "step_roofline": {"gflop_per_pair": gf_pair, "achieved_tflops_per_gpu": value / world * gf_pair / 1e3,
                          "frac_of_tensor_peak": value / world * gf_pair / 1e3 / pk["tensor"]},
        "loss": loss_val,
    }
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
