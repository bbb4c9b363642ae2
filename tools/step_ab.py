"""Device time of the fused training step (CUDA-graph replay, what bench.py's `value` times) under the B200CD_* knobs
of the environment. One JSON line; run it once per setting, alternating settings, on ONE box for an A/B.
    python tools/step_ab.py [config] [precision] [label]"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def build(cfgname: str, precision: str):
    import torch

    import bench
    from multimodal_siamese_cd_b200 import networks
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    mtype, cin, B, kind, alpha, _gf, _yaml = bench.CONFIGS[cfgname]
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    net.module.set_precision(precision)
    ts = TrainStep(net.module, B, 256, 256, kind=kind, alpha=alpha, device=dev, dp_group=None)
    g = torch.Generator(device=dev).manual_seed(7)
    xc = 6 if mtype in bench.TWO_STREAM else cin
    ts.eng.x_t1.copy_(torch.rand(B, xc, 256, 256, device=dev, generator=g))
    ts.eng.x_t2.copy_(torch.rand(B, xc, 256, 256, device=dev, generator=g))
    for t in ts.targets.values():
        t.copy_((torch.rand(t.shape, device=dev, generator=g) > 0.9).float())
    if kind == "mmcr":
        ts.rowmask.copy_(torch.tensor([i % 3 != 2 for i in range(B)], dtype=torch.uint8))
    for _ in range(10):
        ts.run()
    torch.cuda.synchronize()
    return ts, net, B


def timed(ts, n: int) -> float:
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ts.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    import torch

    from multimodal_siamese_cd_b200 import ops, tuning
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    table = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--ab-table=")), None)
    cfgname = args[0] if len(args) > 0 else "dualstream"
    precision = args[1] if len(args) > 1 else "fast"
    label = args[2] if len(args) > 2 else ""
    res = {"label": label, "config": cfgname, "precision": precision,
           "env": {k: v for k, v in os.environ.items() if k.startswith("B200CD_")}}
    if table is None:
        ts, net, B = build(cfgname, precision)
        n = 40 if B <= 16 else 12
        times = [timed(ts, n) for _ in range(4)]
        res.update(ms=[round(t, 4) for t in times], ms_min=round(min(times), 4))
        res["loss"] = float((ts.losses * ts.weights).sum().item())
        res["grad_abs_sum"] = float(ts.eng.grads.flat.double().abs().sum().item())
    else:
        # same-process A/B: one plan built with the library's tile rule, one with the table; measured alternately
        tuning.ENABLED = False
        ts_a, net_a, B = build(cfgname, precision)
        tuning.ENABLED = True
        tuning.load_table(table)
        ts_b, net_b, _ = build(cfgname, precision)
        n = 40 if B <= 16 else 12
        ta, tb = [], []
        for _ in range(4):
            ta.append(timed(ts_a, n))
            tb.append(timed(ts_b, n))
        res.update(base_ms=[round(t, 4) for t in ta], tuned_ms=[round(t, 4) for t in tb], base_min=round(min(ta), 4),
                   tuned_min=round(min(tb), 4), entries=len(tuning.TABLE))
        res["loss"] = [float((t.losses * t.weights).sum().item()) for t in (ts_a, ts_b)]
        ga, gb = ts_a.eng.grads.flat.double(), ts_b.eng.grads.flat.double()
        res["grad_rel_diff"] = float(((ga - gb).norm() / ga.norm()).item())
    ops.device_status()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
