"""Device time of one training step replayed as CUDA graphs (what bench.py times) against the same plan enqueued
eagerly behind a spin kernel (host a whole step ahead), under the engine's stream knobs read from the environment
(B200CD_BRANCH_STREAMS, B200CD_WGRAD_SIDE_STREAM, B200CD_STREAM_PRIO). One JSON line.
    python tools/graph_vs_eager.py [config] [precision]"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    import bench
    from multimodal_siamese_cd_b200 import networks
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    cfgname = sys.argv[1] if len(sys.argv) > 1 else "dualstream"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fast"
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {"config": cfgname, "precision": precision,
           "env": {k: v for k, v in os.environ.items() if k.startswith("B200CD_")}}
    for rep in range(2):
        e0.record()
        for _ in range(40):
            ts.run()
        e1.record()
        torch.cuda.synchronize()
        res[f"graph_ms_{rep}"] = round(e0.elapsed_time(e1) / 40, 4)
        eng = ts.eng
        times = []
        for _ in range(6):
            torch.cuda.synchronize()
            torch.cuda._sleep(int(8e7))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng._run_fwd_eager()
            ts._loss_fwd()
            ts._loss_bwd()
            eng._run_bwd_eager()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        res[f"eager_ms_{rep}"] = round(sorted(times)[len(times) // 2], 4)
        # free-running eager loop (no spin kernel): does the host keep ahead of the device on its own?
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            eng._run_fwd_eager()
            ts._loss_fwd()
            ts._loss_bwd()
            eng._run_bwd_eager()
        e1.record()
        torch.cuda.synchronize()
        res[f"eager_free_ms_{rep}"] = round(e0.elapsed_time(e1) / 20, 4)
    res["loss"] = float((ts.losses * ts.weights).sum().item())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
