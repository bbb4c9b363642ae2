"""End-to-end parity + first step timings on the GPU (development aid; results go to gpurun_out/e2e.json)."""
from __future__ import annotations

import argparse
import json
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import e2e_checks as E  # noqa: E402
from multimodal_siamese_cd_b200 import networks, ops  # noqa: E402
from multimodal_siamese_cd_b200.config import synthetic_cfg  # noqa: E402
from multimodal_siamese_cd_b200.step import TrainStep  # noqa: E402

SMALL = (64, 128)
FULL = (64, 128, 256, 512)
CASES = [
    ("siamese_small_dropin", dict(mtype="siameseunet", cin=4, topo=SMALL, B=3, H=32, W=32, kind="supervised")),
    ("siamese_small_fused", dict(mtype="siameseunet", cin=4, topo=SMALL, B=3, H=32, W=32, kind="supervised", path="fused", steps=3)),
    ("unet_small_dropin", dict(mtype="unet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="supervised")),
    ("dualstream_small_dropin", dict(mtype="dualstreamunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="supervised")),
    ("dtsiamese_small_dropin", dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="dualtask")),
    ("dtsiamese_small_fused", dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="dualtask", path="fused", steps=3)),
    ("whatevernet_small_dropin", dict(mtype="whatevernet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr")),
    ("whatevernet_small_fused", dict(mtype="whatevernet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr", path="fused", steps=3)),
    ("whatevernet2_small_dropin", dict(mtype="whatevernet2", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr")),
    ("siamese_full_64", dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=64, W=64, kind="supervised")),
    ("siamese_full_256", dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", path="fused", steps=3)),
    ("siamese_full_256_corr", dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", path="fused", corr=True)),
    ("dtsiamese_full_128", dict(mtype="dtsiameseunet", cin=6, topo=FULL, B=2, H=128, W=128, kind="dualtask", path="fused")),
]


def timing(mtype, cin, B, kind, steps=10):
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(7)
    net = networks.create_network(cfg).to(dev).train()
    ts = TrainStep(net.module, B, 256, 256, kind=kind, device=dev, dp_group=None)
    for _ in range(4):
        ts.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ts.run()
    e1.record()
    torch.cuda.synchronize()
    ops.device_status(0)
    ms = e0.elapsed_time(e1) / steps
    # forward / backward split
    e0.record()
    for _ in range(steps):
        ts.eng.forward_static()
    e1.record()
    torch.cuda.synchronize()
    fms = e0.elapsed_time(e1) / steps
    out = {"ms_per_step": ms, "fwd_ms": fms, "pairs_per_s": B / ms * 1e3, "mem_gb": ts.eng.mem_bytes / 2 ** 30,
           "launches": ts.eng.launches_per_step()}
    net.module.release_engines()
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--timing", action="store_true")
    args = ap.parse_args()
    only = [n for n in args.only.split(",") if n]
    report = {"cases": {}, "timing": {}}
    for name, kw in CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        try:
            res = E.run_case(**kw)
        except Exception as e:  # noqa: BLE001
            res = {"exception": repr(e), "trace": traceback.format_exc()[-2500:]}
        res["seconds"] = round(time.time() - t0, 2)
        report["cases"][name] = res
        print(json.dumps({name: res}, default=str), flush=True)
    if args.timing:
        for name, kw in (("siamese_b8", dict(mtype="siameseunet", cin=4, B=8, kind="supervised")),
                         ("siamese_b32", dict(mtype="siameseunet", cin=4, B=32, kind="supervised")),
                         ("dualstream_b16", dict(mtype="dualstreamunet", cin=6, B=16, kind="supervised")),
                         ("dtsiamese_b8", dict(mtype="dtsiameseunet", cin=6, B=8, kind="dualtask"))):
            try:
                res = timing(**kw)
            except Exception as e:  # noqa: BLE001
                res = {"exception": repr(e), "trace": traceback.format_exc()[-2500:]}
            report["timing"][name] = res
            print(json.dumps({name: res}, default=str), flush=True)
    out = ROOT / "gpurun_out" / "e2e.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps(report, indent=1, default=str))
    return 0


if __name__ == "__main__":
    sys.exit(main())
