"""Run every per-kernel check (tests/gpu_checks.py) on the GPU, never stop at the first failure, and write a
JSON report to gpurun_out/. Development aid for a box without a local GPU: one gpurun call = one full picture.

usage: python tools/gpu_probe.py [--only name,name] [--perf]
"""
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

import gpu_checks as gc  # noqa: E402
from multimodal_siamese_cd_b200 import ops  # noqa: E402


def time_fn(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def perf() -> dict:
    """First timing of the two tensor-core kernels on the reference's layer shapes (batch 16 image pairs)."""
    out = {}
    dev = "cuda"
    shapes = [  # (name, n_img, H, cin, cout)
        ("inc2_64_64_256", 32, 256, 64, 64),
        ("down1b_128_128_128", 32, 128, 128, 128),
        ("down2b_256_256_64", 32, 64, 256, 256),
        ("down3b_512_512_32", 32, 32, 512, 512),
        ("down4_512_512_16", 32, 16, 512, 512),
        ("up1a_128_64_256", 16, 256, 128, 64),
        ("up4a_1024_256_32", 16, 32, 1024, 256),
        ("up2a_256_64_128", 16, 128, 256, 64),
        ("down1a_64_128_128", 32, 128, 64, 128),
    ]
    for name, n, H, cin, cout in shapes:
        try:
            A = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
            w = torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5)
            Bw = ops.pack_weights(0, w)
            o = torch.empty(n, H, H, cout, device=dev, dtype=torch.bfloat16)
            tiles = ops.conv_gemm_tiles(H, H)
            stats = torch.empty(n * tiles, cout, 2, device=dev)
            flops = 2.0 * n * H * H * cout * cin * 9
            rec = {}
            for halo, wide in ((False, False), (True, False), (False, True)):
                ms = time_fn(lambda: ops.conv_gemm(0, 0, A, Bw, o, stats=stats, halo=halo, wide=wide))
                rec[f"fprop_halo{int(halo)}_wide{int(wide)}_tflops"] = flops / ms / 1e9
            ms = time_fn(lambda: ops.conv_gemm(0, 0, A, Bw, o, stats=stats, pair=True))
            rec["fprop_pair_tflops"] = flops / ms / 1e9
            rec["fprop_pair_us"] = ms * 1e3
            if "--fprop-only" in sys.argv:
                out[name] = rec
                print(json.dumps({name: rec}), flush=True)
                continue
            dr = torch.randn(n, H, H, cout, device=dev).to(torch.bfloat16)
            total = ops.wgrad_tiles(n, H, H)
            for halo in (0, 1):
                if cout >= 128:
                    mt, nt = cout // 128, max(1, cin // 128)
                else:
                    mt, nt = max(1, cin // 128), 1
                splits = max(1, min(total, (148 * 2) // (mt * nt * 3)))
                ws = torch.empty(splits, 9, cout, cin, device=dev)
                if cout >= 128:
                    f = lambda: ops.wgrad_gemm(0, 1, halo, dr, A, ws, splits, 9 * cout * cin, cout * cin, cin, 1)  # noqa: E731
                else:
                    f = lambda: ops.wgrad_gemm(0, -1, halo, A, dr, ws, splits, 9 * cout * cin, cout * cin, 1, cin)  # noqa: E731
                ms = time_fn(f)
                rec[f"wgrad_halo{halo}_ms"] = ms
                rec[f"wgrad_halo{halo}_tflops"] = flops / ms / 1e9
                rec["wgrad_splits"] = splits
            ops.device_status()
            out[name] = rec
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)}
        print(json.dumps({name: out[name]}), flush=True)
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--fprop-only", action="store_true")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "probe.json"))
    args = ap.parse_args()
    names = [n for n in args.only.split(",") if n] or list(gc.ALL_CHECKS)
    report = {"device": torch.cuda.get_device_name(0), "checks": {}}
    failed = []
    for name in names:
        t0 = time.time()
        try:
            res = gc.ALL_CHECKS[name]()
        except Exception as e:  # noqa: BLE001
            res = {"ok": False, "exception": repr(e), "trace": traceback.format_exc()[-1500:]}
            try:
                torch.cuda.synchronize()
            except Exception as e2:  # noqa: BLE001
                res["sync_after"] = repr(e2)
        res["seconds"] = round(time.time() - t0, 3)
        report["checks"][name] = res
        if not res.get("ok"):
            failed.append(name)
        print(json.dumps({name: res}, default=str), flush=True)
    if any(n.startswith(("conv", "convt")) for n in failed):
        try:
            report["decode_fprop"] = gc.decode_fprop()
        except Exception as e:  # noqa: BLE001
            report["decode_fprop"] = {"exception": repr(e)}
        print(json.dumps({"decode_fprop": report["decode_fprop"]}, default=str), flush=True)
    if any(n.startswith("wgrad") for n in failed):
        try:
            report["decode_wgrad"] = gc.decode_wgrad()
        except Exception as e:  # noqa: BLE001
            report["decode_wgrad"] = {"exception": repr(e)}
        print(json.dumps({"decode_wgrad": report["decode_wgrad"]}, default=str), flush=True)
    if args.perf:
        report["perf"] = perf()
    report["failed"] = failed
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(report, indent=1, default=str))
    print("FAILED:", failed, flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
