"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_kernels.py
    compute-sanitizer --tool racecheck python tools/sanitize_kernels.py
    compute-sanitizer --tool synccheck python tools/sanitize_kernels.py

Each family runs once through the same per-kernel checks the GPU tests use (tests/gpu_checks*.py), at their smallest
shapes, and the parity result is printed next to the family name; the sanitizer's own summary follows at exit.
B200CD_SANITIZE_ONLY=substring restricts the run."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import gpu_checks as gc  # noqa: E402
import gpu_checks_hp as gch  # noqa: E402

FAMILIES = {
    # CTA-pair implicit-GEMM kernel (cross-CTA mbarriers, multicast commits, TMEM double buffering)
    "pair_conv3x3_stats": lambda: gc.check_conv3x3_cta_stats(n=2, H=16, W=16, cin=64, cout=128, G=2),
    "pair_dgrad_bnbwd": lambda: gc.check_dgrad_bnbwd(2, 16, 16, 64, 128, G=2),
    "pair_convt_fwd": lambda: gc.check_convt_fwd(n=1, h=8, w_=8, c=64),
    "pair_convt_dgrad": lambda: gc.check_convt_dgrad(n=1, h=8, w_=8, c=64),
    "pair_first_conv": lambda: gc.check_first_conv(B=1, H=16, W=16),
    "pair_conv3x3_split_bf16": lambda: gch.check_hp_conv3x3(n=2, H=16, W=16, cin=64, cout=128, G=2),
    "pair_convt_split_bf16": lambda: gch.check_hp_convt(n=1, h=8, w_=8, c=64),
    # weight-gradient kernel
    "wgrad_pos": lambda: gc.check_wgrad3x3(n=1, H=16, W=16, cin=64, cout=128, halo=1, splits=2),
    "wgrad_mstack_64": lambda: gc.check_wgrad3x3(n=1, H=16, W=16, cin=64, cout=64, halo=1, splits=2),
    "wgrad_convt": lambda: gc.check_wgrad_convt(n=1, h=8, w_=8, c=64, splits=2),
    "wgrad_split_bf16": lambda: gch.check_hp_wgrad3x3(n=1, H=16, W=16, cin=64, cout=128, splits=2),
    "wgrad_reduce_batched": lambda: gc.check_wgrad3x3(n=1, H=16, W=16, cin=64, cout=128, halo=1, splits=2, batched=True),
    # memory-bound kernels
    "bn_apply": lambda: gc.check_bn_apply(n=2, H=16, W=16),
    "bn_apply_pool": lambda: gc.check_bn_apply_pool(n=2, H=16, W=16, G=1),
    "bn_bwd_generic": lambda: gc.check_bn_bwd(n=2, H=16, W=16, order=("skip", "pool", "dir")),
    "bn_bwd_window": lambda: gc.check_bn_bwd(n=2, H=16, W=16, order=("pool", "dir")),
    "bn_bwd_head": lambda: gc.check_bn_bwd(n=2, H=16, W=16, order=("dir", "head"), G=1),
    "bn_apply_split_bf16": lambda: gch.check_hp_bn_apply(n=2, H=16, W=16),
    "bn_bwd_split_bf16": lambda: gch.check_hp_bn_bwd(n=2, H=16, W=16),
    "head_colsum": gc.check_head,
    "head_colsum_pad_split_bf16": gch.check_hp_head,
    "stat_rowsum": lambda: gc.check_stat_rowsum(n=1, H=16, W=16),
    "pack": gc.check_pack_weights,
    "pack_split_bf16": gch.check_hp_pack,
    "power_jaccard": lambda: gc.check_pj(B=2, H=16, W=16),
}

if __name__ == "__main__":
    only = os.environ.get("B200CD_SANITIZE_ONLY", "")
    bad = 0
    for name, fn in FAMILIES.items():
        if only and only not in name:
            continue
        t0 = time.time()
        try:
            r = fn()
            ok = bool(r.get("ok"))
        except Exception as e:  # noqa: BLE001
            ok, r = False, {"error": f"{type(e).__name__}: {e}"}
        bad += 0 if ok else 1
        print(f"{name:32s} {'parity ok' if ok else 'PARITY FAIL ' + str(r)[:300]}  ({time.time() - t0:.1f} s)", flush=True)
    print(f"families run: parity failures = {bad}", flush=True)
    sys.exit(1 if bad else 0)
