"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_kernels.py
    compute-sanitizer --tool racecheck python tools/sanitize_kernels.py
    compute-sanitizer --tool synccheck python tools/sanitize_kernels.py

Each family runs once through the same per-kernel checks the GPU tests use (tests/gpu_checks*.py), at their smallest
shapes, and the parity result is printed next to the family name; the sanitizer's own summary follows at exit.
B200CD_SANITIZE_ONLY=substring restricts the run.

Where compute-sanitizer is not available (it is closed on the pool this repo is measured on):

    python tools/sanitize_kernels.py --guard

runs the same families — plus one whole training step and one odd-sized inference per network family — on a RED-ZONE
allocator (tools/guard_alloc.cpp, compiled here with g++): every torch allocation is its own cudaMalloc between two
4 KiB bands of 0xA5 that are verified after each family and at every free, so a kernel writing before or past any tensor
it was handed is reported. A deliberate 4-byte overrun at the start proves the detector fires."""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GUARD = None
if "--guard" in sys.argv:
    import ctypes

    import torch
    os.environ["B200CD_CUDA_GRAPHS"] = "0"   # graph capture needs the caching allocator's private pools
    so = ROOT / "gpurun_out" / "guard_alloc.so"
    so.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", str(ROOT / "tools" / "guard_alloc.cpp"), "-I/usr/local/cuda/include",
                    "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(so)], check=True)
    torch.cuda.memory.change_current_allocator(
        torch.cuda.memory.CUDAPluggableAllocator(str(so), "guard_malloc", "guard_free"))
    GUARD = ctypes.CDLL(str(so))
    GUARD.guard_check_all.restype = GUARD.guard_allocs.restype = GUARD.guard_blocks_checked.restype = ctypes.c_longlong

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
    "power_jaccard": lambda: gc.check_pj(B=6, H=16, W=16),
}

def _whole_steps() -> dict:
    """Guard mode only: whole training steps (eager) and odd-sized inference through the engine —
    the plan's own buffer arithmetic (concat slices, workspaces, arg-max indices) under the red-zone allocator."""
    import e2e_checks as E
    small = (64, 128)

    def step(**kw):
        r = E.run_case(topo=small, B=3, H=32, W=32, skip_q=True, graphs=False, **kw)
        return {"ok": r["loss_x"] <= 1e-4, **{k: r[k] for k in ("loss_x", "logits_x")}}

    def evalcase(mtype, cin, H, W):
        r = E.run_eval_case(mtype, cin, small, 1, H, W, warm_hw=(32, 32))
        return {"ok": r["logits_q"] <= 1.5e-2, "logits_q": r["logits_q"]}

    return {
        "step_siamese_fused": lambda: step(mtype="siameseunet", cin=4, kind="supervised", path="fused", steps=2),
        "step_dualstream_dropin": lambda: step(mtype="dualstreamunet", cin=6, kind="supervised"),
        "step_dtsiamese_dualtask": lambda: step(mtype="dtsiameseunet", cin=6, kind="dualtask", path="fused"),
        "step_whatevernet_mmcr": lambda: step(mtype="whatevernet", cin=6, kind="mmcr", path="fused"),
        "step_siamese_precise": lambda: step(mtype="siameseunet", cin=4, kind="supervised", path="fused", precision="precise"),
        "eval_siamese_odd_35x50": lambda: evalcase("siameseunet", 4, 35, 50),
        "eval_dualstream_odd_41x41": lambda: evalcase("dualstreamunet", 6, 41, 41),
    }


def _prove_detector() -> bool:
    """A 4-byte write just past a 1000-byte tensor must be reported (and is subtracted from the final count)."""
    import ctypes

    import torch
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rt.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
    t = torch.empty(1000, dtype=torch.uint8, device="cuda")
    before = GUARD.guard_check_all()
    rt.cudaMemset(t.data_ptr() + 1000, 0, 4)
    after = GUARD.guard_check_all()
    rt.cudaMemset(t.data_ptr() + 1000, 0xA5, 4)      # repair the band so the block frees cleanly
    return after == before + 1


if __name__ == "__main__":
    only = os.environ.get("B200CD_SANITIZE_ONLY", "")
    bad = 0
    base_viol = 0
    if GUARD is not None:
        fired = _prove_detector()
        base_viol = GUARD.guard_check_all()
        print(f"red-zone allocator active; deliberate 4-byte overrun detected: {fired}", flush=True)
        bad += 0 if fired else 1
        FAMILIES.update(_whole_steps())
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
        extra = ""
        if GUARD is not None:
            v = GUARD.guard_check_all() - base_viol
            extra = f"  red zones damaged so far: {v}"
        print(f"{name:32s} {'parity ok' if ok else 'PARITY FAIL ' + str(r)[:300]}  ({time.time() - t0:.1f} s){extra}", flush=True)
    print(f"families run: parity failures = {bad}", flush=True)
    if GUARD is not None:
        import gc as _gc
        _gc.collect()
        v = GUARD.guard_check_all() - base_viol
        print(f"red-zone allocator: {GUARD.guard_allocs()} allocations, {GUARD.guard_blocks_checked()} block checks, "
              f"{v} damaged bands (excluding the deliberate one)", flush=True)
        bad += 1 if v else 0
    sys.exit(1 if bad else 0)
