"""Timing of the memory-bound kernels on the bench workload's layer shapes (CUDA events, tensors larger than L2 or
rotated so that nothing is served from L2). usage: python tools/mem_probe.py [names,...]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from multimodal_siamese_cd_b200 import ops  # noqa: E402

DEV = "cuda"
only = set(sys.argv[1].split(",")) if len(sys.argv) > 1 and sys.argv[1] else None


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def report(name, us, nbytes):
    print(json.dumps({"name": name, "us": round(us, 1), "GBs": round(nbytes / us / 1e3, 1)}), flush=True)


def bf(*shape):
    return torch.randn(*shape, device=DEV).to(torch.bfloat16)


def bn_consts(G, C):
    return (torch.randn(G, C, device=DEV) * 0.1, torch.rand(G, C, device=DEV) + 0.5, torch.rand(G, C, device=DEV) + 0.5,
            torch.randn(G, C, device=DEV) * 0.1)


def run_bn_bwd(tag, n, H, C, kinds, G=1):
    r = bf(n, H, H, C)
    mean, invstd, scale, shift = bn_consts(G, C)
    srcs = []
    nb = 2.0 * r.numel() * 2 + r.numel() * 2  # r read twice, dr written
    for k in kinds:
        if k == 1:
            t = bf(n, H, H, C)
            srcs.append({"kind": 1, "t": t})
            nb += 2.0 * t.numel() * 2
        elif k == 2:
            t = bf(n, H // 2, H // 2, C)
            idx = torch.randint(0, 4, (n, H // 2, H // 2, C), device=DEV, dtype=torch.uint8)
            srcs.append({"kind": 2, "t": t, "w": idx})
            nb += 2.0 * (t.numel() * 2 + idx.numel())
        elif k == 3:
            dz = torch.randn(n, 1, H, H, device=DEV)
            w = torch.randn(C, device=DEV)
            srcs.append({"kind": 3, "t": dz, "w": w})
            nb += 2.0 * dz.numel() * 4
    arr = ops.make_srcs(srcs)
    ws = torch.empty(ops.bn_bwd_ws_floats(n, H, H, C, G), device=DEV)
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dr = torch.empty_like(r)
    us = timeit(lambda: ops.bn_bwd(r, mean, invstd, scale, shift, arr, G, ws, dg, db, dr))
    report(f"bn_bwd_{tag}_{n}x{H}x{C}_{kinds}", us, nb)


def run_bn_apply(tag, n, H, C, pool, diff=False, G=1):
    r = bf(n, H, H, C)
    _, _, scale, shift = bn_consts(G, C)
    a = torch.empty_like(r) if not diff else None
    nb = r.numel() * 2.0
    kw = {}
    if a is not None:
        kw["a"] = a
        nb += a.numel() * 2
    if pool:
        kw["pool"] = torch.empty(n, H // 2, H // 2, C, device=DEV, dtype=torch.bfloat16)
        kw["pool_idx"] = torch.empty(n, H // 2, H // 2, C, device=DEV, dtype=torch.uint8)
        nb += kw["pool"].numel() * 3
    if diff:
        kw["dif"] = torch.empty(n // 2, H, H, C, device=DEV, dtype=torch.bfloat16)
        nb += kw["dif"].numel() * 2
    us = timeit(lambda: ops.bn_apply(r, scale, shift, G, diff, **kw))
    report(f"bn_apply_{tag}_{n}x{H}x{C}_pool{int(pool)}_diff{int(diff)}", us, nb)


def run_wgrad_reduce(cout, cin, splits):
    ws = torch.randn(splits, 9, cout, cin, device=DEV)
    grad = torch.empty(cout, cin, 3, 3, device=DEV)
    us = timeit(lambda: ops.wgrad_reduce(ws, splits, 9 * cout * cin, 0, cout, cin, 9, grad))
    report(f"wgrad_reduce_{cout}x{cin}_s{splits}", us, 4.0 * (splits + 1) * 9 * cout * cin)


def run_pack_input(B, cin_total, c_lo, nc, cat_mode):
    x0 = torch.rand(B, cin_total, 256, 256, device=DEV)
    x1 = torch.rand(B, cin_total, 256, 256, device=DEV)
    cin = 2 * nc if cat_mode else nc
    kpad = 64 * ((9 * cin + 63) // 64)
    n_img = B if cat_mode else 2 * B
    out = torch.empty(n_img, 256, 256, kpad, device=DEV, dtype=torch.bfloat16)
    us = timeit(lambda: ops.pack_input(x0, x1, c_lo, nc, cat_mode, kpad, out=out))
    report(f"pack_input_B{B}_nc{nc}_cat{cat_mode}_kpad{kpad}", us, out.numel() * 2 + 4.0 * 2 * B * nc * 65536)


def run_colsum(n, H, C, with_w):
    x = bf(n, H, H, 2 * C)[..., C:]
    w = torch.randn(n * H * H, device=DEV) if with_w else None
    npix = n * H * H
    nblk = max(1, min(1184, npix // 64))
    ws = torch.empty(nblk * C, device=DEV)
    out = torch.empty(C, device=DEV)
    us = timeit(lambda: ops.colsum(x, w, npix, nblk, ws, out))
    report(f"colsum_{n}x{H}x{C}_w{int(with_w)}", us, npix * C * 2.0 + (4.0 * npix if with_w else 0))


TESTS = {
    "bn_bwd": lambda: [run_bn_bwd("a", 16, 256, 64, [1]), run_bn_bwd("a", 16, 256, 64, [2, 1]), run_bn_bwd("a", 16, 256, 64, [3]),
                       run_bn_bwd("b", 16, 128, 128, [1]), run_bn_bwd("b", 16, 128, 128, [2, 1]),
                       run_bn_bwd("c", 16, 64, 256, [1]), run_bn_bwd("d", 16, 32, 512, [1]), run_bn_bwd("e", 16, 16, 512, [1]),
                       run_bn_bwd("s", 32, 256, 64, [2, 1], G=2)],
    "bn_apply": lambda: [run_bn_apply("a", 16, 256, 64, False), run_bn_apply("a", 16, 256, 64, True),
                         run_bn_apply("b", 16, 128, 128, False), run_bn_apply("b", 16, 128, 128, True),
                         run_bn_apply("c", 16, 64, 256, True), run_bn_apply("d", 16, 32, 512, True),
                         run_bn_apply("s", 32, 256, 64, True, diff=True, G=2), run_bn_apply("s", 32, 128, 128, False, diff=True, G=2)],
    "wgrad_reduce": lambda: [run_wgrad_reduce(64, 64, 49), run_wgrad_reduce(128, 128, 49), run_wgrad_reduce(256, 256, 12),
                             run_wgrad_reduce(512, 512, 3)],
    "pack_input": lambda: [run_pack_input(16, 6, 0, 2, 1), run_pack_input(16, 6, 2, 4, 1), run_pack_input(32, 4, 0, 4, 0)],
    "colsum": lambda: [run_colsum(16, 256, 64, False), run_colsum(16, 256, 64, True), run_colsum(16, 32, 512, False)],
}
for name, fn in TESTS.items():
    if only is None or name in only:
        fn()
ops.device_status()


def run_reduce_batched(cout, cin, splits, reps=1):
    specs = []
    keep = []
    for _ in range(reps):
        ws = torch.randn(splits, 9, cout, cin, device=DEV)
        grad = torch.empty(cout, cin, 3, 3, device=DEV)
        keep.append((ws, grad))
        specs.append((ws.view(-1), grad, splits, 9 * cout * cin, 0, cout, cin, 9))
    tab, nj, blocks, nbytes = ops.make_reduce_jobs(specs, DEV)
    us = timeit(lambda: ops.wgrad_reduce_batched(tab, nj, blocks, nbytes))
    report(f"reduce_batched_{cout}x{cin}_s{splits}_x{reps}", us, nbytes)


if only is not None and "reduce_batched" in only:
    run_reduce_batched(64, 64, 49, 16)
    run_reduce_batched(128, 128, 49, 8)
    run_reduce_batched(256, 256, 12, 8)
    run_reduce_batched(512, 512, 3, 8)
    run_reduce_batched(512, 512, 8, 4)
    run_reduce_batched(512, 512, 16, 2)
    ops.device_status()
