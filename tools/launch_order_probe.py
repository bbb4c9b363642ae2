"""Does the time of a conv launch depend on which kernel ran before it? Times launch X after
(a) itself, (b) another CTA-pair variant, (c) the weight-gradient kernel, (d) itself + a small BatchNorm kernel — with the
L2 flushed before the timed launch (data and code cold in L2) and without (warm L2). All launches are enqueued behind a
spin kernel. One JSON line per shape. Measured (profiles/r02_launch_order_probe.jsonl): up to +7 us (256 -> 256 @64x64, 61.6 -> 68.5 us)
when another CTA-pair variant ran in between, nothing after the weight-gradient kernel or a small kernel, nothing on the
other shapes: an order effect tools/tile_sweep.py has to average over (it interleaves its candidates), not a general
instruction-cache cost."""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from multimodal_siamese_cd_b200 import ops, tuning
tuning.ENABLED = False
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def conv(n, H, W, ka, N, bn=None):
    A = (torch.randn(n, H, W, ka, device=dev) * 0.5).to(torch.bfloat16)
    Bw = (torch.randn(N, 9 * ka, device=dev) / (9 * ka) ** 0.5).to(torch.bfloat16)
    out = torch.empty(n, H, W, N, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    rows, _ = ops.conv_stat_rows(n, H, W, ka, N, 1, bn=bn)
    st = torch.zeros(rows * N * 2, device=dev)
    return lambda: ops.conv_gemm(0, 0, A, Bw, out, bias=bias, stats=st, stat_groups=1, bn=bn)


def wgrad(n, H, W, c):
    U = (torch.randn(n, H, W, c, device=dev) * 0.5).to(torch.bfloat16)
    V = (torch.randn(n, H, W, c, device=dev) * 0.5).to(torch.bfloat16)
    ctas = ops.wgrad_ctas_per_split(0, 1, c, c)
    splits = max(1, 148 // ctas)
    ws = torch.empty(splits * 9 * c * c, device=dev)
    return lambda: ops.wgrad_gemm(0, 1, 1, U, V, ws, splits, 9 * c * c, c * c, c, 1, 0)


def small():
    r = (torch.randn(4, 32, 32, 64, device=dev)).to(torch.bfloat16)
    a = torch.empty_like(r)
    sc, sh = torch.ones(1, 64, device=dev), torch.zeros(1, 64, device=dev)
    return lambda: ops.bn_apply(r, sc, sh, 1, False, a=a)


def timed(pre, x, do_flush, reps=15):
    ev = []
    torch.cuda._sleep(int(4e7))
    for _ in range(reps):
        for f in pre:
            f()
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); x(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return round(t[len(t) // 2] * 1e3, 2)


other = {256: conv(16, 64, 64, 128, 128), 128: conv(16, 64, 64, 256, 256), 64: conv(16, 64, 64, 256, 256)}
wg = wgrad(16, 64, 64, 256)
sm = small()
for shape in ((16, 64, 64, 256, 256, 256), (16, 16, 16, 512, 512, 128), (16, 256, 256, 64, 64, 64), (16, 128, 128, 128, 128, 128),
              (16, 32, 32, 512, 512, 256)):
    n, H, W, ka, N, bn = shape
    x = conv(n, H, W, ka, N, bn)
    z = other[bn]
    for f in (x, z, wg, sm):
        f()
    torch.cuda.synchronize()
    res = {"shape": shape}
    for fl in (True, False):
        tag = "flush" if fl else "warmL2"
        res[tag] = {"after_itself": timed([x], x, fl), "after_other_pair_variant": timed([x, z], x, fl),
                    "after_wgrad": timed([x, wg], x, fl), "after_itself_and_small_kernel": timed([x, sm], x, fl),
                    "after_other_and_small": timed([x, z, sm], x, fl)}
    print(json.dumps(res), flush=True)
ops.device_status()
