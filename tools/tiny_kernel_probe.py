"""In-graph cost of the latency-bound kernels that sit between the big launches of a step (BatchNorm statistics,
BatchNorm-backward finalize, row sums, loss finalize): N dependent launches of one kernel captured in ONE CUDA graph,
replayed; time per launch = what the kernel adds to a dependent chain. `pj_loss` (one thread, one load, one store) is
the floor.   python tools/tiny_kernel_probe.py > gpurun_out/tiny_kernels.json"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from multimodal_siamese_cd_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
N = 200


def per_launch_us(fn) -> float:
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * N)


def main():
    res = {}
    sums = torch.rand(3, device=DEV, dtype=torch.float64) + 1
    loss = torch.zeros(1, device=DEV)
    res["pj_loss (floor)"] = per_launch_us(lambda: ops.pj_loss(sums, loss))
    for C_ in (64, 256, 1024):
        for G in (1, 2):
            rows = 296
            part = torch.rand(G, rows, C_, 2, device=DEV)
            f = lambda n: torch.ones(n, device=DEV)  # noqa: E731
            gamma, beta, rm, rv = f(C_), f(C_), f(C_), f(C_)
            nbt = torch.zeros(1, device=DEV, dtype=torch.int64)
            mean, invstd, scale, shift = (torch.zeros(G * C_, device=DEV) for _ in range(4))
            ws = torch.zeros(32 * G * C_ * 2, device=DEV, dtype=torch.float64)
            res[f"bn_stats C{C_} G{G} rows{rows}"] = per_launch_us(
                lambda: ops.bn_stats(part, C_, C_, rows, G, 65536.0, 1, ws, gamma, beta, rm, rv, nbt, 0.1, 1e-5, True, False,
                                     mean, invstd, scale, shift))
    for C_ in (64, 512):
        rows = 296
        st = torch.rand(rows, 2 * C_, 2, device=DEV)
        out = torch.zeros(C_, device=DEV)
        res[f"stat_rowsum C{C_}"] = per_launch_us(lambda: ops.stat_rowsum(st, rows, 2 * C_, C_, C_, out))
    # BatchNorm backward on a tiny tensor: reduce + finalize + dx (3 launches) and finalize + dx from epilogue sums (2)
    for C_ in (64, 512):
        n, H, W, G = 2, 16, 16, 1
        r = torch.randn(n, H, W, C_, device=DEV).bfloat16()
        dy = torch.randn(n, H, W, C_, device=DEV).bfloat16()
        dr = torch.empty_like(r)
        mean, invstd, scale, shift = (torch.rand(G * C_, device=DEV) for _ in range(4))
        dg, db = torch.zeros(C_, device=DEV), torch.zeros(C_, device=DEV)
        ws = torch.zeros(ops.bn_bwd_ws_floats(n, H, W, C_, G), device=DEV)
        srcs = ops.make_srcs([{"kind": 1, "t": dy}])
        res[f"bn_bwd tiny C{C_} (reduce+finalize+dx)"] = per_launch_us(
            lambda: ops.bn_bwd(r, mean, invstd, scale, shift, srcs, G, ws, dg, db, dr)) / 1.0
        rows = 296
        sm = torch.rand(G, rows, C_, 2, device=DEV)
        res[f"bn_bwd tiny C{C_} from_sums (finalize+dx)"] = per_launch_us(
            lambda: ops.bn_bwd(r, mean, invstd, scale, shift, srcs, G, ws, dg, db, dr, sums=sm, sum_rows=rows))
    print(json.dumps({k: round(v, 2) for k, v in res.items()}, indent=1))


if __name__ == "__main__":
    main()
