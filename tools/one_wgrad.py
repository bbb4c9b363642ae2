"""One 3x3 weight-gradient launch, for ncu. usage: python tools/one_wgrad.py n H cin cout [iters]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from multimodal_siamese_cd_b200 import ops
n, H, cin, cout = (int(v) for v in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = "cuda"
x = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
dr = torch.randn(n, H, H, cout, device=dev).to(torch.bfloat16)
total = ops.wgrad_tiles(n, H, H)
ctas = ops.wgrad_ctas_per_split(0, 1, cout, cin) if (cout >= 128 or cin < 128) else ops.wgrad_ctas_per_split(0, 1, cin, cout)
splits = max(1, min(total, 148 // ctas))
ws = torch.empty(splits, 9, cout, cin, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    if cout >= 128 or cin < 128:
        ops.wgrad_gemm(0, 1, 1, dr, x, ws, splits, 9 * cout * cin, cout * cin, cin, 1)
    else:
        ops.wgrad_gemm(0, -1, 1, x, dr, ws, splits, 9 * cout * cin, cout * cin, 1, cin)
e1.record()
torch.cuda.synchronize()
ops.device_status()
ms = e0.elapsed_time(e1)
print(f"wgrad n={n} H={H} {cin}->{cout} splits={splits}: {ms * 1e3:.1f} us, {2.0 * n * H * H * cin * cout * 9 / ms / 1e9:.1f} TFLOP/s")
if "--trace" in sys.argv:  # debug build only (-DB200CD_TRACE)
    import ctypes, json
    lib = ctypes.CDLL(str(ROOT / "multimodal_siamese_cd_b200" / "libb200cd.so"))
    buf = (ctypes.c_longlong * (4 * 4096))()
    rc = lib.b200cd_debug_trace_wgrad(buf)
    tr = [[v for v in buf[r * 4096:(r + 1) * 4096] if v > 0] for r in range(4)]
    t0 = min(v for r in tr for v in r)
    tr = [[v - t0 for v in r] for r in tr]
    m = tr[0]
    its = (len(m) - 1) // 2
    waits = [m[2 * i + 1] - m[2 * i] for i in range(its)]
    iss = [m[2 * i + 2] - m[2 * i + 1] for i in range(its)]
    print("MMA warp: iters", its, "span", m[-1] - m[0], "per iter", (m[-1] - m[0]) / max(its, 1))
    print("  full waits: total", sum(waits), "first 24", waits[:24])
    print("  issue: total", sum(iss), "first 24", iss[:24])
    for role, name in ((1, "prodU"), (2, "prodV")):
        e = tr[role]
        k = len(e) // 2
        w = [e[2 * i + 1] - e[2 * i] for i in range(k)]
        g = [e[2 * i + 2] - e[2 * i + 1] for i in range(k - 1)]
        print(name, "loads", k, "empty-wait total", sum(w), "first 24", w[:24], "issue gaps first 24", g[:24])
    print("epilogue: wait", tr[3][1] - tr[3][0], "drain", tr[3][2] - tr[3][1], "end", tr[3][2], "mma end", m[-1])
