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
if cout >= 128 or cin < 128:
    ctas = ((cout + 127) // 128) * (cin // 128 if cin % 128 == 0 else cin // 64) * 3
else:
    ctas = (cin // 128) * (cout // 128 if cout % 128 == 0 else cout // 64) * 3
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
