"""One 3x3 conv launch per variant, for ncu. usage: python tools/one_conv.py n H cin cout [pair|halo|wide|plain] [iters]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from multimodal_siamese_cd_b200 import ops  # noqa: E402

n, H, cin, cout = (int(v) for v in sys.argv[1:5])
variant = sys.argv[5] if len(sys.argv) > 5 else "pair"
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = "cuda"
A = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5)
Bw = ops.pack_weights(0, w)
o = torch.empty(n, H, H, cout, device=dev, dtype=torch.bfloat16)
stats = torch.empty(n * ops.conv_gemm_tiles(H, H), cout, 2, device=dev)
kw = {"pair": dict(pair=True), "halo": dict(halo=True, wide=False, pair=False), "wide": dict(halo=False, wide=True, pair=False),
      "plain": dict(halo=False, wide=False, pair=False)}[variant]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    ops.conv_gemm(0, 0, A, Bw, o, stats=stats, **kw)
e1.record()
torch.cuda.synchronize()
ops.device_status()
ms = e0.elapsed_time(e1)
print(f"{variant} n={n} H={H} {cin}->{cout}: {ms * 1e3:.1f} us, {2.0 * n * H * H * cin * cout * 9 / ms / 1e9:.1f} TFLOP/s")
