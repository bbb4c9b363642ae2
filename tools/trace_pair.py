"""Timeline of one CTA pair of the fprop pair kernel (debug build: B200CD_NVCC_EXTRA=-DB200CD_TRACE python -m
multimodal_siamese_cd_b200.build --force). usage: python tools/trace_pair.py n H cin cout > gpurun_out/trace.json"""
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from multimodal_siamese_cd_b200 import _lib, ops  # noqa: E402

n, H, cin, cout = (int(v) for v in sys.argv[1:5])
dev = "cuda"
A = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5)
Bw = ops.pack_weights(0, w)
o = torch.empty(n, H, H, cout, device=dev, dtype=torch.bfloat16)
stats = torch.empty(n * ops.conv_gemm_tiles(H, H), cout, 2, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fused = "--bnbwd" in sys.argv   # input-gradient launch with the fused BatchNorm-backward sums
if fused:
    G = 1
    r = torch.randn(n, H, H, cout, device=dev).to(torch.bfloat16)
    scale = torch.rand(G, cout, device=dev) + 0.5
    shift = torch.randn(G, cout, device=dev) * 0.3
    rows, per_cta = ops.conv_stat_rows(n, H, H, cin, cout, G, mode=0)
    sums = torch.empty(G, rows, cout, 2, device=dev)
for i in range(3):
    if i == 2:
        e0.record()
    if fused:
        ops.conv_gemm_bnbwd(0, A, Bw, o, r, scale, shift, sums, G)
    else:
        ops.conv_gemm(0, 0, A, Bw, o, stats=stats, pair=True)
e1.record()
torch.cuda.synchronize()
ops.device_status()
lib = ctypes.CDLL(str(ROOT / "multimodal_siamese_cd_b200" / "libb200cd.so"))
buf = (ctypes.c_longlong * (6 * 4096))()
rc = lib.b200cd_debug_trace(buf)
tr = [list(buf[r * 4096:(r + 1) * 4096]) for r in range(6)]
print(json.dumps({"us": e0.elapsed_time(e1) * 1e3, "shape": [n, H, cin, cout], "rc": rc, "trace": tr}))
