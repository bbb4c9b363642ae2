"""Sustained (power-limited) rate of one weight-gradient shape: python tools/sustained_wgrad.py n H cin cout seconds"""
import subprocess, sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from multimodal_siamese_cd_b200 import ops
n, H, cin, cout = (int(v) for v in sys.argv[1:5])
secs = float(sys.argv[5]) if len(sys.argv) > 5 else 4.0
dev = "cuda"
x = torch.randn(n, H, H, cin, device=dev).to(torch.bfloat16)
dr = torch.randn(n, H, H, cout, device=dev).to(torch.bfloat16)
total = ops.wgrad_tiles(n, H, H)
ctas = ((cout + 127) // 128) * (cin // 128 if cin % 128 == 0 else cin // 64) * 3
splits = max(1, min(total, 148 // ctas))
ws = torch.empty(splits, 9, cout, cin, device=dev)
f = lambda: ops.wgrad_gemm(0, 1, 1, dr, x, ws, splits, 9 * cout * cin, cout * cin, cin, 1)
clocks = []
stop = False
def sample():
    while not stop:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        clocks.append(o)
        time.sleep(0.2)
th = threading.Thread(target=sample); th.start()
flops = 2.0 * n * H * H * cin * cout * 9
t_end = time.time() + secs
while time.time() < t_end:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(f"{ms*1e3:.1f} us {flops/ms/1e9:.0f} TF", clocks[-1] if clocks else "")
stop = True; th.join()
