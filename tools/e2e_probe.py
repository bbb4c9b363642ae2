"""Where does the end-to-end step time go? Variants of the drop-in loop on the bench workload."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from multimodal_siamese_cd_b200 import loss_functions, networks
from multimodal_siamese_cd_b200.config import synthetic_cfg
from multimodal_siamese_cd_b200.data import DevicePrefetcher

dev = torch.device("cuda", 0)
mtype, cin, B, kind, alpha, _, _ = bench.CONFIGS["dualstream"]
torch.manual_seed(7)
net = networks.create_network(synthetic_cfg(mtype, in_channels=cin)).to(dev).train()
g = torch.Generator().manual_seed(7)
host = {"x_t1": torch.rand(B, 6, 256, 256, generator=g).pin_memory(), "x_t2": torch.rand(B, 6, 256, 256, generator=g).pin_memory(),
        "y_change": (torch.rand(B, 1, 256, 256, generator=g) > 0.9).float().pin_memory()}
crit = loss_functions.get_criterion("PowerJaccardLoss")
K = 60

def step(b, item=True, zero=True):
    if zero:
        for p in net.parameters():
            p.grad = None
    out = net(b["x_t1"], b["x_t2"])
    loss = crit(out, b["y_change"])
    loss.backward()
    return loss.item() if item else loss

def timed(name, it, **kw):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in it:
        step(b, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e3
    print(f"{name:50s} {dt:7.2f} ms/step  {B / dt * 1e3:7.1f} pairs/s", flush=True)

devb = {k: v.to(dev) for k, v in host.items()}
for _ in range(5):
    step(devb)
timed("device-resident batch, item()", (devb for _ in range(K)))
timed("device-resident batch, no item()", (devb for _ in range(K)), item=False)
timed("device-resident, no item, no grad reset", (devb for _ in range(K)), item=False, zero=False)
timed("blocking .to(dev) per step, item()", ({k: v.to(dev, non_blocking=True) for k, v in host.items()} for _ in range(K)))
timed("DevicePrefetcher, item()", DevicePrefetcher((host for _ in range(K)), dev))
timed("DevicePrefetcher, no item()", DevicePrefetcher((host for _ in range(K)), dev), item=False)
# host time of one step without waiting for the GPU
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step(devb, item=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host-side launch time per step: {(t1 - t0) / 20 * 1e3:.2f} ms")
