"""Wall time of each of the first drop-in steps (synchronised), to find one-time stalls. usage: e2e_steps.py config"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from multimodal_siamese_cd_b200 import loss_functions, networks
from multimodal_siamese_cd_b200.config import synthetic_cfg
from multimodal_siamese_cd_b200.data import DevicePrefetcher, LossReader
cfgname = sys.argv[1] if len(sys.argv) > 1 else "dtsiamese"
dev = torch.device("cuda", 0)
mtype, cin, B, kind, alpha, _, _ = bench.CONFIGS[cfgname]
torch.manual_seed(7)
net = networks.create_network(synthetic_cfg(mtype, in_channels=cin)).to(dev).train()
g = torch.Generator().manual_seed(7)
xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
host = {"x_t1": torch.rand(B, xc, 256, 256, generator=g).pin_memory(), "x_t2": torch.rand(B, xc, 256, 256, generator=g).pin_memory()}
for k in ("y_change", "y_sem_t1", "y_sem_t2"):
    host[k] = (torch.rand(B, 1, 256, 256, generator=g) > 0.9).float().pin_memory()
crit = loss_functions.get_criterion("PowerJaccardLoss")
reader = LossReader(dev)
def step(b):
    for p in net.parameters():
        p.grad = None
    outs = net(b["x_t1"], b["x_t2"])
    if kind == "supervised":
        loss = crit(outs, b["y_change"])
    else:
        c, s1, s2 = outs
        loss = (crit(c, b["y_change"]) + (crit(s1, b["y_sem_t1"]) + crit(s2, b["y_sem_t2"])) / 2) / 2
    loss.backward()
    reader.push(loss)
ts = []
t_prev = time.perf_counter()
for i, b in enumerate(DevicePrefetcher((host for _ in range(40)), dev)):
    step(b)
    now = time.perf_counter()
    ts.append((now - t_prev) * 1e3)
    t_prev = now
torch.cuda.synchronize()
print(cfgname, "host ms per step:", [round(t, 1) for t in ts])
