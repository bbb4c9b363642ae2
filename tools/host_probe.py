"""Host-side time of the drop-in loop, segment by segment (no synchronisation inside the loop)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from multimodal_siamese_cd_b200 import loss_functions, networks
from multimodal_siamese_cd_b200.config import synthetic_cfg

cfgname = sys.argv[1] if len(sys.argv) > 1 else "dtsiamese"
dev = torch.device("cuda", 0)
mtype, cin, B, kind, alpha, _, _ = bench.CONFIGS[cfgname]
torch.manual_seed(7)
net = networks.create_network(synthetic_cfg(mtype, in_channels=cin)).to(dev).train()
g = torch.Generator().manual_seed(7)
xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
b = {"x_t1": torch.rand(B, xc, 256, 256, generator=g).to(dev), "x_t2": torch.rand(B, xc, 256, 256, generator=g).to(dev)}
for k in ("y_change", "y_sem_t1", "y_sem_t2"):
    b[k] = (torch.rand(B, 1, 256, 256, generator=g) > 0.9).float().to(dev)
crit = loss_functions.get_criterion("PowerJaccardLoss")
T = {"zero": 0.0, "fwd": 0.0, "loss": 0.0, "bwd": 0.0}
K = 40
def step(rec):
    t0 = time.perf_counter()
    for p in net.parameters():
        p.grad = None
    t1 = time.perf_counter()
    outs = net(b["x_t1"], b["x_t2"])
    t2 = time.perf_counter()
    if kind == "supervised":
        loss = crit(outs, b["y_change"])
    else:
        c, s1, s2 = outs
        loss = (crit(c, b["y_change"]) + (crit(s1, b["y_sem_t1"]) + crit(s2, b["y_sem_t2"])) / 2) / 2
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    if rec:
        T["zero"] += t1 - t0; T["fwd"] += t2 - t1; T["loss"] += t3 - t2; T["bwd"] += t4 - t3
for _ in range(5):
    step(False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    step(True)
host = time.perf_counter() - t0
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(cfgname, {k: round(v / K * 1e3, 3) for k, v in T.items()}, "host ms/step", round(host / K * 1e3, 3), "wall ms/step", round(wall / K * 1e3, 3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    step(False)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
