"""Inference on full tiles (utils/evaluation.py:7-23 runs batch 1 over whole AOI tiles): timing + sanity."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from multimodal_siamese_cd_b200 import networks
from multimodal_siamese_cd_b200.config import synthetic_cfg
dev = torch.device("cuda", 0)
for mtype, cin in (("siameseunet", 4), ("dualstreamunet", 6), ("whatevernet", 6)):
    torch.manual_seed(7)
    net = networks.create_network(synthetic_cfg(mtype, in_channels=cin)).to(dev).eval()
    for H, W in ((1024, 1024), (896, 1008)):
        x1 = torch.rand(1, cin, H, W, device=dev); x2 = torch.rand(1, cin, H, W, device=dev)
        with torch.no_grad():
            for _ in range(3):
                out = net(x1, x2)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10):
                out = net(x1, x2)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"{mtype} eval {H}x{W}: {dt*1e3:.2f} ms/tile, out {tuple(out.shape)}, finite {bool(torch.isfinite(out).all())}, "
              f"mean|logit| {out.abs().mean().item():.3f}", flush=True)
    net.module.release_engines()
