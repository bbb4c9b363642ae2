"""Two eager (no CUDA graph) training steps of the bench workload, for `ncu` (launch list / full capture of one kernel).
usage: python tools/ncu_step.py [config] [batch]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
from multimodal_siamese_cd_b200 import networks, ops  # noqa: E402
from multimodal_siamese_cd_b200.config import synthetic_cfg  # noqa: E402
from multimodal_siamese_cd_b200.step import TrainStep  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "dualstream"
mtype, cin, B, kind, alpha, _, _ = bench.CONFIGS[cfgname]
if len(sys.argv) > 2:
    B = int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.manual_seed(7)
net = networks.create_network(synthetic_cfg(mtype, in_channels=cin)).to(dev).train()
net.module.use_cuda_graphs = False
ts = TrainStep(net.module, B, 256, 256, kind=kind, alpha=alpha, device=dev, dp_group=None)
xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
g = torch.Generator().manual_seed(7)
x1, x2 = torch.rand(B, xc, 256, 256, generator=g), torch.rand(B, xc, 256, 256, generator=g)
tg = {k: (torch.rand(B, 1, 256, 256, generator=g) > 0.9).float() for k in ts.targets}
lab = torch.tensor([i % 3 != 2 for i in range(B)])
ts.set_inputs(x1, x2, is_labeled=lab if kind == "mmcr" else None, **tg)
for _ in range(2):
    loss = ts.run()
torch.cuda.synchronize()
ops.device_status(0)
print("loss", loss.item(), "launches", ops.LAUNCHES)
