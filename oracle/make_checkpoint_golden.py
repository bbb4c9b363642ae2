"""Generates tests/golden/ckpt/ref_checkpoint_siamese_tiny.pt: a checkpoint WRITTEN BY THE UNMODIFIED REFERENCE
(utils/networks.py:30-38 save_checkpoint, after one real AdamW step) for a deliberately narrow network
(TOPOLOGY [8, 16], so the file stays ~100 KB). Run in the build container only (/root/reference is not on the GPU box).
The drop-in `networks.load_checkpoint` must read this file as is (tests/test_checkpoint_cpu.py).

usage: python oracle/make_checkpoint_golden.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import shutil
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from multimodal_siamese_cd_b200.config import install_fvcore_stub, synthetic_cfg  # noqa: E402
from oracle.unet_oracle import synthetic_batch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    install_fvcore_stub()
    sys.path.insert(0, args.reference)
    from utils import loss_functions as ref_loss  # noqa: E402  (the reference, imported in place)
    from utils import networks as ref_networks  # noqa: E402

    cfg = synthetic_cfg("siameseunet", in_channels=4, topology=(8, 16))
    tmp = Path(tempfile.mkdtemp())
    cfg.PATHS.OUTPUT = str(tmp)
    cfg.NAME = "tiny"
    torch.manual_seed(cfg.SEED)
    net = ref_networks.create_network(cfg)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=cfg.TRAINER.LR, weight_decay=0.01)   # train_supervised.py:32
    batch = synthetic_batch(2, 4, 16, 16, seed=7)
    crit = ref_loss.get_criterion("PowerJaccardLoss")
    opt.zero_grad()
    loss = crit(net(batch["x_t1"], batch["x_t2"]), batch["y_change"])
    loss.backward()
    opt.step()
    ref_networks.save_checkpoint(net, opt, 3, 17, cfg)                                 # epoch 3, global step 17
    src = tmp / "networks" / "tiny_checkpoint3.pt"
    dst = ROOT / "tests" / "golden" / "ckpt" / "ref_checkpoint_siamese_tiny.pt"
    shutil.copy(src, dst)
    print(dst, dst.stat().st_size, "bytes; loss", float(loss))


if __name__ == "__main__":
    main()
