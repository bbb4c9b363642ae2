"""Generates tests/golden/eval/*.pt from the UNMODIFIED reference (build container only): inference on tiles whose
levels have odd sizes — nn.MaxPool2d(2) floors, Up pads the up-sampled tensor to the skip's size
(utils/networks.py:420, 440-443) — the way utils/evaluation.py:9-23 runs the network (net.eval(), no_grad, whole tile).

usage: python oracle/make_eval_golden.py [--reference /root/reference]

Each fixture: one train-mode forward on an even-sized seed-11 batch (so the BatchNorm running statistics are not the
initial (0, 1)), then the eval-mode logits on a seed-7 batch of the odd size, the thresholded-mask popcount and the
reference's F1 (utils/metrics.py) at threshold 0.5.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from multimodal_siamese_cd_b200.config import install_fvcore_stub, synthetic_cfg  # noqa: E402
from oracle.unet_oracle import synthetic_batch  # noqa: E402

CASES = [
    # name, model type, in_channels, topology, B, (H, W) of the tile, (H, W) of the warm-up batch
    ("siamese_full_eval_67x45", "siameseunet", 4, (64, 128, 256, 512), 1, (67, 45), (32, 32)),
    ("dualstream_small_eval_35x50", "dualstreamunet", 6, (64, 128), 2, (35, 50), (32, 48)),
    ("whatevernet_small_eval_41x41", "whatevernet", 6, (64, 128), 1, (41, 41), (32, 32)),
    ("unet_full_eval_19x33", "unet", 6, (64, 128, 256, 512), 1, (19, 33), (32, 32)),
]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    install_fvcore_stub()
    sys.path.insert(0, args.reference)
    from utils import metrics as ref_metrics  # noqa: E402  (the reference, imported in place)
    from utils import networks as ref_networks  # noqa: E402

    out_dir = ROOT / "tests" / "golden" / "eval"
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for name, mtype, cin, topo, B, (H, W), (Hw, Ww) in CASES:
        cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
        torch.manual_seed(cfg.SEED)
        net = ref_networks.create_network(cfg)
        xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
        warm = synthetic_batch(B, xc, Hw, Ww, seed=11)
        batch = synthetic_batch(B, xc, H, W, seed=7)
        net.train()
        with torch.no_grad():
            net(warm["x_t1"], warm["x_t2"])
        net.eval()
        with torch.no_grad():
            logits = net(batch["x_t1"], batch["x_t2"])
        assert torch.is_tensor(logits) and logits.shape == (B, 1, H, W)
        m = ref_metrics.MultiThresholdMetric(torch.tensor([0.5]))
        m.add_sample(batch["y_change"], torch.sigmoid(logits))
        torch.save({"case": (name, mtype, cin, tuple(topo), B, H, W, Hw, Ww), "logits": logits.clone(),
                    "popcount": int((logits > 0).sum()), "f1": float(m.compute_f1().item()), "torch": torch.__version__},
                   out_dir / f"{name}.pt")
        print(f"{name}: logits {tuple(logits.shape)} popcount {int((logits > 0).sum())}")


if __name__ == "__main__":
    main()
