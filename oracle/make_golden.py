"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the build container only; /root/reference does
not exist on the GPU box). The reference modules are imported in place — nothing is copied — with an in-memory
`fvcore.common.config.CfgNode` stand-in (fvcore is not installed; utils/experiment_manager.py:7 needs it).

usage: python oracle/make_golden.py [--reference /root/reference]

Each fixture holds, for one (network type, config, seed-7 synthetic batch):
  init   fingerprint of every initial parameter (pins default-init RNG order under torch.manual_seed(7))
  outs   the logits returned by the reference module's forward(x_t1, x_t2) in train mode
  loss   the training-loop loss (train_supervised.py:75 / train_supervised_dualtask.py:75-85 / train_semisupervised.py:82-113)
  grads  fingerprint (L2 norm, sum, first 4 values) of every parameter gradient after loss.backward()
  bn     running_mean / running_var fingerprints and num_batches_tracked after the step
  mask_f1  thresholded-mask popcount and F1 from the reference's utils/metrics.py at threshold 0.5
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
    # name, model type, in_channels, topology, B, step kind
    ("unet_small", "unet", 6, (64, 128), 3, "supervised"),
    ("siamese_small", "siameseunet", 4, (64, 128), 3, "supervised"),
    ("dualstream_small", "dualstreamunet", 6, (64, 128), 3, "supervised"),
    ("dtsiamese_small", "dtsiameseunet", 6, (64, 128), 3, "dualtask"),
    ("whatevernet_small", "whatevernet", 6, (64, 128), 3, "mmcr"),
    ("whatevernet2_small", "whatevernet2", 6, (64, 128), 3, "mmcr"),
    ("siamese_full", "siameseunet", 4, (64, 128, 256, 512), 2, "supervised"),
    ("dtsiamese_full", "dtsiameseunet", 6, (64, 128, 256, 512), 2, "dualtask"),
]
H = W = 32
ALPHA = 0.5


def fingerprint(t: torch.Tensor) -> torch.Tensor:
    t = t.detach().double().flatten()
    head = torch.zeros(4, dtype=torch.float64)
    head[: min(4, t.numel())] = t[:4]
    return torch.cat([torch.stack([t.norm(), t.sum()]), head])


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    install_fvcore_stub()
    sys.path.insert(0, args.reference)
    from utils import loss_functions as ref_loss  # noqa: E402  (the reference, imported in place)
    from utils import metrics as ref_metrics  # noqa: E402
    from utils import networks as ref_networks  # noqa: E402

    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for name, mtype, cin, topo, B, kind in CASES:
        cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
        torch.manual_seed(cfg.SEED)
        net = ref_networks.create_network(cfg)     # nn.DataParallel pass-through on CPU
        net.train()
        init = {k: fingerprint(v) for k, v in net.state_dict().items() if v.is_floating_point()}
        batch = synthetic_batch(B, 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin, H, W, seed=7)
        crit = ref_loss.get_criterion("PowerJaccardLoss")
        outs = net(batch["x_t1"], batch["x_t2"])
        if kind == "supervised":
            loss = crit(outs, batch["y_change"])
            out_list = [outs]
        elif kind == "dualtask":
            c, s1, s2 = outs
            sem = (crit(s1, batch["y_sem_t1"]) + crit(s2, batch["y_sem_t2"])) / 2
            loss = (crit(c, batch["y_change"]) + sem) / 2
            out_list = [c, s1, s2]
        else:
            f, s1, s2 = outs
            lab = batch["is_labeled"]
            y = batch["y_change"]
            p2 = torch.sigmoid(s2)
            sup = ALPHA * (crit(f[lab,], y[lab,]) + crit(s1[lab,], y[lab,]) + crit(s2[lab,], y[lab,])) / 3
            unl = torch.logical_not(lab)
            cons = (1 - ALPHA) * crit(s1[unl,], p2[unl,])
            loss = sup + cons
            out_list = [f, s1, s2]
        loss.backward()
        grads = {k: (fingerprint(p.grad) if p.grad is not None else None) for k, p in net.named_parameters()}
        sd = net.state_dict()
        bn = {k: (fingerprint(v) if v.is_floating_point() else v.clone()) for k, v in sd.items()
              if "running_" in k or "num_batches_tracked" in k}
        m = ref_metrics.MultiThresholdMetric(torch.tensor([0.5]))
        m.add_sample(batch["y_change"], torch.sigmoid(out_list[0].detach()))
        fix = {
            "case": (name, mtype, cin, tuple(topo), B, kind, H, W, ALPHA),
            "init": init,
            "outs": [o.detach().clone() for o in out_list],
            "loss": loss.detach().clone(),
            "grads": grads,
            "bn": bn,
            "mask_f1": {"popcount": int((out_list[0] > 0).sum()), "f1": float(m.compute_f1().item()),
                        "tp": float(m.TP.item())},
            "torch": torch.__version__,
        }
        torch.save(fix, out_dir / f"{name}.pt")
        print(f"{name}: loss={loss.item():.8f} outs={[tuple(o.shape) for o in out_list]} "
              f"params={sum(p.numel() for p in net.parameters())}")


if __name__ == "__main__":
    main()
