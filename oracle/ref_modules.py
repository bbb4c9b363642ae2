"""ORACLE — TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference modules (utils/networks.py,
utils/loss_functions.py, utils/metrics.py) from /root/reference or from the staged copy oracle/_ref/reference
(oracle/stage_reference.py), with an in-memory stand-in for the missing fvcore package, and restates the three
training-loop bodies around them. Used by bench.py's baseline legs (--impl reference on the host cores; the
"stock PyTorch eager on B200" library baseline) and by tests — never by the product package.

Loop bodies restated (the arithmetic is the reference's own modules):
  supervised  train_supervised.py:63-77
  dualtask    train_supervised_dualtask.py:68-86
  mmcr        train_semisupervised.py:66-113
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import torch

from . import stage_reference


def load():
    """(networks, loss_functions) of the reference, or None when no reference tree is available."""
    root = stage_reference.staged_root()
    if root is None:
        return None
    repo = Path(__file__).resolve().parent.parent
    if str(repo) not in sys.path:
        sys.path.insert(0, str(repo))
    from multimodal_siamese_cd_b200.config import install_fvcore_stub
    install_fvcore_stub()
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    for name in ("utils", "utils.networks", "utils.loss_functions", "utils.experiment_manager"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", None) or getattr(mod, "__path__", [""])[0]).startswith(str(root)):
            del sys.modules[name]          # a substituted module of the launcher / another tree
    nets = importlib.import_module("utils.networks")
    losses = importlib.import_module("utils.loss_functions")
    return nets, losses


def loss_of(losses, outs, batch: dict, kind: str, alpha: float):
    crit = losses.get_criterion("PowerJaccardLoss")
    if kind == "supervised":
        return crit(outs, batch["y_change"])
    if kind == "dualtask":
        c, s1, s2 = outs
        return (crit(c, batch["y_change"]) + (crit(s1, batch["y_sem_t1"]) + crit(s2, batch["y_sem_t2"])) / 2) / 2
    f, s1, s2 = outs
    lab, y = batch["is_labeled"], batch["y_change"]
    loss = None
    if lab.any():
        loss = alpha * (crit(f[lab,], y[lab,]) + crit(s1[lab,], y[lab,]) + crit(s2[lab,], y[lab,])) / 3
    if not lab.all():
        unl = torch.logical_not(lab)
        cons = (1 - alpha) * crit(s1[unl,], torch.sigmoid(s2)[unl,])
        loss = cons if loss is None else loss + cons
    return loss


def train_step(net, losses, batch: dict, kind: str, alpha: float):
    """zero_grad -> forward -> loss -> backward with the reference's modules. Returns (outs, loss)."""
    for p in net.parameters():
        p.grad = None
    outs = net(batch["x_t1"], batch["x_t2"])
    loss = loss_of(losses, outs, batch, kind, alpha)
    loss.backward()
    return outs, loss
