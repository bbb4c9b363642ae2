"""ORACLE — TEST INFRASTRUCTURE ONLY. Data-parallel semantics of the reference's nn.DataParallel wrap
(utils/networks.py:27) restated two ways:

  dp_emulation_step   single process: split the batch into `world` contiguous chunks (scatter rule), run the network on
                      each chunk in train mode with identical weights (per-replica BatchNorm statistics), concatenate
                      the logits, ONE loss over the global batch, backward => gradients accumulate (= SUM over replicas).
  dp_rank_step        what one rank of a one-process-per-GPU job computes: its own chunk, the power-Jaccard partial
                      sums all-reduced (SUM) between loss forward and backward, gradients all-reduced with SUM.

tests/test_dp_gloo.py checks on CPU (gloo, world_size 2) that both agree — the exchange the B200 path implements
(3 scalars per loss term + SUM gradient all-reduce) reproduces DataParallel exactly.
"""
from __future__ import annotations

import torch

from . import unet_oracle as O


def shard(batch: dict, rows: slice) -> dict:
    return {k: v[rows] for k, v in batch.items()}


def dp_emulation_step(model_type, sd, batch, world, rows_of):
    outs = []
    for r in range(world):
        o = O.forward(model_type, sd, batch["x_t1"][rows_of(r)], batch["x_t2"][rows_of(r)], train=True)
        outs.append(o)
    logits = torch.cat(outs, 0)
    loss = O.power_jaccard_loss(logits, batch["y_change"])
    names = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return {"logits": logits.detach(), "loss": loss.detach(), "grads": dict(zip(names, grads))}


class GlobalPowerJaccard(torch.autograd.Function):
    """power_jaccard_loss (utils/loss_functions.py:141-150) whose three sums are all-reduced across ranks; the
    backward is the closed form of SURVEY §8a L2 evaluated with the GLOBAL sums on the LOCAL rows."""

    @staticmethod
    def forward(ctx, z, t, allreduce):
        p = torch.sigmoid(z)
        sums = torch.stack([(p * t).sum(), (p * p).sum(), (t * t).sum()]).double()
        allreduce(sums)
        inter, denom = sums[0], sums[1] + sums[2] - sums[0] + 1e-6
        ctx.save_for_backward(p, t, sums)
        return (1 - inter / denom).to(z.dtype)

    @staticmethod
    def backward(ctx, g):
        p, t, sums = ctx.saved_tensors
        inter = sums[0].to(p.dtype)
        denom = (sums[1] + sums[2] - sums[0] + 1e-6).to(p.dtype)
        dp = -(t * denom - inter * (2 * p - t)) / (denom * denom)
        return g * dp * p * (1 - p), None, None


def dp_rank_step(model_type, sd, local_batch, allreduce_sums, allreduce_grads):
    out = O.forward(model_type, sd, local_batch["x_t1"], local_batch["x_t2"], train=True)
    loss = GlobalPowerJaccard.apply(out, local_batch["y_change"], allreduce_sums)
    names = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    flat = torch.cat([g.reshape(-1) for g in grads if g is not None])
    allreduce_grads(flat)
    out_grads, off = {}, 0
    for k, g in zip(names, grads):
        if g is None:
            out_grads[k] = None
            continue
        out_grads[k] = flat[off:off + g.numel()].view_as(g)
        off += g.numel()
    return {"logits": out.detach(), "loss": loss.detach(), "grads": out_grads}
