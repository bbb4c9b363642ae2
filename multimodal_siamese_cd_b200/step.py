"""Fused training steps: forward plan -> power-Jaccard loss kernels -> backward plan, without torch autograd in
between, plus the one-process-per-GPU data-parallel exchange (3 scalars per loss term + gradient all-reduce(SUM)).

Loss compositions mirrored from the reference's training loops:
  'supervised'  loss = pj(logits, y_change)                                        train_supervised.py:71-76
  'dualtask'    loss = (pj(c, y_c) + (pj(s1, y_s1) + pj(s2, y_s2)) / 2) / 2        train_supervised_dualtask.py:75-85
  'mmcr'        loss = a*(pj(f[l],y[l]) + pj(s1[l],y[l]) + pj(s2[l],y[l]))/3
                       + (1-a)*pj(s1[u], sigmoid(s2)[u])                           train_semisupervised.py:74-113
                (l = labeled rows, u = unlabeled rows; a term is dropped when its row set is empty; the consistency
                target carries gradient into stream 2)

Data-parallel semantics are those of nn.DataParallel (utils/networks.py:27): the loss is the ratio over
the GLOBAL batch, replica gradients are SUMMED, BatchNorm statistics stay per replica.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .engine import PlanGraph, StepEngine


_DEBUG_SKIP_ALLREDUCE = __import__("os").environ.get("B200CD_DEBUG_SKIP_ALLREDUCE", "0") == "1"
# Data-parallel step as ONE CUDA graph (NCCL only): forward, loss sums, their all-reduce, loss backward, the bucketed
# backward and its gradient all-reduces are captured together, so the host launches one graph per step instead of
# 3 + 2 x buckets graphs with eager collectives in between. B200CD_DP_GRAPH=0 keeps the segmented replay.
_DP_GRAPH = __import__("os").environ.get("B200CD_DP_GRAPH", "1") != "0"


@dataclass
class _Term:
    z: torch.Tensor                 # logits [rows, 1, H, W] fp32 (static engine buffer)
    t_key: str                      # key of the target tensor ('y_change', ...) or '@logit:<i>' for another head
    t_is_logit: bool
    sel: Optional[int]              # None: all rows; 1: labeled rows; 0: unlabeled rows
    weight: float                   # composition weight (gradient scale and loss weight)
    dz: torch.Tensor
    dt: Optional[torch.Tensor]
    accumulate: bool


class TrainStep:
    """One fused fwd+loss+bwd step of `net` (a networks.B200Net) at a fixed (B, H, W)."""

    def __init__(self, net, B: int, H: int, W: int, kind: str = "supervised", alpha: float = 0.5,
                 device: Optional[torch.device] = None, dp_group="auto", grad_buckets: int = 4):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.net, self.kind, self.alpha = net, kind, float(alpha)
        with torch.cuda.device(self.device):
            self.eng: StepEngine = net.engine_for(B, H, W, True, self.device)
        eng = self.eng
        outs = eng.outputs
        dev = self.device
        self.targets = {}

        def tgt(key, rows):
            if key not in self.targets:
                self.targets[key] = torch.zeros(rows, 1, H, W, device=dev)
            return self.targets[key]

        def zdz(i):
            hd, sl = outs[i]
            return (hd.logits, hd.dz) if sl is None else (hd.logits[sl], hd.dz[sl])

        self.terms: list[_Term] = []
        if kind == "supervised":
            assert len(outs) == 1, "supervised step expects a single-output network"
            z, dz = zdz(0)
            tgt("y_change", B)
            self.terms = [_Term(z, "y_change", False, None, 1.0, dz, None, False)]
        elif kind == "dualtask":
            assert len(outs) == 3
            for i, (key, w) in enumerate((("y_change", 0.5), ("y_sem_t1", 0.25), ("y_sem_t2", 0.25))):
                z, dz = zdz(i)
                tgt(key, B)
                self.terms.append(_Term(z, key, False, None, w, dz, None, False))
        elif kind == "mmcr":
            assert len(outs) == 3
            tgt("y_change", B)
            a = self.alpha
            for i in range(3):
                z, dz = zdz(i)
                self.terms.append(_Term(z, "y_change", False, 1, a / 3.0, dz, None, False))
            z1, dz1 = zdz(1)
            _, dz2 = zdz(2)
            self.terms.append(_Term(z1, "@logit:2", True, 0, 1.0 - a, dz1, dz2, True))
        else:
            raise ValueError(f"unknown step kind {kind!r}")
        nt = len(self.terms)
        self.rowmask = torch.ones(B, device=dev, dtype=torch.uint8)
        self.weights = torch.tensor([t.weight for t in self.terms], device=dev, dtype=torch.float32)
        self._w_host = [t.weight for t in self.terms]
        self.sums = torch.zeros(nt, 3, device=dev, dtype=torch.float64)
        self.losses = torch.zeros(nt, device=dev, dtype=torch.float32)
        self.nblk = max(1, min(296, (B * H * W) // 4096))
        self.ws = torch.empty(nt, self.nblk * 3, device=dev, dtype=torch.float64)
        self._g_loss_fwd = None
        self._g_loss_bwd = None
        self._steps = 0
        # data parallel
        self.dp = None
        if dp_group == "auto":
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.dp = dist.group.WORLD
        elif dp_group is not None:
            self.dp = dp_group
        self.grad_buckets = max(1, int(__import__("os").environ.get("B200CD_GRAD_BUCKETS", grad_buckets)))
        self._g_dp = None
        self._dp_graph_ok = False
        if self.dp is not None and _DP_GRAPH:
            import torch.distributed as dist
            self._dp_graph_ok = dist.get_backend(self.dp) == "nccl"   # host-side backends (gloo) cannot be captured

    # ------------------------------------------------------------------------------------------------
    def _target_of(self, term: _Term) -> torch.Tensor:
        if term.t_key.startswith("@logit:"):
            hd, sl = self.eng.outputs[int(term.t_key.split(":")[1])]
            return hd.logits if sl is None else hd.logits[sl]
        return self.targets[term.t_key]

    def _loss_fwd(self) -> None:
        for k, term in enumerate(self.terms):
            mask = None if term.sel is None else self.rowmask
            ops.pj_fwd(term.z, self._target_of(term), term.t_is_logit, mask, term.sel or 0, self.nblk, self.ws[k],
                       self.sums[k])

    def _loss_bwd(self) -> None:
        for k, term in enumerate(self.terms):
            mask = None if term.sel is None else self.rowmask
            ops.pj_loss(self.sums[k], self.losses[k:k + 1])
            ops.pj_bwd(term.z, self._target_of(term), term.t_is_logit, mask, term.sel or 0, self.sums[k],
                       self.weights[k:k + 1], 1.0, term.accumulate, term.dz, term.dt)

    def _graphed(self, attr: str, fn) -> None:
        if not self.eng.use_graphs or self._steps < 2:
            fn()
            return
        g = getattr(self, attr)
        if g is None:
            g = PlanGraph(fn)
            setattr(self, attr, g)
        g.replay()

    # ------------------------------------------------------------------------------------------------
    def set_inputs(self, x_t1: torch.Tensor, x_t2: torch.Tensor, is_labeled=None, **targets) -> None:
        """Stage one batch into the static device buffers (host or device tensors; async from pinned memory)."""
        self.eng.x_t1.copy_(x_t1, non_blocking=True)
        self.eng.x_t2.copy_(x_t2, non_blocking=True)
        for k, v in targets.items():
            self.targets[k].copy_(v.reshape(self.targets[k].shape), non_blocking=True)
        if self.kind == "mmcr":
            assert is_labeled is not None
            lab = is_labeled.to(torch.bool)
            n_lab = int(lab.sum().item()) if not lab.is_cuda else None
            if n_lab is None:
                raise ValueError("is_labeled must be a host tensor (train_semisupervised.py:80)")
            has_l, has_u = n_lab > 0, n_lab < lab.numel()
            if self.dp is not None:
                import torch.distributed as dist
                flags = torch.tensor([float(has_l), float(has_u)], device=self.device)
                dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.dp)
                has_l, has_u = bool(flags[0].item()), bool(flags[1].item())
            a = self.alpha
            w = [a / 3.0 * has_l] * 3 + [(1.0 - a) * has_u]
            if w != self._w_host:
                self._w_host = w
                self.weights.copy_(torch.tensor(w, dtype=torch.float32))
            self.rowmask.copy_(lab.to(torch.uint8), non_blocking=True)

    def _dp_step_eager(self) -> None:
        import torch.distributed as dist
        eng = self.eng
        eng._run_fwd_eager()
        self._loss_fwd()
        self._allreduce_sums()
        self._loss_bwd()
        eng.backward_dp(self.dp, self.grad_buckets, skip_allreduce=_DEBUG_SKIP_ALLREDUCE, inner_graphs=False)

    def _allreduce_sums(self) -> None:
        import torch.distributed as dist

        from . import parallel
        if parallel.native_comm() and self.dp is dist.group.WORLD:
            parallel.allreduce_sum_(self.sums)
        else:
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.dp)

    def run(self) -> torch.Tensor:
        """fwd + loss + bwd on the staged batch. Returns the 0-d loss tensor (device)."""
        eng = self.eng
        if self.dp is not None and self._dp_graph_ok and eng.use_graphs and self._steps >= 2:
            if self._g_dp is None:
                self._g_dp = PlanGraph(self._dp_step_eager)
                from . import parallel
                parallel.register_graph_holder(self)
            self._g_dp.replay()
            self._steps += 1
            return (self.losses * self.weights).sum()
        eng.forward_static()
        self._graphed("_g_loss_fwd", self._loss_fwd)
        if self.dp is not None:
            self._allreduce_sums()
        self._graphed("_g_loss_bwd", self._loss_bwd)
        if self.dp is not None:
            eng.backward_dp(self.dp, self.grad_buckets, skip_allreduce=_DEBUG_SKIP_ALLREDUCE)
        else:
            eng.backward_static()
        self._steps += 1
        return (self.losses * self.weights).sum()

    def release_comm_graphs(self) -> None:
        """Drop the whole-step graph (it holds captured NCCL collectives; the communicator cannot be destroyed while it
        exists). The next run() captures again."""
        self._g_dp = None

    def __call__(self, x_t1, x_t2, is_labeled=None, **targets) -> torch.Tensor:
        self.set_inputs(x_t1, x_t2, is_labeled=is_labeled, **targets)
        return self.run()

    def assign_grads(self) -> None:
        """Point every parameter's .grad at its slice of the flat gradient buffer (no copies)."""
        g = self.eng.grads
        for n, p in g.params:
            p.grad = None if n in g.skip else g.views[n]
