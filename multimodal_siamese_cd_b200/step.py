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
from .engine import StepEngine


_DEBUG_SKIP_ALLREDUCE = __import__("os").environ.get("B200CD_DEBUG_SKIP_ALLREDUCE", "0") == "1"


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
        self._comm_stream = torch.cuda.Stream(device=dev) if self.dp is not None else None
        self._bucket_plan = None
        self._bwd_graphs = None

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
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            setattr(self, attr, g)
        g.replay()

    # ------------------------------------------------------------------------------------------------
    def _plan_buckets(self):
        """Split the backward ops into contiguous segments of roughly equal gradient volume."""
        marks = self.eng.bwd_marks
        total = marks[-1]
        nb = min(self.grad_buckets, len(marks))
        cuts, lo = [], 0
        for b in range(1, nb + 1):
            want = total * b // nb
            i = next(i for i, m in enumerate(marks) if m >= want)
            i = max(i, cuts[-1][1] if cuts else 0)
            cuts.append((cuts[-1][1] if cuts else 0, i + 1, lo, marks[i]))
            lo = marks[i]
        # (op_begin, op_end, grad_lo, grad_hi); drop empty segments
        return [c for c in cuts if c[1] > c[0]]

    def _backward_dp(self) -> None:
        """Backward in segments; each finished gradient prefix is all-reduced (SUM) on a side stream while the
        next segment runs."""
        import torch.distributed as dist
        eng = self.eng
        if self._bucket_plan is None:
            self._bucket_plan = self._plan_buckets()
            self._bwd_graphs = [None] * len(self._bucket_plan)
        main = torch.cuda.current_stream()
        for f in eng.pack_bwd:
            f()
        for bi, (o0, o1, g0, g1) in enumerate(self._bucket_plan):
            def seg(o0=o0, o1=o1):
                eng.run_bwd_range(o0, o1)   # forks / joins the second trunk's stream inside the segment
            if eng.use_graphs and self._steps >= 2:
                if self._bwd_graphs[bi] is None:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        seg()
                    self._bwd_graphs[bi] = g
                self._bwd_graphs[bi].replay()
            else:
                seg()
            ev = torch.cuda.Event()
            ev.record(main)
            self._comm_stream.wait_event(ev)
            if _DEBUG_SKIP_ALLREDUCE:   # measurement aid only (B200CD_DEBUG_SKIP_ALLREDUCE=1): wrong gradients
                continue
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(eng.grads.flat[g0:g1], op=dist.ReduceOp.SUM, group=self.dp)
        main.wait_stream(self._comm_stream)

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

    def run(self) -> torch.Tensor:
        """fwd + loss + bwd on the staged batch. Returns the 0-d loss tensor (device)."""
        eng = self.eng
        eng.forward_static()
        self._graphed("_g_loss_fwd", self._loss_fwd)
        if self.dp is not None:
            import torch.distributed as dist
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.dp)
        self._graphed("_g_loss_bwd", self._loss_bwd)
        if self.dp is not None:
            self._backward_dp()
            eng._runs += 1
        else:
            eng.backward_static()
        self._steps += 1
        return (self.losses * self.weights).sum()

    def __call__(self, x_t1, x_t2, is_labeled=None, **targets) -> torch.Tensor:
        self.set_inputs(x_t1, x_t2, is_labeled=is_labeled, **targets)
        return self.run()

    def assign_grads(self) -> None:
        """Point every parameter's .grad at its slice of the flat gradient buffer (no copies)."""
        g = self.eng.grads
        for n, p in g.params:
            p.grad = None if n in g.skip else g.views[n]
