"""AdamW as the reference constructs it (`optim.AdamW(net.parameters(), lr=cfg.TRAINER.LR, weight_decay=0.01)`,
train_supervised.py:32, train_semisupervised.py:31, train_supervised_dualtask.py:31), stepped by ONE hand-written
kernel over every parameter tensor (include/b200cd.h: b200cd_adamw_step) instead of torch's foreach/fused path.

`FusedAdamW` subclasses `torch.optim.AdamW`, keeps torch's state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter),
so `optimizer.state_dict()` / `load_state_dict()` and the reference's checkpoint files (utils/networks.py:30-56)
round-trip between the two implementations. Parameters whose `.grad` is None are skipped exactly as torch does
(`outc_sem_change`). There is no CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_JOB = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("start", "<i8"), ("vec4", "<i4"),
                 ("reserved", "<i4")], align=True)
assert _JOB.itemsize == 56


class FusedAdamW(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, foreach=False,
                         fused=False)
        self._tables = {}  # group index -> (pointer signature, device table, njobs, blocks)

    def _table(self, gi: int, items):
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
                    for p, st in items)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == sig:
            return hit[1:]
        arr = np.zeros(len(items), dtype=_JOB)
        blocks = 0
        for i, (p, st) in enumerate(items):
            ptrs = sig[i]
            arr[i] = (*ptrs, p.numel(), blocks, int(all(x % 16 == 0 for x in ptrs)), 0)
            blocks += (p.numel() + 1023) // 1024
        table = torch.from_numpy(arr.view(np.uint8).copy()).to(items[0][0].device)
        self._tables[gi] = (sig, table, len(items), blocks)
        return table, len(items), blocks

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise _lib.B200CDError("FusedAdamW: amsgrad / maximize are not part of the reference configuration")
            items = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.B200CDError("FusedAdamW runs on CUDA parameters only (there is no CPU path)")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or \
                        not p.grad.is_contiguous():
                    raise _lib.B200CDError("FusedAdamW expects contiguous fp32 parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                items.append((p, st))
            if not items:
                continue
            steps = {int(float(st["step"])) for _, st in items}
            if len(steps) != 1:
                raise _lib.B200CDError("FusedAdamW: parameters of one group must share their step count")
            t = steps.pop() + 1
            dev = items[0][0].device
            _lib.init(dev.index)
            table, njobs, blocks = self._table(gi, items)
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                _lib.check(lib.b200cd_adamw_step(table.data_ptr(), njobs, blocks, float(group["lr"]), float(b1), float(b2),
                                                 float(group["eps"]), float(group["weight_decay"]), t,
                                                 torch.cuda.current_stream().cuda_stream))
            for _, st in items:
                st["step"] = st["step"] + 1 if torch.is_tensor(st["step"]) else torch.tensor(float(t))
        return loss
