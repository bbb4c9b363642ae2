"""Host -> device staging for the training loops.

The reference moves every batch with a blocking `.to(device)` right before the forward pass
(train_supervised.py:68-75; DataLoader(pin_memory=True) :37-47). `DevicePrefetcher` wraps any iterable of dict batches
(what `MultimodalCDDataset` + `DataLoader` yield: utils/datasets.py:111-181) and copies batch i+1 into a second set of
device buffers on a side stream while the kernels of batch i run, so the PCIe transfer (54 MB per 16 pairs of 6-band
256x256 patches) disappears behind the step. Tensors that must stay on the host (the `is_labeled` row mask of
train_semisupervised.py:80) and non-tensor entries pass through untouched.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence

import torch


class DevicePrefetcher:
    def __init__(self, batches: Iterable[dict], device: torch.device, keep_on_host: Sequence[str] = ("is_labeled",),
                 depth: int = 2):
        if device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches into CUDA memory; there is no CPU path")
        self.batches, self.device, self.keep, self.depth = batches, device, set(keep_on_host), max(2, depth)
        self.stream = torch.cuda.Stream(device=device)
        self._bufs: list[Optional[dict]] = [None] * self.depth
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._free: list[Optional[torch.cuda.Event]] = [None] * self.depth
        self.bytes_staged = 0

    def _stage(self, slot: int, batch: dict) -> None:
        bufs = self._bufs[slot] or {}
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])   # the consumer of this slot's previous batch has finished
            out = {}
            for k, v in batch.items():
                if not torch.is_tensor(v) or k in self.keep:
                    out[k] = v
                    continue
                b = bufs.get(k)
                if b is None or b.shape != v.shape or b.dtype != v.dtype:
                    b = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                b.copy_(v, non_blocking=True)
                if not v.is_cuda:
                    self.bytes_staged += v.numel() * v.element_size()
                out[k] = b
            self._ready[slot].record(self.stream)
        self._bufs[slot] = out

    def __iter__(self) -> Iterator[dict]:
        it = iter(self.batches)
        try:
            self._stage(0, next(it))
        except StopIteration:
            return
        slot = 0
        while True:
            nxt = next(it, None)
            if nxt is not None:
                self._stage((slot + 1) % self.depth, nxt)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            yield self._bufs[slot]
            ev = torch.cuda.Event()
            ev.record(cur)                                  # everything queued on the batch so far
            self._free[slot] = ev
            if nxt is None:
                return
            slot = (slot + 1) % self.depth


class LossReader:
    """Device -> host read-back of the per-step loss without stalling the launch queue.

    The reference appends `loss.item()` every step (train_supervised.py:79), which drains the GPU before the next step
    can be queued. `push(loss)` instead copies the 0-d loss into a pinned host slot asynchronously and returns the
    value of the step pushed `lag` steps earlier (None until then), so the host keeps queueing work; `drain()` returns
    the values still in flight. Every step's loss is read exactly once."""

    def __init__(self, device: torch.device, lag: int = 1):
        self.lag = max(1, lag)
        self._slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(self.lag + 1)]
        self._events = [torch.cuda.Event() for _ in range(self.lag + 1)]
        self._pending: list[int] = []
        self._n = 0
        self.bytes_read = 0
        self.device = device

    def _take(self, slot: int) -> float:
        self._events[slot].synchronize()
        self.bytes_read += 4
        return float(self._slots[slot])

    def push(self, loss: torch.Tensor) -> Optional[float]:
        slot = self._n % (self.lag + 1)
        self._n += 1
        self._slots[slot].copy_(loss.detach().reshape(()), non_blocking=True)
        self._events[slot].record(torch.cuda.current_stream(self.device))
        self._pending.append(slot)
        if len(self._pending) > self.lag:
            return self._take(self._pending.pop(0))
        return None

    def drain(self) -> list:
        out = [self._take(s) for s in self._pending]
        self._pending = []
        return out


class GpuAugmenter:
    """Training-time augmentation + packing on the GPU (next-row N2): what `MultimodalCDDataset.__getitem__` does per
    sample in numpy (utils/datasets.py:111-181 with utils/augmentations.py:6-142 — ImportanceRandomCrop / UniformCrop,
    RandomFlip, RandomRotate, ColorShift, GammaCorrection, Numpy2Torch and the S1 / S2 channel regrouping) as ONE kernel
    launch per batch (b200cd_augment). The random decisions are drawn on the HOST with numpy in the reference's call
    order (`draw`), so `np.random.seed(s)` reproduces the reference's augmentations sample for sample; the pixel work —
    twelve fp32 channels of a full tile per sample — runs on the device.

        aug = GpuAugmenter(cfg, device)
        params = [aug.draw(change_label_hw1) for each sample]            # host, numpy RNG
        x_t1, x_t2, y_change, y_sem = aug.apply(imgs_dev, buildings_dev, change_dev, params)

    imgs_dev[i]: [H0, W0, 2*|S1| + 2*|S2|] fp32 HWC = concat(s1_t1, s1_t2, s2_t1, s2_t2) as datasets.py:147 builds it;
    buildings_dev[i]: [H0, W0, 2]; change_dev[i]: [H0, W0, 1]. Tiles may differ in size from sample to sample."""

    def __init__(self, cfg, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("GpuAugmenter runs on CUDA devices only; there is no CPU path")
        a = cfg.AUGMENTATION
        self.device = device
        self.crop = int(a.CROP_SIZE)
        self.importance = a.IMAGE_OVERSAMPLING_TYPE != "none"
        self.flip, self.rotate = bool(a.RANDOM_FLIP), bool(a.RANDOM_ROTATE)
        self.color, self.gamma = bool(a.COLOR_SHIFT), bool(a.GAMMA_CORRECTION)
        n1, n2 = len(cfg.DATALOADER.S1_BANDS), len(cfg.DATALOADER.S2_BANDS)
        mode = cfg.DATALOADER.INPUT_MODE
        s1_t1, s1_t2 = list(range(0, n1)), list(range(n1, 2 * n1))
        s2_t1, s2_t2 = list(range(2 * n1, 2 * n1 + n2)), list(range(2 * n1 + n2, 2 * n1 + 2 * n2))
        if mode == "s1":
            self.map_t1, self.map_t2 = s1_t1, s1_t2
        elif mode == "s2":
            self.map_t1, self.map_t2 = s2_t1, s2_t2
        else:
            self.map_t1, self.map_t2 = s1_t1 + s2_t1, s1_t2 + s2_t2
        self.c_img = 2 * n1 + 2 * n2
        assert self.c_img <= 16

    # ---- host: the reference's random decisions, in its numpy call order -------------------------------------
    def draw(self, change_label, np_random=None) -> dict:
        """Draws one sample's augmentation (utils/augmentations.py:105-142, 44-101) from numpy's global RNG (or the
        RandomState given). `change_label`: numpy [H0, W0, 1] — only its crop sums are needed (importance sampling)."""
        import numpy as np
        rs = np.random if np_random is None else np_random
        H0, W0 = change_label.shape[:2]
        cs = self.crop

        def random_crop():
            x = rs.randint(0, W0 - cs)
            y = rs.randint(0, H0 - cs)
            return x, y

        if self.importance:
            crops = [random_crop() for _ in range(20)]
            # label sums in float32 with numpy's own reduction: bit-identical weights, so the same candidate is chosen
            weights = np.array([change_label[y:y + cs, x:x + cs, ].sum() for x, y in crops]) + 5
            weights = weights / weights.sum()
            x0, y0 = crops[rs.choice(20, p=weights)]
        else:
            x0, y0 = random_crop()
        p = {"x0": int(x0), "y0": int(y0), "hflip": False, "vflip": False, "rotk": 0, "mul": None, "mul_b": None,
             "gamma": None, "gamma_b": None}
        if self.flip:
            p["hflip"] = bool(rs.choice([True, False]))
            p["vflip"] = bool(rs.choice([True, False]))
        if self.rotate:
            p["rotk"] = int(rs.randint(1, 4))
        if self.color:     # ColorShift draws for its first AND its second tuple member (images, then the building labels)
            p["mul"] = rs.uniform(0.5, 1.5, self.c_img)
            p["mul_b"] = rs.uniform(0.5, 1.5, 2)
        if self.gamma:
            p["gamma"] = rs.uniform(0.25, 2, self.c_img)
            p["gamma_b"] = rs.uniform(0.25, 2, 2)
        return p

    # ---- device -------------------------------------------------------------------------------------------
    def _jobs(self, srcs, params, cmap, which):
        import numpy as np
        dt = np.dtype([("src", "<u8"), ("H0", "<i4"), ("W0", "<i4"), ("C", "<i4"), ("x0", "<i4"), ("y0", "<i4"),
                       ("hflip", "<i4"), ("vflip", "<i4"), ("rotk", "<i4"), ("use_mul", "<i4"), ("use_gamma", "<i4"),
                       ("cmap", "<i4", 16), ("mul", "<f4", 16), ("gamma", "<f4", 16), ("reserved", "<i4")], align=True)
        arr = np.zeros(len(srcs), dtype=dt)
        for i, (t, p) in enumerate(zip(srcs, params)):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 3
            H0, W0, C = t.shape
            assert 0 <= p["x0"] <= W0 - self.crop and 0 <= p["y0"] <= H0 - self.crop
            mul = p["mul"] if which == "img" else (p["mul_b"] if which == "bld" else None)
            gam = p["gamma"] if which == "img" else (p["gamma_b"] if which == "bld" else None)
            cm = np.zeros(16, np.int32)
            cm[:len(cmap)] = cmap
            m, g = np.ones(16, np.float32), np.ones(16, np.float32)
            if mul is not None:
                m[:len(mul)] = mul
            if gam is not None:
                g[:len(gam)] = gam
            arr[i] = (t.data_ptr(), H0, W0, C, p["x0"], p["y0"], int(p["hflip"]), int(p["vflip"]), p["rotk"],
                      int(mul is not None), int(gam is not None), cm, m, g, 0)
        return torch.from_numpy(arr.view(np.uint8).copy()).to(self.device, non_blocking=True)

    def _run(self, srcs, params, cmap, which) -> torch.Tensor:
        from . import _lib, ops
        n, cs = len(srcs), self.crop
        out = torch.empty(n, len(cmap), cs, cs, device=self.device, dtype=torch.float32)
        table = self._jobs(srcs, params, cmap, which)
        _lib.init(self.device.index)
        ops._count(1)
        _lib.check(_lib.load().b200cd_augment(table.data_ptr(), n, cs, len(cmap), out.data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream))
        out._b200cd_keepalive = table     # the job table must outlive the asynchronous launch
        return out

    def apply(self, imgs, buildings, change, params):
        """-> (x_t1, x_t2, y_change, y_sem) as NCHW fp32 device tensors: [n, C, crop, crop], [n, 1, ...], [n, 2, ...]
        (y_sem[:, 0:1] = y_sem_t1, y_sem[:, 1:2] = y_sem_t2, utils/datasets.py:175-178)."""
        with torch.cuda.device(self.device):
            x = self._run(imgs, params, self.map_t1 + self.map_t2, "img")
            y = self._run(change, params, [0], "lbl")
            s = self._run(buildings, params, [0, 1], "bld") if buildings is not None else None
        k = len(self.map_t1)
        return x[:, :k], x[:, k:], y, s
