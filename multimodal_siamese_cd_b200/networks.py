"""Drop-in for the reference's `utils/networks.py`: same config-driven constructors, class and attribute names,
`state_dict` keys (incl. the `module.` prefix), default initialisation order and `forward(x_t1, x_t2)` contract —
but the modules below only HOLD parameters; the arithmetic runs in the sm_100a kernels through `StepEngine`.

Reference surface mirrored (utils/networks.py): create_network :12-27, save_checkpoint :30-38, load_checkpoint :41-56,
UNet :59, DualStreamUNet :82, SiameseUNet :123, DualTaskSiameseUNet :157, WhateverNet :200, WhateverNet2 :266,
Encoder :313, Decoder :346, DoubleConv :386, InConv :405, Down :415, Up :429, OutConv :454.

There is no CPU path: calling a network on CPU tensors raises.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from pathlib import Path

import torch
import torch.nn as nn

from . import parallel
from .engine import StepEngine

__all__ = ["create_network", "save_checkpoint", "load_checkpoint", "UNet", "DualStreamUNet", "SiameseUNet",
           "DualTaskSiameseUNet", "WhateverNet", "WhateverNet2", "Encoder", "Decoder", "DoubleConv", "InConv", "Down",
           "Up", "OutConv", "DataParallelShim"]


# ------------------------------------------------------------------------------------------------------
# parameter containers. Layer objects are real torch.nn layers created in the reference's order, so that
# torch.manual_seed(cfg.SEED) + construction yields bit-identical initial weights and optim.AdamW /
# load_state_dict / .to() behave as usual. Their own forward() is never used.
# ------------------------------------------------------------------------------------------------------
class DoubleConv(nn.Module):
    """Two (3x3 conv, BatchNorm, ReLU) stages; indices 0,1,3,4 of `conv` carry parameters."""

    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        layers = []
        for ci in (in_ch, out_ch):
            layers += [nn.Conv2d(ci, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True)]
        self.conv = nn.Sequential(*layers)


class InConv(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, conv_block=DoubleConv):
        super().__init__()
        self.conv = conv_block(in_ch, out_ch)


class Down(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, conv_block=DoubleConv):
        super().__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), conv_block(in_ch, out_ch))


class Up(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, conv_block=DoubleConv):
        super().__init__()
        half = in_ch // 2
        self.up = nn.ConvTranspose2d(half, half, 2, stride=2)
        self.conv = conv_block(in_ch, out_ch)


class OutConv(nn.Module):
    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, 1)


def _stage_widths(topology) -> list[int]:
    """Output width of every encoder stage; the last stage keeps its width (utils/networks.py:326-330)."""
    topo = list(topology)
    return topo[1:] + topo[-1:]


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        topo = list(cfg.MODEL.TOPOLOGY)
        stages = OrderedDict()
        for i, (cin, cout) in enumerate(zip(topo, _stage_widths(topo)), start=1):
            stages[f"down{i}"] = Down(cin, cout, DoubleConv)
        self.down_seq = nn.ModuleDict(stages)


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        topo = list(cfg.MODEL.TOPOLOGY)
        widths = topo[:1] + _stage_widths(topo)          # width of the feature map at every depth, shallow to deep
        stages = OrderedDict()
        for depth in range(len(topo), 0, -1):            # up{L} .. up1, deepest first (utils/networks.py:364-371)
            below = widths[depth - 1]
            above = widths[depth - 2] if depth > 1 else widths[0]
            stages[f"up{depth}"] = Up(2 * below, above, DoubleConv)
        self.up_seq = nn.ModuleDict(stages)


# ------------------------------------------------------------------------------------------------------
class _StepFunction(torch.autograd.Function):
    """Whole-network autograd node: forward and backward are the engine's (graph-replayed) kernel plans."""

    @staticmethod
    def forward(ctx, eng: StepEngine, x_t1, x_t2, *params):
        eng.forward(x_t1, x_t2)
        ctx.eng = eng
        ctx.n_params = len(params)
        return tuple(o.clone() for o in eng.output_tensors())

    @staticmethod
    def backward(ctx, *grad_outs):
        eng: StepEngine = ctx.eng
        touched = set()
        for (hd, sl), g in zip(eng.outputs, grad_outs):
            tgt = hd.dz if sl is None else hd.dz[sl]
            if g is None:
                tgt.zero_()
            else:
                tgt.copy_(g.reshape(tgt.shape))
            touched.add(id(hd))
        eng.backward_static()
        parallel.allreduce_gradients(eng.grads.flat)   # SUM over replicas when data parallelism is enabled
        return (None, None, None, *eng.grads.detached_copy_views())


class B200Net(nn.Module):
    """Common forward for every network type: route to the engine for this (batch, size, mode)."""

    _MAX_ENGINES = 2

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self._engines: "OrderedDict[tuple, StepEngine]" = OrderedDict()
        self.use_cuda_graphs = True
        # "fast": single-bf16 storage / operands (the throughput mode). "precise": split-bf16 storage and three-MMA
        # products, which meets the reference-fp32 tolerance. Default from B200CD_PRECISION.
        self.precision = os.environ.get("B200CD_PRECISION", "fast")

    def set_precision(self, mode: str) -> "B200Net":
        if mode not in ("fast", "precise"):
            raise ValueError(f"precision must be 'fast' or 'precise' (got {mode!r})")
        self.precision = mode
        return self

    def engine_for(self, B: int, H: int, W: int, train: bool, device: torch.device) -> StepEngine:
        if self.precision not in ("fast", "precise"):
            raise ValueError(f"B200CD_PRECISION / net.precision must be 'fast' or 'precise' (got {self.precision!r})")
        key = (B, H, W, train, device.index, self.precision)
        eng = self._engines.get(key)
        if eng is not None and eng.params_moved():
            del self._engines[key]
            eng = None
        if eng is None:
            while len(self._engines) >= self._MAX_ENGINES:
                self._engines.popitem(last=False)
            with torch.cuda.device(device):
                eng = StepEngine(self, B, H, W, train, device, use_graphs=self.use_cuda_graphs,
                                 precise=self.precision == "precise")
            self._engines[key] = eng
        else:
            self._engines.move_to_end(key)
        return eng

    def release_engines(self) -> None:
        self._engines.clear()

    def _outputs(self, outs: tuple):
        return outs[0] if len(outs) == 1 else tuple(outs)

    def forward(self, x_t1: torch.Tensor, x_t2: torch.Tensor):
        if not (x_t1.is_cuda and x_t2.is_cuda):
            raise RuntimeError("multimodal_siamese_cd_b200 networks run on CUDA (sm_100a) tensors only; "
                               "there is no CPU fallback")
        x_t1 = x_t1.float().contiguous()
        x_t2 = x_t2.float().contiguous()
        B, _, H, W = x_t1.shape
        with torch.cuda.device(x_t1.device):
            eng = self.engine_for(B, H, W, self.training, x_t1.device)
            if self.training and torch.is_grad_enabled():
                outs = _StepFunction.apply(eng, x_t1, x_t2, *[p for _, p in eng.grads.params])
            else:
                eng.forward(x_t1, x_t2)
                outs = tuple(o.clone() for o in eng.output_tensors())
        return self._outputs(outs)

    # engines hold device buffers and graphs: never part of a state_dict / deepcopy / pickle
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engines"] = OrderedDict()
        return d


def _in_conv(n_in: int, cfg) -> InConv:
    return InConv(n_in, cfg.MODEL.TOPOLOGY[0], DoubleConv)


class UNet(B200Net):
    """Early fusion: cat(x_t1, x_t2) -> one U-Net (utils/networks.py:59-79)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.inc = _in_conv(2 * cfg.MODEL.IN_CHANNELS, cfg)
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.outc = OutConv(cfg.MODEL.TOPOLOGY[0], cfg.MODEL.OUT_CHANNELS)


class DualStreamUNet(B200Net):
    """One early-fusion U-Net per modality, decoder outputs concatenated into one head (utils/networks.py:82-120)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        for k, bands in ((1, cfg.DATALOADER.S1_BANDS), (2, cfg.DATALOADER.S2_BANDS)):
            setattr(self, f"inc_stream{k}", _in_conv(2 * len(bands), cfg))
            setattr(self, f"encoder_stream{k}", Encoder(cfg))
            setattr(self, f"decoder_stream{k}", Decoder(cfg))
        self.outc = OutConv(2 * cfg.MODEL.TOPOLOGY[0], cfg.MODEL.OUT_CHANNELS)


class SiameseUNet(B200Net):
    """Shared-weight encoder on t1 and t2, decoder on the feature differences (utils/networks.py:123-154)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.inc = _in_conv(cfg.MODEL.IN_CHANNELS, cfg)
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.outc = OutConv(cfg.MODEL.TOPOLOGY[0], cfg.MODEL.OUT_CHANNELS)


class DualTaskSiameseUNet(B200Net):
    """Siamese trunk + change decoder + semantic decoder applied to t2 then t1 (utils/networks.py:157-197).
    Returns (change, sem_t1, sem_t2). `outc_sem_change` exists for state_dict parity and is never used."""

    def __init__(self, cfg):
        super().__init__(cfg)
        w0, nout = cfg.MODEL.TOPOLOGY[0], cfg.MODEL.OUT_CHANNELS
        self.inc = _in_conv(cfg.MODEL.IN_CHANNELS, cfg)
        self.encoder = Encoder(cfg)
        self.decoder_change = Decoder(cfg)
        self.decoder_sem = Decoder(cfg)
        self.outc_change = OutConv(w0, nout)
        self.outc_sem = OutConv(w0, nout)
        self.outc_sem_change = OutConv(2, 1)


class _TwoStream(B200Net):
    def __init__(self, cfg, fused_inputs: bool):
        super().__init__(cfg)
        w0, nout = cfg.MODEL.TOPOLOGY[0], cfg.MODEL.OUT_CHANNELS
        mult = 2 if fused_inputs else 1
        for k, bands in ((1, cfg.DATALOADER.S1_BANDS), (2, cfg.DATALOADER.S2_BANDS)):
            setattr(self, f"inc_stream{k}", _in_conv(mult * len(bands), cfg))
            setattr(self, f"encoder_stream{k}", Encoder(cfg))
            setattr(self, f"decoder_stream{k}", Decoder(cfg))
            setattr(self, f"outc_stream{k}", OutConv(w0, nout))
        self.outc_fusion = OutConv(2 * w0, nout)

    def _outputs(self, outs: tuple):
        # training: (fusion, stream1, stream2); eval: fusion only (utils/networks.py:260-263, 307-310)
        return tuple(outs) if self.training else outs[0]


class WhateverNet(_TwoStream):
    """A siamese U-Net per modality (SAR, optical) + per-stream heads + fusion head (utils/networks.py:200-263)."""

    def __init__(self, cfg):
        super().__init__(cfg, fused_inputs=False)


class WhateverNet2(_TwoStream):
    """An early-fusion U-Net per modality + per-stream heads + fusion head (utils/networks.py:266-310)."""

    def __init__(self, cfg):
        super().__init__(cfg, fused_inputs=True)


# ------------------------------------------------------------------------------------------------------
class DataParallelShim(nn.Module):
    """What `create_network` returns instead of nn.DataParallel (utils/networks.py:27): same `.module` attribute and
    `module.`-prefixed state_dict, but no per-step parameter broadcast / scatter / gather. Multi-GPU data parallelism
    is one process per GPU (see parallel.py): persistent replicas, NCCL all-reduce(SUM) of the gradients."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


_TYPES = {
    "unet": UNet,
    "dualstreamunet": DualStreamUNet,
    "siameseunet": SiameseUNet,
    "dtsiameseunet": DualTaskSiameseUNet,
    "whatevernet": WhateverNet,
    "whatevernet2": WhateverNet2,
}


def create_network(cfg):
    try:
        cls = _TYPES[cfg.MODEL.TYPE]
    except KeyError:
        raise Exception(f"Unknown network ({cfg.MODEL.TYPE}).") from None
    return DataParallelShim(cls(cfg))


def save_checkpoint(network, optimizer, epoch, step, cfg):
    target = Path(cfg.PATHS.OUTPUT) / "networks" / f"{cfg.NAME}_checkpoint{epoch}.pt"
    target.parent.mkdir(exist_ok=True)
    torch.save({"step": step, "network": network.state_dict(), "optimizer": optimizer.state_dict()}, target)


def load_checkpoint(epoch, cfg, device, net_file: Path = None):
    net = create_network(cfg)
    net.to(device)
    source = net_file if net_file is not None else Path(cfg.PATHS.OUTPUT) / "networks" / f"{cfg.NAME}_checkpoint{epoch}.pt"
    state = torch.load(source, map_location=device)
    optimizer = torch.optim.AdamW(net.parameters(), lr=cfg.TRAINER.LR, weight_decay=0.01)
    net.load_state_dict(state["network"])
    optimizer.load_state_dict(state["optimizer"])
    return net, optimizer, state["step"]
