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
from typing import List, Sequence

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
# torch.library custom ops of the network path (namespace b200cd::): the whole forward plan and the whole
# backward plan of a StepEngine as two ops with fake (meta) implementations, wired together with register_autograd.
# The engine — static buffers, TMA descriptors baked into CUDA graphs — is passed as an integer token.
# ------------------------------------------------------------------------------------------------------
_ENGINE_TOKENS: "dict[int, StepEngine]" = {}
_NEXT_TOKEN = [1]


def _engine_token(eng: StepEngine) -> int:
    tok = getattr(eng, "_token", None)
    if tok is None:
        tok = _NEXT_TOKEN[0]
        _NEXT_TOKEN[0] += 1
        eng._token = tok
        _ENGINE_TOKENS[tok] = eng
    return tok


def _engine_of(token: int) -> StepEngine:
    try:
        return _ENGINE_TOKENS[token]
    except KeyError:
        raise RuntimeError(f"b200cd::network_step: engine token {token} is not alive (released or evicted)") from None


@torch.library.custom_op("b200cd::network_step", mutates_args=(), device_types="cuda")
def network_step(engine: int, x_t1: torch.Tensor, x_t2: torch.Tensor, params: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Forward plan of engine `engine` on (x_t1, x_t2) fp32 NCHW -> the network's logits (fp32 NCHW, one per output).
    `params` are the network's parameters (read through the engine's own views; listed so that autograd tracks them)."""
    eng = _engine_of(engine)
    with torch.cuda.device(x_t1.device):
        eng.forward(x_t1, x_t2)
        if eng.train:
            eng.generation += 1
    return [o.clone() for o in eng.output_tensors()]


@network_step.register_fake
def _network_step_fake(engine, x_t1, x_t2, params):
    eng = _engine_of(engine)
    return [x_t1.new_empty(tuple(o.shape), dtype=torch.float32) for o in eng.output_tensors()]


@torch.library.custom_op("b200cd::network_step_backward", mutates_args=(), device_types="cuda")
def network_step_backward(engine: int, generation: int, grad_outs: Sequence[torch.Tensor]) -> torch.Tensor:
    """Backward plan of engine `engine` from the logit gradients -> ONE fresh flat fp32 buffer holding every parameter
    gradient (GradArena layout), all-reduced (SUM) across the data-parallel group when one is enabled."""
    eng = _engine_of(engine)
    if eng.generation != generation:
        raise RuntimeError(
            "b200cd: backward() of a network call whose engine has run another training forward since "
            f"(forward #{generation}, engine is at #{eng.generation}). The engine keeps ONE set of saved activations "
            "per (batch, size): call backward() before the next forward of the same shape, or sum the losses of one "
            "forward.")
    with torch.cuda.device(eng.device):
        for (hd, sl), g in zip(eng.outputs, grad_outs):
            tgt = hd.dz if sl is None else hd.dz[sl]
            tgt.copy_(g.reshape(tgt.shape))
        if parallel.is_enabled():
            eng.backward_dp(parallel.group(), _GRAD_BUCKETS)
        else:
            eng.backward_static()
    return eng.grads.flat.clone()


@network_step_backward.register_fake
def _network_step_backward_fake(engine, generation, grad_outs):
    eng = _engine_of(engine)
    return grad_outs[0].new_empty((eng.grads.flat.numel(),), dtype=torch.float32)


_GRAD_BUCKETS = int(os.environ.get("B200CD_GRAD_BUCKETS", 4))


def _network_step_setup(ctx, inputs, output):
    engine, x_t1, x_t2, params = inputs
    ctx.engine = engine
    ctx.generation = _engine_of(engine).generation
    ctx.ref = x_t1


def _network_step_bwd(ctx, grads):
    eng = _engine_of(ctx.engine)
    outs = eng.output_tensors()
    gl = [g if g is not None else torch.zeros_like(o) for g, o in zip(grads, outs)]
    flat = torch.ops.b200cd.network_step_backward(ctx.engine, ctx.generation, gl)
    # per-parameter views of the ONE fresh buffer: autograd's AccumulateGrad adopts a gradient it holds the only
    # reference to without a copy kernel (~160 launches per DualStream step otherwise)
    ga = eng.grads
    pg = []
    for n, p in ga.params:
        if n in ga.skip:
            pg.append(None)          # outc_sem_change: grad stays None as in the reference (AdamW skips it)
        else:
            o = ga.offsets[n]
            pg.append(flat[o:o + p.numel()].view(p.shape))
    return None, None, None, pg


torch.library.register_autograd("b200cd::network_step", _network_step_bwd, setup_context=_network_step_setup)


class B200Net(nn.Module):
    """Common forward for every network type: route to the engine for this (batch, size, mode)."""

    # engines are cached per (batch, H, W, mode, device, precision). Training engines and inference engines have
    # separate LRU lists: utils/evaluation.py feeds whole tiles of varying size at batch 1 every LOG_FREQ steps, and
    # that must never evict (and so rebuild / re-capture) the training engine.
    _MAX_TRAIN_ENGINES = 2
    _MAX_EVAL_ENGINES = 8

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self._engines: "OrderedDict[tuple, StepEngine]" = OrderedDict()
        self.use_cuda_graphs = os.environ.get("B200CD_CUDA_GRAPHS", "1") != "0"
        # "fast": single-bf16 storage / operands (the throughput mode). "precise": split-bf16 storage and three-MMA
        # products, which meets the reference-fp32 tolerance. Default from B200CD_PRECISION.
        self.precision = os.environ.get("B200CD_PRECISION", "fast")
        self._check_shapes()

    def _check_shapes(self) -> None:
        """The kernels' channel constraints, reported at construction (a warning: such a network can still be built,
        checkpointed and loaded — it cannot run; the engine raises at the first forward)."""
        topo = list(self.cfg.MODEL.TOPOLOGY)
        if any(c % 64 != 0 for c in topo) or int(getattr(self.cfg.MODEL, "OUT_CHANNELS", 1)) != 1:
            import warnings
            warnings.warn(f"b200cd networks run MODEL.TOPOLOGY widths that are multiples of 64 and OUT_CHANNELS == 1 "
                          f"(got {topo}, {getattr(self.cfg.MODEL, 'OUT_CHANNELS', 1)}): this network can hold and "
                          "load parameters but its forward will raise", stacklevel=3)

    def set_precision(self, mode: str) -> "B200Net":
        if mode not in ("fast", "precise"):
            raise ValueError(f"precision must be 'fast' or 'precise' (got {mode!r})")
        self.precision = mode
        return self

    def _drop(self, key) -> None:
        eng = self._engines.pop(key)
        _ENGINE_TOKENS.pop(getattr(eng, "_token", None), None)

    def engine_for(self, B: int, H: int, W: int, train: bool, device: torch.device) -> StepEngine:
        if self.precision not in ("fast", "precise"):
            raise ValueError(f"B200CD_PRECISION / net.precision must be 'fast' or 'precise' (got {self.precision!r})")
        key = (B, H, W, train, device.index, self.precision)
        eng = self._engines.get(key)
        if eng is not None and eng.params_moved():
            self._drop(key)
            eng = None
        if eng is None:
            same = [k for k in self._engines if k[3] == train]
            cap = self._MAX_TRAIN_ENGINES if train else self._MAX_EVAL_ENGINES
            while len(same) >= cap:
                self._drop(same.pop(0))
            with torch.cuda.device(device):
                eng = StepEngine(self, B, H, W, train, device, use_graphs=self.use_cuda_graphs,
                                 precise=self.precision == "precise")
            self._engines[key] = eng
        else:
            self._engines.move_to_end(key)
        return eng

    def release_engines(self) -> None:
        for key in list(self._engines):
            self._drop(key)

    def _outputs(self, outs: tuple):
        return outs[0] if len(outs) == 1 else tuple(outs)

    def forward(self, x_t1: torch.Tensor, x_t2: torch.Tensor):
        if not (x_t1.is_cuda and x_t2.is_cuda):
            raise RuntimeError("multimodal_siamese_cd_b200 networks run on CUDA (sm_100a) tensors only; "
                               "there is no CPU fallback")
        x_t1 = x_t1.float().contiguous()
        x_t2 = x_t2.float().contiguous()
        B, _, H, W = x_t1.shape
        gather = self.training and parallel.gather_mode()
        if gather:
            # replicated inputs (unchanged reference scripts under torchrun): this rank runs its DataParallel chunk
            rank, world = parallel.rank_world()
            rows = parallel.shard_rows(B, rank, world)
            if rows.stop <= rows.start:
                raise RuntimeError(f"b200cd data parallel: batch of {B} rows leaves rank {rank} of {world} without work")
            x_t1, x_t2 = x_t1[rows].contiguous(), x_t2[rows].contiguous()
        with torch.cuda.device(x_t1.device):
            eng = self.engine_for(x_t1.shape[0], H, W, self.training, x_t1.device)
            if self.training and torch.is_grad_enabled():
                outs = torch.ops.b200cd.network_step(_engine_token(eng), x_t1, x_t2, [p for _, p in eng.grads.params])
            else:
                eng.forward(x_t1, x_t2)
                outs = [o.clone() for o in eng.output_tensors()]
        if gather:
            outs = [parallel.GatherRows.apply(o, B) for o in outs]
        return self._outputs(tuple(outs))

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
