"""Static execution plan of one training (or inference) step of a U-Net-family network on one B200.

The plan is built once per (network, batch, H, W, train/eval): every activation, gradient and workspace
buffer is allocated up front (NHWC bf16, torch's caching allocator owns the memory), every kernel launch is
a closure over those buffers calling the C ABI (include/b200cd.h) on the current stream, and after one eager
run forward and backward are captured into CUDA graphs and replayed. The t1/t2 calls of a shared-weight
encoder (utils/networks.py:141-145) are ONE launch over 2B images with two BatchNorm stat-groups; skip
connections, the t2 - t1 difference, max-pool and the transposed-conv output are written straight into the
buffers their consumers read (no cat / pad / sub kernels).

Reference semantics preserved: separate batch statistics and two sequential running-stat
updates per shared BatchNorm (decoder_sem: t2 first), pre-BN conv biases get exact-zero gradients,
`outc_sem_change` gets no gradient at all.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional

import torch
import torch.nn as nn

from . import ops, parallel

BF16 = torch.bfloat16

# weight-gradient kernels split the pixel dimension until about this many CTAs are in flight (the per-split partial
# results are summed by wgrad_reduce in a fixed order); measured on B200 in profiles/
WGRAD_CTA_TARGET = int(__import__("os").environ.get("B200CD_WGRAD_CTAS", 148))
# BatchNorm backward of stages with a single direct gradient source: accumulate its reduce pass in the epilogue of the
# input-gradient convolution that produces that gradient (ops.conv_gemm_bnbwd)
FUSE_BN_BWD_REDUCE = __import__("os").environ.get("B200CD_FUSE_BN_BWD", "1") != "0"
# Weight-gradient GEMMs run on their own stream (forked after the layer's input-gradient GEMM, joined before each batched
# reduce): nothing downstream of a layer needs its weight gradient, so it fills the ramp-up / tail / launch gaps of the
# dependent BatchNorm-backward -> dgrad chain. Same-box A/B: Siamese B=8 3.83 -> 3.78 ms (+1.4 %), DTSiamese B=8 +0.5 %,
# DualStream B=16 (already two trunk streams) +0.5 %. B200CD_WGRAD_SIDE_STREAM=0 turns it off.
# The two U-Net trunks of DualStreamUNet / WhateverNet / WhateverNet2 are independent until the fusion head: with
# B200CD_BRANCH_STREAMS=1 (default) the second trunk's launches go to a second stream (forked after the weight packing /
# the head gradients, joined before the head / every batched split reduce), so the ramp-up, tail and launch gaps of one
# trunk's kernels are filled by the other's. Each trunk has its own statistics / reduction workspaces.
# Measured (same box): DualStream B=16 9.95 -> 9.62 ms (+3.5 %, value and e2e), WhateverNet B=16 +3.7 %, WhateverNet
# B=64 -0.9 % (long launches have no tails worth filling, and two concurrent working sets share the L2) — so "1" enables
# it only below BRANCH_MAX_PIXELS full-resolution pixels per trunk launch; "2" always, "0" never.
_BRANCH_MODE = __import__("os").environ.get("B200CD_BRANCH_STREAMS", "1")
BRANCH_STREAMS = _BRANCH_MODE != "0"
BRANCH_MAX_PIXELS = 4 << 20
WGRAD_SIDE_STREAM = __import__("os").environ.get("B200CD_WGRAD_SIDE_STREAM", "1") != "0"
# Stream priorities: the dependent chain of a plan (convolutions, BatchNorm passes, input gradients — both trunk
# streams) runs on HIGH-priority streams, the weight-gradient GEMMs that only have to finish by the end of a backward
# segment on a default-priority one, so whenever both have a kernel ready the chain's CTAs are dispatched first. The
# priorities are kernel-node attributes inside the captured graphs too.
# tail flush of the split-K reduction when at most 1 / TAIL_FLUSH_DIV of the weight-gradient volume is left (0 = off)
TAIL_FLUSH_DIV = int(__import__("os").environ.get("B200CD_TAIL_FLUSH_DIV", "24"))
STREAM_PRIO = __import__("os").environ.get("B200CD_STREAM_PRIO", "1") != "0"
# one weight-gradient side stream per trunk of a two-trunk plan (else both trunks' weight gradients share one)
WGRAD_SIDE_PER_BRANCH = __import__("os").environ.get("B200CD_WGRAD_SIDE_PER_BRANCH", "0") != "0"
_PRIO_WGRAD_HIGH = __import__("os").environ.get("B200CD_STREAM_PRIO", "1") == "2"   # experiment: the reverse assignment
# Graphs instantiated and launched by libb200cd (b200cd_graph_instantiate, include/b200cd.h) with
# cudaGraphInstantiateFlagUseNodePriority, so the stream priorities above survive inside a replayed plan: torch's own
# instantiate passes no such flag and every node would run at the launch stream's priority.
GRAPH_NODE_PRIO = __import__("os").environ.get("B200CD_GRAPH_NODE_PRIO", "1") != "0"


class PlanGraph:
    """One captured plan: torch.cuda.graph captures (thread_local error mode: a DataLoader pin-memory thread calling
    cudaHostAlloc / event functions during a global-mode capture would invalidate it), the library instantiates and
    launches when node priorities are wanted; else torch replays."""

    def __init__(self, fn):
        import ctypes as C

        from . import _lib
        self._exec = None
        native = GRAPH_NODE_PRIO and STREAM_PRIO
        torch.cuda.synchronize()
        try:
            self.g = torch.cuda.CUDAGraph(keep_graph=True) if native else torch.cuda.CUDAGraph()
        except TypeError:                      # a torch without keep_graph: torch instantiates and replays
            native = False
            self.g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g, capture_error_mode="thread_local"):
            fn()
        if native:
            out = C.c_void_p()
            _lib.check(_lib.load().b200cd_graph_instantiate(C.c_void_p(int(self.g.raw_cuda_graph())), 1, C.byref(out)))
            self._exec = out.value

    def replay(self) -> None:
        if self._exec is not None:
            from . import _lib
            _lib.check(_lib.load().b200cd_graph_launch(self._exec, torch.cuda.current_stream().cuda_stream))
        else:
            self.g.replay()

    def __del__(self):
        if getattr(self, "_exec", None) is not None:
            try:
                from . import _lib
                _lib.load().b200cd_graph_exec_destroy(self._exec)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass
            self._exec = None
# transposed-conv bias gradient from the per-CTA channel sums of the dgrad launch that writes the concat-buffer gradient
UP_BIAS_FROM_STATS = __import__("os").environ.get("B200CD_UP_BIAS_FROM_STATS", "1") != "0"
FUSE_BN_BWD_MIN_PIXELS = int(__import__("os").environ.get("B200CD_FUSE_BN_BWD_MIN_PIXELS", 32768))
# 64-wide weight-gradient tiles: one CTA may own two kx columns (gemm_wgrad.cu, NKX = 2). Measured slower than one kx
# per CTA once tap pairs are issued as N = 128 MMAs (the kernel is bound by MN-major operand reads from shared memory,
# not by the L2 -> SM stream): off by default, kept for the parity tests and for re-measurement.
WGRAD_TWO_KX = __import__("os").environ.get("B200CD_WGRAD_TWO_KX", "0") != "0"


def _kpad(cin: int) -> int:
    k = 9 * cin
    return 64 * ((k + 63) // 64)


@dataclass
class Stage:
    """conv3x3 (+bias) -> BatchNorm -> ReLU, the unit DoubleConv is made of (utils/networks.py:391-398)."""
    name: str
    conv: nn.Conv2d
    bn: nn.BatchNorm2d
    n_img: int
    H: int
    W: int
    G: int
    order_rev: bool
    first: bool
    in_view: torch.Tensor
    r: torch.Tensor = None
    outs: dict = field(default_factory=dict)
    srcs: list = field(default_factory=list)
    d_in: Optional[torch.Tensor] = None
    dr: Optional[torch.Tensor] = None
    Wf: torch.Tensor = None
    Wd: Optional[torch.Tensor] = None
    mean: torch.Tensor = None
    invstd: torch.Tensor = None
    scale: torch.Tensor = None
    shift: torch.Tensor = None
    stat_rows: int = 0
    stat_per_cta: bool = False
    bwd_sums: Optional[torch.Tensor] = None   # [G][rows][C][2]: BatchNorm-backward sums accumulated by the dgrad epilogue
    bwd_sum_rows: int = 0
    branch: int = 0          # trunk (stream) the stage belongs to
    join_before_bwd: bool = False   # its backward needs the gradients of both branches (first encoder op after two decoders)
    fused_eval: bool = False        # inference: conv + folded BatchNorm + ReLU in one launch

    @property
    def cin(self) -> int:
        return self.conv.in_channels

    @property
    def cout(self) -> int:
        return self.conv.out_channels


@dataclass
class UpConv:
    """ConvTranspose2d(c, c, 2, stride=2) writing into the upper half of a concat buffer (utils/networks.py:433-449)."""
    name: str
    up: nn.ConvTranspose2d
    x: torch.Tensor          # [nb, h, w, c] input (dense)
    out: torch.Tensor        # cat[..., c:] view at 2x resolution
    d_out: Optional[torch.Tensor] = None   # d_cat[..., c:] view
    d_x: Optional[torch.Tensor] = None     # [nb, h, w, c]
    Wf: torch.Tensor = None
    Wd: Optional[torch.Tensor] = None
    bias_rows: int = 0       # > 0: the dgrad that writes d_cat also wrote per-CTA channel sums (bias gradient for free)
    # inference on odd-sized levels: the 2h x 2w output goes to `dense`, then ops.pad_copy centres it in `out`
    dense: Optional[torch.Tensor] = None
    pad: tuple = (0, 0)
    branch: int = 0
    fork_before_fwd: bool = False   # first op of a branch that consumes what the main stream produced so far


@dataclass
class Head:
    """OutConv 1x1 (utils/networks.py:454-461) over one or two decoder outputs."""
    name: str
    conv: nn.Conv2d
    inputs: list            # one or two [nb, H, W, 64] tensors
    logits: torch.Tensor    # [nb, 1, H, W] fp32
    dz: Optional[torch.Tensor] = None


class GradArena:
    """One flat fp32 buffer holding every parameter gradient, laid out in the order the backward plan produces
    them (heads, then decoders deepest-last, then the encoder), so a data-parallel caller can all-reduce a
    completed prefix while the rest of backward is still running."""

    def __init__(self, module: nn.Module, device: torch.device, write_order: list, skip: set):
        self.params = [(n, p) for n, p in module.named_parameters()]
        names = {id(p): n for n, p in self.params}
        total = 0
        self.offsets = {}
        self.views = {}
        seen = set()
        for p in write_order:
            n = names[id(p)]
            if n in seen or n in skip:
                continue
            seen.add(n)
            self.offsets[n] = total
            total += (p.numel() + 3) // 4 * 4  # keep every slice 16-byte aligned
        missing = [n for n, _ in self.params if n not in seen and n not in skip]
        assert not missing, f"parameters without a gradient producer: {missing}"
        self.flat = torch.zeros(total, device=device, dtype=torch.float32)
        for n, p in self.params:
            if n in self.offsets:
                o = self.offsets[n]
                self.views[n] = self.flat[o:o + p.numel()].view(p.shape)
        self.skip = skip
        self._names = names

    def view_of(self, p: nn.Parameter) -> torch.Tensor:
        return self.views[self._names[id(p)]]

    def detached_copy_views(self) -> list:
        """Per-parameter views (in `self.params` order, None for skipped parameters) of ONE fresh copy of the flat
        buffer. autograd's AccumulateGrad deep-copies a returned gradient unless it holds the only reference to it —
        for views kept in `self.views` that is one small copy kernel per parameter (~160 launches per DualStream step);
        fresh views of a fresh copy are adopted as `p.grad` without further kernels, and `p.grad` never aliases the
        buffer the next backward overwrites (gradient accumulation over several backward calls stays correct)."""
        flat = self.flat.clone()
        out = []
        for n, p in self.params:
            if n in self.skip:
                out.append(None)
            else:
                o = self.offsets[n]
                out.append(flat[o:o + p.numel()].view(p.shape))
        return out

    def end_of(self, p: nn.Parameter) -> int:
        n = self._names[id(p)]
        return self.offsets[n] + (p.numel() + 3) // 4 * 4


class StepEngine:
    """Builds and runs the plan for `net` (one of the drop-in classes in networks.py)."""

    def __init__(self, net: nn.Module, B: int, H: int, W: int, train: bool, device: torch.device,
                 use_graphs: bool = True, precise: bool = False):
        if train and (H % 16 != 0 or W % 16 != 0):
            raise ValueError(f"b200cd engine: training tiles must be multiples of 16 (got {H}x{W})")
        if H < 16 or W < 16:
            raise ValueError(f"b200cd engine: tiles smaller than 16x16 vanish in the four MaxPool levels (got {H}x{W})")
        self.net, self.B, self.H, self.W, self.train, self.device = net, B, H, W, train, device
        self.use_graphs = use_graphs
        # precise: split-bf16 storage (hi + lo halves, 16 mantissa bits) and three-MMA products — the mode that meets
        # the reference-fp32 tolerance (include/b200cd.h, ABI version 2); default is single-bf16 storage (fast)
        self.precise = bool(precise)
        self.stages: list[Stage] = []
        self.upconvs: list[UpConv] = []
        self.heads: list[Head] = []
        self.fwd_ops: list[Callable[[], None]] = []
        self.bwd_ops: list[Callable[[], None]] = []
        self._side_streams: dict = {}             # weight-gradient side stream per executing trunk (0 / 1)
        self._side_dirty: set = set()
        self._exec_branch = 0                     # trunk whose op is being enqueued (_run_ops)
        self._cur_branch = 0                      # trunk being built (0, or 1 for the second stream's trunk)
        self.fwd_branch: list[int] = []           # per forward op: 0 / 1 = trunk, -1 = needs both trunks (heads)
        self.bwd_branch: list[int] = []
        self._branch_stream = None
        self._chain_stream = None
        self._in_chain = False
        self._branch_main = None
        self._branch_active = False
        self.branch_streams = BRANCH_STREAMS      # bench.profile_step turns both off to time launches one by one
        self.wgrad_side = WGRAD_SIDE_STREAM
        self.pack_fwd: list[Callable[[], None]] = []
        self.pack_bwd: list[Callable[[], None]] = []
        self.bwd_marks: list[int] = []  # per backward op: length of the flat-gradient prefix complete after it
        self.mem_bytes = 0
        cin_total = self._input_channels()
        self.x_t1 = torch.zeros(B, cin_total, H, W, device=device)
        self.x_t2 = torch.zeros(B, cin_total, H, W, device=device)
        self.grads: Optional[GradArena] = None
        self._ws_need = {"stats": 0, "stats2": 0, "wgrad": 0, "bnbwd": 0, "colsum": 0}
        self._reduce_specs: list[tuple] = []   # (backward op index, ws offset, grad view, splits, stride, layout, d0, d1, taps)
        self._reduce_tables: list[tuple] = []  # per flush: (device job table, njobs, blocks, bytes)
        self._build()
        self._alloc_ws()
        self._param_ptrs = self._ptr_signature()
        self._g_fwd = None
        self._g_bwd = None
        self._runs = 0
        self.generation = 0          # training forwards run so far (the drop-in autograd node checks it in backward)
        self._dp_plan = None         # data-parallel backward: bucket plan, per-bucket graphs, communication stream
        self._dp_graphs = None
        self._dp_stream = None
        self._dp_runs = 0

    # ------------------------------------------------------------------------------------------------
    def _input_channels(self) -> int:
        cfg = self.net.cfg
        t = cfg.MODEL.TYPE
        if t in ("dualstreamunet", "whatevernet", "whatevernet2"):
            return len(cfg.DATALOADER.S1_BANDS) + len(cfg.DATALOADER.S2_BANDS)
        return cfg.MODEL.IN_CHANNELS

    def _ptr_signature(self):
        # the module tree is walked once; afterwards only the storage addresses are compared (this runs every forward)
        if getattr(self, "_sig_tensors", None) is None:
            self._sig_tensors = list(self.net.parameters()) + list(self.net.buffers())
        return tuple(t.data_ptr() for t in self._sig_tensors)

    def params_moved(self) -> bool:
        return self._ptr_signature() != self._param_ptrs

    def _new(self, *shape, dtype=BF16, zero=False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(*shape, device=self.device, dtype=dtype)
        self.mem_bytes += t.numel() * t.element_size()
        return t

    def _act(self, n: int, H: int, W: int, C: int) -> torch.Tensor:
        """An NHWC bf16 activation / gradient buffer; in the precise mode the hi half of a [hi | lo] row."""
        if not self.precise:
            return self._new(n, H, W, C)
        self.mem_bytes += 4 * n * H * W * C
        return ops.split_alloc((n, H, W, C), self.device)

    # ------------------------------------------------------------------------------------------------
    # building blocks
    # ------------------------------------------------------------------------------------------------
    def _stage(self, name, conv, bn, in_view, n_img, H, W, G, order_rev=False, first=False) -> Stage:
        st = Stage(name, conv, bn, n_img, H, W, G, order_rev, first, in_view)
        st.branch = self._cur_branch
        C = conv.out_channels
        if C % 64 != 0 or (not first and conv.in_channels % 64 != 0):
            raise ValueError(f"b200cd engine: channel counts must be multiples of 64 ({name}: {conv.in_channels}->{C})")
        km = 3 if self.precise else 1   # K-tripled [hi | lo | hi] weight operands in the precise mode
        st.r = self._act(n_img, H, W, C)
        for k in ("mean", "invstd", "scale", "shift"):
            setattr(st, k, self._new(G, C, dtype=torch.float32))
        if first:
            st.Wf = self._new(C, km * in_view.shape[3])
        else:
            st.Wf = self._new(C, km * 9 * conv.in_channels)
            if self.train:
                st.Wd = self._new(conv.in_channels, km * 9 * C)
        if self.train:
            st.dr = self._act(n_img, H, W, C)
        tiles = ops.conv_gemm_tiles(H, W)
        # statistics rows per stat-group: one per 128-pixel tile, or (CTA-pair kernel) one per CTA and epilogue group
        st.stat_rows, st.stat_per_cta = (n_img // G) * tiles, False
        if self.device.type == "cuda":
            ka = in_view.shape[3] if first else conv.in_channels
            st.stat_rows, st.stat_per_cta = ops.conv_stat_rows(n_img, H, W, ka, C, G, mode=1 if first else 0,
                                                               prec=self.precise)
        self._ws_need["stats"] = max(self._ws_need["stats"], G * st.stat_rows * C * 2)
        self._ws_need["stats2"] = max(self._ws_need["stats2"], 32 * G * C * 2)
        if self.train:
            self._ws_need["bnbwd"] = max(self._ws_need["bnbwd"], ops.bn_bwd_ws_floats(n_img, H, W, C, G))
        self.stages.append(st)
        return st

    def _double_conv(self, name, dc: nn.Module, in_view, n_img, H, W, G, order_rev=False, first=False):
        """dc.conv = Sequential(conv, bn, relu, conv, bn, relu). Returns (stage1, stage2); stage2.outs is set by the caller."""
        seq = dc.conv
        s1 = self._stage(f"{name}.0", seq[0], seq[1], in_view, n_img, H, W, G, order_rev, first)
        a1 = self._act(n_img, H, W, s1.cout)
        s1.outs = {"a": a1}
        s2 = self._stage(f"{name}.3", seq[3], seq[4], a1, n_img, H, W, G, order_rev)
        if self.train:
            s2.d_in = self._act(n_img, H, W, s1.cout)
            s1.srcs = [{"kind": 1, "t": s2.d_in}]
        return s1, s2

    def _encoder(self, tag: str, inc: nn.Module, encoder: nn.Module, c_lo: int, nc: int, siamese: bool):
        """inc + encoder (utils/networks.py:405-412, 313-343). Returns the list of second stages per level
        (their apply outputs / gradient sources are wired by the decoders) and the level geometry."""
        B, H, W = self.B, self.H, self.W
        n_img, G = (2 * B, 2) if siamese else (B, 1)
        cin = nc if siamese else 2 * nc
        assert inc.conv.conv[0].in_channels == cin, f"{tag}: first conv expects {inc.conv.conv[0].in_channels} channels, data has {cin}"
        kpad = _kpad(cin)
        cols = self._act(n_img, H, W, kpad)
        self.fwd_ops.append(lambda: ops.pack_input(self.x_t1, self.x_t2, c_lo, nc, 0 if siamese else 1, kpad, out=cols,
                                                   prec=self.precise))
        self.fwd_branch.append(self._cur_branch)
        levels = []
        s1, s2 = self._double_conv(f"{tag}.inc", inc.conv, cols, n_img, H, W, G, first=True)
        levels.append(s2)
        h, w = H, W
        prev = s2
        for lname, down in encoder.down_seq.items():
            pool = self._act(n_img, h // 2, w // 2, prev.cout)
            prev.outs["pool"] = pool
            pidx = self._new(n_img, h // 2, w // 2, prev.cout, dtype=torch.uint8) if self.train else None
            prev.outs["pool_idx"] = pidx
            h, w = h // 2, w // 2
            d1, d2 = self._double_conv(f"{tag}.{lname}", down.mpconv[1], pool, n_img, h, w, G)
            if self.train:
                d1.d_in = self._act(n_img, h, w, prev.cout)  # gradient w.r.t. the pooled tensor
                prev.srcs.append({"kind": 2, "t": d1.d_in, "w": pidx})
            levels.append(d2)
            prev = d2
        return levels

    def _decoder(self, tag: str, decoder: nn.Module, levels: list, mode: str, order_rev: bool = False):
        """Decoder (utils/networks.py:346-382) over encoder `levels`.
        mode: 'plain' (skip = encoder activation, same batch), 'diff' (skip = t2 - t1, batch B from a 2B encoder),
              'copy' (skip = encoder activation of a 2B shared-weight encoder, e.g. decoder_sem)."""
        enc_n = levels[0].n_img
        nb = enc_n // 2 if mode == "diff" else enc_n
        G = 2 if mode == "copy" else 1
        deep = levels[-1]
        # x entering the first Up: deepest feature (or its difference)
        if mode == "diff":
            x = self._act(nb, deep.H, deep.W, deep.cout)
            deep.outs["dif"] = x
            deep.outs["diff"] = True
            deep.outs.setdefault("a", None)
        else:
            x = deep.outs.get("a")
            if x is None:
                x = self._act(enc_n, deep.H, deep.W, deep.cout)
                deep.outs["a"] = x
        last = None
        d_x_prev_consumer = None  # (stage whose srcs receive d_x)
        producer = ("enc", deep)
        ups = list(decoder.up_seq.items())
        for i, (uname, up) in enumerate(ups):
            skip_stage = levels[len(levels) - 2 - i]
            c = up.up.in_channels
            assert skip_stage.cout == c, f"{tag}.{uname}: skip has {skip_stage.cout} channels, Up expects {c}"
            Hs, Ws = skip_stage.H, skip_stage.W
            cat = self._act(nb, Hs, Ws, 2 * c)
            # skip half of the concat buffer, written by the encoder's apply kernel
            if mode == "diff":
                # the activation itself is not materialised: pool and t2 - t1 are produced from registers
                skip_stage.outs["dif"] = cat[..., :c]
                skip_stage.outs["diff"] = True
            elif mode == "copy":
                if skip_stage.outs.get("a") is None:
                    skip_stage.outs["a"] = cat[..., :c]
                else:
                    assert "a2" not in skip_stage.outs
                    skip_stage.outs["a2"] = cat[..., :c]
            else:
                assert "a" not in skip_stage.outs, "plain skip is written straight into the concat buffer"
                skip_stage.outs["a"] = cat[..., :c]
            uc = UpConv(f"{tag}.{uname}.up", up.up, x, cat[..., c:])
            uc.branch = self._cur_branch
            hx, wx = x.shape[1], x.shape[2]
            if (2 * hx, 2 * wx) != (Hs, Ws):
                # MaxPool floored an odd level: the reference pads the up-sampled tensor to the skip's size with
                # diff // 2 on the top / left (utils/networks.py:440-443)
                assert not self.train and 0 <= Hs - 2 * hx <= 1 and 0 <= Ws - 2 * wx <= 1
                uc.dense = self._act(nb, 2 * hx, 2 * wx, c)
                uc.pad = ((Hs - 2 * hx) // 2, (Ws - 2 * wx) // 2)
            km = 3 if self.precise else 1
            uc.Wf = self._new(4 * c, km * c)
            self.upconvs.append(uc)
            d_cat = None
            if self.train:
                d_cat = self._act(nb, Hs, Ws, 2 * c)
                uc.Wd = self._new(c, km * 4 * c)
                uc.d_out = d_cat[..., c:]
                uc.d_x = self._act(nb, Hs // 2, Ws // 2, c)
                # gradient of the skip half flows back into the encoder stage
                if mode == "diff":
                    skip_stage.srcs.append({"kind": 1, "t": d_cat[..., :c], "n_mod": nb, "scale_lo": -1.0, "scale_hi": 1.0})
                else:
                    skip_stage.srcs.append({"kind": 1, "t": d_cat[..., :c]})
                # gradient w.r.t. x flows into whoever produced x
                kind, pst = producer
                if kind == "enc" and mode == "diff":
                    pst.srcs.append({"kind": 1, "t": uc.d_x, "n_mod": nb, "scale_lo": -1.0, "scale_hi": 1.0})
                else:
                    pst.srcs.append({"kind": 1, "t": uc.d_x})
            s1, s2 = self._double_conv(f"{tag}.{uname}", up.conv, cat, nb, Hs, Ws, G, order_rev)
            if self.train:
                s1.d_in = d_cat
            xo = self._act(nb, Hs, Ws, s2.cout)
            s2.outs = {"a": xo}
            self._up_plan.append((uc, s1, s2))
            x = xo
            producer = ("dec", s2)
            last = s2
        return last

    def _head(self, name: str, conv: nn.Conv2d, dec_stages: list) -> Head:
        ins = [s.outs["a"] for s in dec_stages]
        nb, H, W, C = ins[0].shape
        assert conv.in_channels == C * len(ins) and conv.out_channels == 1, \
            f"{name}: only single-channel heads over 64-channel decoder outputs are supported"
        hd = Head(name, conv, ins, self._new(nb, 1, H, W, dtype=torch.float32))
        if self.train:
            hd.dz = self._new(nb, 1, H, W, dtype=torch.float32, zero=True)
            wflat = conv.weight.view(-1)
            for i, s in enumerate(dec_stages):
                s.srcs.append({"kind": 3, "t": hd.dz, "w": wflat[i * C:(i + 1) * C]})
            self._ws_need["colsum"] = max(self._ws_need["colsum"], 1184 * C)
        self.heads.append(hd)
        return hd

    # ------------------------------------------------------------------------------------------------
    def _build(self) -> None:
        net, cfg = self.net, self.net.cfg
        t = cfg.MODEL.TYPE
        self._up_plan = []
        ns1 = len(cfg.DATALOADER.S1_BANDS) if hasattr(cfg, "DATALOADER") and hasattr(cfg.DATALOADER, "S1_BANDS") else 0
        ns2 = len(cfg.DATALOADER.S2_BANDS) if ns1 else 0
        self.outputs = []  # (head, row slice or None) in the order forward() returns them
        if t == "unet":
            lv = self._encoder("u", net.inc, net.encoder, 0, cfg.MODEL.IN_CHANNELS, siamese=False)
            d = self._decoder("u.dec", net.decoder, lv, "plain")
            self.outputs = [(self._head("outc", net.outc.conv, [d]), None)]
        elif t == "siameseunet":
            lv = self._encoder("s", net.inc, net.encoder, 0, cfg.MODEL.IN_CHANNELS, siamese=True)
            d = self._decoder("s.dec", net.decoder, lv, "diff")
            self.outputs = [(self._head("outc", net.outc.conv, [d]), None)]
        elif t == "dtsiameseunet":
            lv = self._encoder("s", net.inc, net.encoder, 0, cfg.MODEL.IN_CHANNELS, siamese=True)
            # the two decoders are independent between the shared encoder and the heads: decoder_sem is branch 1, forked
            # right after the encoder (its ops precede decoder_change's in the plan so that the fork does not wait for
            # them) and joined before the heads (forward) / before the encoder's backward
            n_up = len(self.upconvs)
            self._cur_branch = 1
            ds = self._decoder("s.dec_sem", net.decoder_sem, lv, "copy", order_rev=True)
            self._cur_branch = 0
            self.upconvs[n_up].fork_before_fwd = True
            lv[-1].join_before_bwd = True
            dc = self._decoder("s.dec_change", net.decoder_change, lv, "diff")
            hc = self._head("outc_change", net.outc_change.conv, [dc])
            hs = self._head("outc_sem", net.outc_sem.conv, [ds])
            B = self.B
            self.outputs = [(hc, None), (hs, slice(0, B)), (hs, slice(B, 2 * B))]  # (change, sem_t1, sem_t2)
        elif t in ("dualstreamunet", "whatevernet2"):
            lv1 = self._encoder("s1", net.inc_stream1, net.encoder_stream1, 0, ns1, siamese=False)
            d1 = self._decoder("s1.dec", net.decoder_stream1, lv1, "plain")
            self._cur_branch = 1
            lv2 = self._encoder("s2", net.inc_stream2, net.encoder_stream2, ns1, ns2, siamese=False)
            d2 = self._decoder("s2.dec", net.decoder_stream2, lv2, "plain")
            self._cur_branch = 0
            if t == "dualstreamunet":
                self.outputs = [(self._head("outc", net.outc.conv, [d1, d2]), None)]
            else:
                hf = self._head("outc_fusion", net.outc_fusion.conv, [d1, d2])
                h1 = self._head("outc_stream1", net.outc_stream1.conv, [d1])
                h2 = self._head("outc_stream2", net.outc_stream2.conv, [d2])
                self.outputs = [(hf, None), (h1, None), (h2, None)]
        elif t == "whatevernet":
            lv1 = self._encoder("s1", net.inc_stream1, net.encoder_stream1, 0, ns1, siamese=True)
            d1 = self._decoder("s1.dec", net.decoder_stream1, lv1, "diff")
            self._cur_branch = 1
            lv2 = self._encoder("s2", net.inc_stream2, net.encoder_stream2, ns1, ns2, siamese=True)
            d2 = self._decoder("s2.dec", net.decoder_stream2, lv2, "diff")
            self._cur_branch = 0
            hf = self._head("outc_fusion", net.outc_fusion.conv, [d1, d2])
            h1 = self._head("outc_stream1", net.outc_stream1.conv, [d1])
            h2 = self._head("outc_stream2", net.outc_stream2.conv, [d2])
            self.outputs = [(hf, None), (h1, None), (h2, None)]
        else:
            raise ValueError(f"Unknown network ({t}).")
        self._pack_specs = []
        self._emit_forward()
        if self.train:
            self._emit_backward()
        # every weight of the step is converted to its bf16 GEMM operand layouts (forward and, when training, input
        # gradient) by ONE launch at the start of the forward pass: the fp32 master weights belong to the optimizer and
        # change every step, and they are read once
        if self.device.type == "cuda" and self._pack_specs:
            tab, nj, blocks, elems = ops.make_pack_jobs(self._pack_specs, self.device, prec=self.precise)
            src = sum(sp[1].numel() for sp in self._pack_specs)
            self.pack_fwd.append(lambda: ops.pack_weights_batched(tab, nj, blocks, elems, src, prec=self.precise))
            if not self.train:
                btab, bnj, bblocks = ops.make_bn_eval_jobs([(st.bn, st.mean, st.invstd, st.scale, st.shift)
                                                            for st in self.stages], self.device)
                self.pack_fwd.append(lambda: ops.bn_eval_affine_batched(btab, bnj, bblocks))

    # ------------------------------------------------------------------------------------------------
    def _alloc_ws(self) -> None:
        n = self._ws_need
        nbr = 2 if any(st.branch == 1 for st in self.stages) else 1   # one set of scratch buffers per trunk
        self.ws_stats = [self._new(max(n["stats"], 2), dtype=torch.float32) for _ in range(nbr)]
        self.ws_stats2 = [self._new(max(n["stats2"], 2), dtype=torch.float64) for _ in range(nbr)]
        if self.train:
            self.ws_wgrad = self._new(max(n["wgrad"], 4), dtype=torch.float32)
            for group in getattr(self, "_flush_groups", []):
                jobs = [(self.ws_wgrad.narrow(0, off, splits * stride), gw, splits, stride, layout, d0, d1, taps, s2)
                        for (_, off, gw, splits, stride, layout, d0, d1, taps, s2) in group]
                self._reduce_tables.append(ops.make_reduce_jobs(jobs, self.device))
            self.ws_bnbwd = [self._new(max(n["bnbwd"], 4), dtype=torch.float32) for _ in range(nbr)]
            self.ws_colsum = [self._new(max(n["colsum"], 4), dtype=torch.float32) for _ in range(nbr)]
            self.ws_upstats = [self._new(max(n.get("upstats", 0), 4), dtype=torch.float32) for _ in range(nbr)]

    # ------------------------------------------------------------------------------------------------
    # forward emission: stages were created in execution order, transposed convs are interleaved by name order
    # ------------------------------------------------------------------------------------------------
    def _emit_stage_fwd(self, st: Stage) -> None:
        eng = self
        conv, bn = st.conv, st.bn
        tiles = ops.conv_gemm_tiles(st.H, st.W)
        C = st.cout
        if st.first:
            eng._pack_specs.append((2, conv.weight, st.Wf, st.in_view.shape[3]))
        elif eng.train:
            eng._pack_specs.append((0, conv.weight, st.Wf, 0, 1, st.Wd))
        else:
            eng._pack_specs.append((0, conv.weight, st.Wf, 0))
        mode = 1 if st.first else 0
        train = eng.train
        count = (st.n_img // st.G) * st.H * st.W
        tpg = st.stat_rows
        spl = max(1, min(32, tpg // 64))
        sg = st.G if st.stat_per_cta else 0
        outs = st.outs
        prec = eng.precise

        # inference: BatchNorm is a fixed affine (computed for all stages by ONE launch at the start of the forward plan,
        # ops.bn_eval_affine_batched). A stage whose only product is its activation runs as ONE launch — conv with the
        # affine + ReLU folded into the epilogue, no pre-BN tensor, no apply pass; stages that also pool / difference /
        # copy keep conv -> apply (two launches).
        live = {k for k in ("a", "a2", "pool", "dif") if outs.get(k) is not None}
        fused_eval = (not train) and live == {"a"} and eng.device.type == "cuda" and ops.FPROP_PAIR
        st.fused_eval = fused_eval

        def run():
            if fused_eval:
                ops.conv_gemm_affine(mode, st.in_view, st.Wf, outs["a"], conv.bias, st.scale[0], st.shift[0], True, prec=prec)
                return
            stats = eng.ws_stats[st.branch] if train else None
            ops.conv_gemm(mode, 0, st.in_view, st.Wf, st.r, bias=conv.bias, stats=stats, stat_groups=sg, prec=prec)
            if train:
                ops.bn_stats(stats, C, C, tpg, st.G, count, spl, eng.ws_stats2[st.branch], bn.weight, bn.bias, bn.running_mean,
                             bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps, True,
                             st.order_rev, st.mean, st.invstd, st.scale, st.shift)
            ops.bn_apply(st.r, st.scale, st.shift, st.G, bool(outs.get("diff", False)), a=outs.get("a"),
                         a2=outs.get("a2"), pool=outs.get("pool"), dif=outs.get("dif"), pool_idx=outs.get("pool_idx"),
                         prec=prec)

        eng.fwd_ops.append(run)
        eng.fwd_branch.append(st.branch)

    def _emit_forward(self) -> None:
        up_by_first_stage = {id(s1): uc for (uc, s1, s2) in self._up_plan}
        for st in self.stages:
            uc = up_by_first_stage.get(id(st))
            if uc is not None:
                if self.train:
                    self._pack_specs.append((3, uc.up.weight, uc.Wf, 0, 4, uc.Wd))
                else:
                    self._pack_specs.append((3, uc.up.weight, uc.Wf, 0))
                if uc.dense is None:
                    self.fwd_ops.append(lambda uc=uc: ops.conv_gemm(1, 1, uc.x, uc.Wf, uc.out, bias=uc.up.bias,
                                                                    prec=self.precise))
                    self.fwd_branch.append(10 if uc.fork_before_fwd else uc.branch)
                else:
                    def run_up(uc=uc):
                        ops.conv_gemm(1, 1, uc.x, uc.Wf, uc.dense, bias=uc.up.bias, prec=self.precise)
                        ops.pad_copy(uc.dense, uc.out, uc.pad[0], uc.pad[1], prec=self.precise)
                    self.fwd_ops.append(run_up)
                    self.fwd_branch.append(10 if uc.fork_before_fwd else uc.branch)
            self._emit_stage_fwd(st)
        for hd in self.heads:
            def run(hd=hd):
                a1 = hd.inputs[1] if len(hd.inputs) > 1 else None
                ops.head_fwd(hd.inputs[0], a1, hd.conv.weight.view(-1), hd.conv.bias, hd.logits, prec=self.precise)
            self.fwd_ops.append(run)
            self.fwd_branch.append(-1)

    # ------------------------------------------------------------------------------------------------
    # backward emission: reverse order of the forward stages
    # ------------------------------------------------------------------------------------------------
    def _side(self):
        """Context manager: the enclosed launches go to the side stream, ordered after everything already on the
        current stream (fork). No-op on CPU or with B200CD_WGRAD_SIDE_STREAM=0."""
        eng = self

        class _Fork:
            def __enter__(self_inner):
                self_inner.ctx = None
                if eng.device.type != "cuda" or not eng.wgrad_side:
                    return
                key = eng._exec_branch if WGRAD_SIDE_PER_BRANCH else 0
                side = eng._side_streams.get(key)
                if side is None:
                    side = eng._side_streams[key] = torch.cuda.Stream(device=eng.device,
                                                                      priority=-1 if _PRIO_WGRAD_HIGH else 0)
                side.wait_stream(torch.cuda.current_stream())
                self_inner.ctx = torch.cuda.stream(side)
                self_inner.ctx.__enter__()
                eng._side_dirty.add(key)

            def __exit__(self_inner, *exc):
                if self_inner.ctx is not None:
                    self_inner.ctx.__exit__(*exc)
                return False

        return _Fork()

    def _join_side(self) -> None:
        """The current stream waits for everything launched on the side stream so far."""
        for key in sorted(self._side_dirty):
            torch.cuda.current_stream().wait_stream(self._side_streams[key])
        self._side_dirty.clear()

    def _ws_region(self, floats: int) -> int:
        """Every layer owns a region of the split workspace: the per-split partial weight gradients of a whole
        backward segment are summed by ONE batched launch (ops.wgrad_reduce_batched)."""
        off = self._ws_need["wgrad"]
        self._ws_need["wgrad"] = off + (floats + 3) // 4 * 4
        return off

    def _wgrad_plan(self, st: Stage):
        """Choose operand roles / split count for the 3x3 weight gradient and reserve its workspace region."""
        cin, cout = st.cin, st.cout
        total = ops.wgrad_tiles(st.n_img, st.H, st.W)
        if st.first:
            kp = st.in_view.shape[3]
            ctas = max(1, kp // 128)
            splits = max(1, min(total, WGRAD_CTA_TARGET // ctas))
            return ("first", splits, self._ws_region(splits * cout * kp), 0)
        if cout >= 128 or cin < 128:
            role = "pos"   # M <-> cout (U = dr), N <-> cin
            ctas = ops.wgrad_ctas_per_split(0, 1, cout, cin)
        else:
            role = "neg"   # M <-> cin (U = input), N <-> cout
            ctas = ops.wgrad_ctas_per_split(0, 1, cin, cout)
        nwide = cin if role == "pos" else cout      # width of the N side of the GEMM
        if WGRAD_TWO_KX and nwide % 128 != 0:
            # 64-wide N tiles: CTAs own two kx columns (splits of them) or the third one (splits2 = splits / 2)
            per_xy = max(1, ctas // 3)
            splits2 = max(1, min(total, WGRAD_CTA_TARGET // (3 * per_xy)))
            splits = min(total, 2 * splits2)
            return (role, splits, self._ws_region(splits * 9 * cout * cin), splits2)
        splits = max(1, min(total, max(1, WGRAD_CTA_TARGET // ctas)))
        return (role, splits, self._ws_region(splits * 9 * cout * cin), 0)

    def _fusable_producer(self, grad: torch.Tensor, mode: int = 0, ka: int = 64) -> Optional[Stage]:
        """The stage whose BatchNorm backward has `grad` as its one and only (direct, unscaled) gradient source and the
        same dense layout: its reduce pass can run in the epilogue of the kernel that writes `grad`. Allocates the
        stage's partial-sum buffer on first use."""
        if not FUSE_BN_BWD_REDUCE or self.device.type != "cuda" or not ops.FPROP_PAIR or self.precise:
            return None
        for ps in self.stages:
            if len(ps.srcs) == 1 and ps.srcs[0]["kind"] == 1 and ps.srcs[0]["t"] is grad and not ps.srcs[0].get("n_mod"):
                if tuple(grad.shape) != tuple(ps.r.shape) or ps.cout % 64 != 0:
                    return None
                if ps.n_img * ps.H * ps.W < FUSE_BN_BWD_MIN_PIXELS:
                    return None   # few work items per CTA pair: the longer epilogue is not hidden (measured)
                if ps.bwd_sums is None:
                    rows, per_cta = ops.conv_stat_rows(ps.n_img, ps.H, ps.W, ka, ps.cout, ps.G, mode=mode, variant="bnbwd")
                    if not per_cta:
                        return None
                    ps.bwd_sum_rows = rows
                    ps.bwd_sums = self._new(ps.G, rows, ps.cout, 2, dtype=torch.float32)
                return ps
        return None

    def _emit_stage_bwd(self, st: Stage) -> None:
        eng = self
        g = self.grads
        conv, bn = st.conv, st.bn
        gw, ggam, gbet = g.view_of(conv.weight), g.view_of(bn.weight), g.view_of(bn.bias)
        role, splits, off, splits2 = self._wgrad_plan(st)
        cin, cout = st.cin, st.cout
        srcs = st.srcs
        assert 1 <= len(srcs) <= 3, f"{st.name}: {len(srcs)} gradient sources"
        op_index = len(eng.bwd_ops)
        if role == "first":
            kp = st.in_view.shape[3]
            size = splits * cout * kp       # tiny (Cin <= 8): reduced by its own launch right away
        else:
            size = splits * 9 * cout * cin
            eng._reduce_specs.append((op_index, off, gw, splits, 9 * cout * cin, 0, cout, cin, 9, splits2))

        # the stage whose output feeds this conv: if this conv's input gradient is its ONLY gradient source, the reduce
        # pass of its BatchNorm backward runs in this dgrad's epilogue (ops.conv_gemm_bnbwd)
        prod = eng._fusable_producer(st.d_in, 0, st.cout) if st.d_in is not None else None
        up_rows = 0
        uc_of = {id(s1): uc for (uc, s1, s2) in eng._up_plan}.get(id(st))
        if uc_of is not None and prod is None and eng.device.type == "cuda" and ops.FPROP_PAIR and UP_BIAS_FROM_STATS:
            rows, per_cta = ops.conv_stat_rows(st.n_img, st.H, st.W, st.cout, st.cin, 1, prec=eng.precise)
            if per_cta:
                up_rows = uc_of.bias_rows = rows
                eng._ws_need["upstats"] = max(eng._ws_need.get("upstats", 0), rows * st.cin * 2)

        prec = eng.precise

        def run():
            ops.bn_bwd(st.r, st.mean, st.invstd, st.scale, st.shift, ops.make_srcs(srcs), st.G, eng.ws_bnbwd[st.branch], ggam, gbet,
                       st.dr, sums=st.bwd_sums, sum_rows=st.bwd_sum_rows, prec=prec)
            ws = eng.ws_wgrad.narrow(0, off, size)
            # input gradient first (the next layer's BatchNorm backward waits for it), then the weight gradient on the
            # side stream: it starts when the dgrad kernel leaves the SMs and runs under the next HBM-bound kernels
            if role != "first" and st.d_in is not None:
                if prod is not None:
                    ops.conv_gemm_bnbwd(0, st.dr, st.Wd, st.d_in, prod.r, prod.scale, prod.shift, prod.bwd_sums, prod.G)
                elif up_rows:
                    # d_in is the concat-buffer gradient of an Up: its per-channel pixel sums (upper half = the
                    # transposed-conv bias gradient) come out of this launch's per-CTA statistics
                    ops.conv_gemm(0, 0, st.dr, st.Wd, st.d_in, stats=eng.ws_upstats[st.branch], stat_groups=1, prec=prec)
                else:
                    ops.conv_gemm(0, 0, st.dr, st.Wd, st.d_in, prec=prec)
            with eng._side():
                if role == "first":
                    ops.wgrad_gemm(1, 1, 0, st.dr, st.in_view, ws, splits, cout * kp, 0, kp, 1, prec=prec)
                    ops.wgrad_reduce(ws, splits, cout * kp, 1, cout, cin, 9, gw)
                elif role == "pos":
                    ops.wgrad_gemm(0, 1, 1, st.dr, st.in_view, ws, splits, 9 * cout * cin, cout * cin, cin, 1, splits2,
                                   prec=prec)
                else:
                    ops.wgrad_gemm(0, -1, 1, st.in_view, st.dr, ws, splits, 9 * cout * cin, cout * cin, 1, cin, splits2,
                                   prec=prec)

        eng.bwd_ops.append(run)
        eng.bwd_branch.append(-10 if st.join_before_bwd else st.branch)
        eng.bwd_marks.append(g.end_of(conv.weight))

    def _emit_up_bwd(self, uc: UpConv) -> None:
        eng = self
        g = self.grads
        gw, gb = g.view_of(uc.up.weight), g.view_of(uc.up.bias)
        c = uc.up.in_channels
        nb, h, w, _ = uc.x.shape
        total = ops.wgrad_tiles(nb, h, w)
        ctas = ((c + 127) // 128) * (c // 128 if c % 128 == 0 else c // 64)
        splits = max(1, min(total, max(1, WGRAD_CTA_TARGET // ctas)))
        off = self._ws_region(splits * 4 * c * c)
        size = splits * 4 * c * c
        npix = nb * 4 * h * w
        nblk = max(1, min(1184, npix // 64))
        self._ws_need["colsum"] = max(self._ws_need["colsum"], nblk * c)
        eng._reduce_specs.append((len(eng.bwd_ops), off, gw, splits, 4 * c * c, 0, c, c, 4, 0))

        prod = eng._fusable_producer(uc.d_x, 2, c)

        def run():
            if uc.bias_rows:
                ops.stat_rowsum(eng.ws_upstats[uc.branch], uc.bias_rows, 2 * c, c, c, gb)
            else:
                ops.colsum(uc.d_out, None, npix, nblk, eng.ws_colsum[uc.branch], gb, prec=eng.precise)
            if prod is not None:
                ops.conv_gemm_bnbwd(2, uc.d_out, uc.Wd, uc.d_x, prod.r, prod.scale, prod.shift, prod.bwd_sums, prod.G)
            else:
                ops.conv_gemm(2, 0, uc.d_out, uc.Wd, uc.d_x, prec=eng.precise)
            with eng._side():
                ops.wgrad_gemm(2, 1, 0, uc.x, uc.d_out, eng.ws_wgrad.narrow(0, off, size), splits, 4 * c * c, c * c, c, 1,
                               prec=eng.precise)

        eng.bwd_ops.append(run)
        eng.bwd_branch.append(uc.branch)
        eng.bwd_marks.append(g.end_of(uc.up.weight))

    def _emit_backward(self) -> None:
        up_by_first = {id(s1): uc for (uc, s1, s2) in self._up_plan}
        order = []
        for hd in self.heads:
            order += [hd.conv.weight, hd.conv.bias]
        for st in reversed(self.stages):
            order += [st.bn.bias, st.bn.weight, st.conv.bias, st.conv.weight]
            uc = up_by_first.get(id(st))
            if uc is not None:
                order += [uc.up.bias, uc.up.weight]
        skip = {n for n, _ in self.net.named_parameters() if n.startswith("outc_sem_change")}
        self.grads = GradArena(self.net, self.device, order, skip)
        g = self.grads
        for hd in self.heads:
            C = hd.inputs[0].shape[3]
            npix = hd.logits.numel()
            nblk = max(1, min(1184, npix // 64))
            gw, gb = g.view_of(hd.conv.weight).view(-1), g.view_of(hd.conv.bias)

            def run(hd=hd, C=C, npix=npix, nblk=nblk, gw=gw, gb=gb):
                dz = hd.dz.view(-1)
                for i, a in enumerate(hd.inputs):
                    ops.colsum(a, dz, npix, nblk, self.ws_colsum[0], gw[i * C:(i + 1) * C], prec=self.precise)
                ops.colsum(None, dz, npix, nblk, self.ws_colsum[0], gb)
            self.bwd_ops.append(run)
            self.bwd_branch.append(-1)
            self.bwd_marks.append(g.end_of(hd.conv.bias))
        up_by_first_stage = {id(s1): uc for (uc, s1, s2) in self._up_plan}
        for st in reversed(self.stages):
            self._emit_stage_bwd(st)
            uc = up_by_first_stage.get(id(st))
            if uc is not None:
                self._emit_up_bwd(uc)
        self._plan_reduce_flushes()

    def _plan_reduce_flushes(self, segments: int = 4) -> None:
        """The split partials of consecutive layers are reduced together: a flush after the backward ops at which the
        accumulated weight-gradient volume crosses k / segments of the total (and after the last op). Until a layer's
        flush has run, its weight gradient is incomplete, so bwd_marks (the finished prefix of the flat gradient buffer
        that a data-parallel caller may all-reduce) only advance at flush points."""
        specs = self._reduce_specs
        if not specs:
            return
        total = sum(sp[2].numel() for sp in specs)
        flush_after, acc, k = [], 0, 1
        for sp in specs:
            acc += sp[2].numel()
            if acc * segments >= k * total:
                flush_after.append(sp[0])
                while acc * segments >= k * total:
                    k += 1
        if flush_after[-1] != specs[-1][0]:
            flush_after.append(specs[-1][0])
        # Tail flush: backward ends with the shallow encoder layers — the longest-running convolutions of the plan with
        # the smallest weight gradients (64 / 128 channels). A flush right before them completes all but a few MB of
        # the gradient buffer ~1.5 ms before the plan ends, so a data-parallel caller's last exposed all-reduce is tiny
        # (plan_buckets cuts there; measured: tools/dp_timeline.py).
        self._tail_mark = None
        rest = 0
        for i in range(len(specs) - 1, 0, -1):
            rest += specs[i][2].numel()
            if (rest + specs[i - 1][2].numel()) * TAIL_FLUSH_DIV > total:
                if i < len(specs) - 1 and specs[i - 1][0] not in flush_after and rest * TAIL_FLUSH_DIV <= total:
                    flush_after.append(specs[i - 1][0])
                    flush_after.sort()
                    self._tail_mark = specs[i - 1][0]
                break
        self._flush_groups = []
        lo = 0
        for fi, last_op in enumerate(flush_after):
            group = [sp for sp in specs[lo:] if sp[0] <= last_op]
            lo += len(group)
            self._flush_groups.append(group)
            orig = self.bwd_ops[last_op]

            def run(orig=orig, fi=fi):
                orig()
                self._join_side()
                self._join_branches()     # the group may hold layers of both trunks
                tab, nj, blocks, nbytes = self._reduce_tables[fi]
                ops.wgrad_reduce_batched(tab, nj, blocks, nbytes)

            self.bwd_ops[last_op] = run
        # marks: heads are complete as produced; stage ops only complete the prefix up to the last flushed layer
        first_stage_op = specs[0][0]
        flushed = self.bwd_marks[first_stage_op - 1] if first_stage_op > 0 else 0
        done = set(flush_after)
        for i in range(first_stage_op, len(self.bwd_ops)):
            if i in done:
                flushed = self.bwd_marks[i]
            self.bwd_marks[i] = flushed
        self.bwd_marks[-1] = self.grads.flat.numel()   # the last flush is not after the last op: everything is complete

    # ------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------
    def _two_streams(self, branches: list) -> bool:
        if not (self.branch_streams and self.device.type == "cuda" and any(b in (1, 10) for b in branches)):
            return False
        px = max(st.n_img * st.H * st.W for st in self.stages)
        return _BRANCH_MODE == "2" or px <= BRANCH_MAX_PIXELS

    def _join_branches(self) -> None:
        """The current stream waits for everything queued so far on the other trunk's stream (no-op when the plan runs
        on one stream)."""
        if self._branch_active:
            cur = torch.cuda.current_stream()
            other = self._branch_stream if cur != self._branch_stream else self._branch_main
            cur.wait_stream(other)

    def _run_ops(self, ops_list: list, branches: list) -> None:
        """Run plan ops. Branch codes: 0 main stream; 1 branch stream; 10 branch stream after waiting for the main
        stream (first op of a branch that consumes main-stream results); -10 main stream after waiting for the branch
        stream; -1 needs both (join, run on main, fork again: heads). Forks at the start, joins at the end."""
        if not self._two_streams(branches):
            for f in ops_list:
                f()
            return
        if self._branch_stream is None:
            self._branch_stream = torch.cuda.Stream(device=self.device,
                                                    priority=-1 if STREAM_PRIO and not _PRIO_WGRAD_HIGH else 0)
        main, side = torch.cuda.current_stream(), self._branch_stream
        self._branch_main = main
        side.wait_stream(main)                       # fork: everything queued so far is visible to both trunks
        self._branch_active = True
        try:
            for b, f in zip(branches, ops_list):
                self._exec_branch = 1 if b in (1, 10) else 0
                if b == 1 or b == 10:
                    if b == 10:                      # consumes what the main stream has produced so far
                        side.wait_stream(main)
                    with torch.cuda.stream(side):
                        f()
                elif b == 0:
                    f()
                elif b == -10:                       # main-stream op that consumes the branch stream's results
                    main.wait_stream(side)
                    f()
                else:                                # needs both trunks, and later trunk ops need it
                    main.wait_stream(side)
                    f()
                    side.wait_stream(main)
        finally:
            self._branch_active = False
            self._exec_branch = 0
        main.wait_stream(side)                       # join

    def run_bwd_range(self, o0: int, o1: int) -> None:
        """Backward ops [o0, o1) (a data-parallel caller runs the plan in segments, step.py)."""
        self._run_ops(self.bwd_ops[o0:o1], self.bwd_branch[o0:o1])
        self._join_side()

    def _on_chain_stream(self, fn) -> None:
        """Run `fn` on the plan's high-priority stream (forked from / joined into the current stream) when
        B200CD_STREAM_PRIO is on; else right here."""
        if not STREAM_PRIO or self.device.type != "cuda" or self._in_chain:
            fn()
            return
        if self._chain_stream is None:
            self._chain_stream = torch.cuda.Stream(device=self.device, priority=0 if _PRIO_WGRAD_HIGH else -1)
        cur = torch.cuda.current_stream()
        self._chain_stream.wait_stream(cur)
        self._in_chain = True
        try:
            with torch.cuda.stream(self._chain_stream):
                fn()
        finally:
            self._in_chain = False
        cur.wait_stream(self._chain_stream)

    def _run_fwd_eager(self) -> None:
        def body():
            for f in self.pack_fwd:
                f()
            self._run_ops(self.fwd_ops, self.fwd_branch)
        self._on_chain_stream(body)

    def _run_bwd_eager(self) -> None:
        def body():
            for f in self.pack_bwd:
                f()
            self.run_bwd_range(0, len(self.bwd_ops))
        self._on_chain_stream(body)

    def forward(self, x_t1: torch.Tensor, x_t2: torch.Tensor) -> None:
        """Copies the inputs into the static buffers and runs the forward plan; logits land in head.logits."""
        self.x_t1.copy_(x_t1, non_blocking=True)
        self.x_t2.copy_(x_t2, non_blocking=True)
        self.forward_static()

    def forward_static(self) -> None:
        if self.use_graphs and self._runs >= 1:
            if self._g_fwd is None:
                self._g_fwd = PlanGraph(self._run_fwd_eager)
            self._g_fwd.replay()
        else:
            self._run_fwd_eager()
        self._runs += 1

    def backward_static(self) -> None:
        """Runs the backward plan from the dz buffers of the heads; gradients land in self.grads.flat."""
        if self.use_graphs and self._runs >= 2:
            if self._g_bwd is None:
                self._g_bwd = PlanGraph(self._run_bwd_eager)
            self._g_bwd.replay()
        else:
            self._run_bwd_eager()
        self._runs += 1

    # ------------------------------------------------------------------------------------------------
    # data-parallel backward (one process per GPU): the plan runs in segments of roughly equal gradient volume and each
    # finished prefix of the flat gradient buffer is all-reduced (SUM, nn.DataParallel's reduce-add) on a side stream
    # while the next segment runs. Used by the fused TrainStep and by the drop-in autograd node alike.
    # ------------------------------------------------------------------------------------------------
    def plan_buckets(self, nbuckets: int) -> list:
        marks = self.bwd_marks
        total = marks[-1]
        nb = max(1, min(nbuckets, len(marks)))
        cuts, lo = [], 0
        for b in range(1, nb + 1):
            want = total * b // nb
            i = next(i for i, m in enumerate(marks) if m >= want)
            i = max(i, cuts[-1][1] if cuts else 0)
            cuts.append((cuts[-1][1] if cuts else 0, i + 1, lo, marks[i]))
            lo = marks[i]
        cuts = [c for c in cuts if c[1] > c[0]]     # (op_begin, op_end, grad_lo, grad_hi); empty segments dropped
        # the tail flush (see _plan_reduce_flushes) splits the last bucket: everything but the shallow layers' few MB
        # goes out while those layers still run
        tm = getattr(self, "_tail_mark", None)
        if tm is not None and cuts:
            o0, o1, g0, g1 = cuts[-1]
            if o0 <= tm < o1 - 1 and g0 < marks[tm] < g1:
                cuts[-1:] = [(o0, tm + 1, g0, marks[tm]), (tm + 1, o1, marks[tm], g1)]
        return cuts

    def backward_dp(self, group, nbuckets: int = 4, skip_allreduce: bool = False, inner_graphs: bool = True,
                    timeline: Optional[list] = None) -> None:
        """inner_graphs=False: every segment runs eagerly (the caller is capturing the whole step, collectives
        included, into ONE graph). timeline: a list that receives (label, timing event) pairs — segment ends on the
        compute stream, all-reduce start / end on the communication stream (tools/dp_timeline.py)."""
        import torch.distributed as dist
        if self._dp_plan is None:
            self._dp_plan = self.plan_buckets(nbuckets)
            self._dp_graphs = [None] * len(self._dp_plan)
            self._dp_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream()
        comm = self._dp_stream
        for f in self.pack_bwd:
            f()
        for bi, (o0, o1, g0, g1) in enumerate(self._dp_plan):
            if self.use_graphs and inner_graphs and self._dp_runs >= 1:
                if self._dp_graphs[bi] is None:
                    # forks / joins the second trunk's stream inside the segment
                    self._dp_graphs[bi] = PlanGraph(lambda o0=o0, o1=o1: self._on_chain_stream(lambda: self.run_bwd_range(o0, o1)))
                self._dp_graphs[bi].replay()
            else:
                self._on_chain_stream(lambda o0=o0, o1=o1: self.run_bwd_range(o0, o1))
            ev = torch.cuda.Event(enable_timing=timeline is not None)
            ev.record(main)
            comm.wait_event(ev)
            if timeline is not None:
                timeline.append((f"seg{bi}_done", ev))
            if skip_allreduce or g1 <= g0:   # skip: measurement aid only (B200CD_DEBUG_SKIP_ALLREDUCE=1), wrong gradients
                continue                     # g1 == g0: the segment completed no gradient (same on every rank)
            with torch.cuda.stream(comm):
                if timeline is not None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(comm)
                    timeline.append((f"ar{bi}_start[{4 * (g1 - g0)} B]", e))
                if parallel.native_comm():     # the library's own NCCL communicator (b200cd_allreduce_bucket)
                    parallel.allreduce_sum_(self.grads.flat[g0:g1])
                else:
                    dist.all_reduce(self.grads.flat[g0:g1], op=dist.ReduceOp.SUM, group=group)
                if timeline is not None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(comm)
                    timeline.append((f"ar{bi}_end", e))
        main.wait_stream(comm)
        self._dp_runs += 1
        self._runs += 1

    def output_tensors(self) -> list[torch.Tensor]:
        outs = []
        for hd, sl in self.outputs:
            outs.append(hd.logits if sl is None else hd.logits[sl])
        return outs

    def launches_per_step(self) -> dict:
        """Kernel launches of one forward / backward (counted from the plan, not measured)."""
        n_first = sum(1 for s in self.stages if s.first)
        n_st = len(self.stages)
        if not self.train:
            # pack_input per encoder, weight pack + BatchNorm affine (pack_fwd), one launch per fused stage, two otherwise
            fwd = len(self.pack_fwd) + n_first + sum(1 if s.fused_eval else 2 for s in self.stages) + \
                len(self.upconvs) + len(self.heads)
            return {"forward": fwd, "backward": 0}
        fwd = len(self.pack_fwd) + n_first + n_st * 4 + len(self.upconvs) + len(self.heads)  # conv, 2x stats, apply
        bwd = len(self.pack_bwd) + n_st * 5 + sum(1 for s in self.stages if s.d_in is not None) + \
            len(self.upconvs) * 5 + sum(2 * (len(h.inputs) + 1) for h in self.heads)
        return {"forward": fwd, "backward": bwd}
