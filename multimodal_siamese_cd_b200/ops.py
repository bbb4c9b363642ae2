"""Tensor-level wrappers over the C ABI (include/b200cd.h).

Activations are torch bf16 tensors of logical shape [n, H, W, C] whose last dim is contiguous and whose
pixel stride `ld = t.stride(2)` may exceed C (a channel slice of a concat buffer). These functions only
extract pointers/strides and launch on torch's current CUDA stream; they never compute on the CPU and
raise on CPU tensors.

`prec=True` selects the split-bf16 ("precise") kernels: the tensor passed is the hi half of a [hi | lo] row and its
lo half lies ld / 2 elements behind it (see include/b200cd.h, ABI version 2, and `split_alloc` / `lo_half` below).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib, tuning
from ._lib import GradSrc


# Kernel launches issued through this module since import (our own kernels only; bench.py reports the
# per-step delta as `gpu_launches`).
LAUNCHES = 0


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


# Optional per-call profiling (bench.py's roofline leg): when PROFILE is a list, every wrapper appends
# (kernel family, algorithmic flops, algorithmic bytes, start event, end event) around its launch(es).
PROFILE = None


class _Prof:
    __slots__ = ("name", "flops", "bytes", "e0", "tag")

    def __init__(self, name: str, flops: float = 0.0, nbytes: float = 0.0, tag: str = ""):
        self.name, self.flops, self.bytes, self.tag = name, flops, nbytes, tag

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.flops, self.bytes, self.e0, e1, self.tag))
        return False


def _nbytes(*ts) -> float:
    return float(sum(t.shape.numel() * t.element_size() for t in ts if t is not None))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: Optional[torch.Tensor]) -> int:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.B200CDError("b200cd ops run on CUDA tensors only (there is no CPU path)")
        dev = t.device.index if dev is None else dev
    if dev is None:
        raise _lib.B200CDError("no tensor given")
    _lib.init(dev)
    return dev


def _nhwc(t: torch.Tensor) -> tuple[int, int, int, int, int]:
    """(n, H, W, C, ld) of an NHWC view."""
    assert t.dim() == 4 and t.dtype == torch.bfloat16 and t.stride(3) == 1, "expected an NHWC bf16 view"
    n, H, W, Cc = t.shape
    ld = t.stride(2)
    assert t.stride(1) == W * ld and (n == 1 or t.stride(0) == H * W * ld), "pixel stride must be uniform"
    return n, H, W, Cc, ld


def split_alloc(shape, device, zero: bool = False) -> torch.Tensor:
    """A split-bf16 activation [n, H, W, C]: allocates [n, H, W, 2C] and returns the hi half (a view with pixel stride
    2C); the lo half is `lo_half(view)`. Channel slices of the returned view keep the ld / 2 relation."""
    n, H, W, Cc = shape
    buf = (torch.zeros if zero else torch.empty)((n, H, W, 2 * Cc), device=device, dtype=torch.bfloat16)
    return buf[..., :Cc]


def lo_half(t: torch.Tensor) -> torch.Tensor:
    """The lo half of a split-bf16 view (same shape / strides, ld / 2 elements further)."""
    return t.as_strided(t.shape, t.stride(), t.storage_offset() + t.stride(2) // 2)


def split_from_float(x: torch.Tensor) -> torch.Tensor:
    """fp32 [n, H, W, C] -> split-bf16 view holding hi = bf16(x), lo = bf16(x - hi) (tests and tools)."""
    out = split_alloc(tuple(x.shape), x.device)
    hi = x.to(torch.bfloat16)
    out.copy_(hi)
    lo_half(out).copy_((x - hi.float()).to(torch.bfloat16))
    return out


def split_to_float(t: torch.Tensor) -> torch.Tensor:
    return t.float() + lo_half(t).float()


def device_status(dev: Optional[int] = None) -> None:
    """Synchronise and raise if any tensor-core kernel reported a pipeline time-out."""
    dev = torch.cuda.current_device() if dev is None else dev
    _lib.init(dev)
    _lib.device_status(dev, _stream())


# ---------------------------------------------------------------------------------------------------------
def pack_input(x0: torch.Tensor, x1: torch.Tensor, c_lo: int, nc: int, cat_mode: int, kpad: int,
               out: Optional[torch.Tensor] = None, prec: bool = False) -> torch.Tensor:
    _require_cuda(x0, x1)
    assert x0.dtype == torch.float32 and x0.is_contiguous() and x1.is_contiguous() and x0.shape == x1.shape
    B, cs, H, W = x0.shape
    n_img = B if cat_mode else 2 * B
    if out is None:
        out = split_alloc((n_img, H, W, kpad), x0.device) if prec else \
            torch.empty((n_img, H, W, kpad), device=x0.device, dtype=torch.bfloat16)
    if prec:
        assert _nhwc(out)[4] == 2 * kpad, "split im2col rows are [hi kpad | lo kpad]"
    else:
        assert out.is_contiguous()
    _count(1)
    fn = _lib.load().b200cd_pack_input_hp if prec else _lib.load().b200cd_pack_input
    with _Prof("pack_input", 0.0, _nbytes(out) * (2 if prec else 1) + 4.0 * 2 * B * nc * H * W):
        _lib.check(fn(x0.data_ptr(), x1.data_ptr(), cs, c_lo, nc, cat_mode, B, H, W, kpad, out.data_ptr(), _stream()))
    return out


def pack_weights(mode: int, w: torch.Tensor, kpad: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous()
    d0, d1 = w.shape[0], w.shape[1]
    if mode == 0:
        shape = (d0, 9 * d1)
    elif mode == 1:
        shape = (d1, 9 * d0)
    elif mode == 2:
        shape = (d0, kpad)
    elif mode == 3:
        shape = (4 * d1, d0)
    else:
        shape = (d0, 4 * d1)
    if out is None:
        out = torch.empty(shape, device=w.device, dtype=torch.bfloat16)
    assert tuple(out.shape) == shape and out.is_contiguous()
    _count(1)
    with _Prof("pack_weights", 0.0, _nbytes(w, out)):
        _lib.check(_lib.load().b200cd_pack_weights(mode, w.data_ptr(), out.data_ptr(), d0, d1, kpad, _stream()))
    return out


_PACK_JOB_DTYPE = None


def packed_weight_shape(mode: int, d0: int, d1: int, kpad: int = 0, prec: bool = False) -> tuple[int, int]:
    """Shape of the bf16 GEMM operand `mode` of a weight with leading dims (d0, d1); K is tripled ([hi | lo | hi] per
    tap) in the split-bf16 mode."""
    m = 3 if prec else 1
    if mode == 0:
        return (d0, 9 * d1 * m)
    if mode == 1:
        return (d1, 9 * d0 * m)
    if mode == 2:
        return (d0, kpad * m)
    if mode == 3:
        return (4 * d1, d0 * m)
    return (d0, 4 * d1 * m)


def make_pack_jobs(specs: Sequence[tuple], device, prec: bool = False) -> tuple[torch.Tensor, int, int, int]:
    """specs: (mode, weight fp32 tensor, out bf16 tensor, kpad[, mode2, out2 bf16 tensor]). Returns (device job table,
    njobs, total thread blocks, packed elements) for pack_weights_batched (b200cd_pack_job layout: three pointers,
    six int32, one int64 = 56 bytes)."""
    import numpy as np
    global _PACK_JOB_DTYPE
    if _PACK_JOB_DTYPE is None:
        _PACK_JOB_DTYPE = np.dtype([("w", "<u8"), ("out", "<u8"), ("out2", "<u8"), ("mode", "<i4"), ("mode2", "<i4"),
                                    ("d0", "<i4"), ("d1", "<i4"), ("kpad", "<i4"), ("reserved", "<i4"), ("start", "<i8")],
                                   align=True)
        assert _PACK_JOB_DTYPE.itemsize == 56
    lib = _lib.load()
    arr = np.zeros(len(specs), dtype=_PACK_JOB_DTYPE)
    blocks = 0
    elems = 0
    for i, spec in enumerate(specs):
        mode, w, out, kpad = spec[:4]
        mode2, out2 = (spec[4], spec[5]) if len(spec) > 4 else (0, None)
        assert w.dtype == torch.float32 and w.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
        assert out2 is None or (out2.dtype == torch.bfloat16 and out2.is_contiguous() and out2.numel() == out.numel())
        assert out.data_ptr() % 4 == 0 and (out2 is None or out2.data_ptr() % 4 == 0), "packed operands are written 4 bytes at a time"
        arr[i] = (w.data_ptr(), out.data_ptr(), 0 if out2 is None else out2.data_ptr(), mode, mode2, w.shape[0],
                  w.shape[1], kpad, 0, blocks)
        assert tuple(out.shape) == packed_weight_shape(mode, w.shape[0], w.shape[1], kpad, prec), (mode, out.shape)
        nb = (lib.b200cd_pack_job_blocks_hp if prec else lib.b200cd_pack_job_blocks)(mode, w.shape[0], w.shape[1], kpad)
        assert nb > 0
        blocks += nb
        elems += out.numel() * (1 if out2 is None else 2)
    table = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return table, len(specs), blocks, elems


def pack_weights_batched(table: torch.Tensor, njobs: int, blocks: int, elems: int = 0, src_elems: int = 0,
                         prec: bool = False) -> None:
    _require_cuda(table)
    _count(1)
    fn = _lib.load().b200cd_pack_weights_hp_batched if prec else _lib.load().b200cd_pack_weights_batched
    with _Prof("pack_weights", 0.0, 2.0 * elems + 4.0 * src_elems):
        _lib.check(fn(table.data_ptr(), njobs, blocks, _stream()))


def conv_gemm_tiles(H: int, W: int) -> int:
    return _lib.load().b200cd_conv_gemm_tiles(H, W)


FPROP_HALO_POLICY = "auto"  # 3x3 convs using the halo variant: "auto" (N % 128 != 0 or K-chunk count >= 8), "n64", "all", "none"
FPROP_WIDE_TILES = True     # 128 x 256 output tiles where N % 256 == 0 and the halo variant is not used
FPROP_PAIR = True           # 3x3 convs on CTA pairs (cta_group::2, persistent, gemm_fprop2.cu)


def _conv_flags(mode: int, out_mode: int, N: int, ka: int, halo, wide, pair, stat_groups: int, bn=None) -> int:
    """Tile policy (measured on B200, profiles/): 3x3 convs on CTA pairs; otherwise 128x256 tiles where N % 256 == 0,
    the halo variant when the N tile is 64 wide or the K loop is long (>= 8 chunks of 64 channels). bn: N tile of the
    CTA-pair kernel (64 / 128 / 256; None = the library's rule) — tuning.py holds the measured per-shape choices."""
    pair_ok = out_mode == 0 or mode == 1
    if pair is None:
        pair = FPROP_PAIR and pair_ok and halo is None and wide is None
    pair = bool(pair) and pair_ok
    if wide is None:
        wide = FPROP_WIDE_TILES and N % 256 == 0 and halo is not True
    if halo is None:
        pol = FPROP_HALO_POLICY
        halo = mode == 0 and not wide and (pol == "all" or (pol in ("n64", "auto") and N % 128 != 0) or
                                           (pol == "auto" and ka >= 512))
    if pair:
        return 4 | ((8 | (stat_groups << 8)) if stat_groups > 0 else 0) | tuning.flag_bits(bn)
    return (1 if (halo and mode == 0) else 0) | (2 if wide else 0)


def conv_stat_rows(n_img: int, H: int, W: int, ka: int, N: int, stat_groups: int, mode: int = 0,
                   prec: bool = False, variant: str = "stats", bn=-1) -> tuple[int, bool]:
    """(rows per stat-group of the statistics buffer a 3x3 conv of this shape writes, per-CTA layout?). With the
    CTA-pair kernel the rows are per CTA and epilogue group (<= 296); otherwise one row per 128-pixel tile.
    variant / bn: as the launch that will write the buffer ("stats": conv_gemm with statistics, "bnbwd":
    conv_gemm_bnbwd; bn = -1: the tuned tile of that launch shape, None: the library's rule)."""
    if bn == -1:
        bn = tuning.tile(variant, mode, 0, n_img, H, W, ka, N, prec)
    flags = _conv_flags(mode, 0, N, ka, None, None, None, stat_groups, bn) | (16 if prec else 0)
    rows = _lib.load().b200cd_conv_gemm_stat_rows(mode, 0, flags, n_img, H, W, ka, N)
    if rows < 0:
        raise _lib.B200CDError("conv_gemm_stat_rows: unsupported shape")
    per_cta = bool(flags & 8)
    return (rows if per_cta else rows // stat_groups), per_cta


def conv_gemm(mode: int, out_mode: int, A: torch.Tensor, Bw: torch.Tensor, out: torch.Tensor,
              bias: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
              halo: Optional[bool] = None, wide: Optional[bool] = None, pair: Optional[bool] = None,
              stat_groups: int = 0, prec: bool = False, bn=-1) -> None:
    """G1. A: NHWC view (mode 2: at 2x the GEMM resolution); out: NHWC view (out_mode 1: at 2x).
    prec: A / out are split-bf16 views, Bw is the K-tripled [hi | lo | hi] operand.
    bn: N tile of the CTA-pair kernel (-1: tuning.py's entry for this launch shape, None: the library's rule)."""
    _require_cuda(A, Bw, out)
    n, Ha, Wa, ka, a_ld = _nhwc(A)
    H, W = (Ha // 2, Wa // 2) if mode == 2 else (Ha, Wa)
    no, Ho, Wo, Co, o_ld = _nhwc(out)
    N = Bw.shape[0]
    taps = 9 if mode == 0 else (1 if mode == 1 else 4)
    assert Bw.dtype == torch.bfloat16 and Bw.is_contiguous() and Bw.shape[1] == taps * ka * (3 if prec else 1)
    if out_mode == 1:
        assert (no, Ho, Wo) == (n, 2 * H, 2 * W) and N == 4 * Co
        cout = Co
    else:
        assert (no, Ho, Wo, Co) == (n, H, W, N)
        cout = 0
    sg = stat_groups if stats is not None else 0
    if bn == -1:
        bn = tuning.tile("stats" if sg > 0 else "plain", mode, out_mode, n, H, W, ka, N, prec)
    flags = _conv_flags(mode, out_mode, N, ka, halo, wide, pair, sg, bn)
    if prec:
        assert flags & 4, "split-bf16 operands need the CTA-pair kernel"
        flags |= 16
    _count(1)
    fam = "fprop3x3" if mode == 0 else ("gemm1tap" if mode == 1 else "convT_dgrad")
    # algorithmic flops: the reference's contraction (the three split MMAs are how it is computed, not extra work)
    with _Prof(fam, 2.0 * n * H * W * N * taps * ka, _nbytes(A, out) * (2 if prec else 1) + _nbytes(Bw),
               f"{n}x{H}x{W} k{ka}->n{N} om{out_mode} halo{flags}"):
        _lib.check(_lib.load().b200cd_conv_gemm(mode, out_mode, flags, A.data_ptr(), a_ld, n, H, W, ka, Bw.data_ptr(), N, cout,
                                                out.data_ptr(), o_ld, _ptr(bias), _ptr(stats), _stream()))


def conv_gemm_affine(mode: int, A: torch.Tensor, Bw: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor],
                     scale: torch.Tensor, shift: torch.Tensor, relu: bool = True, prec: bool = False, bn=-1) -> None:
    """Inference: conv + folded BatchNorm (+ ReLU) in one launch, out = act((A * Bw + bias) * scale + shift)."""
    _require_cuda(A, Bw, out, scale, shift)
    n, H, W, ka, a_ld = _nhwc(A)
    no, Ho, Wo, Co, o_ld = _nhwc(out)
    N = Bw.shape[0]
    taps = 9 if mode == 0 else 1
    assert mode in (0, 1) and Bw.shape[1] == taps * ka * (3 if prec else 1) and (no, Ho, Wo, Co) == (n, H, W, N)
    assert scale.dtype == torch.float32 and scale.numel() >= N and shift.numel() >= N
    if bn == -1:
        bn = tuning.tile("affine", mode, 0, n, H, W, ka, N, prec)
    flags = 4 | (16 if prec else 0) | tuning.flag_bits(bn)
    _count(1)
    fam = "fprop3x3" if mode == 0 else "gemm1tap"
    with _Prof(fam, 2.0 * n * H * W * N * taps * ka, _nbytes(A, out) * (2 if prec else 1) + _nbytes(Bw),
               f"{n}x{H}x{W} k{ka}->n{N} affine"):
        _lib.check(_lib.load().b200cd_conv_gemm_affine(mode, flags, A.data_ptr(), a_ld, n, H, W, ka, Bw.data_ptr(), N,
                                                       out.data_ptr(), o_ld, _ptr(bias), scale.data_ptr(), shift.data_ptr(),
                                                       int(relu), _stream()))


_BN_EVAL_JOB_DTYPE = None


def make_bn_eval_jobs(specs: Sequence[tuple], device) -> tuple[torch.Tensor, int, int]:
    """specs: (bn module, mean, invstd, scale, shift) with the four [G, C] fp32 outputs. Returns (device job table,
    njobs, total thread blocks) for bn_eval_affine_batched (b200cd_bn_eval_job: 8 pointers + 4 x 4 bytes = 80 bytes)."""
    import numpy as np
    global _BN_EVAL_JOB_DTYPE
    if _BN_EVAL_JOB_DTYPE is None:
        _BN_EVAL_JOB_DTYPE = np.dtype([("gamma", "<u8"), ("beta", "<u8"), ("rm", "<u8"), ("rv", "<u8"), ("mean", "<u8"),
                                       ("invstd", "<u8"), ("scale", "<u8"), ("shift", "<u8"), ("C", "<i4"), ("G", "<i4"),
                                       ("eps", "<f4"), ("start", "<i4")], align=True)
        assert _BN_EVAL_JOB_DTYPE.itemsize == 80
    arr = np.zeros(len(specs), dtype=_BN_EVAL_JOB_DTYPE)
    blocks = 0
    for i, (bn, mean, invstd, scale, shift) in enumerate(specs):
        G, Cc = mean.shape
        for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var, mean, invstd, scale, shift):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        arr[i] = (bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                  mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), Cc, G, bn.eps, blocks)
        blocks += (Cc + 255) // 256
    table = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return table, len(specs), blocks


def bn_eval_affine_batched(table: torch.Tensor, njobs: int, blocks: int) -> None:
    _require_cuda(table)
    _count(1)
    with _Prof("bn_stats", 0.0, 0.0):
        _lib.check(_lib.load().b200cd_bn_eval_affine_batched(table.data_ptr(), njobs, blocks, _stream()))


def conv_gemm_bnbwd(mode: int, A: torch.Tensor, Bw: torch.Tensor, out: torch.Tensor, r: torch.Tensor, scale: torch.Tensor,
                    shift: torch.Tensor, sums: torch.Tensor, stat_groups: int, bn=-1) -> None:
    """Input-gradient convolution (mode 0 with dgrad weights / mode 2) whose epilogue also accumulates the
    BatchNorm-backward sums of the gradient it stores (see include/b200cd.h: b200cd_conv_gemm_bnbwd)."""
    _require_cuda(A, Bw, out, r, sums)
    n, Ha, Wa, ka, a_ld = _nhwc(A)
    H, W = (Ha // 2, Wa // 2) if mode == 2 else (Ha, Wa)
    no, Ho, Wo, Co, o_ld = _nhwc(out)
    nr, Hr, Wr, Cr, r_ld = _nhwc(r)
    N = Bw.shape[0]
    taps = 9 if mode == 0 else 4
    assert mode in (0, 2) and Bw.shape[1] == taps * ka and (no, Ho, Wo, Co) == (n, H, W, N) == (nr, Hr, Wr, Cr)
    assert sums.dtype == torch.float32 and scale.shape[-1] == N
    if bn == -1:
        bn = tuning.tile("bnbwd", mode, 0, n, H, W, ka, N, False)
    flags = 4 | 8 | (stat_groups << 8) | tuning.flag_bits(bn)
    _count(1)
    # own profile families: these launches also do the BatchNorm-backward reduce pass of the gradient they store
    fam = "dgrad3x3_bnbwd" if mode == 0 else "convT_dgrad_bnbwd"
    with _Prof(fam, 2.0 * n * H * W * N * taps * ka, _nbytes(A, out, Bw, r), f"{n}x{H}x{W} k{ka}->n{N} bnbwd"):
        _lib.check(_lib.load().b200cd_conv_gemm_bnbwd(mode, flags, A.data_ptr(), a_ld, n, H, W, ka, Bw.data_ptr(), N,
                                                      out.data_ptr(), o_ld, r.data_ptr(), r_ld, scale.data_ptr(),
                                                      shift.data_ptr(), sums.data_ptr(), _stream()))


def wgrad_tiles(n: int, H: int, W: int) -> int:
    return _lib.load().b200cd_wgrad_tiles(n, H, W)


def wgrad_ctas_per_split(mode: int, halo: int, cu: int, cv: int) -> int:
    """CTAs one pixel-range split of wgrad_gemm launches (b200cd_wgrad_ctas_per_split)."""
    c = _lib.load().b200cd_wgrad_ctas_per_split(mode, halo, cu, cv)
    if c <= 0:
        raise ValueError(f"wgrad_ctas_per_split({mode}, {halo}, {cu}, {cv})")
    return c


def wgrad_gemm(mode: int, sign: int, halo: int, U: torch.Tensor, V: torch.Tensor, ws: torch.Tensor, splits: int,
               split_stride: int, tap_stride: int, m_stride: int, n_stride: int, splits2: int = 0,
               prec: bool = False) -> None:
    """G2. U: NHWC view at the GEMM resolution; V: NHWC view (mode 2: at 2x). prec: split-bf16 views."""
    _require_cuda(U, V, ws)
    n, H, W, cu, u_ld = _nhwc(U)
    nv, Hv, Wv, cv, v_ld = _nhwc(V)
    assert nv == n and ((Hv, Wv) == (2 * H, 2 * W) if mode == 2 else (Hv, Wv) == (H, W))
    assert ws.dtype == torch.float32
    _count(1)
    taps = 9 if mode == 0 else (1 if mode == 1 else 4)
    fn = _lib.load().b200cd_wgrad_gemm_hp if prec else _lib.load().b200cd_wgrad_gemm
    with _Prof("wgrad", 2.0 * n * H * W * cu * cv * taps, _nbytes(U, V) * (2 if prec else 1) + 4.0 * splits * taps * cu * cv,
               f"{n}x{H}x{W} m{cu} n{cv} mode{mode} sign{sign} splits{splits}+{splits2}"):
        _lib.check(fn(mode, sign, halo, U.data_ptr(), u_ld, cu, V.data_ptr(), v_ld, cv, n, H, W, ws.data_ptr(), splits,
                      splits2, split_stride, tap_stride, m_stride, n_stride, _stream()))


def wgrad_reduce(ws: torch.Tensor, splits: int, split_stride: int, layout: int, d0: int, d1: int, taps: int,
                 grad: torch.Tensor) -> None:
    _require_cuda(ws, grad)
    assert grad.dtype == torch.float32 and grad.is_contiguous() and grad.numel() == d0 * d1 * taps
    _count(1)
    with _Prof("wgrad_reduce", 0.0, 4.0 * (splits + 1) * d0 * d1 * taps):
        _lib.check(_lib.load().b200cd_wgrad_reduce(ws.data_ptr(), splits, split_stride, layout, d0, d1, taps,
                                                   grad.data_ptr(), _stream()))


_REDUCE_JOB_DTYPE = None


def make_reduce_jobs(specs: Sequence[tuple], device) -> tuple[torch.Tensor, int, int, float]:
    """specs: (ws fp32 view, grad fp32 tensor, splits, split_stride, layout, d0, d1, taps[, splits2]). Returns (device
    job table, njobs, total thread blocks, algorithmic bytes) for wgrad_reduce_batched (b200cd_reduce_job: 64 bytes)."""
    import numpy as np
    global _REDUCE_JOB_DTYPE
    if _REDUCE_JOB_DTYPE is None:
        _REDUCE_JOB_DTYPE = np.dtype([("ws", "<u8"), ("grad", "<u8"), ("split_stride", "<i8"), ("start", "<i8"),
                                      ("splits", "<i4"), ("layout", "<i4"), ("d0", "<i4"), ("d1", "<i4"), ("taps", "<i4"),
                                      ("parts", "<i4"), ("splits2", "<i4"), ("reserved", "<i4")], align=True)
        assert _REDUCE_JOB_DTYPE.itemsize == 64
    lib = _lib.load()
    arr = np.zeros(len(specs), dtype=_REDUCE_JOB_DTYPE)
    blocks = 0
    nbytes = 0.0
    for i, spec in enumerate(specs):
        ws, grad, splits, split_stride, layout, d0, d1, taps = spec[:8]
        splits2 = spec[8] if len(spec) > 8 else 0
        total = d0 * d1 * taps
        assert ws.dtype == torch.float32 and grad.dtype == torch.float32 and grad.is_contiguous() and grad.numel() == total
        assert layout == 0 and d1 % 4 == 0 and split_stride % 4 == 0 and ws.data_ptr() % 16 == 0 and \
            grad.data_ptr() % 16 == 0
        nb = lib.b200cd_reduce_job_blocks(splits, d0, d1, taps)
        assert nb > 0
        arr[i] = (ws.data_ptr(), grad.data_ptr(), split_stride, blocks, splits, layout, d0, d1, taps,
                  lib.b200cd_reduce_job_parts(splits, d1, taps), splits2, 0)
        blocks += nb
        nbytes += 4.0 * (splits + 1) * total - (4.0 * (splits - splits2) * total / 3 if splits2 else 0.0)
    table = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return table, len(specs), blocks, nbytes


def wgrad_reduce_batched(table: torch.Tensor, njobs: int, blocks: int, nbytes: float = 0.0) -> None:
    _require_cuda(table)
    _count(1)
    with _Prof("wgrad_reduce", 0.0, nbytes):
        _lib.check(_lib.load().b200cd_wgrad_reduce_batched(table.data_ptr(), njobs, blocks, _stream()))


def bn_stats(partial: Optional[torch.Tensor], ld: int, C_: int, tiles_per_group: int, G: int, count: float, spl: int,
             ws: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, running_mean: torch.Tensor,
             running_var: torch.Tensor, nbt: Optional[torch.Tensor], momentum: float, eps: float, train: bool,
             order_rev: bool, mean: torch.Tensor, invstd: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> None:
    _require_cuda(gamma, beta, running_mean, running_var, mean)
    _count(2 if train else 1)
    with _Prof("bn_stats", 0.0, 8.0 * tiles_per_group * G * C_ if train else 0.0):
        _lib.check(_lib.load().b200cd_bn_stats(_ptr(partial), ld, C_, tiles_per_group, G, float(count), spl, _ptr(ws),
                                               gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                                               running_var.data_ptr(), _ptr(nbt), momentum, eps, int(train),
                                               int(order_rev), mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(),
                                               shift.data_ptr(), _stream()))


def bn_apply(r: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, G: int, diff: bool,
             a: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None, pool: Optional[torch.Tensor] = None,
             dif: Optional[torch.Tensor] = None, pool_idx: Optional[torch.Tensor] = None, prec: bool = False) -> None:
    _require_cuda(r, scale, shift)
    n, H, W, Cc, ld_r = _nhwc(r)

    def ld(t):
        return 0 if t is None else _nhwc(t)[4]

    _count(1)
    fn = _lib.load().b200cd_bn_apply_hp if prec else _lib.load().b200cd_bn_apply
    with _Prof("bn_apply", 0.0, _nbytes(r, a, a2, pool, dif) * (2 if prec else 1),
               f"{n}x{H}x{W}x{Cc} G{G} diff{int(diff)} pool{int(pool is not None)}"):
        _lib.check(fn(r.data_ptr(), ld_r, scale.data_ptr(), shift.data_ptr(), n, H, W, Cc, G, int(diff), _ptr(a), ld(a),
                      _ptr(a2), ld(a2), _ptr(pool), ld(pool), _ptr(dif), ld(dif), _ptr(pool_idx), _stream()))


def make_srcs(srcs: Sequence[dict]) -> C.Array:
    """Build the b200cd_grad_src[3] array. Each dict: kind, t (tensor), w (tensor, kind 3), n_mod, scale_lo, scale_hi."""
    arr = (GradSrc * 3)()
    assert len(srcs) <= 3
    for i, s in enumerate(srcs):
        t = s["t"]
        arr[i].kind = s["kind"]
        arr[i].ptr = t.data_ptr()
        arr[i].w = s["w"].data_ptr() if s.get("w") is not None else None   # head weights (3) / pool arg-max index (2)
        arr[i].ld = t.stride(2) if s["kind"] in (1, 2) else 0
        arr[i].n_mod = s.get("n_mod", 0)
        arr[i].scale_lo = s.get("scale_lo", 1.0)
        arr[i].scale_hi = s.get("scale_hi", 1.0)
    return arr


def bn_bwd_ws_floats(n: int, H: int, W: int, C_: int, G: int) -> int:
    return _lib.load().b200cd_bn_bwd_ws_floats(n, H, W, C_, G)


def bn_bwd(r: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
           srcs: C.Array, G: int, ws: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor, dr: torch.Tensor,
           sums: Optional[torch.Tensor] = None, sum_rows: int = 0, prec: bool = False) -> None:
    _require_cuda(r, dr, ws)
    n, H, W, Cc, ld_r = _nhwc(r)
    nsrc = sum(1.0 if s.kind == 1 else (0.25 if s.kind == 2 else 0.0) for s in srcs)
    if prec:
        assert sums is None, "the fused BatchNorm-backward reduce is a bf16-storage feature"
        _count(3)
        with _Prof("bn_bwd", 0.0, 2.0 * _nbytes(r) * (2.0 + 2.0 * nsrc + 1.0),
                   f"{n}x{H}x{W}x{Cc} G{G} srcs{[s.kind for s in srcs]} hp"):
            _lib.check(_lib.load().b200cd_bn_bwd_hp(r.data_ptr(), ld_r, mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(),
                                                    shift.data_ptr(), srcs, n, H, W, Cc, G, ws.data_ptr(),
                                                    dgamma.data_ptr(), dbeta.data_ptr(), dr.data_ptr(), _nhwc(dr)[4],
                                                    _stream()))
        return
    if sums is not None:
        # the reduce pass ran in the epilogue of the convolution that produced the gradient: finalize + dx only
        _count(2)
        with _Prof("bn_bwd", 0.0, _nbytes(r) * (1.0 + nsrc + 1.0), f"{n}x{H}x{W}x{Cc} G{G} from_sums"):
            _lib.check(_lib.load().b200cd_bn_bwd_from_sums(r.data_ptr(), ld_r, mean.data_ptr(), invstd.data_ptr(),
                                                           scale.data_ptr(), shift.data_ptr(), srcs, n, H, W, Cc, G,
                                                           sums.data_ptr(), sum_rows, ws.data_ptr(), dgamma.data_ptr(),
                                                           dbeta.data_ptr(), dr.data_ptr(), _nhwc(dr)[4], _stream()))
        return
    _count(3)
    # two passes: each reads r and every gradient source; the second writes dr
    with _Prof("bn_bwd", 0.0, _nbytes(r) * (2.0 + 2.0 * nsrc + 1.0),
               f"{n}x{H}x{W}x{Cc} G{G} srcs{[s.kind for s in srcs]}"):
        _lib.check(_lib.load().b200cd_bn_bwd(r.data_ptr(), ld_r, mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(),
                                             shift.data_ptr(), srcs, n, H, W, Cc, G, ws.data_ptr(), dgamma.data_ptr(),
                                             dbeta.data_ptr(), dr.data_ptr(), _nhwc(dr)[4], _stream()))


def head_fwd(a0: torch.Tensor, a1: Optional[torch.Tensor], w: torch.Tensor, b: torch.Tensor,
             logits: torch.Tensor, prec: bool = False) -> None:
    _require_cuda(a0, w, b, logits)
    n, H, W, Cc, ld0 = _nhwc(a0)
    ld1 = 0 if a1 is None else _nhwc(a1)[4]
    _count(1)
    fn = _lib.load().b200cd_head_fwd_hp if prec else _lib.load().b200cd_head_fwd
    with _Prof("head_fwd", 0.0, _nbytes(a0, a1) * (2 if prec else 1) + _nbytes(logits)):
        _lib.check(fn(a0.data_ptr(), ld0, _ptr(a1), ld1, Cc, w.data_ptr(), b.data_ptr(), n * H * W, logits.data_ptr(),
                      _stream()))


def pad_copy(src: torch.Tensor, dst: torch.Tensor, top: int, left: int, prec: bool = False) -> None:
    """dst (NHWC view, e.g. the upper half of a concat buffer) = src placed at (top, left), zero border:
    Up's centre pad, utils/networks.py:440-443."""
    _require_cuda(src, dst)
    if prec:   # both halves of the split tensors: two plain copies
        pad_copy(src, dst, top, left)
        pad_copy(lo_half(src), lo_half(dst), top, left)
        return
    n, h, w, Cc, ld_s = _nhwc(src)
    n2, H, W, C2, ld_d = _nhwc(dst)
    assert n == n2 and Cc == C2
    _count(1)
    with _Prof("pad_copy", 0.0, 2.0 * n * (h * w + H * W) * Cc):
        _lib.check(_lib.load().b200cd_pad_copy(src.data_ptr(), ld_s, n, h, w, Cc, dst.data_ptr(), ld_d, H, W, top, left,
                                               _stream()))


def colsum(x: Optional[torch.Tensor], wgt: Optional[torch.Tensor], npix: int, nblk: int, ws: torch.Tensor,
           out: torch.Tensor, prec: bool = False) -> None:
    _require_cuda(ws, out)
    if x is None:
        Cc, ld = 1, 0
    else:
        Cc, ld = x.shape[3], x.stride(2)
    _count(2)
    hp = prec and x is not None
    fn = _lib.load().b200cd_colsum_hp if hp else _lib.load().b200cd_colsum
    with _Prof("colsum", 0.0, npix * ((4.0 if hp else 2.0) * Cc if x is not None else 0.0) + (4.0 * npix if wgt is not None else 0.0)):
        _lib.check(fn(_ptr(x), ld, Cc, _ptr(wgt), npix, nblk, ws.data_ptr(), out.data_ptr(), _stream()))


def stat_rowsum(stats: torch.Tensor, rows: int, ld: int, c_off: int, Cc: int, out: torch.Tensor) -> None:
    """out[j] = sum over the per-CTA statistics rows of channel c_off + j (transposed-conv bias gradient)."""
    _require_cuda(stats, out)
    assert stats.dtype == torch.float32 and out.dtype == torch.float32 and out.numel() == Cc
    _count(1)
    with _Prof("colsum", 0.0, 8.0 * rows * Cc):
        _lib.check(_lib.load().b200cd_stat_rowsum(stats.data_ptr(), rows, ld, c_off, Cc, out.data_ptr(), _stream()))


def pj_fwd(z: torch.Tensor, t: torch.Tensor, t_is_logit: bool, rowmask: Optional[torch.Tensor], sel: int, nblk: int,
           ws: torch.Tensor, sums: torch.Tensor) -> None:
    _require_cuda(z, t, ws, sums)
    rows = z.shape[0]
    per_row = z.numel() // rows
    _count(2)
    with _Prof("pj_fwd", 0.0, 8.0 * z.numel()):
        _lib.check(_lib.load().b200cd_pj_fwd(z.data_ptr(), t.data_ptr(), int(t_is_logit), _ptr(rowmask), sel, rows,
                                             per_row, nblk, ws.data_ptr(), sums.data_ptr(), _stream()))


def pj_loss(sums: torch.Tensor, loss: torch.Tensor) -> None:
    _require_cuda(sums, loss)
    _count(1)
    _lib.check(_lib.load().b200cd_pj_loss(sums.data_ptr(), loss.data_ptr(), _stream()))


def pj_bwd(z: torch.Tensor, t: torch.Tensor, t_is_logit: bool, rowmask: Optional[torch.Tensor], sel: int,
           sums: torch.Tensor, gptr: Optional[torch.Tensor], gmul: float, accumulate: bool, dz: torch.Tensor,
           dt: Optional[torch.Tensor]) -> None:
    _require_cuda(z, t, dz)
    rows = z.shape[0]
    per_row = z.numel() // rows
    _count(1)
    with _Prof("pj_bwd", 0.0, (12.0 + (4.0 if dt is not None else 0.0)) * z.numel()):
        _lib.check(_lib.load().b200cd_pj_bwd(z.data_ptr(), t.data_ptr(), int(t_is_logit), _ptr(rowmask), sel, rows,
                                             per_row, sums.data_ptr(), _ptr(gptr), gmul, int(accumulate), dz.data_ptr(),
                                             _ptr(dt), _stream()))
