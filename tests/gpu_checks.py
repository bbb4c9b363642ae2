"""Per-kernel parity checks of the CUDA path (through the C ABI) against plain torch fp32 on the same
bf16-rounded inputs. Imported by tests/test_gpu_ops.py (pytest -m gpu) and tools/gpu_probe.py.

Each check returns a dict of metrics; `ok` says whether the tolerance written next to it was met.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from multimodal_siamese_cd_b200 import ops

DEV = "cuda"


def bf16r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def nhwc(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def err(got: torch.Tensor, ref: torch.Tensor, bf16_out: bool = False) -> dict:
    """rel_l2 against the fp32 reference; for kernels whose OUTPUT is stored in bf16 also the comparison against
    the reference rounded to bf16 (what a perfect kernel would store): `ulp_frac` = fraction of elements that differ
    from it at all, `rel_l2_rounded` their size (an element differs only when the fp32 sums straddle a rounding
    boundary, and then by exactly one bf16 ulp)."""
    got = got.double()
    out = {}
    if bf16_out:
        rr = ref.to(torch.bfloat16).double()
        out["ulp_frac"] = (got != rr).double().mean().item()
        out["rel_l2_rounded"] = ((got - rr).norm() / rr.norm().clamp_min(1e-30)).item()
    ref = ref.double()
    d = (got - ref)
    out.update({
        "max_abs": d.abs().max().item(),
        "rel_l2": (d.norm() / ref.norm().clamp_min(1e-30)).item(),
        "ref_absmax": ref.abs().max().item(),
        "finite": bool(torch.isfinite(got).all().item()),
    })
    return out


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    return g


# ----------------------------------------------------------------------------------------------------
def check_pack_weights() -> dict:
    g = _gen(1)
    w = torch.randn(128, 64, 3, 3, device=DEV, generator=g)
    wt = torch.randn(128, 64, 2, 2, device=DEV, generator=g)  # convT: [ci][co][2][2]
    w1 = torch.randn(64, 6, 3, 3, device=DEV, generator=g)
    out = {}
    p0 = ops.pack_weights(0, w)
    ref0 = w.permute(0, 2, 3, 1).reshape(128, 9 * 64).to(torch.bfloat16)
    out["m0"] = bool(torch.equal(p0, ref0))
    p1 = ops.pack_weights(1, w)
    ref1 = w.flip(2, 3).permute(1, 2, 3, 0).reshape(64, 9 * 128).to(torch.bfloat16)
    out["m1"] = bool(torch.equal(p1, ref1))
    p2 = ops.pack_weights(2, w1, kpad=64)
    ref2 = torch.zeros(64, 64, device=DEV)
    ref2[:, :54] = w1.permute(0, 2, 3, 1).reshape(64, 54)
    out["m2"] = bool(torch.equal(p2, ref2.to(torch.bfloat16)))
    p3 = ops.pack_weights(3, wt)
    ref3 = wt.permute(2, 3, 1, 0).reshape(4 * 64, 128).to(torch.bfloat16)
    out["m3"] = bool(torch.equal(p3, ref3))
    p4 = ops.pack_weights(4, wt)
    ref4 = wt.permute(0, 2, 3, 1).reshape(128, 4 * 64).to(torch.bfloat16)
    out["m4"] = bool(torch.equal(p4, ref4))
    # every job of a network in one launch, forward + input-gradient layouts written from one read of the weights
    w_big = torch.randn(256, 128, 3, 3, device=DEV, generator=g)
    b0, b1 = torch.empty_like(p0), torch.empty_like(p1)
    b3, b4 = torch.empty_like(p3), torch.empty_like(p4)
    b2 = torch.empty_like(p2)
    bb0 = torch.empty(256, 9 * 128, device=DEV, dtype=torch.bfloat16)
    bb1 = torch.empty(128, 9 * 256, device=DEV, dtype=torch.bfloat16)
    only1 = torch.empty_like(p1)
    tab, nj, blocks, elems = ops.make_pack_jobs([(0, w, b0, 0, 1, b1), (2, w1, b2, 64), (3, wt, b3, 0, 4, b4),
                                                 (0, w_big, bb0, 0, 1, bb1), (1, w, only1, 0)], DEV)
    ops.pack_weights_batched(tab, nj, blocks, elems)
    out["batched"] = bool(torch.equal(b0, ref0) and torch.equal(b1, ref1) and torch.equal(b2, ref2.to(torch.bfloat16)) and
                          torch.equal(b3, ref3) and torch.equal(b4, ref4) and torch.equal(only1, ref1) and
                          torch.equal(bb0, w_big.permute(0, 2, 3, 1).reshape(256, 9 * 128).to(torch.bfloat16)) and
                          torch.equal(bb1, w_big.flip(2, 3).permute(1, 2, 3, 0).reshape(128, 9 * 256).to(torch.bfloat16)))
    out["ok"] = all(out.values())
    return out


def im2col_ref(x: torch.Tensor, kpad: int) -> torch.Tensor:
    """x: [n, C, H, W] fp32 -> [n, H, W, kpad] with k = tap*C + c."""
    n, Cc, H, W = x.shape
    cols = F.unfold(x, 3, padding=1).view(n, Cc, 9, H, W)  # [n, c, tap, H, W]
    cols = cols.permute(0, 3, 4, 2, 1).reshape(n, H, W, 9 * Cc)
    out = torch.zeros(n, H, W, kpad, device=x.device)
    out[..., : 9 * Cc] = cols
    return out.to(torch.bfloat16)


def check_pack_input() -> dict:
    g = _gen(2)
    x1 = torch.rand(2, 6, 24, 40, device=DEV, generator=g)
    x2 = torch.rand(2, 6, 24, 40, device=DEV, generator=g)
    out = {}
    got = ops.pack_input(x1, x2, 2, 4, 0, 64)  # siamese on the S2 bands
    ref = im2col_ref(torch.cat([x1[:, 2:6], x2[:, 2:6]], 0), 64)
    out["cat_batch"] = bool(torch.equal(got, ref))
    got = ops.pack_input(x1, x2, 0, 2, 1, 64)  # early fusion of the S1 bands
    ref = im2col_ref(torch.cat([x1[:, 0:2], x2[:, 0:2]], 1), 64)
    out["cat_chan"] = bool(torch.equal(got, ref))
    got = ops.pack_input(x1, x2, 2, 4, 1, 128)  # 8 channels -> K = 72 -> kpad 128
    ref = im2col_ref(torch.cat([x1[:, 2:6], x2[:, 2:6]], 1), 128)
    out["cat_chan8"] = bool(torch.equal(got, ref))
    out["ok"] = all(out.values())
    return out


# ----------------------------------------------------------------------------------------------------
def check_conv3x3(n=2, H=32, W=32, cin=64, cout=64, bias=True, slice_in=False, slice_out=False, seed=3,
                  tol=6e-3, halo=None, wide=None, pair=None, compare_single=False) -> dict:
    """conv_gemm mode 0 vs F.conv2d (fp32 math on the bf16-rounded operands); output is bf16 so the bound is
    one bf16 ulp (2^-8 relative) plus accumulation-order noise."""
    g = _gen(seed)
    x = bf16r(torch.randn(n, cin, H, W, device=DEV, generator=g))
    w = bf16r(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(cout, device=DEV, generator=g) if bias else None
    if slice_in:
        buf = torch.zeros(n, H, W, cin + 64, device=DEV, dtype=torch.bfloat16)
        A = buf[..., 64:]
        A.copy_(nhwc(x))
    else:
        A = nhwc(x).to(torch.bfloat16)
    if slice_out:
        obuf = torch.full((n, H, W, cout + 64), 7.0, device=DEV, dtype=torch.bfloat16)
        out = obuf[..., :cout]
    else:
        obuf = None
        out = torch.empty(n, H, W, cout, device=DEV, dtype=torch.bfloat16)
    tiles = ops.conv_gemm_tiles(H, W)
    stats = torch.zeros(n * tiles, cout, 2, device=DEV)
    Bw = ops.pack_weights(0, w)
    ops.conv_gemm(0, 0, A, Bw, out, bias=b, stats=stats, halo=halo, wide=wide, pair=pair)
    ops.device_status()
    ref = F.conv2d(x, w, b, padding=1)
    res = err(nchw(out.float()), ref, bf16_out=True)
    if compare_single:  # the CTA-pair kernel runs the K loop in the same order as the single-CTA halo kernel
        out1 = torch.empty(n, H, W, cout, device=DEV, dtype=torch.bfloat16)
        stats1 = torch.zeros_like(stats)
        ops.conv_gemm(0, 0, A, Bw, out1, bias=b, stats=stats1, halo=True, wide=False, pair=False)
        ops.device_status()
        # outputs bit-identical; the per-tile statistics are summed in a different (fixed) order by the two kernels
        res["same_as_single"] = bool(torch.equal(out1, out.contiguous()) and
                                     torch.allclose(stats1, stats, rtol=2e-5, atol=1e-4))
    got_sum = stats[..., 0].sum(0)
    got_sq = stats[..., 1].sum(0)
    o32 = out.float()
    res["stats_sum_rel"] = ((got_sum - o32.sum((0, 1, 2))).norm() / o32.sum((0, 1, 2)).norm().clamp_min(1e-6)).item()
    res["stats_sq_rel"] = ((got_sq - (o32 * o32).sum((0, 1, 2))).norm() / (o32 * o32).sum((0, 1, 2)).norm()).item()
    if obuf is not None:
        res["untouched"] = bool((obuf[..., cout:] == 7.0).all().item())
    # the share of outputs that land on the other side of a bf16 rounding boundary than the fp32 reference grows with
    # the reduction length (fp32 summation-order noise): 5e-3 up to K = 4608, proportionally more beyond
    res["ulp_tol"] = 5e-3 * max(1.0, 9 * cin / 4608)
    res["ok"] = res["finite"] and res["rel_l2"] < tol and res["stats_sum_rel"] < 1e-3 and res["stats_sq_rel"] < 1e-3 \
        and res.get("untouched", True) and res.get("same_as_single", True)
    return res


def check_conv3x3_cta_stats(n=4, H=32, W=32, cin=64, cout=128, G=2, seed=48) -> dict:
    """CTA-pair conv with per-CTA running BatchNorm statistics (flags bit 3): stats[g][row][n] summed over the rows must
    equal the per-group sums of the stored output, and the fused statistics kernel must reproduce BatchNorm's batch
    statistics and running-stat update per group."""
    g = _gen(seed)
    x = bf16r(torch.randn(n, cin, H, W, device=DEV, generator=g))
    w = bf16r(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(cout, device=DEV, generator=g)
    A = nhwc(x).to(torch.bfloat16)
    out = torch.empty(n, H, W, cout, device=DEV, dtype=torch.bfloat16)
    rows, per_cta = ops.conv_stat_rows(n, H, W, cin, cout, G)
    stats = torch.full((G, rows, cout, 2), float("nan"), device=DEV)   # the kernel must define every row it owns
    Bw = ops.pack_weights(0, w)
    ops.conv_gemm(0, 0, A, Bw, out, bias=b, stats=stats, stat_groups=G)
    ops.device_status()
    res = {"per_cta": per_cta, "rows": rows, "finite": bool(torch.isfinite(stats).all().item())}
    o32 = out.float().view(G, n // G, H, W, cout)
    ref_sum = o32.sum((1, 2, 3))
    ref_sq = (o32 * o32).sum((1, 2, 3))
    res["sum_rel"] = ((stats[..., 0].sum(1) - ref_sum).norm() / ref_sum.norm()).item()
    res["sq_rel"] = ((stats[..., 1].sum(1) - ref_sq).norm() / ref_sq.norm()).item()
    gamma = torch.rand(cout, device=DEV, generator=g) + 0.5
    beta = torch.randn(cout, device=DEV, generator=g) * 0.2
    mean = torch.empty(G, cout, device=DEV)
    invstd, scale, shift = torch.empty_like(mean), torch.empty_like(mean), torch.empty_like(mean)
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    nbt = torch.zeros((), device=DEV, dtype=torch.int64)
    ws = torch.empty(4, G, cout, 2, device=DEV, dtype=torch.float64)
    ops.bn_stats(stats, cout, cout, rows, G, (n // G) * H * W, 4, ws, gamma, beta, rm, rv, nbt, 0.1, 1e-5, True, False,
                 mean, invstd, scale, shift)
    torch.cuda.synchronize()
    bn = torch.nn.BatchNorm2d(cout).to(DEV).train()
    xs = nchw(out.float())
    for gi in range(G):
        bn(xs[gi * (n // G):(gi + 1) * (n // G)])
        mu = xs[gi * (n // G):(gi + 1) * (n // G)].mean((0, 2, 3))
        res[f"mean_g{gi}"] = (mean[gi] - mu).abs().max().item()
    res["running_mean"] = (rm - bn.running_mean).abs().max().item()
    res["running_var"] = (rv - bn.running_var).abs().max().item()
    res["nbt"] = int(nbt.item())
    res["ok"] = (res["per_cta"] and res["finite"] and res["sum_rel"] < 1e-5 and res["sq_rel"] < 1e-5 and
                 res["running_mean"] < 1e-5 and res["running_var"] < 1e-4 and res["nbt"] == G and
                 all(res[f"mean_g{gi}"] < 1e-5 for gi in range(G)))
    return res


def check_conv3x3_dgrad(n=2, H=32, W=32, cin=64, cout=128, seed=4, tol=6e-3) -> dict:
    g = _gen(seed)
    w = bf16r(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cout ** 0.5))
    dr = bf16r(torch.randn(n, cout, H, W, device=DEV, generator=g))
    Bd = ops.pack_weights(1, w)  # [cin][9*cout]
    out = torch.empty(n, H, W, cin, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm(0, 0, nhwc(dr).to(torch.bfloat16), Bd, out)
    ops.device_status()
    ref = F.conv_transpose2d(dr, w, padding=1)  # = input gradient of conv2d(x, w, padding=1)
    res = err(nchw(out.float()), ref, bf16_out=True)
    res["ok"] = res["finite"] and res["rel_l2"] < tol
    return res


def check_stat_rowsum(n=2, H=32, W=32, c=64, cout=128, seed=71, prec=False) -> dict:
    """The transposed-conv BIAS gradient as the engine computes it by default (engine.py: UP_BIAS_FROM_STATS): the input
    gradient convolution that writes the concat-buffer gradient d_cat [n, H, W, 2c] runs with per-CTA statistics
    (stat_groups = 1, the `up_rows` side output), and b200cd_stat_rowsum sums the rows of the UPPER c channels:
    out[j] = sum over pixels of d_cat[..., c + j]. Checked against the pixel sums of the stored gradient (all 2c
    channels of the side output, then the rowsum of the upper half), in both numerics modes."""
    g = _gen(seed)
    w = torch.randn(cout, 2 * c, 3, 3, device=DEV, generator=g) / (3.0 * cout ** 0.5)
    dr = torch.randn(n, H, W, cout, device=DEV, generator=g)
    if prec:
        import gpu_checks_hp as hp
        Bd = hp.pack_hp(1, w)
        A = ops.split_from_float(dr)
        d_cat = ops.split_alloc((n, H, W, 2 * c), DEV)
    else:
        Bd = ops.pack_weights(1, bf16r(w))
        A = dr.to(torch.bfloat16)
        d_cat = torch.empty(n, H, W, 2 * c, device=DEV, dtype=torch.bfloat16)
    rows, per_cta = ops.conv_stat_rows(n, H, W, cout, 2 * c, 1, prec=prec)
    stats = torch.full((rows, 2 * c, 2), float("nan"), device=DEV)
    ops.conv_gemm(0, 0, A, Bd, d_cat, stats=stats, stat_groups=1, prec=prec)
    gb = torch.full((c,), float("nan"), device=DEV)
    ops.stat_rowsum(stats, rows, 2 * c, c, c, gb)
    ops.device_status()
    stored = (ops.split_to_float(d_cat) if prec else d_cat.float()).double()
    ref_all = stored.sum((0, 1, 2))
    scale = stored.abs().sum((0, 1, 2)).max().item()     # sums cancel: compare against the size of what was summed
    res = {"per_cta": per_cta, "rows": rows,
           "side_output_abs": ((stats[..., 0].double().sum(0) - ref_all).abs().max().item()) / scale,
           "bias_grad_abs": ((gb.double() - ref_all[c:]).abs().max().item()) / scale,
           "finite": bool(torch.isfinite(gb).all().item())}
    # against what autograd computes for ConvTranspose2d.bias: the pixel sum of the exact upstream gradient
    exact = F.conv_transpose2d(nchw(dr).double(), w.double(), padding=1).sum((0, 2, 3))[c:]
    res["vs_exact_rel"] = ((gb.double() - exact).norm() / exact.norm()).item()
    res["ok"] = res["per_cta"] and res["finite"] and res["side_output_abs"] < 2e-6 and res["bias_grad_abs"] < 2e-6 and \
        res["vs_exact_rel"] < (1e-4 if prec else 2e-2)
    return res


def check_dgrad_bnbwd(n=4, H=32, W=32, cin=64, cout=128, G=2, seed=66, mode=0) -> dict:
    """Input-gradient convolution with the BatchNorm-backward reduce pass fused into its epilogue: the gradient it
    stores must equal the plain dgrad launch bit for bit, the per-CTA sums (S1, S2) must equal sum dy*m and sum dy*m*r
    of the stored gradient, and bn_bwd_from_sums must reproduce bn_bwd."""
    g = _gen(seed)
    if mode == 0:
        w = bf16r(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cout ** 0.5))
        dout = bf16r(torch.randn(n, cout, H, W, device=DEV, generator=g))
        Bd = ops.pack_weights(1, w)                                   # [cin][9*cout]
        A = nhwc(dout).to(torch.bfloat16)
    else:
        w = bf16r(torch.randn(cin, cin, 2, 2, device=DEV, generator=g) / (2.0 * cin ** 0.5))
        dout = bf16r(torch.randn(n, cin, 2 * H, 2 * W, device=DEV, generator=g))
        Bd = ops.pack_weights(4, w)                                   # [cin][4*cin]
        A = nhwc(dout).to(torch.bfloat16)
    r = bf16r(torch.randn(n, H, W, cin, device=DEV, generator=g)).to(torch.bfloat16)   # pre-BN tensor of the producer
    scale = torch.rand(G, cin, device=DEV, generator=g) + 0.5
    scale[:, ::7] *= -1.0                                                               # negative gamma too
    shift = torch.randn(G, cin, device=DEV, generator=g) * 0.3
    mean = torch.randn(G, cin, device=DEV, generator=g) * 0.1
    invstd = torch.rand(G, cin, device=DEV, generator=g) + 0.5
    rows, per_cta = ops.conv_stat_rows(n, H, W, cout if mode == 0 else cin, cin, G, mode=mode)
    sums = torch.full((G, rows, cin, 2), float("nan"), device=DEV)
    d_fused = torch.empty(n, H, W, cin, device=DEV, dtype=torch.bfloat16)
    d_plain = torch.empty_like(d_fused)
    ops.conv_gemm_bnbwd(mode, A, Bd, d_fused, r, scale, shift, sums, G)
    ops.conv_gemm(mode, 0, A, Bd, d_plain)
    ops.device_status()
    res = {"per_cta": per_cta, "same_gradient": bool(torch.equal(d_fused, d_plain)),
           "finite": bool(torch.isfinite(sums).all().item())}
    dy = d_fused.float().view(G, n // G, H, W, cin)
    rr = r.float().view(G, n // G, H, W, cin)
    m = (torch.addcmul(shift.view(G, 1, 1, 1, cin), rr, scale.view(G, 1, 1, 1, cin)) > 0).float()
    s1 = (dy * m).sum((1, 2, 3))
    s2 = (dy * m * rr).sum((1, 2, 3))
    res["s1_rel"] = ((sums[..., 0].sum(1) - s1).norm() / s1.norm()).item()
    res["s2_rel"] = ((sums[..., 1].sum(1) - s2).norm() / s2.norm()).item()
    # bn_bwd from the fused sums vs the two-pass bn_bwd
    srcs = ops.make_srcs([{"kind": 1, "t": d_fused}])
    ws = torch.empty(ops.bn_bwd_ws_floats(n, H, W, cin, G), device=DEV)
    out = {}
    for tag, kw in (("two_pass", {}), ("from_sums", {"sums": sums, "sum_rows": rows})):
        dg, db = torch.empty(cin, device=DEV), torch.empty(cin, device=DEV)
        dr = torch.empty(n, H, W, cin, device=DEV, dtype=torch.bfloat16)
        ops.bn_bwd(r, mean, invstd, scale, shift, srcs, G, ws, dg, db, dr, **kw)
        out[tag] = (dg, db, dr.float())
    torch.cuda.synchronize()
    res["dgamma_rel"] = ((out["from_sums"][0] - out["two_pass"][0]).norm() / out["two_pass"][0].norm()).item()
    res["dbeta_rel"] = ((out["from_sums"][1] - out["two_pass"][1]).norm() / out["two_pass"][1].norm()).item()
    res["dr_rel"] = ((out["from_sums"][2] - out["two_pass"][2]).norm() / out["two_pass"][2].norm()).item()
    res["ok"] = (res["per_cta"] and res["same_gradient"] and res["finite"] and res["s1_rel"] < 1e-5 and res["s2_rel"] < 1e-5
                 and res["dgamma_rel"] < 1e-5 and res["dbeta_rel"] < 1e-5 and res["dr_rel"] < 2e-3)
    return res


def check_forced_tiles(n=4, H=32, W=32, cin=64, cout=256, G=2, seed=120) -> dict:
    """The N tile of the CTA-pair kernel chosen by the caller (flags bits 5..6 of b200cd_conv_gemm, tuning.py): every
    tile width that divides N must give the SAME output bit for bit as the library's own choice (the K loop runs in the
    same order) — forward with per-CTA statistics, plain, the fused BatchNorm-backward dgrad and the transposed-conv
    scatter epilogue — and statistics / sums whose column totals agree."""
    g = _gen(seed)
    x = bf16r(torch.randn(n, cin, H, W, device=DEV, generator=g))
    w = bf16r(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(cout, device=DEV, generator=g)
    A = nhwc(x).to(torch.bfloat16)
    Bw = ops.pack_weights(0, w)
    res = {"ok": True}
    base = {}
    for bn in [None] + [t for t in (64, 128, 256) if cout % t == 0]:
        tag = "auto" if bn is None else str(bn)
        rows, per_cta = ops.conv_stat_rows(n, H, W, cin, cout, G, bn=bn)
        stats = torch.full((G, rows, cout, 2), float("nan"), device=DEV)
        o_s = torch.empty(n, H, W, cout, device=DEV, dtype=torch.bfloat16)
        o_p = torch.empty_like(o_s)
        ops.conv_gemm(0, 0, A, Bw, o_s, bias=b, stats=stats, stat_groups=G, bn=bn)
        ops.conv_gemm(0, 0, A, Bw, o_p, bias=b, bn=bn)
        # fused BatchNorm-backward dgrad: A plays the gradient, the output has `cout` channels
        r = bf16r(torch.randn(n, H, W, cout, device=DEV, generator=_gen(seed + 1))).to(torch.bfloat16)
        sc = torch.rand(G, cout, device=DEV, generator=_gen(seed + 2)) + 0.5
        sh = torch.randn(G, cout, device=DEV, generator=_gen(seed + 3)) * 0.3
        rows_b, _ = ops.conv_stat_rows(n, H, W, cin, cout, G, variant="bnbwd", bn=bn)
        sums = torch.full((G, rows_b, cout, 2), float("nan"), device=DEV)
        o_b = torch.empty_like(o_s)
        ops.conv_gemm_bnbwd(0, A, Bw, o_b, r, sc, sh, sums, G, bn=bn)
        ops.device_status()
        cur = {"stats_out": o_s, "plain_out": o_p, "bnbwd_out": o_b, "stat_tot": stats.double().sum(1), "sum_tot": sums.double().sum(1),
               "finite": bool(torch.isfinite(stats).all().item() and torch.isfinite(sums).all().item())}
        res[f"rows_{tag}"] = rows
        if bn is None:
            base = cur
            res["ok"] = res["ok"] and per_cta and cur["finite"] and bool(torch.equal(o_s, o_p))
            continue
        same = all(bool(torch.equal(cur[k], base[k])) for k in ("stats_out", "plain_out", "bnbwd_out"))
        st_rel = ((cur["stat_tot"] - base["stat_tot"]).norm() / base["stat_tot"].norm()).item()
        su_rel = ((cur["sum_tot"] - base["sum_tot"]).norm() / base["sum_tot"].norm()).item()
        res[f"same_{tag}"], res[f"stat_rel_{tag}"], res[f"sums_rel_{tag}"] = same, st_rel, su_rel
        res["ok"] = res["ok"] and same and cur["finite"] and st_rel < 1e-6 and su_rel < 1e-5
    # transposed-conv forward (single tap, scatter epilogue): N = 4 * c
    c = cin
    wt = bf16r(torch.randn(c, c, 2, 2, device=DEV, generator=g) / (2.0 * c ** 0.5))
    bt = torch.randn(c, device=DEV, generator=g)
    Bt = ops.pack_weights(3, wt)
    outs = []
    for bn in [None] + [t for t in (64, 128, 256) if (4 * c) % t == 0]:
        cat = torch.zeros(n, 2 * H, 2 * W, 2 * c, device=DEV, dtype=torch.bfloat16)
        ops.conv_gemm(1, 1, A, Bt, cat[..., c:], bias=bt, bn=bn)
        ops.device_status()
        outs.append(cat)
    res["convt_same"] = all(bool(torch.equal(o, outs[0])) for o in outs[1:])
    res["ok"] = res["ok"] and res["convt_same"]
    return res


def check_pad_copy(n=2, h=8, w_=10, c=64, H=9, W=11, top=0, left=0, seed=90) -> dict:
    """Up's centre pad (utils/networks.py:440-443): dense tensor placed in the upper half of a concat buffer."""
    g = _gen(seed)
    src = torch.randn(n, h, w_, c, device=DEV, generator=g).to(torch.bfloat16)
    cat = torch.full((n, H, W, 2 * c), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.pad_copy(src, cat[..., c:], top, left)
    ops.device_status()
    ref = torch.nn.functional.pad(src.permute(0, 3, 1, 2), (left, W - w_ - left, top, H - h - top)).permute(0, 2, 3, 1)
    ok = bool(torch.equal(cat[..., c:], ref)) and bool((cat[..., :c] == 3.0).all())
    return {"ok": ok}


def check_convt_fwd(n=2, h=16, w_=16, c=128, seed=5, tol=6e-3, pair=None) -> dict:
    """ConvTranspose2d(c, c, 2, stride=2) forward, scattered into the second half of a 2c concat buffer."""
    g = _gen(seed)
    x = bf16r(torch.randn(n, c, h, w_, device=DEV, generator=g))
    wt = bf16r(torch.randn(c, c, 2, 2, device=DEV, generator=g) / (c ** 0.5))
    b = torch.randn(c, device=DEV, generator=g)
    cat = torch.full((n, 2 * h, 2 * w_, 2 * c), 3.0, device=DEV, dtype=torch.bfloat16)
    Bw = ops.pack_weights(3, wt)  # [4c][c]
    ops.conv_gemm(1, 1, nhwc(x).to(torch.bfloat16), Bw, cat[..., c:], bias=b, pair=pair)
    ops.device_status()
    ref = F.conv_transpose2d(x, wt, b, stride=2)
    res = err(nchw(cat[..., c:].float()), ref, bf16_out=True)
    res["untouched"] = bool((cat[..., :c] == 3.0).all().item())
    res["ok"] = res["finite"] and res["rel_l2"] < tol and res["untouched"]
    return res


def check_first_conv(B=3, H=32, W=32, nc=4, cat_mode=0, seed=55, pair=None, G=2) -> dict:
    """First-layer 3x3 conv (Cin = 2..8) as a single-tap GEMM over the im2col rows the input packer writes, with the
    BatchNorm partial statistics (per tile, or per CTA with the CTA-pair kernel)."""
    g = _gen(seed)
    x1 = torch.rand(B, 6, H, W, device=DEV, generator=g)
    x2 = torch.rand(B, 6, H, W, device=DEV, generator=g)
    cin = 2 * nc if cat_mode else nc
    kpad = 64 * ((9 * cin + 63) // 64)
    w = bf16r(torch.randn(64, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(64, device=DEV, generator=g)
    cols = ops.pack_input(x1, x2, 2, nc, cat_mode, kpad)
    n_img = cols.shape[0]
    Bw = ops.pack_weights(2, w, kpad=kpad)
    out = torch.empty(n_img, H, W, 64, device=DEV, dtype=torch.bfloat16)
    Gs = G if (pair is not False and n_img % G == 0) else 0
    if Gs:
        rows, per_cta = ops.conv_stat_rows(n_img, H, W, kpad, 64, Gs, mode=1)
    else:
        rows, per_cta = n_img * ops.conv_gemm_tiles(H, W), False
    stats = torch.zeros((Gs if per_cta else 1), rows if per_cta else n_img * ops.conv_gemm_tiles(H, W), 64, 2, device=DEV)
    ops.conv_gemm(1, 0, cols, Bw, out, bias=b, stats=stats, pair=pair, stat_groups=Gs if per_cta else 0)
    ops.device_status()
    xin = torch.cat([x1[:, 2:2 + nc], x2[:, 2:2 + nc]], 1 if cat_mode else 0)
    ref = F.conv2d(bf16r(xin), w, b, padding=1)
    res = err(nchw(out.float()), ref, bf16_out=True)
    o32 = out.float()
    res["stats_sum_rel"] = ((stats[..., 0].sum((0, 1)) - o32.sum((0, 1, 2))).norm() / o32.sum((0, 1, 2)).norm()).item()
    res["per_cta"] = per_cta
    res["ok"] = res["finite"] and res["rel_l2"] < 6e-3 and res["stats_sum_rel"] < 1e-4
    return res


def check_convt_dgrad(n=2, h=16, w_=16, c=128, seed=6, tol=6e-3, pair=None) -> dict:
    g = _gen(seed)
    wt = bf16r(torch.randn(c, c, 2, 2, device=DEV, generator=g) / (2.0 * c ** 0.5))
    dcat = torch.zeros(n, 2 * h, 2 * w_, 2 * c, device=DEV, dtype=torch.bfloat16)
    dout = bf16r(torch.randn(n, c, 2 * h, 2 * w_, device=DEV, generator=g))
    dcat[..., c:] = nhwc(dout).to(torch.bfloat16)
    Bd = ops.pack_weights(4, wt)  # [c][4c]
    out = torch.empty(n, h, w_, c, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm(2, 0, dcat[..., c:], Bd, out, pair=pair)
    ops.device_status()
    ref = F.conv2d(dout, wt, stride=2)  # input gradient of conv_transpose2d(x, wt, stride=2)
    res = err(nchw(out.float()), ref, bf16_out=True)
    res["ok"] = res["finite"] and res["rel_l2"] < tol
    return res


def _splits_for(total: int, want: int) -> int:
    return max(1, min(total, want))


def check_wgrad3x3(n=2, H=32, W=32, cin=64, cout=128, sign=1, halo=0, splits=5, seed=7, tol=2e-3, splits2=0,
                   batched=False) -> dict:
    """wgrad_gemm mode 0 + wgrad_reduce vs autograd's weight gradient (fp32 accumulation on both sides).
    splits2 > 0: CTAs own two kx columns (64-wide N tiles); the kx = 2 taps then have splits2 partials and the
    workspace slots beyond them stay NaN — the batched reduction must not touch them."""
    g = _gen(seed)
    x = bf16r(torch.randn(n, cin, H, W, device=DEV, generator=g))
    dr = bf16r(torch.randn(n, cout, H, W, device=DEV, generator=g))
    xa = nhwc(x).to(torch.bfloat16)
    da = nhwc(dr).to(torch.bfloat16)
    splits = _splits_for(ops.wgrad_tiles(n, H, W), splits)
    ws = torch.full((splits, 9, cout, cin), float("nan"), device=DEV)
    if sign == 1:   # M <-> cout (U = dOut), N <-> cin (V = input, shifted)
        ops.wgrad_gemm(0, 1, halo, da, xa, ws, splits, 9 * cout * cin, cout * cin, cin, 1, splits2)
    else:           # M <-> cin (U = input), N <-> cout (V = dOut, shifted the other way)
        ops.wgrad_gemm(0, -1, halo, xa, da, ws, splits, 9 * cout * cin, cout * cin, 1, cin, splits2)
    grad = torch.empty(cout, cin, 3, 3, device=DEV)
    if batched or splits2:
        other = torch.empty(64, 64, 3, 3, device=DEV)   # a second job in the same launch
        ws2 = torch.randn(3, 9, 64, 64, device=DEV, generator=g)
        tab, nj, blocks, nbytes = ops.make_reduce_jobs(
            [(ws2.view(-1), other, 3, 9 * 64 * 64, 0, 64, 64, 9), (ws.view(-1), grad, splits, 9 * cout * cin, 0, cout, cin, 9, splits2)], DEV)
        ops.wgrad_reduce_batched(tab, nj, blocks, nbytes)
    else:
        ops.wgrad_reduce(ws, splits, 9 * cout * cin, 0, cout, cin, 9, grad)
    ops.device_status()
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), dr, padding=1)
    res = err(grad, ref)
    res["ok"] = res["finite"] and res["rel_l2"] < tol
    if batched or splits2:
        res["other_job"] = err(other, ws2.sum(0).permute(1, 2, 0).reshape(64, 64, 3, 3))
        res["ok"] = res["ok"] and res["other_job"]["rel_l2"] < 1e-6
    return res


def check_wgrad_convt(n=2, h=16, w_=16, c=128, splits=3, seed=8, tol=2e-3) -> dict:
    g = _gen(seed)
    x = bf16r(torch.randn(n, c, h, w_, device=DEV, generator=g))
    dout = bf16r(torch.randn(n, c, 2 * h, 2 * w_, device=DEV, generator=g))
    dcat = torch.zeros(n, 2 * h, 2 * w_, 2 * c, device=DEV, dtype=torch.bfloat16)
    dcat[..., c:] = nhwc(dout).to(torch.bfloat16)
    splits = _splits_for(ops.wgrad_tiles(n, h, w_), splits)
    ws = torch.full((splits, 4, c, c), float("nan"), device=DEV)  # [s][tap][ci][co]
    ops.wgrad_gemm(2, 1, 0, nhwc(x).to(torch.bfloat16), dcat[..., c:], ws, splits, 4 * c * c, c * c, c, 1)
    grad = torch.empty(c, c, 2, 2, device=DEV)
    ops.wgrad_reduce(ws, splits, 4 * c * c, 0, c, c, 4, grad)
    ops.device_status()
    xr = x.clone().requires_grad_(False)
    wt = torch.zeros(c, c, 2, 2, device=DEV, requires_grad=True)
    F.conv_transpose2d(xr, wt, stride=2).backward(dout)
    res = err(grad, wt.grad)
    res["ok"] = res["finite"] and res["rel_l2"] < tol
    return res


def check_wgrad_first(n=2, H=32, W=32, cin=6, splits=4, seed=9, tol=2e-3) -> dict:
    """First-layer conv: forward as a 1-tap GEMM over im2col rows, weight gradient with mode 1."""
    g = _gen(seed)
    x1 = torch.rand(n, cin, H, W, device=DEV, generator=g)
    x2 = torch.rand(n, cin, H, W, device=DEV, generator=g)
    w = bf16r(torch.randn(64, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(64, device=DEV, generator=g)
    cols = ops.pack_input(x1, x2, 0, cin, 0, 64)  # [2n, H, W, 64]
    Bw = ops.pack_weights(2, w, kpad=64)
    out = torch.empty(2 * n, H, W, 64, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm(1, 0, cols, Bw, out, bias=b)
    xcat = bf16r(torch.cat([x1, x2], 0))
    ref = F.conv2d(xcat, w, b, padding=1)
    res = {"fwd": err(nchw(out.float()), ref)}
    dr = bf16r(torch.randn(2 * n, 64, H, W, device=DEV, generator=g))
    splits = _splits_for(ops.wgrad_tiles(2 * n, H, W), splits)
    ws = torch.full((splits, 64, 64), float("nan"), device=DEV)  # [s][co][kpad]
    ops.wgrad_gemm(1, 1, 0, nhwc(dr).to(torch.bfloat16), cols, ws, splits, 64 * 64, 0, 64, 1)
    grad = torch.empty(64, cin, 3, 3, device=DEV)
    ops.wgrad_reduce(ws, splits, 64 * 64, 1, 64, cin, 9, grad)
    ops.device_status()
    refg = torch.nn.grad.conv2d_weight(xcat, (64, cin, 3, 3), dr, padding=1)
    res["wgrad"] = err(grad, refg)
    res["ok"] = res["fwd"]["finite"] and res["fwd"]["rel_l2"] < 6e-3 and res["wgrad"]["finite"] and \
        res["wgrad"]["rel_l2"] < tol
    return res


# ----------------------------------------------------------------------------------------------------
def _bn_setup(n, H, W, Cc, G, seed):
    g = _gen(seed)
    r = (torch.randn(n, H, W, Cc, device=DEV, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(Cc, device=DEV, generator=g) + 0.5
    beta = torch.randn(Cc, device=DEV, generator=g) * 0.2
    return g, r, gamma, beta


def _bn_forward_cuda(r, gamma, beta, G, train=True, rm=None, rv=None, order_rev=False):
    """Statistics from a pass-through conv_gemm is overkill here: emulate the epilogue partials with torch
    (per 128-pixel tile sums), then run the CUDA stats/finalize kernels."""
    n, H, W, Cc = r.shape
    tiles = ops.conv_gemm_tiles(H, W)
    tw, th = (16, 8) if W >= 16 else (8, 16)
    r32 = r.float()
    # [n, ty, th, tx, tw, C] -> per tile sums
    rt = r32.view(n, H // th, th, W // tw, tw, Cc)
    part = torch.stack([rt.sum((2, 4)), (rt * rt).sum((2, 4))], -1).reshape(n * tiles, Cc, 2).contiguous()
    mean = torch.empty(G, Cc, device=DEV)
    invstd = torch.empty_like(mean)
    scale = torch.empty_like(mean)
    shift = torch.empty_like(mean)
    spl = 4
    ws = torch.empty(spl, G, Cc, 2, device=DEV, dtype=torch.float64)
    rm = torch.zeros(Cc, device=DEV) if rm is None else rm
    rv = torch.ones(Cc, device=DEV) if rv is None else rv
    nbt = torch.zeros((), device=DEV, dtype=torch.int64)
    ops.bn_stats(part, Cc, Cc, (n // G) * tiles, G, (n // G) * H * W, spl, ws, gamma, beta, rm, rv, nbt, 0.1, 1e-5,
                 train, order_rev, mean, invstd, scale, shift)
    return mean, invstd, scale, shift, rm, rv, nbt


def check_bn_apply(n=4, H=16, W=32, Cc=64, seed=10) -> dict:
    G = 2
    g, r, gamma, beta = _bn_setup(n, H, W, Cc, G, seed)
    mean, invstd, scale, shift, rm, rv, nbt = _bn_forward_cuda(r, gamma, beta, G)
    a = torch.empty(n, H, W, Cc, device=DEV, dtype=torch.bfloat16)
    cat = torch.zeros(n // 2, H, W, 2 * Cc, device=DEV, dtype=torch.bfloat16)
    pool = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.bfloat16)
    pidx = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.uint8)
    ops.bn_apply(r, scale, shift, G, True, a=a, pool=pool, dif=cat[..., :Cc], pool_idx=pidx)
    torch.cuda.synchronize()
    res = {}
    # arg-max index: first maximum of the stored window in row-major order
    win = a.float().view(n, H // 2, 2, W // 2, 2, Cc).permute(0, 1, 3, 5, 2, 4).reshape(n, H // 2, W // 2, Cc, 4)
    res["pool_idx_ok"] = bool(torch.equal(pidx.long(), win.argmax(-1)) or
                              torch.equal(win.gather(-1, pidx.long().unsqueeze(-1)).squeeze(-1), win.max(-1).values))
    bn = torch.nn.BatchNorm2d(Cc).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    x = nchw(r.float())
    h = n // 2
    y1 = F.relu(bn(x[:h]))  # two calls, separate statistics, two running-stat updates (SURVEY §0 finding 1)
    y2 = F.relu(bn(x[h:]))
    ref_a = torch.cat([y1, y2], 0)
    res["a"] = err(nchw(a.float()), ref_a)
    res["pool"] = err(nchw(pool.float()), F.max_pool2d(ref_a, 2))
    res["diff"] = err(nchw(cat[..., :Cc].float()), y2 - y1)
    res["running_mean"] = err(rm, bn.running_mean)
    res["running_var"] = err(rv, bn.running_var)
    res["nbt"] = int(nbt.item())
    res["ok"] = (res["a"]["rel_l2"] < 4e-3 and res["pool"]["rel_l2"] < 4e-3 and res["diff"]["rel_l2"] < 6e-3 and
                 res["running_mean"]["max_abs"] < 1e-5 and res["running_var"]["max_abs"] < 1e-4 and res["nbt"] == 2 and
                 res["pool_idx_ok"])
    return res


def check_bn_bwd(n=4, H=16, W=32, Cc=64, seed=11, order=("skip", "pool", "dir"), G=2) -> dict:
    """BN+ReLU backward with up to four kinds of gradient source, in any order: `skip` (concat-buffer gradient of the
    t2 - t1 difference: +/- by timestamp), `pool` (max-pool routing through the stored arg-max), `dir` (direct),
    `head` (dz * w of a 1x1 head). The source order selects the kernel variant (generic / per-kind / 2x2-window)."""
    g, r, gamma, beta = _bn_setup(n, H, W, Cc, G, seed)
    mean, invstd, scale, shift, *_ = _bn_forward_cuda(r, gamma, beta, G)
    h = n // 2
    d_skip = bf16r(torch.randn(h, H, W, 2 * Cc, device=DEV, generator=g))  # concat-buffer gradient, first C = skip
    d_pool = bf16r(torch.randn(n, H // 2, W // 2, Cc, device=DEV, generator=g))
    d_dir = bf16r(torch.randn(n, H, W, Cc, device=DEV, generator=g))
    dz = torch.randn(n, 1, H, W, device=DEV, generator=g)
    w_head = torch.randn(Cc, device=DEV, generator=g)
    d_skip_b = d_skip.to(torch.bfloat16)
    pidx = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.uint8)
    ops.bn_apply(r, scale, shift, G, False, pool=torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.bfloat16),
                 pool_idx=pidx)
    table = {
        "skip": {"kind": 1, "t": d_skip_b[..., :Cc], "n_mod": h, "scale_lo": -1.0, "scale_hi": 1.0},
        "pool": {"kind": 2, "t": d_pool.to(torch.bfloat16), "w": pidx},
        "dir": {"kind": 1, "t": d_dir.to(torch.bfloat16)},
        "head": {"kind": 3, "t": dz, "w": w_head},
    }
    srcs = ops.make_srcs([table[k] for k in order])
    ws = torch.empty(ops.bn_bwd_ws_floats(n, H, W, Cc, G), device=DEV)
    dgamma = torch.empty(Cc, device=DEV)
    dbeta = torch.empty(Cc, device=DEV)
    dr = torch.empty(n, H, W, Cc, device=DEV, dtype=torch.bfloat16)
    ops.bn_bwd(r, mean, invstd, scale, shift, srcs, G, ws, dgamma, dbeta, dr)
    torch.cuda.synchronize()
    # torch reference: same graph in fp32
    x = nchw(r.float()).requires_grad_(True)
    ga = gamma.clone().requires_grad_(True)
    be = beta.clone().requires_grad_(True)
    ys = []
    per = n // G
    for gi in range(G):
        xs = x[gi * per:(gi + 1) * per]
        ys.append(F.relu(F.batch_norm(xs, None, None, ga, be, True, 0.1, 1e-5)))
    aa = torch.cat(ys, 0)
    # forward stores a in bf16 and pools the rounded values: route through the same rounding for the arg-max
    aq = aa + (bf16r(aa.detach()) - aa.detach())
    loss = 0.0
    if "skip" in order:
        loss = loss + ((aa[h:] - aa[:h]) * nchw(d_skip[..., :Cc])).sum()
    if "pool" in order:
        loss = loss + (F.max_pool2d(aq, 2) * nchw(d_pool)).sum()
    if "dir" in order:
        loss = loss + (aa * nchw(d_dir)).sum()
    if "head" in order:
        loss = loss + ((aa * w_head.view(1, Cc, 1, 1)).sum(1, keepdim=True) * dz).sum()
    loss.backward()
    res = {"dr": err(nchw(dr.float()), x.grad), "dgamma": err(dgamma, ga.grad), "dbeta": err(dbeta, be.grad)}
    res["ok"] = res["dr"]["rel_l2"] < 8e-3 and res["dgamma"]["rel_l2"] < 2e-3 and res["dbeta"]["rel_l2"] < 2e-3
    return res


def check_bn_apply_pool(n=4, H=16, W=32, Cc=64, G=2, seed=14) -> dict:
    """BN-apply + ReLU + MaxPool2d without the t2 - t1 difference (plain encoders), one or two stat-groups, with a second
    copy into a concat slice: the 2x2-window kernel."""
    g, r, gamma, beta = _bn_setup(n, H, W, Cc, G, seed)
    mean, invstd, scale, shift, *_ = _bn_forward_cuda(r, gamma, beta, G)
    a = torch.empty(n, H, W, Cc, device=DEV, dtype=torch.bfloat16)
    cat = torch.zeros(n, H, W, 2 * Cc, device=DEV, dtype=torch.bfloat16)
    pool = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.bfloat16)
    pidx = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.uint8)
    ops.bn_apply(r, scale, shift, G, False, a=a, a2=cat[..., :Cc], pool=pool, pool_idx=pidx)
    torch.cuda.synchronize()
    per = n // G
    x = nchw(r.float())
    ys = [F.relu(F.batch_norm(x[gi * per:(gi + 1) * per], None, None, gamma, beta, True, 0.1, 1e-5)) for gi in range(G)]
    ref_a = torch.cat(ys, 0)
    res = {"a": err(nchw(a.float()), ref_a), "pool": err(nchw(pool.float()), F.max_pool2d(ref_a, 2))}
    res["a2_same"] = bool(torch.equal(cat[..., :Cc], a) and (cat[..., Cc:] == 0).all().item())
    win = a.float().view(n, H // 2, 2, W // 2, 2, Cc).permute(0, 1, 3, 5, 2, 4).reshape(n, H // 2, W // 2, Cc, 4)
    res["pool_idx_ok"] = bool(torch.equal(pidx.long(), win.argmax(-1)) or
                              torch.equal(win.gather(-1, pidx.long().unsqueeze(-1)).squeeze(-1), win.max(-1).values))
    res["pool_is_max_of_stored"] = bool(torch.equal(pool.float(), win.max(-1).values))
    res["ok"] = (res["a"]["rel_l2"] < 4e-3 and res["pool"]["rel_l2"] < 4e-3 and res["a2_same"] and res["pool_idx_ok"] and
                 res["pool_is_max_of_stored"])
    return res


def check_head(n=2, H=16, W=16, seed=12) -> dict:
    g = _gen(seed)
    a0 = bf16r(torch.randn(n, H, W, 64, device=DEV, generator=g))
    a1 = bf16r(torch.randn(n, H, W, 64, device=DEV, generator=g))
    w = torch.randn(128, device=DEV, generator=g)
    b = torch.randn(1, device=DEV, generator=g)
    logits = torch.empty(n, 1, H, W, device=DEV)
    ops.head_fwd(a0.to(torch.bfloat16), a1.to(torch.bfloat16), w, b, logits)
    ref = (torch.cat([a0, a1], -1) * w).sum(-1) + b
    res = {"fusion": err(logits.view(n, H, W), ref)}
    ops.head_fwd(a0.to(torch.bfloat16), None, w[:64].contiguous(), b, logits)
    res["single"] = err(logits.view(n, H, W), (a0 * w[:64]).sum(-1) + b)
    # weight / bias gradient through colsum
    dz = torch.randn(n * H * W, device=DEV, generator=g)
    nblk = 8
    ws = torch.empty(nblk * 64, device=DEV)
    dw = torch.empty(64, device=DEV)
    ops.colsum(a0.to(torch.bfloat16), dz, n * H * W, nblk, ws, dw)
    res["dw"] = err(dw, (a0.view(-1, 64) * dz[:, None]).sum(0))
    db = torch.empty(1, device=DEV)
    ops.colsum(None, dz, n * H * W, nblk, ws, db)
    res["db"] = err(db, dz.sum().view(1))
    cs = torch.empty(64, device=DEV)
    ops.colsum(a1.to(torch.bfloat16), None, n * H * W, nblk, ws, cs)
    res["colsum"] = err(cs, a1.view(-1, 64).sum(0))
    torch.cuda.synchronize()
    res["ok"] = all(res[k]["rel_l2"] < 1e-5 for k in ("fusion", "single", "dw", "colsum")) and res["db"]["max_abs"] < 1e-3
    return res


def power_jaccard_ref(z, t):
    p = torch.sigmoid(z)
    i = (p.flatten() * t.flatten()).sum()
    d = (p.flatten() ** 2 + t.flatten() ** 2).sum() - i + 1e-6
    return 1 - i / d


def check_pj(B=6, H=32, W=32, seed=13) -> dict:
    g = _gen(seed)
    z = torch.randn(B, 1, H, W, device=DEV, generator=g) * 2
    t = (torch.rand(B, 1, H, W, device=DEV, generator=g) > 0.9).float()
    z2 = torch.randn(B, 1, H, W, device=DEV, generator=g)
    mask = torch.tensor([1, 1, 0, 1, 1, 0], device=DEV, dtype=torch.uint8)
    nblk = 16
    ws = torch.empty(nblk * 3, device=DEV, dtype=torch.float64)
    sums = torch.empty(3, device=DEV, dtype=torch.float64)
    loss = torch.empty((), device=DEV)
    res = {}
    # (1) plain, all rows
    ops.pj_fwd(z, t, False, None, 0, nblk, ws, sums)
    ops.pj_loss(sums, loss)
    dz = torch.empty_like(z)
    gout = torch.tensor(0.5, device=DEV)
    ops.pj_bwd(z, t, False, None, 0, sums, gout, 2.0, False, dz, None)
    zr = z.clone().requires_grad_(True)
    lr = power_jaccard_ref(zr, t)
    lr.backward()
    res["loss"] = abs(loss.item() - lr.item())
    res["dz"] = err(dz, zr.grad)  # g = 0.5 * 2.0 = 1
    # (2) labeled rows only
    ops.pj_fwd(z, t, False, mask, 1, nblk, ws, sums)
    ops.pj_loss(sums, loss)
    ops.pj_bwd(z, t, False, mask, 1, sums, None, 1.0, False, dz, None)
    zr = z.clone().requires_grad_(True)
    mb = mask.bool()
    lr = power_jaccard_ref(zr[mb], t[mb])
    lr.backward()
    res["loss_masked"] = abs(loss.item() - lr.item())
    res["dz_masked"] = err(dz, zr.grad)
    # (3) consistency term on the unlabeled rows: target = sigmoid(z2), gradient to both
    ops.pj_fwd(z, z2, True, mask, 0, nblk, ws, sums)
    ops.pj_loss(sums, loss)
    dz2 = torch.empty_like(z2)
    ops.pj_bwd(z, z2, True, mask, 0, sums, None, 1.0, False, dz, dz2)
    zr = z.clone().requires_grad_(True)
    z2r = z2.clone().requires_grad_(True)
    lr = power_jaccard_ref(zr[~mb], torch.sigmoid(z2r)[~mb])
    lr.backward()
    res["loss_cons"] = abs(loss.item() - lr.item())
    res["dz_cons"] = err(dz, zr.grad)
    res["dz2_cons"] = err(dz2, z2r.grad)
    # known answers: pj(0, 1) = 1/3 ; pj(z, 0) = 1 with zero gradient
    z0 = torch.zeros(2, 1, 8, 8, device=DEV)
    ops.pj_fwd(z0, torch.ones_like(z0), False, None, 0, 4, ws, sums)
    ops.pj_loss(sums, loss)
    res["ka_third"] = abs(loss.item() - 1.0 / 3.0)
    ops.pj_fwd(z[:2], torch.zeros_like(z[:2]), False, None, 0, 4, ws, sums)
    ops.pj_loss(sums, loss)
    dzz = torch.empty_like(z[:2])
    ops.pj_bwd(z[:2].contiguous(), torch.zeros_like(z[:2]), False, None, 0, sums, None, 1.0, False, dzz, None)
    res["ka_one"] = abs(loss.item() - 1.0)
    res["ka_zero_grad"] = dzz.abs().max().item()
    torch.cuda.synchronize()
    res["ok"] = (res["loss"] < 1e-6 and res["loss_masked"] < 1e-6 and res["loss_cons"] < 1e-6 and
                 res["dz"]["rel_l2"] < 1e-4 and res["dz_masked"]["rel_l2"] < 1e-4 and res["dz_cons"]["rel_l2"] < 1e-4 and
                 res["dz2_cons"]["rel_l2"] < 1e-4 and res["ka_third"] < 1e-6 and res["ka_one"] < 1e-7 and
                 res["ka_zero_grad"] == 0.0)
    return res


# ----------------------------------------------------------------------------------------------------
# Layout decoders: structured operands that reveal which (row, k) element each output used. Run by the
# probe when a GEMM check fails (blind debugging aid; not part of the pytest suite).
def decode_fprop() -> dict:
    n, H, W = 1, 8, 16  # exactly one 128-pixel tile (tw=16, th=8)
    A = torch.zeros(n, H, W, 64, device=DEV)
    pix = torch.arange(H * W, device=DEV)
    A.view(-1, 64)[pix, pix % 64] = 1.0  # pixel p selects k = p % 64
    out = {}
    for name, fn in (("n", lambda nn, kk: nn), ("k", lambda nn, kk: kk)):
        nn, kk = torch.meshgrid(torch.arange(64, device=DEV), torch.arange(64, device=DEV), indexing="ij")
        Bw = fn(nn, kk).float().to(torch.bfloat16).contiguous()  # [N=64][K=64]
        o = torch.empty(n, H, W, 64, device=DEV, dtype=torch.bfloat16)
        ops.conv_gemm(1, 0, A.to(torch.bfloat16), Bw, o)
        ops.device_status()
        out[name] = o.float().view(128, 64)[:, :].cpu()
    # expected: out["n"][p][c] = c ; out["k"][p][c] = p % 64
    exp_n = torch.arange(64).float().expand(128, 64)
    exp_k = (torch.arange(128) % 64).float().view(128, 1).expand(128, 64)
    return {"n_ok": bool(torch.equal(out["n"], exp_n)), "k_ok": bool(torch.equal(out["k"], exp_k)),
            "n_sample": out["n"][:10, :10].tolist(), "k_sample": out["k"][:10, :10].tolist(),
            "k_col0": out["k"][:, 0].tolist()}


def decode_wgrad() -> dict:
    n, H, W = 1, 8, 8  # one 64-pixel tile
    U = torch.zeros(n, H, W, 128, device=DEV)
    pix = torch.arange(64, device=DEV)
    U.view(-1, 128)[pix, pix] = 1.0  # pixel p selects m = p
    res = {}
    for name in ("p", "n"):
        V = torch.zeros(64, 64, device=DEV)
        if name == "p":
            V[:] = torch.arange(64, device=DEV).float().view(64, 1)
        else:
            V[:] = torch.arange(64, device=DEV).float().view(1, 64)
        ws = torch.full((1, 1, 128, 64), float("nan"), device=DEV)
        ops.wgrad_gemm(1, 1, 0, U.to(torch.bfloat16), V.view(1, 8, 8, 64).to(torch.bfloat16), ws, 1, 128 * 64, 0, 64, 1)
        ops.device_status()
        res[name] = ws[0, 0].cpu()
    exp_p = torch.zeros(128, 64)
    exp_p[:64] = torch.arange(64).float().view(64, 1)
    exp_n = torch.zeros(128, 64)
    exp_n[:64] = torch.arange(64).float().view(1, 64)
    return {"p_ok": bool(torch.equal(res["p"], exp_p)), "n_ok": bool(torch.equal(res["n"], exp_n)),
            "p_sample": res["p"][:10, :10].tolist(), "n_sample": res["n"][:10, :10].tolist(),
            "p_col0": res["p"][:, 0].tolist()}


ALL_CHECKS = {
    "stat_rowsum_up_bias": check_stat_rowsum,
    "stat_rowsum_up_bias_256": lambda: check_stat_rowsum(n=3, H=16, W=16, c=128, cout=128, seed=72),
    "stat_rowsum_up_bias_split_bf16": lambda: check_stat_rowsum(prec=True, seed=73),
    "pack_weights": check_pack_weights,
    "pack_input": check_pack_input,
    "pj": check_pj,
    "head": check_head,
    "bn_apply": check_bn_apply,
    "bn_bwd": check_bn_bwd,
    "bn_bwd_dir": lambda: check_bn_bwd(order=("dir",), seed=15),
    "bn_bwd_dir_G1_128": lambda: check_bn_bwd(2, 16, 16, 128, order=("dir",), G=1, seed=16),
    "bn_bwd_skip_dir": lambda: check_bn_bwd(order=("skip", "dir"), seed=17),
    "bn_bwd_pool_dir_window": lambda: check_bn_bwd(order=("pool", "dir"), seed=18),
    "bn_bwd_pool_dir_window_G1_256": lambda: check_bn_bwd(2, 16, 16, 256, order=("pool", "dir"), G=1, seed=19),
    "bn_bwd_pool_skip_window": lambda: check_bn_bwd(order=("pool", "skip"), seed=20),
    "bn_bwd_pool_skip_dir_window": lambda: check_bn_bwd(order=("pool", "skip", "dir"), seed=21),
    "bn_bwd_pool_only_window": lambda: check_bn_bwd(order=("pool",), seed=22),
    "bn_bwd_head": lambda: check_bn_bwd(order=("head",), G=1, seed=23),
    "bn_bwd_dir_head": lambda: check_bn_bwd(order=("dir", "head"), seed=24),
    "bn_apply_pool_G2": check_bn_apply_pool,
    "bn_apply_pool_G1_512": lambda: check_bn_apply_pool(2, 16, 16, 512, G=1, seed=25),
    "conv3x3_64_64": lambda: check_conv3x3(2, 32, 32, 64, 64),
    "conv3x3_128_128_16": lambda: check_conv3x3(2, 16, 16, 128, 128, seed=31),
    "conv3x3_64_128_64": lambda: check_conv3x3(1, 64, 64, 64, 128, bias=False, seed=32),
    "conv3x3_slices": lambda: check_conv3x3(2, 32, 32, 192, 256, slice_in=True, slice_out=True, seed=33),
    "conv3x3_512_512_16": lambda: check_conv3x3(2, 16, 16, 512, 512, seed=34),
    "conv3x3_ragged_8x8": lambda: check_conv3x3(3, 8, 8, 128, 128, seed=35),
    "conv3x3_ragged_4x4": lambda: check_conv3x3(3, 4, 4, 512, 512, seed=36),
    "conv3x3_ragged_24x40": lambda: check_conv3x3(2, 24, 40, 64, 64, seed=37),
    "conv3x3_wide_128_256": lambda: check_conv3x3(2, 32, 32, 128, 256, seed=39, halo=False, wide=True),
    "conv3x3_wide_slices_192_512": lambda: check_conv3x3(2, 16, 16, 192, 512, slice_in=True, slice_out=True, seed=40, halo=False, wide=True),
    "conv3x3_narrow_128_256": lambda: check_conv3x3(2, 32, 32, 128, 256, seed=39, halo=False, wide=False),
    "conv3x3_halo_64_64": lambda: check_conv3x3(2, 32, 32, 64, 64, halo=True),
    "conv3x3_nohalo_64_64": lambda: check_conv3x3(2, 32, 32, 64, 64, halo=False),
    "conv3x3_halo_128_128": lambda: check_conv3x3(2, 16, 16, 128, 128, seed=31, halo=True),
    "conv3x3_halo_slices_192_256": lambda: check_conv3x3(2, 32, 32, 192, 256, slice_in=True, slice_out=True, seed=33, halo=True),
    "conv3x3_halo_ragged_8x8": lambda: check_conv3x3(3, 8, 8, 128, 64, seed=38, halo=True),
    "conv3x3_halo_ragged_4x4": lambda: check_conv3x3(3, 4, 4, 512, 512, seed=36, halo=True),
    "conv3x3_halo_ragged_24x40": lambda: check_conv3x3(2, 24, 40, 64, 64, seed=37, halo=True),
    "conv3x3_pair_64_64": lambda: check_conv3x3(2, 32, 32, 64, 64, pair=True, compare_single=True),
    "conv3x3_pair_64_64_many": lambda: check_conv3x3(16, 64, 64, 64, 64, seed=43, pair=True, compare_single=True),
    "conv3x3_pair_128_64_resident": lambda: check_conv3x3(8, 64, 64, 128, 64, seed=44, pair=True, compare_single=True),
    "conv3x3_pair_64_128_resident": lambda: check_conv3x3(8, 64, 64, 64, 128, seed=45, pair=True, compare_single=True),
    "conv3x3_pair_128_128": lambda: check_conv3x3(4, 32, 32, 128, 128, seed=31, pair=True, compare_single=True),
    "conv3x3_pair_256_256_many": lambda: check_conv3x3(16, 32, 32, 256, 256, seed=46, pair=True),
    "conv3x3_pair_slices_192_512": lambda: check_conv3x3(2, 16, 16, 192, 512, slice_in=True, slice_out=True, seed=40, pair=True),
    "conv3x3_pair_512_512_16": lambda: check_conv3x3(2, 16, 16, 512, 512, seed=34, pair=True),
    "conv3x3_pair_odd_tiles_8x8": lambda: check_conv3x3(3, 8, 8, 128, 64, seed=38, pair=True, compare_single=True),
    "conv3x3_pair_ragged_4x4": lambda: check_conv3x3(3, 4, 4, 512, 512, seed=36, pair=True),
    "conv3x3_pair_ragged_24x40": lambda: check_conv3x3(2, 24, 40, 64, 64, seed=37, pair=True, compare_single=True),
    "conv3x3_pair_1024_256": lambda: check_conv3x3(4, 32, 32, 1024, 256, seed=47, pair=True),
    "conv3x3_cta_stats_G2": check_conv3x3_cta_stats,
    "conv3x3_cta_stats_G1_many": lambda: check_conv3x3_cta_stats(16, 64, 64, 64, 64, G=1, seed=49),
    "conv3x3_cta_stats_G2_512": lambda: check_conv3x3_cta_stats(4, 16, 16, 256, 512, G=2, seed=50),
    "conv3x3_cta_stats_odd_tiles": lambda: check_conv3x3_cta_stats(6, 8, 8, 128, 64, G=2, seed=51),
    "forced_tiles_64_256": check_forced_tiles,
    "forced_tiles_256_512_deep": lambda: check_forced_tiles(4, 16, 16, 256, 512, G=1, seed=121),
    "forced_tiles_128_128_ragged": lambda: check_forced_tiles(2, 24, 40, 128, 128, G=2, seed=122),
    "dgrad_bnbwd_G2": check_dgrad_bnbwd,
    "dgrad_bnbwd_G1_many": lambda: check_dgrad_bnbwd(16, 64, 64, 64, 64, G=1, seed=67),
    "dgrad_bnbwd_256_512": lambda: check_dgrad_bnbwd(4, 16, 16, 256, 512, G=1, seed=68),
    "dgrad_bnbwd_odd_tiles_8x8": lambda: check_dgrad_bnbwd(6, 8, 8, 128, 64, G=2, seed=69),
    "dgrad_bnbwd_ragged_24x40": lambda: check_dgrad_bnbwd(2, 24, 40, 64, 64, G=1, seed=70),
    "convt_dgrad_bnbwd": lambda: check_dgrad_bnbwd(4, 16, 16, 128, 128, G=1, seed=71, mode=2),
    "convt_dgrad_bnbwd_64_many": lambda: check_dgrad_bnbwd(8, 64, 64, 64, 64, G=1, seed=72, mode=2),
    "conv3x3_dgrad": check_conv3x3_dgrad,
    "conv3x3_dgrad_ragged_8x8": lambda: check_conv3x3_dgrad(3, 8, 8, 128, 128, seed=42),
    "conv3x3_dgrad_512_256": lambda: check_conv3x3_dgrad(2, 32, 32, 512, 256, seed=41),
    "convt_fwd": check_convt_fwd,
    "convt_fwd_64": lambda: check_convt_fwd(2, 16, 32, 64, seed=51),
    "convt_fwd_tiny_8x8": lambda: check_convt_fwd(3, 8, 8, 128, seed=52),
    "convt_fwd_tiny_4x4": lambda: check_convt_fwd(3, 4, 4, 512, seed=53),
    "convt_fwd_odd_6x12": lambda: check_convt_fwd(2, 6, 12, 64, seed=54),
    "convt_fwd_single_cta": lambda: check_convt_fwd(pair=False),
    "convt_fwd_single_cta_odd_6x12": lambda: check_convt_fwd(2, 6, 12, 64, seed=54, pair=False),
    "convt_fwd_many_256": lambda: check_convt_fwd(8, 32, 32, 256, seed=56),
    "convt_fwd_many_64": lambda: check_convt_fwd(16, 64, 64, 64, seed=57),
    "first_conv_siamese_pair": check_first_conv,
    "first_conv_early_fusion_8ch_pair": lambda: check_first_conv(4, 32, 48, 4, cat_mode=1, seed=58, G=1),
    "first_conv_single_cta": lambda: check_first_conv(pair=False),
    "first_conv_many": lambda: check_first_conv(16, 64, 64, 2, cat_mode=1, seed=59, G=1),
    "convt_dgrad": check_convt_dgrad,
    "convt_dgrad_single_cta": lambda: check_convt_dgrad(pair=False),
    "convt_dgrad_single_cta_tiny_4x4": lambda: check_convt_dgrad(3, 4, 4, 512, seed=63, pair=False),
    "convt_dgrad_many_256": lambda: check_convt_dgrad(8, 32, 32, 256, seed=64),
    "convt_dgrad_many_64": lambda: check_convt_dgrad(16, 64, 64, 64, seed=65),
    "convt_dgrad_tiny_8x8": lambda: check_convt_dgrad(3, 8, 8, 128, seed=62),
    "convt_dgrad_tiny_4x4": lambda: check_convt_dgrad(3, 4, 4, 512, seed=63),
    "convt_dgrad_64": lambda: check_convt_dgrad(2, 16, 32, 64, seed=61),
    "wgrad3x3_pos": lambda: check_wgrad3x3(sign=1, halo=0),
    "wgrad3x3_pos_halo": lambda: check_wgrad3x3(sign=1, halo=1),
    "wgrad3x3_neg": lambda: check_wgrad3x3(2, 32, 32, 128, 64, sign=-1, halo=0, seed=71),
    "wgrad3x3_neg_halo": lambda: check_wgrad3x3(2, 32, 32, 128, 64, sign=-1, halo=1, seed=71),
    "wgrad3x3_64_64_halo": lambda: check_wgrad3x3(2, 32, 32, 64, 64, sign=1, halo=1, seed=72),
    "wgrad3x3_64_64_many_splits": lambda: check_wgrad3x3(2, 32, 32, 64, 64, sign=1, halo=1, splits=27, seed=76),
    "wgrad3x3_64_64_mstack_neg": lambda: check_wgrad3x3(2, 32, 32, 64, 64, sign=-1, halo=1, splits=9, seed=83),
    "wgrad3x3_64_64_mstack_ragged_24x40": lambda: check_wgrad3x3(2, 24, 40, 64, 64, sign=1, halo=1, splits=7, seed=84),
    "wgrad3x3_64_64_mstack_1px_edge": lambda: check_wgrad3x3(1, 9, 9, 64, 64, sign=1, halo=1, splits=4, seed=85),
    "wgrad3x3_64_64_two_kx": lambda: check_wgrad3x3(2, 32, 32, 64, 64, sign=1, halo=1, splits=10, splits2=5, seed=77),
    "wgrad3x3_64_128_two_kx": lambda: check_wgrad3x3(2, 32, 32, 64, 128, sign=1, halo=1, splits=24, splits2=12, seed=78),
    "wgrad3x3_256_64_two_kx_neg": lambda: check_wgrad3x3(2, 32, 32, 256, 64, sign=-1, halo=1, splits=6, splits2=3, seed=79),
    "wgrad3x3_two_kx_ragged_24x40": lambda: check_wgrad3x3(2, 24, 40, 64, 64, sign=1, halo=1, splits=7, splits2=2, seed=80),
    "wgrad3x3_batched_reduce_many": lambda: check_wgrad3x3(2, 32, 32, 128, 128, sign=1, halo=1, splits=27, seed=81, batched=True),
    "wgrad3x3_batched_reduce_few_512": lambda: check_wgrad3x3(4, 16, 16, 512, 512, sign=1, halo=1, splits=3, seed=82, batched=True),
    "wgrad3x3_512_512_halo": lambda: check_wgrad3x3(4, 16, 16, 512, 512, sign=1, halo=1, splits=8, seed=73),
    "wgrad3x3_tiny_4x4": lambda: check_wgrad3x3(3, 4, 4, 512, 512, sign=1, halo=1, splits=2, seed=74),
    "wgrad3x3_ragged_24x40": lambda: check_wgrad3x3(2, 24, 40, 64, 128, sign=1, halo=1, splits=7, seed=75),
    "pad_copy_bottom_right": check_pad_copy,
    "pad_copy_centre_2px": lambda: check_pad_copy(3, 6, 6, 128, 8, 9, top=1, left=1, seed=91),
    "wgrad_convt": check_wgrad_convt,
    "wgrad_convt_tiny_4x4": lambda: check_wgrad_convt(3, 4, 4, 512, splits=2, seed=82),
    "wgrad_convt_64": lambda: check_wgrad_convt(2, 16, 32, 64, seed=81),
    "wgrad_first": check_wgrad_first,
}
