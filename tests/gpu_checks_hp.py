"""Per-kernel parity checks of the split-bf16 ("precise") kernels (include/b200cd.h, ABI version 2) against torch in
fp64 on the values the split tensors actually hold (hi + lo). Imported by tests/test_gpu_ops.py (pytest -m gpu).

Expected error of a split-bf16 product chain: each operand carries 16 mantissa bits (2^-17 relative) and the lo*lo
term (2^-18) is dropped, outputs are stored with 16 bits again -> rel L2 ~1e-5 on tensor-core outputs that are stored
split, ~3e-6 on fp32 outputs; the bounds below are those times ~3."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from multimodal_siamese_cd_b200 import ops
from gpu_checks import DEV, _gen, err, nchw, nhwc, _splits_for, _bn_forward_cuda

TOL_SPLIT_OUT = 3e-5     # tensor-core result stored as hi + lo
TOL_F32_OUT = 1e-5       # tensor-core result stored as fp32 (weight gradients)
TOL_ELEM = 2e-5          # element-wise kernels: fp32 arithmetic on 16-bit inputs, 16-bit outputs


def q16(x: torch.Tensor) -> torch.Tensor:
    """What a split-bf16 store of x reads back as."""
    hi = x.to(torch.bfloat16).float()
    return hi + (x - hi).to(torch.bfloat16).float()


def split(x_nhwc: torch.Tensor) -> torch.Tensor:
    return ops.split_from_float(x_nhwc.contiguous())


def pack_hp(mode: int, w: torch.Tensor, kpad: int = 0, mode2: int = None):
    shape = ops.packed_weight_shape(mode, w.shape[0], w.shape[1], kpad, prec=True)
    out = torch.empty(shape, device=w.device, dtype=torch.bfloat16)
    spec = (mode, w.contiguous(), out, kpad)
    out2 = None
    if mode2 is not None:
        out2 = torch.empty(ops.packed_weight_shape(mode2, w.shape[0], w.shape[1], kpad, prec=True), device=w.device,
                           dtype=torch.bfloat16)
        spec = spec + (mode2, out2)
    tab, nj, blocks, elems = ops.make_pack_jobs([spec], w.device, prec=True)
    ops.pack_weights_batched(tab, nj, blocks, elems, w.numel(), prec=True)
    return out if mode2 is None else (out, out2)


def _expand_ref(m: torch.Tensor, taps: int) -> torch.Tensor:
    """fp32 [rows][taps*K] -> the K-tripled split layout [rows][taps][hi | lo | hi] in bf16."""
    rows = m.shape[0]
    K = m.shape[1] // taps
    m3 = m.view(rows, taps, K)
    hi = m3.to(torch.bfloat16)
    lo = (m3 - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo, hi], 2).reshape(rows, taps * 3 * K)


def check_hp_pack() -> dict:
    g = _gen(101)
    w = torch.randn(128, 64, 3, 3, device=DEV, generator=g)
    wt = torch.randn(128, 64, 2, 2, device=DEV, generator=g)
    w1 = torch.randn(64, 6, 3, 3, device=DEV, generator=g)
    out = {}
    p0, p1 = pack_hp(0, w, mode2=1)
    out["m0"] = bool(torch.equal(p0, _expand_ref(w.permute(0, 2, 3, 1).reshape(128, 9 * 64), 9)))
    out["m1"] = bool(torch.equal(p1, _expand_ref(w.flip(2, 3).permute(1, 2, 3, 0).reshape(64, 9 * 128), 9)))
    ref2 = torch.zeros(64, 64, device=DEV)
    ref2[:, :54] = w1.permute(0, 2, 3, 1).reshape(64, 54)
    out["m2"] = bool(torch.equal(pack_hp(2, w1, kpad=64), _expand_ref(ref2, 1)))
    p3, p4 = pack_hp(3, wt, mode2=4)
    out["m3"] = bool(torch.equal(p3, _expand_ref(wt.permute(2, 3, 1, 0).reshape(4 * 64, 128), 1)))
    out["m4"] = bool(torch.equal(p4, _expand_ref(wt.permute(0, 2, 3, 1).reshape(128, 4 * 64), 4)))
    # input packing: split im2col rows
    x1 = torch.rand(2, 6, 16, 16, device=DEV, generator=g)
    x2 = torch.rand(2, 6, 16, 16, device=DEV, generator=g)
    for cat_mode, nc, c_lo in ((0, 4, 2), (1, 2, 0)):
        cin = 2 * nc if cat_mode else nc
        cols = ops.pack_input(x1, x2, c_lo, nc, cat_mode, 64, prec=True)
        xin = torch.cat([x1[:, c_lo:c_lo + nc], x2[:, c_lo:c_lo + nc]], 1 if cat_mode else 0)
        un = F.unfold(xin, 3, padding=1).view(xin.shape[0], cin, 9, 16, 16).permute(0, 3, 4, 2, 1).reshape(xin.shape[0], 16, 16, 9 * cin)
        ref = torch.zeros(xin.shape[0], 16, 16, 64, device=DEV)
        ref[..., :9 * cin] = un
        hi = ref.to(torch.bfloat16)
        out[f"cols{cat_mode}"] = bool(torch.equal(cols, hi) and
                                      torch.equal(ops.lo_half(cols), (ref - hi.float()).to(torch.bfloat16)))
    ops.device_status()
    out["ok"] = all(out.values())
    return out


def check_hp_conv3x3(n=2, H=32, W=32, cin=64, cout=64, G=0, slice_in=False, slice_out=False, seed=103) -> dict:
    """3x3 conv on split tensors: out = conv(x, w) + b within the split-bf16 product error of the fp64 result, the
    stored statistics equal the sums of the stored (hi + lo) output, slices of concat buffers are respected."""
    g = _gen(seed)
    x = q16(torch.randn(n, cin, H, W, device=DEV, generator=g))
    w = q16(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(cout, device=DEV, generator=g)
    if slice_in:    # A = channels [64, 64 + cin) of a wider split buffer
        big = ops.split_alloc((n, H, W, cin + 64), DEV, zero=True)
        A = big[..., 64:]
        A.copy_(split(nhwc(x)))
        ops.lo_half(A).copy_(ops.lo_half(split(nhwc(x))))
    else:
        A = split(nhwc(x))
    if slice_out:
        obig = ops.split_alloc((n, H, W, cout + 64), DEV)
        obig.fill_(7.0)
        ops.lo_half(obig).fill_(7.0)
        out = obig[..., :cout]
    else:
        obig = None
        out = ops.split_alloc((n, H, W, cout), DEV)
    Bw = pack_hp(0, w)
    if G:
        rows, per_cta = ops.conv_stat_rows(n, H, W, cin, cout, G, prec=True)
        stats = torch.full((G, rows, cout, 2), float("nan"), device=DEV)
        ops.conv_gemm(0, 0, A, Bw, out, bias=b, stats=stats, stat_groups=G, prec=True)
    else:
        stats = torch.zeros(n * ops.conv_gemm_tiles(H, W), cout, 2, device=DEV)
        ops.conv_gemm(0, 0, A, Bw, out, bias=b, stats=stats, prec=True)
    ops.device_status()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    got = ops.split_to_float(out)
    res = err(nchw(got), ref)
    o = got.double()
    if G:
        og = o.view(G, n // G, H, W, cout)
        rs, rq = og.sum((1, 2, 3)), (og * og).sum((1, 2, 3))
        gs, gq = stats[..., 0].double().sum(1), stats[..., 1].double().sum(1)
    else:
        rs, rq = o.sum((0, 1, 2)), (o * o).sum((0, 1, 2))
        gs, gq = stats[..., 0].double().sum(0), stats[..., 1].double().sum(0)
    res["stats_sum_abs"] = ((gs - rs).abs().max() / rq.sqrt().max()).item()
    res["stats_sq_rel"] = ((gq - rq).norm() / rq.norm()).item()
    if obig is not None:
        res["untouched"] = bool((obig[..., cout:] == 7.0).all().item() and (ops.lo_half(obig)[..., cout:] == 7.0).all().item())
    res["ok"] = res["finite"] and res["rel_l2"] < TOL_SPLIT_OUT and res["stats_sum_abs"] < 1e-5 and \
        res["stats_sq_rel"] < 1e-5 and res.get("untouched", True)
    return res


def check_hp_dgrad(n=2, H=32, W=32, cin=64, cout=128, seed=104) -> dict:
    g = _gen(seed)
    w = q16(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) / (3.0 * cout ** 0.5))
    dr = q16(torch.randn(n, cout, H, W, device=DEV, generator=g))
    out = ops.split_alloc((n, H, W, cin), DEV)
    ops.conv_gemm(0, 0, split(nhwc(dr)), pack_hp(1, w), out, prec=True)
    ops.device_status()
    ref = F.conv_transpose2d(dr.double(), w.double(), padding=1)
    res = err(nchw(ops.split_to_float(out)), ref)
    res["ok"] = res["finite"] and res["rel_l2"] < TOL_SPLIT_OUT
    return res


def check_hp_convt(n=2, h=16, w_=16, c=128, seed=105) -> dict:
    """ConvTranspose2d(c, c, 2, stride 2) on split tensors: forward scattered into the upper half of a split concat
    buffer, its input gradient gathered from it, and its weight gradient."""
    g = _gen(seed)
    x = q16(torch.randn(n, c, h, w_, device=DEV, generator=g))
    wt = q16(torch.randn(c, c, 2, 2, device=DEV, generator=g) / (c ** 0.5))
    b = torch.randn(c, device=DEV, generator=g)
    cat = ops.split_alloc((n, 2 * h, 2 * w_, 2 * c), DEV)
    cat.fill_(3.0)
    ops.lo_half(cat).fill_(3.0)
    Bf, Bd = pack_hp(3, wt, mode2=4)
    xs = split(nhwc(x))
    ops.conv_gemm(1, 1, xs, Bf, cat[..., c:], bias=b, prec=True)
    ops.device_status()
    ref = F.conv_transpose2d(x.double(), wt.double(), b.double(), stride=2)
    res = {"fwd": err(nchw(ops.split_to_float(cat[..., c:])), ref)}
    res["untouched"] = bool((cat[..., :c] == 3.0).all().item() and (ops.lo_half(cat)[..., :c] == 3.0).all().item())
    dout = q16(torch.randn(n, c, 2 * h, 2 * w_, device=DEV, generator=g))
    dcat = ops.split_alloc((n, 2 * h, 2 * w_, 2 * c), DEV, zero=True)
    ds = split(nhwc(dout))
    dcat[..., c:].copy_(ds)
    ops.lo_half(dcat[..., c:]).copy_(ops.lo_half(ds))
    dx = ops.split_alloc((n, h, w_, c), DEV)
    ops.conv_gemm(2, 0, dcat[..., c:], Bd, dx, prec=True)
    res["dgrad"] = err(nchw(ops.split_to_float(dx)), F.conv2d(dout.double(), wt.double(), stride=2))
    splits = _splits_for(ops.wgrad_tiles(n, h, w_), 3)
    ws = torch.full((splits, 4, c, c), float("nan"), device=DEV)
    ops.wgrad_gemm(2, 1, 0, xs, dcat[..., c:], ws, splits, 4 * c * c, c * c, c, 1, prec=True)
    grad = torch.empty(c, c, 2, 2, device=DEV)
    ops.wgrad_reduce(ws, splits, 4 * c * c, 0, c, c, 4, grad)
    ops.device_status()
    wz = torch.zeros(c, c, 2, 2, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv_transpose2d(x.double(), wz, stride=2).backward(dout.double())
    res["wgrad"] = err(grad, wz.grad)
    res["ok"] = res["untouched"] and res["fwd"]["rel_l2"] < TOL_SPLIT_OUT and res["dgrad"]["rel_l2"] < TOL_SPLIT_OUT and \
        res["wgrad"]["rel_l2"] < TOL_F32_OUT and all(res[k]["finite"] for k in ("fwd", "dgrad", "wgrad"))
    return res


def check_hp_first_conv(B=3, H=32, W=32, nc=4, cat_mode=0, G=2, seed=106) -> dict:
    """First-layer conv as a single-tap split GEMM over the split im2col rows, with per-CTA statistics, and its weight
    gradient (wgrad mode 1)."""
    g = _gen(seed)
    x1 = torch.rand(B, 6, H, W, device=DEV, generator=g)
    x2 = torch.rand(B, 6, H, W, device=DEV, generator=g)
    cin = 2 * nc if cat_mode else nc
    kpad = 64 * ((9 * cin + 63) // 64)
    w = q16(torch.randn(64, cin, 3, 3, device=DEV, generator=g) / (3.0 * cin ** 0.5))
    b = torch.randn(64, device=DEV, generator=g)
    cols = ops.pack_input(x1, x2, 2, nc, cat_mode, kpad, prec=True)
    n_img = cols.shape[0]
    out = ops.split_alloc((n_img, H, W, 64), DEV)
    Gs = G if n_img % G == 0 else 1
    rows, per_cta = ops.conv_stat_rows(n_img, H, W, kpad, 64, Gs, mode=1, prec=True)
    stats = torch.full((Gs, rows, 64, 2), float("nan"), device=DEV)
    ops.conv_gemm(1, 0, cols, pack_hp(2, w, kpad=kpad), out, bias=b, stats=stats, stat_groups=Gs, prec=True)
    ops.device_status()
    xin = q16(torch.cat([x1[:, 2:2 + nc], x2[:, 2:2 + nc]], 1 if cat_mode else 0))
    ref = F.conv2d(xin.double(), w.double(), b.double(), padding=1)
    got = ops.split_to_float(out)
    res = {"fwd": err(nchw(got), ref), "per_cta": per_cta}
    res["stats_sum_rel"] = ((stats[..., 0].double().sum((0, 1)) - got.double().sum((0, 1, 2))).norm() /
                            got.double().sum((0, 1, 2)).norm()).item()
    dr = q16(torch.randn(n_img, 64, H, W, device=DEV, generator=g))
    splits = _splits_for(ops.wgrad_tiles(n_img, H, W), 4)
    ws = torch.full((splits, 64, kpad), float("nan"), device=DEV)
    ops.wgrad_gemm(1, 1, 0, split(nhwc(dr)), cols, ws, splits, 64 * kpad, 0, kpad, 1, prec=True)
    grad = torch.empty(64, cin, 3, 3, device=DEV)
    ops.wgrad_reduce(ws, splits, 64 * kpad, 1, 64, cin, 9, grad)
    ops.device_status()
    res["wgrad"] = err(grad, torch.nn.grad.conv2d_weight(xin.double(), (64, cin, 3, 3), dr.double(), padding=1))
    res["ok"] = res["fwd"]["finite"] and res["fwd"]["rel_l2"] < TOL_SPLIT_OUT and res["stats_sum_rel"] < 1e-5 and \
        res["wgrad"]["finite"] and res["wgrad"]["rel_l2"] < TOL_F32_OUT
    return res


def check_hp_wgrad3x3(n=2, H=32, W=32, cin=64, cout=128, sign=1, splits=5, seed=107) -> dict:
    g = _gen(seed)
    x = q16(torch.randn(n, cin, H, W, device=DEV, generator=g))
    dr = q16(torch.randn(n, cout, H, W, device=DEV, generator=g))
    xa, da = split(nhwc(x)), split(nhwc(dr))
    splits = _splits_for(ops.wgrad_tiles(n, H, W), splits)
    ws = torch.full((splits, 9, cout, cin), float("nan"), device=DEV)
    if sign == 1:
        ops.wgrad_gemm(0, 1, 1, da, xa, ws, splits, 9 * cout * cin, cout * cin, cin, 1, prec=True)
    else:
        ops.wgrad_gemm(0, -1, 1, xa, da, ws, splits, 9 * cout * cin, cout * cin, 1, cin, prec=True)
    grad = torch.empty(cout, cin, 3, 3, device=DEV)
    ops.wgrad_reduce(ws, splits, 9 * cout * cin, 0, cout, cin, 9, grad)
    ops.device_status()
    ref = torch.nn.grad.conv2d_weight(x.double(), (cout, cin, 3, 3), dr.double(), padding=1)
    res = err(grad, ref)
    res["ok"] = res["finite"] and res["rel_l2"] < TOL_F32_OUT
    return res


def _bn_setup_hp(n, H, W, Cc, seed):
    g = _gen(seed)
    r = q16(torch.randn(n, H, W, Cc, device=DEV, generator=g) * 1.5 + 0.3)
    gamma = torch.rand(Cc, device=DEV, generator=g) + 0.5
    beta = torch.randn(Cc, device=DEV, generator=g) * 0.2
    return g, r, gamma, beta


def _bn_stats_hp(r, gamma, beta, G):
    """mean / invstd / scale / shift through the CUDA statistics kernels from fp32 per-tile partials of r."""
    return _bn_forward_cuda(r, gamma, beta, G)   # only sums r (fp32): works on the fp32 values of the split tensor


def check_hp_bn_apply(n=4, H=16, W=32, Cc=64, seed=110, odd=False) -> dict:
    """BN-apply + ReLU + pool + arg-max + t2 - t1 + second copy on split tensors (odd: H, W not even — eval tiles)."""
    G = 2
    g, r, gamma, beta = _bn_setup_hp(n, H, W, Cc, seed)
    mean, invstd, scale, shift, *_ = _bn_stats_hp(r, gamma, beta, G)
    if odd:
        r = r[:, :H - 1, :W - 1].contiguous()
        H, W = H - 1, W - 1
    rs = split(r)
    a = ops.split_alloc((n, H, W, Cc), DEV)
    cat = ops.split_alloc((n // 2, H, W, 2 * Cc), DEV, zero=True)
    pool = ops.split_alloc((n, H // 2, W // 2, Cc), DEV)
    pidx = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.uint8)
    ops.bn_apply(rs, scale, shift, G, True, a=a, pool=pool, dif=cat[..., :Cc], pool_idx=pidx, prec=True)
    ops.device_status()
    per = n // G
    y = torch.relu(r.double().view(G, per, H, W, Cc) * scale.double().view(G, 1, 1, 1, Cc) + shift.double().view(G, 1, 1, 1, Cc))
    y = y.view(n, H, W, Cc)
    res = {"a": err(ops.split_to_float(a), y)}
    res["pool"] = err(nchw(ops.split_to_float(pool)), F.max_pool2d(nchw(y), 2))
    res["diff"] = err(ops.split_to_float(cat[..., :Cc]), y[n // 2:] - y[:n // 2])
    af = ops.split_to_float(a)[:, :2 * (H // 2), :2 * (W // 2)]
    win = af.reshape(n, H // 2, 2, W // 2, 2, Cc).permute(0, 1, 3, 5, 2, 4).reshape(n, H // 2, W // 2, Cc, 4)
    res["pool_idx_ok"] = bool(torch.equal(pidx.long(), win.argmax(-1)) or
                              torch.equal(win.gather(-1, pidx.long().unsqueeze(-1)).squeeze(-1), win.max(-1).values))
    res["cat_upper_untouched"] = bool((cat[..., Cc:] == 0).all().item() and (ops.lo_half(cat)[..., Cc:] == 0).all().item())
    res["ok"] = res["a"]["rel_l2"] < TOL_ELEM and res["pool"]["rel_l2"] < TOL_ELEM and res["diff"]["rel_l2"] < 4 * TOL_ELEM and \
        res["pool_idx_ok"] and res["cat_upper_untouched"]
    return res


def check_hp_bn_bwd(n=4, H=16, W=32, Cc=64, seed=111, order=("skip", "pool", "dir"), G=2) -> dict:
    """BN + ReLU backward on split tensors with every kind of gradient source (see gpu_checks.check_bn_bwd)."""
    g, r, gamma, beta = _bn_setup_hp(n, H, W, Cc, seed)
    mean, invstd, scale, shift, *_ = _bn_stats_hp(r, gamma, beta, G)
    rs = split(r)
    h = n // 2
    d_skip = q16(torch.randn(h, H, W, 2 * Cc, device=DEV, generator=g))
    d_pool = q16(torch.randn(n, H // 2, W // 2, Cc, device=DEV, generator=g))
    d_dir = q16(torch.randn(n, H, W, Cc, device=DEV, generator=g))
    dz = torch.randn(n, 1, H, W, device=DEV, generator=g)
    w_head = torch.randn(Cc, device=DEV, generator=g)
    pidx = torch.empty(n, H // 2, W // 2, Cc, device=DEV, dtype=torch.uint8)
    a = ops.split_alloc((n, H, W, Cc), DEV)
    ops.bn_apply(rs, scale, shift, G, False, a=a, pool=ops.split_alloc((n, H // 2, W // 2, Cc), DEV), pool_idx=pidx, prec=True)
    table = {
        "skip": {"kind": 1, "t": split(d_skip)[..., :Cc], "n_mod": h, "scale_lo": -1.0, "scale_hi": 1.0},
        "pool": {"kind": 2, "t": split(d_pool), "w": pidx},
        "dir": {"kind": 1, "t": split(d_dir)},
        "head": {"kind": 3, "t": dz, "w": w_head},
    }
    srcs = ops.make_srcs([table[k] for k in order])
    ws = torch.empty(ops.bn_bwd_ws_floats(n, H, W, Cc, G), device=DEV)
    dgamma = torch.empty(Cc, device=DEV)
    dbeta = torch.empty(Cc, device=DEV)
    dr = ops.split_alloc((n, H, W, Cc), DEV)
    ops.bn_bwd(rs, mean, invstd, scale, shift, srcs, G, ws, dgamma, dbeta, dr, prec=True)
    ops.device_status()
    x = nchw(r.double()).requires_grad_(True)
    ga = gamma.double().requires_grad_(True)
    be = beta.double().requires_grad_(True)
    per = n // G
    aa = torch.cat([F.relu(F.batch_norm(x[gi * per:(gi + 1) * per], None, None, ga, be, True, 0.1, 1e-5)) for gi in range(G)], 0)
    # the forward pools the STORED activations: route the arg-max through the stored values
    aq = aa + (nchw(ops.split_to_float(a)).double() - aa.detach())
    loss = 0.0
    if "skip" in order:
        loss = loss + ((aa[h:] - aa[:h]) * nchw(d_skip[..., :Cc]).double()).sum()
    if "pool" in order:
        loss = loss + (F.max_pool2d(aq, 2) * nchw(d_pool).double()).sum()
    if "dir" in order:
        loss = loss + (aa * nchw(d_dir).double()).sum()
    if "head" in order:
        loss = loss + ((aa * w_head.double().view(1, Cc, 1, 1)).sum(1, keepdim=True) * dz.double()).sum()
    loss.backward()
    res = {"dr": err(nchw(ops.split_to_float(dr)), x.grad), "dgamma": err(dgamma, ga.grad), "dbeta": err(dbeta, be.grad)}
    # the CUDA statistics are fp32 sums of squares (mean / invstd ~1e-6), and the ReLU mask of elements within that of
    # zero may differ from the fp64 graph: bounds are a few 1e-5, three orders below the bf16-storage kernels' 8e-3
    res["ok"] = res["dr"]["rel_l2"] < 1e-4 and res["dgamma"]["rel_l2"] < 5e-5 and res["dbeta"]["rel_l2"] < 5e-5
    return res


def check_hp_head(n=2, H=16, W=16, seed=112) -> dict:
    g = _gen(seed)
    a0 = q16(torch.randn(n, H, W, 64, device=DEV, generator=g))
    a1 = q16(torch.randn(n, H, W, 64, device=DEV, generator=g))
    w = torch.randn(128, device=DEV, generator=g)
    b = torch.randn(1, device=DEV, generator=g)
    logits = torch.empty(n, 1, H, W, device=DEV)
    s0, s1 = split(a0), split(a1)
    ops.head_fwd(s0, s1, w, b, logits, prec=True)
    res = {"fusion": err(logits.view(n, H, W), (torch.cat([a0, a1], -1).double() * w.double()).sum(-1) + b.double())}
    ops.head_fwd(s0, None, w[:64].contiguous(), b, logits, prec=True)
    res["single"] = err(logits.view(n, H, W), (a0.double() * w[:64].double()).sum(-1) + b.double())
    dz = torch.randn(n * H * W, device=DEV, generator=g)
    nblk = 8
    ws = torch.empty(nblk * 64, device=DEV)
    dw = torch.empty(64, device=DEV)
    ops.colsum(s0, dz, n * H * W, nblk, ws, dw, prec=True)
    res["dw"] = err(dw, (a0.double().view(-1, 64) * dz.double()[:, None]).sum(0))
    cs = torch.empty(64, device=DEV)
    ops.colsum(s1, None, n * H * W, nblk, ws, cs, prec=True)
    res["colsum"] = err(cs, a1.double().view(-1, 64).sum(0))
    # centre pad of Up on split tensors
    src = split(torch.randn(n, 8, 10, 64, device=DEV, generator=g))
    cat = ops.split_alloc((n, 9, 11, 128), DEV)
    cat.fill_(3.0)
    ops.lo_half(cat).fill_(3.0)
    ops.pad_copy(src, cat[..., 64:], 0, 1, prec=True)
    ops.device_status()
    refp = F.pad(ops.split_to_float(src).permute(0, 3, 1, 2), (1, 0, 0, 1)).permute(0, 2, 3, 1)
    res["pad_ok"] = bool(torch.equal(ops.split_to_float(cat[..., 64:]), refp) and (cat[..., :64] == 3.0).all().item())
    res["ok"] = all(res[k]["rel_l2"] < 2e-6 for k in ("fusion", "single", "dw")) and res["colsum"]["rel_l2"] < 1e-5 and res["pad_ok"]
    return res


HP_CHECKS = {
    "hp_pack": check_hp_pack,
    "hp_conv3x3_64_64": lambda: check_hp_conv3x3(),
    "hp_conv3x3_64_128_stats2": lambda: check_hp_conv3x3(n=4, cin=64, cout=128, G=2),
    "hp_conv3x3_128_256_stats1": lambda: check_hp_conv3x3(n=2, H=16, W=16, cin=128, cout=256, G=1),
    "hp_conv3x3_512_512_8x8": lambda: check_hp_conv3x3(n=3, H=8, W=8, cin=512, cout=512, G=1, seed=113),
    "hp_conv3x3_slices": lambda: check_hp_conv3x3(cin=128, cout=64, slice_in=True, slice_out=True, G=1, seed=114),
    "hp_conv3x3_odd_19x21": lambda: check_hp_conv3x3(n=1, H=19, W=21, cin=64, cout=64, seed=115),
    "hp_dgrad_128_64": lambda: check_hp_dgrad(),
    "hp_dgrad_256_256": lambda: check_hp_dgrad(n=2, H=16, W=16, cin=256, cout=256, seed=116),
    "hp_convt_128": lambda: check_hp_convt(),
    "hp_convt_64": lambda: check_hp_convt(n=2, h=16, w_=16, c=64, seed=117),
    "hp_first_conv_siamese4": lambda: check_hp_first_conv(),
    "hp_first_conv_cat8": lambda: check_hp_first_conv(B=2, nc=4, cat_mode=1, G=1, seed=118),
    "hp_first_conv_sar2": lambda: check_hp_first_conv(B=2, nc=2, cat_mode=0, seed=119),
    "hp_wgrad_pos_128_64": lambda: check_hp_wgrad3x3(),
    "hp_wgrad_neg_128_64": lambda: check_hp_wgrad3x3(cin=128, cout=64, sign=-1, seed=120),
    "hp_wgrad_64_64_mstack": lambda: check_hp_wgrad3x3(cin=64, cout=64, seed=121),
    "hp_wgrad_256_256": lambda: check_hp_wgrad3x3(n=2, H=16, W=16, cin=256, cout=256, splits=3, seed=122),
    "hp_bn_apply": lambda: check_hp_bn_apply(),
    "hp_bn_apply_128": lambda: check_hp_bn_apply(n=2, H=16, W=16, Cc=128, seed=123),
    "hp_bn_apply_odd": lambda: check_hp_bn_apply(odd=True, seed=124),
    "hp_bn_bwd_skip_pool_dir": lambda: check_hp_bn_bwd(),
    "hp_bn_bwd_pool_skip": lambda: check_hp_bn_bwd(order=("pool", "skip"), seed=125),
    "hp_bn_bwd_dir_head": lambda: check_hp_bn_bwd(order=("dir", "head"), G=1, seed=126),
    "hp_bn_bwd_head_head_128": lambda: check_hp_bn_bwd(Cc=128, order=("head",), G=1, seed=127),
    "hp_head_colsum_pad": check_hp_head,
}
