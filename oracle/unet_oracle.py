"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's training-step hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package (multimodal_siamese_cd_b200/) never does. It is plain torch on CPU tensors: functional
re-statements of the reference modules driven by a reference-format `state_dict`, written from the reference's
behaviour, not copied from it. Every function cites the reference lines it follows (paths relative to the reference
root, /root/reference in the build container).

Where the arithmetic lives: in torch (third-party, unpinned by the reference — it ships no requirements file;
SURVEY.md §8c). This restatement therefore calls the same torch CPU kernels (F.conv2d, F.batch_norm, ...) in the
reference's order. Parity pin: tests/golden/*.pt hold outputs of the UNMODIFIED reference modules
(utils/networks.py, utils/loss_functions.py imported from /root/reference by oracle/make_golden.py in the build
container); tests/test_oracle_golden.py checks this file against them bit-for-bit-level (<= 1e-6) on CPU.
tests/golden/eval/*.pt (oracle/make_eval_golden.py) pin the inference path on tiles with odd-sized levels the same way.

Two modes:
  q=False  exact fp32 (or fp64 via dtype) restatement == the reference.
  q=True   the same algorithm with values rounded to bf16 exactly where the B200 pipeline stores bf16
           (north_star: "TF32/bf16 inputs, fp32 accumulation"): conv inputs/weights/outputs, activations,
           feature differences, and the gradients that the pipeline materialises (dr, d_activation).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------
# bf16 storage emulation
# ----------------------------------------------------------------------------------------------------------
# Storage format the q=True mode emulates: "bf16" (the fast mode: one bf16 per element) or "split" (the precise mode:
# hi = bf16(x), lo = bf16(x - hi), 16 mantissa bits; include/b200cd.h ABI version 2).
_STORAGE = ["bf16"]


def set_storage(fmt: str) -> None:
    assert fmt in ("bf16", "split"), fmt
    _STORAGE[0] = fmt


def _store(x):
    hi = x.to(torch.bfloat16).to(x.dtype)
    if _STORAGE[0] == "bf16":
        return hi
    return hi + (x - hi).to(torch.bfloat16).to(x.dtype)


class _RoundBoth(torch.autograd.Function):
    """value and incoming gradient both rounded to the storage format (a tensor the pipeline stores in bf16 / split-bf16
    whose gradient is stored the same way)."""

    @staticmethod
    def forward(ctx, x):
        return _store(x)

    @staticmethod
    def backward(ctx, g):
        return _store(g)


class _RoundFwd(torch.autograd.Function):
    """value rounded to the storage format, gradient passed through (weights; inputs of the fp32 1x1 heads)."""

    @staticmethod
    def forward(ctx, x):
        return _store(x)

    @staticmethod
    def backward(ctx, g):
        return g


def rb(x, q):
    return _RoundBoth.apply(x) if q else x


def rf(x, q):
    return _RoundFwd.apply(x) if q else x


# ----------------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------------
class BNState:
    """Running statistics, updated in place exactly like nn.BatchNorm2d in train mode."""

    def __init__(self, sd: dict, prefix: str):
        self.mean = sd[prefix + ".running_mean"]
        self.var = sd[prefix + ".running_var"]
        self.nbt = sd[prefix + ".num_batches_tracked"]


def conv_bn_relu(x, sd, conv_p, bn_p, train, q, x_is_rounded=False):
    """nn.Conv2d(.,.,3,padding=1) -> nn.BatchNorm2d -> nn.ReLU  (utils/networks.py:392-394 / 395-397).
    Returns the fp32 activation (callers round per consumer, see module docstring)."""
    w, b = sd[conv_p + ".weight"], sd[conv_p + ".bias"]
    xin = x if x_is_rounded else rb(x, q)
    r = F.conv2d(xin, rf(w, q), b, padding=1)
    r = rb(r, q)                                            # conv output is stored in bf16
    st = BNState(sd, bn_p)
    if train:
        st.nbt += 1                                         # nn.BatchNorm2d.forward: num_batches_tracked += 1
    y = F.batch_norm(r, st.mean, st.var, sd[bn_p + ".weight"], sd[bn_p + ".bias"], train, 0.1, 1e-5)
    return F.relu(y)


def double_conv(x, sd, p, train, q, x_is_rounded=False):
    """DoubleConv.forward (utils/networks.py:386-402); p = '<...>.conv' (the nn.Sequential)."""
    a1 = conv_bn_relu(x, sd, f"{p}.0", f"{p}.1", train, q, x_is_rounded)
    return conv_bn_relu(a1, sd, f"{p}.3", f"{p}.4", train, q)


def encoder(x, sd, inc_p, enc_p, n_levels, train, q):
    """InConv + Encoder.forward (utils/networks.py:405-412, 334-343). Returns features shallow -> deep (fp32)."""
    feats = [double_conv(x, sd, f"{inc_p}.conv.conv", train, q)]
    for i in range(1, n_levels + 1):
        a = feats[-1]
        pooled = F.max_pool2d(rb(a, q), 2)                  # MaxPool2d(2) of the stored activation (:420)
        feats.append(double_conv(pooled, sd, f"{enc_p}.down_seq.down{i}.mpconv.1.conv", train, q, x_is_rounded=q))
    return feats


def decoder(feats, sd, dec_p, train, q, feats_rounded=False):
    """Decoder.forward (utils/networks.py:375-382) with Up.forward (:436-451): ConvTranspose2d(2, stride 2) ->
    centre pad -> cat([skip, up]) -> DoubleConv. `feats` shallow -> deep; returns the fp32 decoder output."""
    n = len(feats) - 1
    x = feats[-1] if feats_rounded else rb(feats[-1], q)
    for k in range(n, 0, -1):
        p = f"{dec_p}.up_seq.up{k}"
        up = F.conv_transpose2d(x, rf(sd[p + ".up.weight"], q), sd[p + ".up.bias"], stride=2)
        up = rb(up, q)
        skip = feats[k - 1] if feats_rounded else rb(feats[k - 1], q)
        dy, dx = skip.shape[2] - up.shape[2], skip.shape[3] - up.shape[3]
        up = F.pad(up, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        a = double_conv(torch.cat([skip, up], 1), sd, f"{p}.conv.conv", train, q, x_is_rounded=q)
        x = rb(a, q) if k > 1 else a
    return x                                                 # fp32; heads read it rounded forward-only


def head(xs, sd, p, q):
    """OutConv (utils/networks.py:454-461) on cat(xs, 1); inputs are read from bf16 storage, math in fp32."""
    x = torch.cat([rf(x, q) for x in xs], 1)
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"])


def _diff(f1, f2, q):
    """torch.sub(f_t2, f_t1) per level (utils/networks.py:147-150); the difference is stored in bf16."""
    return [rb(b - a, q) for a, b in zip(f1, f2)]


def _levels(sd, enc_p):
    i = 0
    while f"{enc_p}.down_seq.down{i + 1}.mpconv.1.conv.0.weight" in sd:
        i += 1
    return i


# ----------------------------------------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------------------------------------
def forward(model_type: str, sd: dict, x_t1, x_t2, train: bool = True, q: bool = False, n_s1: int = 2):
    """forward(x_t1, x_t2) of the six network types. `sd`: reference state_dict WITHOUT the 'module.' prefix
    (BN buffers are updated in place when train=True). Returns a tensor or the reference's tuple."""
    if model_type == "unet":                                 # UNet.forward utils/networks.py:73-79
        L = _levels(sd, "encoder")
        f = encoder(torch.cat((x_t1, x_t2), 1), sd, "inc", "encoder", L, train, q)
        return head([decoder(f, sd, "decoder", train, q)], sd, "outc", q)
    if model_type == "siameseunet":                          # SiameseUNet.forward :139-154
        L = _levels(sd, "encoder")
        f1 = encoder(x_t1, sd, "inc", "encoder", L, train, q)
        f2 = encoder(x_t2, sd, "inc", "encoder", L, train, q)
        return head([decoder(_diff(f1, f2, q), sd, "decoder", train, q, feats_rounded=q)], sd, "outc", q)
    if model_type == "dtsiameseunet":                        # DualTaskSiameseUNet.forward :176-197
        L = _levels(sd, "encoder")
        f1 = encoder(x_t1, sd, "inc", "encoder", L, train, q)
        f2 = encoder(x_t2, sd, "inc", "encoder", L, train, q)
        out_change = head([decoder(_diff(f1, f2, q), sd, "decoder_change", train, q, feats_rounded=q)], sd,
                          "outc_change", q)
        out_sem_t2 = head([decoder(f2, sd, "decoder_sem", train, q)], sd, "outc_sem", q)   # t2 FIRST (:191-195)
        out_sem_t1 = head([decoder(f1, sd, "decoder_sem", train, q)], sd, "outc_sem", q)
        return out_change, out_sem_t1, out_sem_t2
    if model_type in ("dualstreamunet", "whatevernet2"):     # :103-120, :288-310
        xs = []
        for k, sl in ((1, slice(0, n_s1)), (2, slice(n_s1, None))):
            L = _levels(sd, f"encoder_stream{k}")
            f = encoder(torch.cat((x_t1[:, sl], x_t2[:, sl]), 1), sd, f"inc_stream{k}", f"encoder_stream{k}", L, train, q)
            xs.append(decoder(f, sd, f"decoder_stream{k}", train, q))
        if model_type == "dualstreamunet":
            return head(xs, sd, "outc", q)
        outs = (head(xs, sd, "outc_fusion", q), head(xs[:1], sd, "outc_stream1", q), head(xs[1:], sd, "outc_stream2", q))
        return outs if train else outs[0]
    if model_type == "whatevernet":                          # WhateverNet.forward :231-263
        xs = []
        for k, sl in ((1, slice(0, n_s1)), (2, slice(n_s1, None))):
            L = _levels(sd, f"encoder_stream{k}")
            f1 = encoder(x_t1[:, sl], sd, f"inc_stream{k}", f"encoder_stream{k}", L, train, q)
            f2 = encoder(x_t2[:, sl], sd, f"inc_stream{k}", f"encoder_stream{k}", L, train, q)
            xs.append(decoder(_diff(f1, f2, q), sd, f"decoder_stream{k}", train, q, feats_rounded=q))
        outs = (head(xs, sd, "outc_fusion", q), head(xs[:1], sd, "outc_stream1", q), head(xs[1:], sd, "outc_stream2", q))
        return outs if train else outs[0]
    raise Exception(f"Unknown network ({model_type}).")


# ----------------------------------------------------------------------------------------------------------
# losses and their compositions
# ----------------------------------------------------------------------------------------------------------
def power_jaccard_loss(logits, target):
    """utils/loss_functions.py:141-150."""
    p = torch.sigmoid(logits).flatten()
    t = target.flatten()
    inter = (p * t).sum()
    denom = (p ** 2 + t ** 2).sum() - inter + 1e-6
    return 1 - inter / denom


def supervised_loss(out, y_change):
    """train_supervised.py:71-76."""
    return power_jaccard_loss(out, y_change)


def dualtask_loss(outs, y_change, y_sem_t1, y_sem_t2):
    """train_supervised_dualtask.py:75-85."""
    c, s1, s2 = outs
    sem = (power_jaccard_loss(s1, y_sem_t1) + power_jaccard_loss(s2, y_sem_t2)) / 2
    return (power_jaccard_loss(c, y_change) + sem) / 2


def mmcr_loss(outs, y_change, is_labeled, alpha):
    """train_semisupervised.py:74-113 (PowerJaccard consistency variant)."""
    f, s1, s2 = outs
    p2 = torch.sigmoid(s2)
    sup = cons = None
    if is_labeled.any():
        sup = alpha * (power_jaccard_loss(f[is_labeled], y_change[is_labeled]) +
                       power_jaccard_loss(s1[is_labeled], y_change[is_labeled]) +
                       power_jaccard_loss(s2[is_labeled], y_change[is_labeled])) / 3
    if not is_labeled.all():
        unl = torch.logical_not(is_labeled)
        cons = (1 - alpha) * power_jaccard_loss(s1[unl], p2[unl])
    if sup is None:
        return cons
    return sup if cons is None else sup + cons


def change_mask_f1(logits, y_true):
    """Thresholded mask and F1 as utils/evaluation.py:12-40 + utils/metrics.py:23-66 compute them at threshold 0.5:
    mask = (sigmoid(z) - 0.5 + 0.5).round().bool(); F1 from TP / 'FP' / 'FN' (the swap in :30-31 leaves F1 unchanged)."""
    pred = (torch.sigmoid(logits) - 0.5 + 0.5).round().bool()
    yt = y_true.bool()
    tp = (yt & pred).sum().float()
    fp = (yt & ~pred).sum().float()      # reference naming (swapped)
    fn = (~yt & pred).sum().float()
    prec = tp / (tp + fp).clamp(10e-05)
    rec = tp / (tp + fn).clamp(10e-05)
    f1 = 2 * prec * rec / (prec + rec).clamp(10e-05)
    return pred, f1


# ----------------------------------------------------------------------------------------------------------
# default initialisation (so the CPU baseline needs nothing from the product package)
# ----------------------------------------------------------------------------------------------------------
def reference_state_dict(model_type, in_channels=6, topology=(64, 128, 256, 512), n_s1=2, n_s2=4, seed=7) -> dict:
    """state_dict (no 'module.' prefix) with torch's default init drawn in the reference's construction order
    (utils/networks.py:132-137, 166-174, 209-220, 326-332, 364-371, 391-398, 433-434): nn.Conv2d /
    nn.ConvTranspose2d draw weight then bias, BatchNorm draws nothing."""
    import torch.nn as nn
    topo = list(topology)
    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v

    def double_conv(prefix, cin, cout):
        put(f"{prefix}.conv.0", nn.Conv2d(cin, cout, 3, padding=1))
        put(f"{prefix}.conv.1", nn.BatchNorm2d(cout))
        put(f"{prefix}.conv.3", nn.Conv2d(cout, cout, 3, padding=1))
        put(f"{prefix}.conv.4", nn.BatchNorm2d(cout))

    widths = topo[1:] + topo[-1:]

    def enc(prefix):
        for i, (ci, co) in enumerate(zip(topo, widths), start=1):
            double_conv(f"{prefix}.down_seq.down{i}.mpconv.1", ci, co)

    def dec(prefix):
        w = topo[:1] + widths
        for depth in range(len(topo), 0, -1):
            below, above = w[depth - 1], (w[depth - 2] if depth > 1 else w[0])
            put(f"{prefix}.up_seq.up{depth}.up", nn.ConvTranspose2d(below, below, 2, stride=2))
            double_conv(f"{prefix}.up_seq.up{depth}.conv", 2 * below, above)

    def outc(prefix, cin):
        put(f"{prefix}.conv", nn.Conv2d(cin, 1, 1))

    torch.manual_seed(seed)
    t0 = topo[0]
    if model_type in ("unet", "siameseunet"):
        double_conv("inc.conv", in_channels * (2 if model_type == "unet" else 1), t0)
        enc("encoder"); dec("decoder"); outc("outc", t0)
    elif model_type == "dtsiameseunet":
        double_conv("inc.conv", in_channels, t0)
        enc("encoder"); dec("decoder_change"); dec("decoder_sem")
        outc("outc_change", t0); outc("outc_sem", t0); outc("outc_sem_change", 2)
    elif model_type in ("dualstreamunet", "whatevernet", "whatevernet2"):
        mult = 1 if model_type == "whatevernet" else 2
        for k, nb in ((1, n_s1), (2, n_s2)):
            double_conv(f"inc_stream{k}.conv", mult * nb, t0)
            enc(f"encoder_stream{k}"); dec(f"decoder_stream{k}")
            if model_type != "dualstreamunet":
                outc(f"outc_stream{k}", t0)
        outc("outc" if model_type == "dualstreamunet" else "outc_fusion", 2 * t0)
    else:
        raise Exception(f"Unknown network ({model_type}).")
    return sd


# ----------------------------------------------------------------------------------------------------------
# one full step (what parity tests and the CPU baseline run)
# ----------------------------------------------------------------------------------------------------------
def clone_state(sd: dict, dtype=torch.float32, requires_grad: bool = True) -> dict:
    out = {}
    for k, v in sd.items():
        k = k[len("module."):] if k.startswith("module.") else k
        if v.is_floating_point():
            t = v.detach().to("cpu", dtype).clone()
            if requires_grad and not (k.endswith("running_mean") or k.endswith("running_var")):
                t.requires_grad_(True)
            out[k] = t
        else:
            out[k] = v.detach().to("cpu").clone()
    return out


def train_step(model_type, sd, batch, kind="supervised", alpha=0.5, q=False, n_s1=2):
    """zero_grad -> forward -> loss -> backward of one batch. Returns dict(outs, loss, grads{name: tensor})."""
    outs = forward(model_type, sd, batch["x_t1"], batch["x_t2"], train=True, q=q, n_s1=n_s1)
    if kind == "supervised":
        loss = supervised_loss(outs, batch["y_change"])
    elif kind == "dualtask":
        loss = dualtask_loss(outs, batch["y_change"], batch["y_sem_t1"], batch["y_sem_t2"])
    elif kind == "mmcr":
        loss = mmcr_loss(outs, batch["y_change"], batch["is_labeled"], alpha)
    else:
        raise ValueError(kind)
    names = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return {"outs": outs, "loss": loss.detach(), "grads": dict(zip(names, grads))}


def synthetic_batch(B, C, H, W, seed=7, corr=False, dtype=torch.float32):
    """Deterministic synthetic batch of SURVEY.md §8d (CPU generator so every box sees the same numbers)."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(B, C, H, W, generator=g)
    u = torch.rand(B, C, H, W, generator=g)
    x2 = (x1 + 0.1 * u).clamp(0, 1) if corr else u
    y = (torch.rand(B, 1, H, W, generator=g) > 0.9).float()
    ys1 = (torch.rand(B, 1, H, W, generator=g) > 0.8).float()
    ys2 = (torch.rand(B, 1, H, W, generator=g) > 0.8).float()
    lab = torch.tensor([i % 3 != 2 for i in range(B)])
    return {"x_t1": x1.to(dtype), "x_t2": x2.to(dtype), "y_change": y.to(dtype), "y_sem_t1": ys1.to(dtype),
            "y_sem_t2": ys2.to(dtype), "is_labeled": lab}
