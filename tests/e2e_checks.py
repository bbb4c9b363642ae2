"""End-to-end parity of the CUDA path (drop-in modules and fused TrainStep) against the oracle on the same seeded
weights and batch. Shared by tests/test_gpu_e2e.py and tools/gpu_e2e_probe.py."""
from __future__ import annotations

import torch

from multimodal_siamese_cd_b200 import loss_functions, networks
from multimodal_siamese_cd_b200.config import synthetic_cfg
from multimodal_siamese_cd_b200.step import TrainStep
from oracle import unet_oracle as O

TWO_STREAM = ("dualstreamunet", "whatevernet", "whatevernet2")
_ORACLE_MEMO: dict = {}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def is_prebn_bias(name: str) -> bool:
    """Bias of a 3x3 conv that feeds a train-mode BatchNorm (DoubleConv's Sequential indices 0 and 3)."""
    return name.endswith((".conv.0.bias", ".conv.3.bias"))


def grad_report(got: dict, ref: dict) -> dict:
    """global relL2 over all parameters (pre-BN conv biases excluded: analytically zero) + per-parameter stats."""
    num = den = 0.0
    per = {}
    for k, r in ref.items():
        g = got.get(k)
        if r is None:
            assert g is None, f"{k}: reference has no gradient, CUDA path produced one"
            continue
        if is_prebn_bias(k):
            continue
        d = (g.double().cpu() - r.double()).norm().item()
        n = r.double().norm().item()
        num += d * d
        den += n * n
        per[k] = d / max(n, 1e-30)
    vals = sorted(per.values())
    worst = max(per, key=per.get)
    return {"global": (num / max(den, 1e-60)) ** 0.5, "median": vals[len(vals) // 2], "max": vals[-1], "worst": worst}


def run_case(mtype: str, cin: int, topo, B: int, H: int, W: int, kind: str, alpha: float = 0.5, path: str = "dropin",
             corr: bool = False, graphs: bool = True, steps: int = 1, precision: str = "fast", fp64: bool = False,
             skip_q: bool = False, format_floor: bool = False) -> dict:
    """Returns error metrics of the CUDA path vs the bf16-storage oracle (q) and the exact fp32 oracle (x = the
    reference's arithmetic). fp64=True adds (d): the same step in float64, plus the reference's OWN fp32-vs-fp64
    error (`floor_*`) — the yardstick north_star's tolerance has to be read against."""
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg)
    net.module.set_precision(precision)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    xc = 6 if mtype in TWO_STREAM else cin
    batch = O.synthetic_batch(B, xc, H, W, seed=7, corr=corr)
    net.to(dev)
    net.train()
    net.module.use_cuda_graphs = graphs
    gb = {k: v.to(dev) for k, v in batch.items() if k != "is_labeled"}
    got_grads = None
    for it in range(steps):
        if it > 0:  # repeat the same step from the same state (exercises graph replay)
            net.load_state_dict(sd0)
        if path == "dropin":
            for p in net.parameters():
                p.grad = None
            crit = loss_functions.get_criterion("PowerJaccardLoss")
            outs = net(gb["x_t1"], gb["x_t2"])
            if kind == "supervised":
                loss = crit(outs, gb["y_change"])
                out_list = [outs]
            elif kind == "dualtask":
                c, s1, s2 = outs
                loss = (crit(c, gb["y_change"]) + (crit(s1, gb["y_sem_t1"]) + crit(s2, gb["y_sem_t2"])) / 2) / 2
                out_list = [c, s1, s2]
            else:
                f, s1, s2 = outs
                lab = batch["is_labeled"]
                y = gb["y_change"]
                p2 = torch.sigmoid(s2)
                sup = alpha * (crit(f[lab,], y[lab,]) + crit(s1[lab,], y[lab,]) + crit(s2[lab,], y[lab,])) / 3
                unl = torch.logical_not(lab)
                loss = sup + (1 - alpha) * crit(s1[unl,], p2[unl,])
                out_list = [f, s1, s2]
            loss.backward()
            got_grads = {n: (p.grad.detach().clone() if p.grad is not None else None)
                         for n, p in net.module.named_parameters()}
            got_outs = [o.detach().clone() for o in out_list]
            got_loss = loss.item()
        else:
            ts = TrainStep(net.module, B, H, W, kind=kind, alpha=alpha, device=dev, dp_group=None) if it == 0 else ts
            tg = {k: gb[k] for k in ts.targets}
            loss = ts(gb["x_t1"], gb["x_t2"], is_labeled=batch["is_labeled"] if kind == "mmcr" else None, **tg)
            got_loss = loss.item()
            got_outs = [o.detach().clone() for o in ts.eng.output_tensors()]
            g = ts.eng.grads
            got_grads = {n: (None if n in g.skip else g.views[n].detach().clone()) for n, _ in g.params}
    torch.cuda.synchronize()
    from multimodal_siamese_cd_b200 import ops
    ops.device_status(0)
    res = {"loss_cuda": got_loss}
    # analytically-zero gradients are emitted as exact zeros; the reference carries ~1e-9 noise there
    res["prebn_bias_grad_max"] = max(g.abs().max().item() for n, g in got_grads.items() if is_prebn_bias(n))
    sd_after = {k[len("module."):]: v.detach().cpu() for k, v in net.state_dict().items()}
    ref_x = None
    for tag, q in (("q", True), ("x", False), ("d", False)):
        if (tag == "q" and skip_q) or (tag == "d" and not fp64):
            continue
        # the oracle runs are the slow part at BASELINE shapes: shared between the numerics modes of one case
        memo_key = (mtype, cin, tuple(topo), B, H, W, kind, alpha, corr, tag)
        if memo_key in _ORACLE_MEMO:
            ref, sd = _ORACLE_MEMO[memo_key]
        elif tag == "d":
            sd = O.clone_state(sd0, dtype=torch.float64)
            b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
            ref = O.train_step(mtype, sd, b64, kind=kind, alpha=alpha, q=False)
        else:
            sd = O.clone_state(sd0)
            ref = O.train_step(mtype, sd, batch, kind=kind, alpha=alpha, q=q)
        if tag != "q":
            _ORACLE_MEMO[memo_key] = (ref, sd)
        ro = ref["outs"] if isinstance(ref["outs"], tuple) else (ref["outs"],)
        if tag == "x":
            ref_x = (ro, ref)
        if tag == "d" and ref_x is not None:   # the reference's own fp32 arithmetic against fp64
            res["floor_logits"] = max(rel(a.detach(), b.detach()) for a, b in zip(ref_x[0], ro))
            res["floor_loss"] = abs(ref_x[1]["loss"].item() - ref["loss"].item())
            res["floor_grads"] = grad_report(ref_x[1]["grads"], ref["grads"])
            res["floor_mask_flips"] = int(((ref_x[0][0].detach() > 0) != (ro[0].detach() > 0)).sum())
        res[f"logits_{tag}"] = max(rel(g, r.detach()) for g, r in zip(got_outs, ro))
        res[f"loss_{tag}"] = abs(got_loss - ref["loss"].item())
        res[f"grads_{tag}"] = grad_report(got_grads, ref["grads"])
        pm, _ = O.change_mask_f1(ro[0].detach(), batch["y_change"])
        gm, gf1 = O.change_mask_f1(got_outs[0].cpu(), batch["y_change"])
        _, rf1 = O.change_mask_f1(ro[0].detach(), batch["y_change"])
        res[f"mask_flips_{tag}"] = int((pm != gm).sum())
        # flips on pixels whose reference logit has margin (identity is only meaningful away from 0)
        res[f"margin_flips_{tag}"] = int(((pm != gm) & (ro[0].detach().abs() >= 0.05)).sum())
        res[f"margin3_flips_{tag}"] = int(((pm != gm) & (ro[0].detach().abs() >= 1e-3)).sum())
        res[f"f1_diff_{tag}"] = abs(gf1.item() - rf1.item())
        if tag == ("x" if precision == "precise" else "q"):   # running statistics against the oracle of the same storage
            bn_err = 0.0
            for k, v in sd.items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    bn_err = max(bn_err, (sd_after[k] - v).abs().max().item())
                elif k.endswith("num_batches_tracked"):
                    assert int(sd_after[k]) == int(v), f"{k}: {int(sd_after[k])} != {int(v)}"
            res["bn_running_maxabs"] = bn_err
    if format_floor and fp64:
        # the storage format's own floor: the oracle with split-bf16 storage points and EXACT (fp64) arithmetic against
        # the exact fp64 run — what a perfect kernel set with this storage would measure as grads_d
        memo_key = (mtype, cin, tuple(topo), B, H, W, kind, alpha, corr, "fmt")
        if memo_key not in _ORACLE_MEMO:
            O.set_storage("split")
            try:
                b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
                rq = O.train_step(mtype, O.clone_state(sd0, dtype=torch.float64), b64, kind=kind, alpha=alpha, q=True)
            finally:
                O.set_storage("bf16")
            _ORACLE_MEMO[memo_key] = rq
        rq = _ORACLE_MEMO[memo_key]
        r64 = _ORACLE_MEMO[(mtype, cin, tuple(topo), B, H, W, kind, alpha, corr, "d")][0]
        res["format_floor_grads"] = grad_report(rq["grads"], r64["grads"])
        ro64 = r64["outs"] if isinstance(r64["outs"], tuple) else (r64["outs"],)
        rqo = rq["outs"] if isinstance(rq["outs"], tuple) else (rq["outs"],)
        res["format_floor_logits"] = max(rel(a.detach(), b.detach()) for a, b in zip(rqo, ro64))
    res["n_pixels"] = got_outs[0].numel()
    net.module.release_engines()
    return res


def run_eval_case(mtype: str, cin: int, topo, B: int, H: int, W: int, warm_steps: int = 1, warm_hw=None,
                  precision: str = "fast") -> dict:
    """Inference path of utils/evaluation.py:7-23: net.eval(), no_grad, sigmoid(logits) > 0.5, F1 — after `warm_steps`
    training steps so that the BatchNorm running statistics are not the initial (0, 1)."""
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg)
    net.module.set_precision(precision)
    qq = precision == "fast"       # the oracle that shares the training warm-up emulates the mode's storage
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    xc = 6 if mtype in TWO_STREAM else cin
    batch = O.synthetic_batch(B, xc, H, W, seed=7)
    net.to(dev)
    gb = {k: v.to(dev) for k, v in batch.items() if k != "is_labeled"}
    sd = O.clone_state(sd0, requires_grad=False)
    net.train()
    # training tiles are multiples of 16 (the reference trains on 256 x 256 crops): odd-sized inference cases warm the
    # running statistics on a separate even-sized batch
    wb = batch if warm_hw is None else O.synthetic_batch(B, xc, warm_hw[0], warm_hw[1], seed=11)
    for _ in range(warm_steps):                      # forward only: updates the running statistics on both sides
        with torch.no_grad():
            net(wb["x_t1"].to(dev), wb["x_t2"].to(dev))
            O.forward(mtype, sd, wb["x_t1"], wb["x_t2"], train=True, q=qq)
    net.eval()
    with torch.no_grad():
        out = net(gb["x_t1"], gb["x_t2"])
        out_again = net(gb["x_t1"], gb["x_t2"])      # second call replays the CUDA graph
        ref_q = O.forward(mtype, sd, batch["x_t1"], batch["x_t2"], train=False, q=True)
        ref_x = O.forward(mtype, sd, batch["x_t1"], batch["x_t2"], train=False, q=False)
    torch.cuda.synchronize()
    assert torch.is_tensor(out), "eval returns a single tensor for every network type but dtsiamese"
    res = {"logits_q": rel(out, ref_q), "logits_x": rel(out, ref_x), "replay_equal": bool(torch.equal(out, out_again))}
    gm, gf1 = O.change_mask_f1(out.cpu(), batch["y_change"])
    pm, rf1 = O.change_mask_f1(ref_x, batch["y_change"])
    res["margin_flips_x"] = int(((pm != gm) & (ref_x.abs() >= 0.05)).sum())
    res["margin3_flips_x"] = int(((pm != gm) & (ref_x.abs() >= 1e-3)).sum())
    res["f1_diff_x"] = abs(gf1.item() - rf1.item())
    sd_after = {k[len("module."):]: v.detach().cpu() for k, v in net.state_dict().items()}
    res["bn_unchanged_in_eval"] = all(
        torch.allclose(sd_after[k], v, atol=2e-3) for k, v in sd.items() if k.endswith(("running_mean", "running_var")))
    net.module.release_engines()
    return res


def run_training_loop(mtype: str = "siameseunet", cin: int = 4, topo=(64, 128), B: int = 4, H: int = 32, W: int = 32,
                      steps: int = 6, lr: float = 1e-3, optimizer: str = "torch", kind: str = "supervised") -> dict:
    """The body of the reference's training loop (train_supervised.py:63-79) for several optimizer steps: drop-in
    modules + AdamW on the GPU against the oracle + torch.optim.AdamW on the CPU, different batch every step. Loss
    trajectories must track each other: parameters, BatchNorm running statistics and optimizer state all evolve."""
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg)
    sd = O.clone_state(net.state_dict())
    net.to(dev).train()
    if optimizer == "fused":
        from multimodal_siamese_cd_b200.optim import FusedAdamW
        opt = FusedAdamW(net.parameters(), lr=lr, weight_decay=0.01)
    else:
        opt = torch.optim.AdamW(net.parameters(), lr=lr, weight_decay=0.01)
    ref_params = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    ref_names = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    ref_opt = torch.optim.AdamW(ref_params, lr=lr, weight_decay=0.01)
    crit = loss_functions.get_criterion("PowerJaccardLoss")
    xc = 6 if mtype in TWO_STREAM else cin
    got, ref = [], []
    for it in range(steps):
        batch = O.synthetic_batch(B, xc, H, W, seed=100 + it)
        opt.zero_grad()
        loss = crit(net(batch["x_t1"].to(dev), batch["x_t2"].to(dev)), batch["y_change"].to(dev))
        loss.backward()
        opt.step()
        got.append(loss.item())
        r = O.train_step(mtype, sd, batch, kind=kind, q=True)
        for n, p in zip(ref_names, ref_params):
            p.grad = r["grads"][n]
        ref_opt.step()
        ref_opt.zero_grad()
        ref.append(r["loss"].item())
    torch.cuda.synchronize()
    sd_after = {k[len("module."):]: v.detach().cpu() for k, v in net.state_dict().items()}
    # weights only: Adam turns every gradient into a step of ~lr regardless of its magnitude, so parameters that start
    # at zero (BatchNorm / conv biases) follow the sign of noise-level gradients and are not comparable element-wise
    prm = max(((sd_after[n] - p.detach()).norm() / p.detach().norm().clamp_min(1e-12)).item()
              for n, p in zip(ref_names, ref_params) if p.dim() >= 2)
    net.module.release_engines()
    return {"loss_cuda": got, "loss_ref": ref, "max_loss_diff": max(abs(a - b) for a, b in zip(got, ref)),
            "loss_moved": abs(ref[0] - ref[-1]), "max_param_rel": prm}


def run_baseline_size_properties(mtype: str = "dualstreamunet", cin: int = 6, B: int = 16, H: int = 256, W: int = 256) -> dict:
    """BASELINE config #2 at its full size (16 patch pairs of 256 x 256, topology [64,128,256,512]) — too large for the
    CPU oracle inside a unit test, so the step is checked through properties that do not need one:
      * determinism: the same step from the same state twice (eager, then CUDA-graph replay) gives bit-identical loss,
        logits and gradients (split-K partials and BatchNorm sums are reduced in a fixed order, no float atomics);
      * loss consistency: the fused loss equals the power-Jaccard formula evaluated in fp64 on the step's own logits;
      * finite, non-trivial gradients for every parameter that has one; exact zeros on pre-BN conv biases."""
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    batch = O.synthetic_batch(B, 6 if mtype in TWO_STREAM else cin, H, W, seed=7)
    gb = {k: v.to(dev) for k, v in batch.items() if k != "is_labeled"}
    ts = TrainStep(net.module, B, H, W, kind="supervised", device=dev, dp_group=None)
    runs = []
    for it in range(3):                      # run 0 eager, runs 1-2 capture + replay the graphs
        net.load_state_dict(sd0)
        loss = ts(gb["x_t1"], gb["x_t2"], y_change=gb["y_change"])
        runs.append((loss.item(), ts.eng.output_tensors()[0].detach().clone(), ts.eng.grads.flat.detach().clone()))
    torch.cuda.synchronize()
    from multimodal_siamese_cd_b200 import ops
    ops.device_status(0)
    res = {"loss": runs[0][0]}
    res["deterministic"] = all(runs[i][0] == runs[0][0] and torch.equal(runs[i][1], runs[0][1]) and
                               torch.equal(runs[i][2], runs[0][2]) for i in (1, 2))
    z = runs[0][1].double().flatten()
    t = gb["y_change"].double().flatten()
    pr = torch.sigmoid(z)
    inter = (pr * t).sum()
    ref_loss = 1.0 - inter / ((pr * pr).sum() + (t * t).sum() - inter + 1e-6)
    res["loss_vs_formula"] = abs(ref_loss.item() - runs[0][0])
    g = ts.eng.grads
    res["all_finite"] = bool(torch.isfinite(g.flat).all())
    res["zero_grad_tensors"] = [n for n, _ in g.params if n not in g.skip and not is_prebn_bias(n)
                                and float(g.views[n].abs().max()) == 0.0]
    res["prebn_bias_grad_max"] = max(float(g.views[n].abs().max()) for n, _ in g.params if is_prebn_bias(n))
    net.module.release_engines()
    return res


def run_golden_train_case(fixture_path, precision: str = "fast") -> dict:
    """The drop-in step on the GPU against a fixture written by the UNMODIFIED reference (oracle/make_golden.py): same
    seed-7 default init, same synthetic batch; logits and loss compared directly with the reference's fp32 values."""
    fix = torch.load(fixture_path, weights_only=False)
    name, mtype, cin, topo, B, kind, H, W, alpha = fix["case"]
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    net.module.set_precision(precision)
    batch = O.synthetic_batch(B, 6 if mtype in TWO_STREAM else cin, H, W, seed=7)
    gb = {k: v.to(dev) for k, v in batch.items() if k != "is_labeled"}
    crit = loss_functions.get_criterion("PowerJaccardLoss")
    outs = net(gb["x_t1"], gb["x_t2"])
    if kind == "supervised":
        loss, out_list = crit(outs, gb["y_change"]), [outs]
    elif kind == "dualtask":
        c, s1, s2 = outs
        loss = (crit(c, gb["y_change"]) + (crit(s1, gb["y_sem_t1"]) + crit(s2, gb["y_sem_t2"])) / 2) / 2
        out_list = [c, s1, s2]
    else:
        f, s1, s2 = outs
        lab, y = batch["is_labeled"], gb["y_change"]
        unl = torch.logical_not(lab)
        loss = alpha * (crit(f[lab,], y[lab,]) + crit(s1[lab,], y[lab,]) + crit(s2[lab,], y[lab,])) / 3 + \
            (1 - alpha) * crit(s1[unl,], torch.sigmoid(s2)[unl,])
        out_list = [f, s1, s2]
    loss.backward()
    torch.cuda.synchronize()
    res = {"loss_diff": abs(loss.item() - fix["loss"].item()),
           "logits_rel": max(rel(o, g) for o, g in zip(out_list, fix["outs"]))}
    g_rel = []
    worst = 0.0
    for k, p in net.named_parameters():
        ref = fix["grads"][k]
        if ref is None:
            assert p.grad is None, k
            continue
        if is_prebn_bias(k[len("module."):] if k.startswith("module.") else k):
            continue                                   # analytically zero; the reference carries ~1e-9 noise
        got = p.grad.detach().double().flatten().cpu()
        g_rel.append((abs(got.norm().item() - ref[0].item()), ref[0].item()))
        # fingerprint = (L2 norm, sum, first 4 values) of the reference's gradient: per-tensor checks, so that a wrong
        # SMALL tensor (a transposed-conv bias, a deep BatchNorm beta) cannot hide behind the large ones
        worst = max(worst, abs(got.norm().item() - ref[0].item()) / max(ref[0].item(), 1e-30),
                    (got[:4] - ref[2:2 + min(4, got.numel())]).norm().item() / max(ref[0].item(), 1e-30))
    # fingerprints hold the L2 norm of every gradient: relative error of the norms, norm-weighted
    res["grad_norm_rel"] = sum(d for d, _ in g_rel) / max(sum(n for _, n in g_rel), 1e-30)
    res["grad_fingerprint_worst"] = worst
    gm = (out_list[0].detach() > 0)
    res["popcount_diff"] = abs(int(gm.sum()) - fix["mask_f1"]["popcount"])
    res["mask_flips"] = int((gm.cpu() != (fix["outs"][0] > 0)).sum())
    res["n_pixels"] = gm.numel()
    net.module.release_engines()
    return res


def run_golden_eval_case(fixture_path, precision: str = "fast") -> dict:
    """Odd-sized inference on the GPU against a fixture written by the UNMODIFIED reference
    (oracle/make_eval_golden.py)."""
    fix = torch.load(fixture_path, weights_only=False)
    name, mtype, cin, topo, B, H, W, Hw, Ww = fix["case"]
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev)
    net.module.set_precision(precision)
    xc = 6 if mtype in TWO_STREAM else cin
    warm = O.synthetic_batch(B, xc, Hw, Ww, seed=11)
    batch = O.synthetic_batch(B, xc, H, W, seed=7)
    net.train()
    with torch.no_grad():
        net(warm["x_t1"].to(dev), warm["x_t2"].to(dev))
    net.eval()
    with torch.no_grad():
        out = net(batch["x_t1"].to(dev), batch["x_t2"].to(dev))
    torch.cuda.synchronize()
    ref = fix["logits"]
    res = {"logits_rel": rel(out, ref)}
    flips = ((out.cpu() > 0) != (ref > 0)) & (ref.abs() >= 0.05)
    res["margin_flips"] = int(flips.sum())
    res["margin3_flips"] = int((((out.cpu() > 0) != (ref > 0)) & (ref.abs() >= 1e-3)).sum())
    _, f1 = O.change_mask_f1(out.cpu(), batch["y_change"])
    res["f1_diff"] = abs(float(f1) - fix["f1"])
    net.module.release_engines()
    return res

