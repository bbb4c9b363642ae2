"""CPU: pins oracle/unet_oracle.py (and the drop-in constructors' default-init order) against outputs of the
UNMODIFIED reference stored in tests/golden/ (made by oracle/make_golden.py in the build container)."""
from __future__ import annotations

from pathlib import Path

import pytest
import torch

from multimodal_siamese_cd_b200 import networks
from multimodal_siamese_cd_b200.config import synthetic_cfg
from oracle import unet_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLD.glob("*.pt"))


def fingerprint(t):
    t = t.detach().double().flatten()
    head = torch.zeros(4, dtype=torch.float64)
    head[: min(4, t.numel())] = t[:4]
    return torch.cat([torch.stack([t.norm(), t.sum()]), head])


def build(case):
    name, mtype, cin, topo, B, kind, H, W, alpha = case
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg)
    return cfg, net


@pytest.mark.parametrize("name", CASES)
def test_default_init_matches_reference(name):
    """Same constructor order => same RNG consumption => bit-identical initial weights and state_dict keys."""
    fix = torch.load(GOLD / f"{name}.pt", weights_only=False)
    cfg, net = build(fix["case"])
    sd = net.state_dict()
    ref_keys = set(fix["init"]) | {k for k in fix["bn"] if k.endswith("num_batches_tracked")}
    assert set(sd.keys()) == ref_keys
    for k, fp in fix["init"].items():
        got = fingerprint(sd[k])
        assert torch.equal(got[2:], fp[2:]), k                       # leading values: bit-identical draws
        assert torch.allclose(got[:2], fp[:2], rtol=1e-10, atol=1e-12), k   # norm / sum (reduction order may differ)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    fix = torch.load(GOLD / f"{name}.pt", weights_only=False)
    cname, mtype, cin, topo, B, kind, H, W, alpha = fix["case"]
    cfg, net = build(fix["case"])
    sd = O.clone_state(net.state_dict())
    xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
    batch = O.synthetic_batch(B, xc, H, W, seed=7)
    torch.set_num_threads(8)
    res = O.train_step(mtype, sd, batch, kind=kind, alpha=alpha, q=False)
    outs = res["outs"] if isinstance(res["outs"], tuple) else (res["outs"],)
    assert len(outs) == len(fix["outs"])
    for o, g in zip(outs, fix["outs"]):
        assert (o.detach() - g).abs().max().item() <= 1e-5
    assert abs(res["loss"].item() - fix["loss"].item()) <= 1e-6
    for k, gfp in fix["grads"].items():
        k2 = k[len("module."):]
        got = res["grads"][k2]
        if gfp is None:
            assert got is None, k            # outc_sem_change: never used => grad None
            continue
        fp = fingerprint(got)
        if k.endswith((".conv.0.bias", ".conv.3.bias")):
            # a conv bias feeding train-mode BatchNorm has a mathematically zero gradient; the reference yields
            # summation noise (|g| <= 6e-9 measured): compare with an absolute bound instead
            assert fp[0].item() <= 1e-6 and gfp[0].item() <= 1e-6, k
            continue
        scale = max(gfp[0].item(), 1e-7)
        # L2 norm and leading values; fp32 summation order is identical (same torch kernels), so this is tight
        assert abs(fp[0] - gfp[0]).item() <= 1e-4 * scale + 1e-9, (k, fp[0].item(), gfp[0].item())
        assert (fp[2:] - gfp[2:]).abs().max().item() <= 1e-4 * scale + 1e-8, k
    for k, v in fix["bn"].items():
        k2 = k[len("module."):]
        if k.endswith("num_batches_tracked"):
            assert int(sd[k2]) == int(v), k   # shared-weight BNs are updated twice per step (SURVEY §0 finding 1)
        else:
            assert (fingerprint(sd[k2]) - v).abs().max().item() <= 1e-4 * max(v[0].item(), 1.0), k
    pred, f1 = O.change_mask_f1(outs[0].detach(), batch["y_change"])
    assert int(pred.sum()) == fix["mask_f1"]["popcount"]
    assert abs(f1.item() - fix["mask_f1"]["f1"]) <= 1e-6


def test_known_answers():
    """SURVEY §8c known answers for power_jaccard_loss."""
    z = torch.zeros(2, 1, 8, 8)
    assert abs(O.power_jaccard_loss(z, torch.ones_like(z)).item() - 1 / 3) < 1e-6
    zr = torch.randn(2, 1, 8, 8, requires_grad=True)
    loss = O.power_jaccard_loss(zr, torch.zeros_like(zr))
    loss.backward()
    assert loss.item() == 1.0 and zr.grad.abs().max().item() == 0.0
    big = torch.full((1, 1, 4, 4), 40.0)
    assert O.power_jaccard_loss(big, torch.ones_like(big)).item() < 1e-6


def test_quantised_mode_is_close_to_exact():
    """q=True only adds bf16 storage rounding: logits stay within a few 1e-2 relative of the exact oracle."""
    cfg = synthetic_cfg("siameseunet", in_channels=4, topology=(64, 128))
    torch.manual_seed(7)
    net = networks.create_network(cfg)
    batch = O.synthetic_batch(2, 4, 32, 32, seed=7)
    a = O.train_step("siameseunet", O.clone_state(net.state_dict()), batch, q=False)
    b = O.train_step("siameseunet", O.clone_state(net.state_dict()), batch, q=True)
    rel = (a["outs"] - b["outs"]).norm() / a["outs"].norm()
    assert rel.item() < 5e-2
    assert abs(a["loss"].item() - b["loss"].item()) < 1e-3


@pytest.mark.parametrize("mtype,cin", [("unet", 6), ("siameseunet", 4), ("dualstreamunet", 6), ("dtsiameseunet", 6),
                                       ("whatevernet", 6), ("whatevernet2", 6)])
def test_oracle_init_equals_dropin_init(mtype, cin):
    """oracle.reference_state_dict and the drop-in constructors draw the same initial weights (both pinned to the
    reference by the golden init fingerprints above)."""
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=(64, 128))
    torch.manual_seed(7)
    sd_net = {k[len("module."):]: v for k, v in networks.create_network(cfg).state_dict().items()}
    sd_or = O.reference_state_dict(mtype, in_channels=cin, topology=(64, 128), seed=7)
    assert set(sd_net) == set(sd_or)
    for k in sd_net:
        assert torch.equal(sd_net[k], sd_or[k]), k


EVAL_GOLD = GOLD / "eval"
EVAL_CASES = sorted(p.stem for p in EVAL_GOLD.glob("*.pt"))


@pytest.mark.parametrize("name", EVAL_CASES)
def test_oracle_eval_odd_tiles_match_reference(name):
    """Inference on tiles with odd-sized levels (utils/evaluation.py:9-23; MaxPool2d floor utils/networks.py:420, Up
    centre pad :440-443): the oracle's eval-mode logits, mask and F1 against the unmodified reference
    (oracle/make_eval_golden.py)."""
    fix = torch.load(EVAL_GOLD / f"{name}.pt", weights_only=False)
    cname, mtype, cin, topo, B, H, W, Hw, Ww = fix["case"]
    cfg = synthetic_cfg(mtype, in_channels=cin, topology=topo)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg)
    sd = O.clone_state(net.state_dict(), requires_grad=False)
    xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
    warm = O.synthetic_batch(B, xc, Hw, Ww, seed=11)
    batch = O.synthetic_batch(B, xc, H, W, seed=7)
    torch.set_num_threads(8)
    with torch.no_grad():
        O.forward(mtype, sd, warm["x_t1"], warm["x_t2"], train=True, q=False)      # moves the running statistics
        logits = O.forward(mtype, sd, batch["x_t1"], batch["x_t2"], train=False, q=False)
    assert torch.is_tensor(logits) and tuple(logits.shape) == (B, 1, H, W)
    assert (logits - fix["logits"]).abs().max().item() <= 1e-5
    mask, f1 = O.change_mask_f1(logits, batch["y_change"])
    assert int(mask.sum()) == fix["popcount"] and abs(float(f1) - fix["f1"]) <= 1e-6

