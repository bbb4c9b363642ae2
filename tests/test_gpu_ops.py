"""GPU: every kernel through the C ABI against torch fp32 on identical (bf16-rounded) inputs. Tolerances live next to
each check in tests/gpu_checks.py: bf16-output kernels within one bf16 rounding of the fp32 result (rel L2 < 6e-3,
and identical to the rounded reference on all but ~1e-3 of the elements), fp32-output kernels < 2e-3 ... 1e-5."""
import pytest
import torch

import gpu_checks as gc
import gpu_checks_hp as gch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gc.ALL_CHECKS))
def test_kernel(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    res = gc.ALL_CHECKS[name]()
    assert res["ok"], res
    if "ulp_frac" in res:
        assert res["ulp_frac"] < res.get("ulp_tol", 5e-3) and res["rel_l2_rounded"] < 5e-4, res


@pytest.mark.parametrize("name", list(gch.HP_CHECKS))
def test_kernel_split_bf16(name):
    """The "precise" (split-bf16, ABI version 2) kernels against torch fp64 on the values the split tensors hold."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    res = gch.HP_CHECKS[name]()
    assert res["ok"], res


@pytest.mark.parametrize("variant", range(4))
def test_gpu_augmenter_matches_reference_transforms(variant):
    """data.GpuAugmenter (b200cd_augment: crop + flips + rot90 + colour shift + gamma + channel regrouping + CHW packing,
    one launch per batch) against the numpy pipeline of utils/augmentations.py as restated in oracle/augment_oracle.py
    (pinned bit-exactly against the unmodified reference classes by tests/test_augment_cpu.py), same numpy seed.
    Geometry is exact; the photometric ops are fp32 on the device vs float64 in numpy (<= 2e-6)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np

    from multimodal_siamese_cd_b200.data import GpuAugmenter
    from oracle import augment_oracle
    from test_augment_cpu import VARIANTS, aug_cfg, sample
    dev = torch.device("cuda", 0)
    cfg = aug_cfg(**VARIANTS[variant])
    aug = GpuAugmenter(cfg, dev)
    host = [sample(seed, H=50 + 3 * seed, W=61 - 2 * seed) for seed in range(5)]    # tiles of different sizes
    want, params = [], []
    for i, (imgs, bld, chg) in enumerate(host):
        np.random.seed(300 + i)
        want.append(augment_oracle.transform(cfg, imgs, bld, chg))
        np.random.seed(300 + i)
        params.append(aug.draw(chg))
    x1, x2, y, s = aug.apply([torch.from_numpy(h[0]).to(dev) for h in host], [torch.from_numpy(h[1]).to(dev) for h in host],
                             [torch.from_numpy(h[2]).to(dev) for h in host], params)
    torch.cuda.synchronize()
    photometric = cfg.AUGMENTATION.COLOR_SHIFT or cfg.AUGMENTATION.GAMMA_CORRECTION
    for i, (wi, wb, wc) in enumerate(want):
        # x_t1 = [S1 t1 | S2 t1], x_t2 = [S1 t2 | S2 t2] (utils/datasets.py:151-162)
        w1 = np.concatenate([wi[0:2], wi[4:8]], 0)
        w2 = np.concatenate([wi[2:4], wi[8:12]], 0)
        for got, ref in ((x1[i], w1), (x2[i], w2), (s[i], wb)):
            d = np.abs(got.cpu().numpy() - ref).max()
            assert d <= (2e-6 if photometric else 0.0), (variant, i, d)
        assert np.array_equal(y[i].cpu().numpy(), wc)
    with pytest.raises(RuntimeError):
        GpuAugmenter(cfg, torch.device("cpu"))


def test_device_prefetcher_order_and_contents():
    """data.DevicePrefetcher yields every batch once, in order, with the host contents, while reusing two buffer sets."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multimodal_siamese_cd_b200.data import DevicePrefetcher
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    host = [{"x_t1": torch.rand(2, 6, 32, 32, generator=g).pin_memory(), "x_t2": torch.rand(2, 6, 32, 32, generator=g),
             "y_change": (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float(), "is_labeled": torch.tensor([True, False]),
             "aoi_id": f"aoi{i}"} for i in range(7)]
    pf = DevicePrefetcher(iter(host), dev)
    seen = 0
    acc = []
    for i, b in enumerate(pf):
        assert b["x_t1"].is_cuda and b["x_t2"].is_cuda and b["y_change"].is_cuda
        assert not b["is_labeled"].is_cuda and b["aoi_id"] == f"aoi{i}"
        acc.append((b["x_t1"] * 2).sum())          # queue work on the batch before the next one is staged
        assert torch.equal(b["x_t1"].cpu(), host[i]["x_t1"]) and torch.equal(b["y_change"].cpu(), host[i]["y_change"])
        seen += 1
    assert seen == 7
    for i, a in enumerate(acc):
        assert abs(a.item() - 2 * host[i]["x_t1"].sum().item()) < 1e-2
    assert pf.bytes_staged == sum(v.numel() * 4 for h in host for k, v in h.items() if k in ("x_t1", "x_t2", "y_change"))
    with pytest.raises(RuntimeError):
        DevicePrefetcher(iter(host), torch.device("cpu"))


def test_fused_adamw_matches_torch():
    """optim.FusedAdamW (one kernel over all parameters) against torch.optim.AdamW on the same parameters and
    gradients for several steps, including a parameter without gradient and a state_dict round trip."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multimodal_siamese_cd_b200.optim import FusedAdamW
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    shapes = [(64, 6, 3, 3), (64,), (128, 64, 3, 3), (1, 128, 1, 1), (1,), (37,), (512, 512, 2, 2)]
    ours = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    frozen_o, frozen_r = torch.nn.Parameter(torch.ones(3, device=dev)), torch.nn.Parameter(torch.ones(3, device=dev))
    o1 = FusedAdamW(ours + [frozen_o], lr=1e-2, weight_decay=0.01)
    o2 = torch.optim.AdamW(ref + [frozen_r], lr=1e-2, weight_decay=0.01)
    for step in range(5):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, device=dev, generator=g) * (0.1 + step)
            a.grad = gr.clone()
            b.grad = gr.clone()
        o1.step()
        o2.step()
        if step == 2:   # checkpoint round trip through torch's format (utils/networks.py:30-56)
            sd = o1.state_dict()
            o1 = FusedAdamW(ours + [frozen_o], lr=1e-2, weight_decay=0.01)
            o1.load_state_dict(sd)
    torch.cuda.synchronize()
    assert torch.equal(frozen_o, frozen_r) and len(o1.state[frozen_o]) == 0
    for a, b in zip(ours, ref):
        rel = ((a - b).norm() / b.norm()).item()
        assert rel < 2e-6, (tuple(a.shape), rel)
        sa, sb = o1.state[a], o2.state[b]
        assert float(sa["step"]) == float(sb["step"]) == 5.0
        assert ((sa["exp_avg"] - sb["exp_avg"]).norm() / sb["exp_avg"].norm()).item() < 2e-6
        assert ((sa["exp_avg_sq"] - sb["exp_avg_sq"]).norm() / sb["exp_avg_sq"].norm()).item() < 2e-6
    cpu_p = torch.nn.Parameter(torch.ones(3))
    cpu_p.grad = torch.ones(3)
    with pytest.raises(Exception, match="CUDA parameters only"):
        FusedAdamW([cpu_p], lr=1e-3).step()


def _metric_ref(y_true, y_pred, thr):
    """utils/metrics.py:23-31 restated with plain torch ops (reference naming of FP / FN included)."""
    yt = y_true.bool()[None, ...]
    off = (y_pred[None, ...] - thr[:, None, None, None, None] + 0.5).round().bool()
    dims = (-1, -2, -3, -4)
    return torch.stack([(yt & off).sum(dims), (~yt & ~off).sum(dims), (yt & ~off).sum(dims), (~yt & off).sum(dims)], 1)


def test_confusion_counts_match_reference_metric():
    """metrics.MultiThresholdMetric (one kernel) against the restated reference metric on random data with several
    thresholds and exact ties, and against the TP / F1 the UNMODIFIED reference produced for the golden fixtures."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pathlib import Path

    from multimodal_siamese_cd_b200.metrics import MultiThresholdMetric
    from oracle import unet_oracle as O
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    thr = torch.tensor([0.5, 0.25, 0.75, 0.9], device=dev)
    m = MultiThresholdMetric(thr)
    want = torch.zeros(4, 4, dtype=torch.int64, device=dev)
    for _ in range(3):
        p = torch.rand(3, 1, 64, 48, device=dev, generator=g)
        p.view(-1)[::7] = 0.5                      # exact ties: round-half-to-even makes them negative at thr 0.5
        p.view(-1)[::11] = 0.75
        y = (torch.rand(3, 1, 64, 48, device=dev, generator=g) > 0.7).float()
        m.add_sample(y, p)
        want += _metric_ref(y, p, thr)
    assert torch.equal(m._counts, want)
    assert torch.equal(torch.stack([m.TP, m.TN, m.FP, m.FN], 1), want.float())
    # logits path == sigmoid + probability path
    z = torch.randn(2, 1, 32, 32, device=dev, generator=g) * 3
    y = (torch.rand(2, 1, 32, 32, device=dev, generator=g) > 0.5).float()
    a, b = MultiThresholdMetric(thr[:1]), MultiThresholdMetric(thr[:1])
    a.add_logits(y, z)
    b.add_sample(y, torch.sigmoid(z))
    assert torch.equal(a._counts, b._counts)
    # golden fixtures: TP and F1 computed by the reference's own utils/metrics.py on the reference's logits
    for f in sorted((Path(__file__).parent / "golden").glob("*.pt")):
        fix = torch.load(f, map_location="cpu", weights_only=False)   # our own fixtures (tuples, dicts, tensors)
        name, mtype, cin, topo, B, kind, H, W, alpha = fix["case"]
        xc = 6 if mtype in ("dualstreamunet", "whatevernet", "whatevernet2") else cin
        batch = O.synthetic_batch(B, xc, H, W, seed=7)
        mm = MultiThresholdMetric(torch.tensor([0.5], device=dev))
        mm.add_logits(batch["y_change"].to(dev), fix["outs"][0].to(dev))
        assert float(mm.TP.item()) == fix["mask_f1"]["tp"], name
        assert abs(float(mm.compute_f1().item()) - fix["mask_f1"]["f1"]) < 1e-6, name
    with pytest.raises(RuntimeError):
        MultiThresholdMetric(torch.tensor([0.5]))
