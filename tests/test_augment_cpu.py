"""CPU: the augmentation oracle (oracle/augment_oracle.py) against the UNMODIFIED reference transforms
(utils/augmentations.py) under the same numpy seed, and the host half of data.GpuAugmenter (`draw`) against the oracle's
random decisions. The device half is checked in tests/test_gpu_ops.py."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from multimodal_siamese_cd_b200.config import new_config  # noqa: E402
from oracle import augment_oracle, stage_reference  # noqa: E402


def aug_cfg(crop=32, importance=True, flip=True, rot=True, color=False, gamma=False):
    from multimodal_siamese_cd_b200.config import CfgNode
    cfg = new_config()
    cfg.AUGMENTATION = CfgNode()
    cfg.AUGMENTATION.CROP_SIZE = crop
    cfg.AUGMENTATION.IMAGE_OVERSAMPLING_TYPE = "importance" if importance else "none"
    cfg.AUGMENTATION.RANDOM_FLIP = flip
    cfg.AUGMENTATION.RANDOM_ROTATE = rot
    cfg.AUGMENTATION.COLOR_SHIFT = color
    cfg.AUGMENTATION.GAMMA_CORRECTION = gamma
    cfg.DATALOADER.S1_BANDS = [0, 1]
    cfg.DATALOADER.S2_BANDS = [2, 1, 0, 3]
    cfg.DATALOADER.INPUT_MODE = "s1s2"
    return cfg


def sample(seed, H=50, W=61):
    r = np.random.RandomState(seed)
    return (r.rand(H, W, 12).astype(np.float32), (r.rand(H, W, 2) > 0.7).astype(np.float32),
            (r.rand(H, W, 1) > 0.9).astype(np.float32))


VARIANTS = [dict(), dict(importance=False), dict(color=True, gamma=True), dict(flip=False, rot=False, gamma=True)]


@pytest.mark.parametrize("kw", VARIANTS)
def test_oracle_matches_reference_transforms(kw):
    ref = stage_reference.staged_root()
    if ref is None:
        pytest.skip("no reference tree")
    sys.path.insert(0, str(ref))
    try:
        import importlib
        for name in ("utils", "utils.augmentations"):
            m = sys.modules.get(name)
            if m is not None and not str(getattr(m, "__file__", None) or getattr(m, "__path__", [""])[0]).startswith(str(ref)):
                del sys.modules[name]
        ref_aug = importlib.import_module("utils.augmentations")
    finally:
        sys.path.remove(str(ref))
    cfg = aug_cfg(**kw)
    tf = ref_aug.compose_transformations(cfg, no_augmentations=False)
    for seed in range(4):
        imgs, bld, chg = sample(seed)
        np.random.seed(100 + seed)
        want = [t.numpy() for t in tf((imgs, bld, chg))]
        np.random.seed(100 + seed)
        got = augment_oracle.transform(cfg, imgs, bld, chg)
        for a, b in zip(got, want):
            assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("kw", VARIANTS)
def test_draw_replays_the_oracles_random_decisions(kw):
    """GpuAugmenter.draw consumes numpy's RNG exactly like the transforms do: after drawing, the generator state is the
    same as after the oracle's transform, and the drawn crop / flips / rotation reproduce its geometry."""
    import torch

    from multimodal_siamese_cd_b200.data import GpuAugmenter
    cfg = aug_cfg(**kw)
    aug = GpuAugmenter.__new__(GpuAugmenter)          # host half only: no device needed
    aug.device = torch.device("cpu")
    a = cfg.AUGMENTATION
    aug.crop, aug.importance = a.CROP_SIZE, a.IMAGE_OVERSAMPLING_TYPE != "none"
    aug.flip, aug.rotate, aug.color, aug.gamma, aug.c_img = a.RANDOM_FLIP, a.RANDOM_ROTATE, a.COLOR_SHIFT, a.GAMMA_CORRECTION, 12
    for seed in range(4):
        imgs, bld, chg = sample(seed)
        np.random.seed(7 + seed)
        want = augment_oracle.transform(cfg, imgs, bld, chg)
        state_after = np.random.get_state()[1].copy()
        np.random.seed(7 + seed)
        p = aug.draw(chg)
        assert np.array_equal(np.random.get_state()[1], state_after)
        cs = aug.crop
        t = chg[p["y0"]:p["y0"] + cs, p["x0"]:p["x0"] + cs]
        if p["hflip"]:
            t = np.flip(t, axis=1)
        if p["vflip"]:
            t = np.flip(t, axis=0)
        t = np.rot90(t, p["rotk"], axes=(0, 1))
        assert np.array_equal(t.transpose(2, 0, 1), want[2])
