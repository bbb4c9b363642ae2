"""ORACLE — TEST INFRASTRUCTURE ONLY. numpy restatement of the reference's training-time augmentation pipeline
(utils/augmentations.py:6-142 as composed by compose_transformations :6-32 and applied by
utils/datasets.py:149-162), written from its behaviour. tests/test_augment_cpu.py pins it against the UNMODIFIED
reference classes (when a reference tree is available) under the same numpy seed; tests/test_gpu_ops.py compares the
CUDA kernel (b200cd_augment, data.GpuAugmenter) with it."""
from __future__ import annotations

import numpy as np


def transform(cfg, imgs: np.ndarray, buildings: np.ndarray, change: np.ndarray):
    """(imgs [H,W,C], buildings [H,W,2], change [H,W,1]) -> CHW float32 arrays, drawing from numpy's global RNG in the
    reference's order: 20 candidate crops (or one), the importance choice, two flips, the rotation count, the colour
    factors of the first and second tuple member, the gamma exponents of the first and second tuple member."""
    a = cfg.AUGMENTATION
    cs = a.CROP_SIZE
    H, W = change.shape[:2]

    def random_crop():                                           # UniformCrop.random_crop :109-120
        x = np.random.randint(0, W - cs)
        y = np.random.randint(0, H - cs)
        return x, y

    if a.IMAGE_OVERSAMPLING_TYPE == "none":
        x, y = random_crop()
    else:                                                        # ImportanceRandomCrop.__call__ :128-142
        crops = [random_crop() for _ in range(20)]
        w = np.array([change[yy:yy + cs, xx:xx + cs, ].sum() for xx, yy in crops]) + 5
        w = w / w.sum()
        x, y = crops[np.random.choice(20, p=w)]
    t = [arr[y:y + cs, x:x + cs, ] for arr in (imgs, buildings, change)]
    if a.RANDOM_FLIP:                                            # RandomFlip :44-62
        hf = np.random.choice([True, False])
        vf = np.random.choice([True, False])
        if hf:
            t = [np.flip(v, axis=1) for v in t]
        if vf:
            t = [np.flip(v, axis=0) for v in t]
    if a.RANDOM_ROTATE:                                          # RandomRotate :65-72
        k = np.random.randint(1, 4)
        t = [np.rot90(v, k, axes=(0, 1)) for v in t]
    if a.COLOR_SHIFT:                                            # ColorShift :75-86 (first two tuple members)
        for i in (0, 1):
            f = np.random.uniform(0.5, 1.5, t[i].shape[-1])
            t[i] = np.clip(t[i] * f[np.newaxis, np.newaxis, :], 0, 1).astype(np.float32)
    if a.GAMMA_CORRECTION:                                       # GammaCorrection :89-101
        for i in (0, 1):
            g = np.random.uniform(0.25, 2, t[i].shape[-1])
            t[i] = np.clip(np.power(t[i], g[np.newaxis, np.newaxis, :]), 0, 1).astype(np.float32)
    return tuple(np.ascontiguousarray(v.transpose(2, 0, 1)).astype(np.float32) for v in t)   # Numpy2Torch :35-41
