"""TEST-ONLY stand-in for `rasterio` (not installed in this image; imported by the reference's utils/geofiles.py).
Injected through PYTHONPATH by tests/test_reference_scripts.py only. `open(file).read()` serves a deterministic
synthetic raster keyed by the file name, shaped like the SpaceNet-7 derived files the reference's dataset reads
(utils/datasets.py:29-45): s1_* 2 bands, s2_* 4 bands, buildings_* 1 band; t1 / t2 of one site are correlated."""
import os
import zlib
from pathlib import Path

import numpy as np

TILE = int(os.environ.get("B200CD_STUB_TILE", "272"))


class _Dataset:
    transform = None
    crs = None

    def __init__(self, path):
        self.name = Path(str(path)).name

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def read(self):
        stem = self.name.rsplit(".", 1)[0]
        kind, rest = stem.split("_", 1)
        site = rest.rsplit("_", 2)[0]                      # <aoi>_<year>_<month>
        base = np.random.default_rng(zlib.crc32(f"{kind}:{site}".encode()))
        noise = np.random.default_rng(zlib.crc32(stem.encode()))
        if kind in ("s1", "s2"):
            bands = 2 if kind == "s1" else 4
            img = base.random((bands, TILE, TILE), dtype=np.float32) + 0.1 * noise.random((bands, TILE, TILE), dtype=np.float32)
            return np.clip(img, 0, 1).astype(np.float32)
        # building footprints: 16x16 blocks, more of them at later dates
        coarse = base.random((1, (TILE + 15) // 16, (TILE + 15) // 16))
        year, month = (int(v) for v in rest.rsplit("_", 2)[1:])
        thr = 0.9 - 0.02 * ((year - 2018) * 12 + month)
        return np.kron(coarse < 1 - thr, np.ones((1, 16, 16)))[:, :TILE, :TILE].astype(np.float32)


def open(file, *args, **kwargs):  # noqa: A001
    return _Dataset(file)
