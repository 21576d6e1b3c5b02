"""Deterministic synthetic WorldCover / HYSOGs rasters for GCN10 blocks.

The real inputs (ESA WorldCover 10 m tiles behind /vsicurl, the HYSOGs250m GeoTIFF;
/root/reference/landcover/esa_worldcover_2021.vrt, /root/reference/hsg/readme.txt) need the
network, so tests and the benchmark synthesise rasters with the same value alphabets
(SURVEY.md section 8a row a9):

* land cover: ESA classes 10..100 (+95) and nodata 0, laid out in spatially coherent,
  irregular patches a few hundred pixels across with a sprinkle of small inclusions;
* HSG: 1..4 (A..D), dual groups 11..14 (A/D..D/D) and nodata (255, 0) in patches of a few
  250 m cells.

Every value is a pure integer hash of (seed, x, y), evaluated identically by numpy (CPU,
small cases, oracle inputs) and by torch on the GPU (full 36000 x 36000 tiles, where numpy
would take minutes), so the same seed gives byte-identical rasters on both.  Generation is
done in row bands to bound temporary memory.

Profiles: ``worldcover`` (default, coherent), ``random`` (i.i.d. per pixel: the cache- and
bank-unfriendly adversarial case), ``coastal`` (BASELINE config 3: more than half of the land
cover is nodata 0 / water 80, >=30 % of HSG cells are dual groups, >=20 % are nodata 255).
"""
from __future__ import annotations

import numpy as np

ESA_ALPHABET = (0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 100)

_M32 = 0xFFFFFFFF


def _mix(h):
    """32-bit integer finaliser on int64 arrays/tensors (values stay below 2**62)."""
    h = h & _M32
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & _M32
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & _M32
    h = h ^ (h >> 15)
    return h


def _hash2(a, b, seed):
    return _mix(a * 0x9E3779B1 + b * 0x85EBCA77 + (int(seed) & _M32) * 0x27D4EB2F + 0x165667B1)


class _NP:
    int64 = np.int64

    @staticmethod
    def arange(n):
        return np.arange(n, dtype=np.int64)

    @staticmethod
    def table(values):
        return np.asarray(values, dtype=np.uint8)

    @staticmethod
    def empty_u8(h, w):
        return np.empty((h, w), dtype=np.uint8)


class _Torch:
    def __init__(self, device):
        import torch
        self.torch = torch
        self.device = device

    def arange(self, n):
        return self.torch.arange(n, dtype=self.torch.int64, device=self.device)

    def table(self, values):
        return self.torch.tensor(list(values), dtype=self.torch.uint8, device=self.device)

    def empty_u8(self, h, w):
        return self.torch.empty((h, w), dtype=self.torch.uint8, device=self.device)


def _backend(device):
    return _NP() if device is None else _Torch(device)


def _esa_alphabet(profile):
    if profile == "coastal":
        # 16 slots: 0 (nodata) x5, 80 (water) x5, the rest land -> > 60 % masked / water
        return (0, 0, 0, 0, 0, 80, 80, 80, 80, 80, 10, 30, 40, 60, 90, 95)
    # 16 slots so that the modulo is unbiased; common classes repeated
    return (0, 10, 10, 20, 30, 30, 40, 40, 50, 60, 70, 80, 90, 95, 100, 100)


def esa_tile(w: int, h: int, seed: int, profile: str = "worldcover", device=None,
             patch: int = 192, band_rows: int = 1024, out=None):
    """uint8 [h, w] land-cover tile.  ``device=None`` -> numpy, else a torch device."""
    B = _backend(device)
    alpha = B.table(_esa_alphabet(profile))
    res = B.empty_u8(h, w) if out is None else out
    x = B.arange(w)
    # separable boundary wobble: column offset depends on the 16-row strip, row offset on the
    # 16-column strip, so patch edges wander by up to +-patch/4 and cross 16-px vector borders
    wob = max(patch // 2, 1)
    big = 1 << 20
    wx = _hash2(x >> 4, x * 0 + 7, seed) % wob                   # per column, added to y
    x_fine = x // 6
    for y0 in range(0, h, band_rows):
        y1 = min(h, y0 + band_rows)
        y = B.arange(y1 - y0) + y0
        if profile == "random":
            hh = _hash2(x[None, :], y[:, None], seed)
            res[y0:y1] = alpha[hh & 15]
            continue
        wy = _hash2(y >> 4, y * 0 + 3, seed) % wob               # per row, added to x
        cx = (x[None, :] + wy[:, None] + big) // patch
        cy = (y[:, None] + wx[None, :] + big) // patch
        hh = _hash2(cx, cy, seed)
        # small inclusions: 1 in 16 of the 6x6 px cells takes its own class
        hf = _hash2(x_fine[None, :], (y // 6)[:, None], seed ^ 0x5bd1e995)
        hh = hh ^ ((hf & 15) == 0) * (hf >> 8)
        res[y0:y1] = alpha[hh & 15]
    return res


def hsg_tile(w: int, h: int, seed: int, profile: str = "worldcover", device=None, patch: int = 9):
    """uint8 [h, w] hydrologic-soil-group window (250 m cells)."""
    B = _backend(device)
    if profile == "coastal":
        # 16 slots: 5 dual (>=30 %), 4 nodata 255 (>=20 %), 1 nodata 0, 6 single groups
        alpha = (11, 12, 13, 14, 11, 255, 255, 255, 255, 0, 1, 2, 3, 4, 4, 2)
    else:
        alpha = (1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 11, 12, 13, 14, 255, 0)
    alpha = B.table(alpha)
    x = B.arange(w)
    y = B.arange(h)
    if profile == "random":
        hh = _hash2(x[None, :], y[:, None], seed ^ 0x1234567)
    else:
        wob = max(patch // 2, 1)
        wy = _hash2(y >> 1, y * 0 + 11, seed) % wob
        wx = _hash2(x >> 1, x * 0 + 13, seed) % wob
        cx = (x[None, :] + wy[:, None]) // patch
        cy = (y[:, None] + wx[None, :]) // patch
        hh = _hash2(cx, cy, seed ^ 0x1234567)
    return alpha[hh & 15]


def block_geometry(lon0: float, lat0: float, w: int = 36000, h: int = 36000,
                   px: float = 1.0 / 12000.0, hsg_px: float = 1.0 / 480.0,
                   hsg_origin_shift=(0.0, 0.0), margin_cells: int = 0):
    """Geotransforms of a block window whose NW corner is (lon0, lat0).

    Returns (gt, soil_gt, hsx, hsy): the ESA window geotransform, the HSG window
    geotransform and the HSG window size that covers the block (what
    /root/reference/src/raster.c:126-162 would produce for a HSG raster whose grid is
    shifted by ``hsg_origin_shift`` degrees against the block corner).
    """
    gt = (lon0, px, 0.0, lat0, 0.0, -px)
    sx, sy = hsg_origin_shift
    s0 = lon0 - sx - margin_cells * hsg_px
    s3 = lat0 + sy + margin_cells * hsg_px
    soil_gt = (s0, hsg_px, 0.0, s3, 0.0, -hsg_px)
    import math
    hsx = int(math.ceil((lon0 + w * px - s0) / hsg_px)) + margin_cells
    hsy = int(math.ceil((s3 - (lat0 - h * px)) / hsg_px)) + margin_cells
    return gt, soil_gt, hsx, hsy
