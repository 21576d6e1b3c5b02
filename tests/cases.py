"""Shared block geometries and input builders for the parity tests."""
from __future__ import annotations

import numpy as np

from gcn10_b200 import synth

PX = 1.0 / 12000.0                      # synthetic tile pixel
PX_VRT = 8.3333333333330430e-05         # /root/reference/landcover/esa_worldcover_2021.vrt:3
HSG_PX = 1.0 / 480.0                    # 250 m HYSOGs grid in degrees (25:1)


def make_block(w, h, *, lon0=-114.0, lat0=42.0, px=PX, hsg_px=HSG_PX, shift=(0.0, 0.0), margin=0,
               seed=1, profile="worldcover", esa_patch=48, hsg_patch=3, hsx=None, hsy=None):
    """Returns dict(esa, gt, hsg, soil_gt) for a block window of w x h pixels."""
    gt, soil_gt, nx, ny = synth.block_geometry(lon0, lat0, w, h, px, hsg_px, shift, margin)
    hsx = nx if hsx is None else hsx
    hsy = ny if hsy is None else hsy
    esa = synth.esa_tile(w, h, seed, profile, patch=esa_patch)
    hsg = synth.hsg_tile(hsx, hsy, seed + 1000, profile, patch=hsg_patch)
    return dict(esa=esa, gt=gt, hsg=hsg, soil_gt=soil_gt)


# (id, kwargs) -- sizes cover: multiples of 16, ragged right edges, tiny blocks, > one 4096-px strip,
# more rows than one CTA chunk, the real-VRT pixel size, shifted HSG origins, non-25 ratios,
# HSG windows that are too small (clamp path of cn.c:228-229) and the three data profiles.
SMALL_CASES = [
    ("w16_h1", dict(w=16, h=1)),
    ("w1_h1", dict(w=1, h=1)),
    ("w15_h7", dict(w=15, h=7)),
    ("w17_h33", dict(w=17, h=33)),
    ("w640_h480", dict(w=640, h=480)),
    ("w1000_h300_shift", dict(w=1000, h=300, shift=(0.00071, 0.00113), margin=1)),
    ("w4096_h130", dict(w=4096, h=130)),
    ("w4111_h260", dict(w=4111, h=260, seed=3)),
    ("w9000_h64_vrtpx", dict(w=9000, h=64, px=PX_VRT, lon0=-3.0, lat0=3.0)),
    ("w3001_h517_vrtpx_shift", dict(w=3001, h=517, px=PX_VRT, lon0=33.0, lat0=-57.0, shift=(0.0011, 0.0004), margin=2)),
    ("ratio10", dict(w=2500, h=200, hsg_px=PX * 10)),
    ("ratio3_wide_span", dict(w=5000, h=100, hsg_px=PX * 3)),
    ("ratio1", dict(w=700, h=90, hsg_px=PX)),
    ("ratio_frac", dict(w=1999, h=257, hsg_px=PX * 7.3, shift=(0.0002, 0.0001), margin=1)),
    ("hsg_too_small_clamp", dict(w=2000, h=400, hsx=20, hsy=5)),
    ("hsg_1x1", dict(w=333, h=77, hsx=1, hsy=1)),
    ("random_profile", dict(w=2048, h=300, profile="random", seed=11)),
    ("coastal_profile", dict(w=3000, h=300, profile="coastal", seed=12)),
    ("tall", dict(w=48, h=3000, seed=13)),
]
