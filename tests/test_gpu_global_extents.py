"""BASELINE config 5 in miniature: the geometry of EVERY block of the reference's shapefile (2651 extents),
cut from the real WorldCover VRT grid (4 320 000 x 1 728 000 px at 8.333...e-05 deg, so windows are
36001 px) and a global 250 m HSG grid, goes through the GPU's fp64 index-map kernel and must equal the
oracle's maps exactly -- tie columns included.  Pixel planes are spot-checked on a subset."""
import numpy as np
import pytest

from gcn10_b200 import capi, hostlib, synth
from tests import golden_io

pytestmark = pytest.mark.gpu

PX_VRT = 8.3333333333330430e-05
VRT_T = (-180.0, PX_VRT, 0.0, 84.0, 0.0, -PX_VRT)          # landcover/esa_worldcover_2021.vrt:3
VRT_W, VRT_H = 4320000, 1728000
HSG_PX = 1.0 / 480.0
HSG_T = (-180.0, HSG_PX, 0.0, 84.0, 0.0, -HSG_PX)           # a global 250 m grid in EPSG:4326 (hsg/readme.txt)
HSG_W, HSG_H = 172800, 69120


def _windows(west, north):
    bbox = (float(west), float(north - 3), float(west + 3), float(north))
    return hostlib.raster_window(VRT_W, VRT_H, VRT_T, bbox), hostlib.raster_window(HSG_W, HSG_H, HSG_T, bbox)


def test_index_maps_for_all_2651_block_extents(gpu_ctx, port):
    extents = golden_io.block_extents()
    assert len(extents) == 2651
    shapes = {}
    for bid, west, north in extents:
        we, wh = _windows(west, north)
        assert we is not None and wh is not None, bid
        _, _, w, h, gt = we
        _, _, hsx, hsy, sgt = wh
        shapes[(w, h)] = shapes.get((w, h), 0) + 1
        ci, cj = gpu_ctx.index_maps(w, h, gt, hsx, hsy, sgt)
        assert np.array_equal(ci, port.col_index(w, gt, sgt, hsx)), f"block {bid}: column map"
        assert np.array_equal(cj, port.row_index(h, gt, sgt, hsy)), f"block {bid}: row map"
    # 2638 blocks are 36001 x 36001, the rest lose a pixel at a raster edge
    assert shapes.get((36001, 36001), 0) == 2638 and sum(shapes.values()) == 2651, shapes


def test_planes_for_sampled_real_extents(gpu_ctx, port, tables):
    """Full 36001-wide rows (ragged: 36001 = 16 * 2250 + 1) of a few real extents, all 18 planes."""
    extents = golden_io.block_extents()
    rng = np.random.default_rng(5)
    for idx in rng.choice(len(extents), size=4, replace=False):
        bid, west, north = extents[idx]
        we, wh = _windows(west, north)
        _, _, w, h, gt = we
        _, _, hsx, hsy, sgt = wh
        hsg = synth.hsg_tile(hsx, hsy, seed=bid, profile="coastal")
        row0 = int(rng.integers(0, h - 300)) // 256 * 256
        n = 300
        esa_rows = synth.esa_tile(w, n, seed=bid, patch=160)
        want = port.block_rows(esa_rows, gt, hsg, sgt, tables, y0=row0, y1=row0 + n, h=h)
        got = gpu_ctx.block_rows(esa_rows, h, row0, gt, hsg, sgt)
        assert np.array_equal(got, want), f"block {bid} rows {row0}..{row0 + n}"
