#!/usr/bin/env python
"""Smallest end-to-end case for compute-sanitizer: one ragged block through all three host-buffer entry
points (raw planes, band form, compressed tiles) with both kernels of the vector path, the byte-wise
kernel and the tile encoder launched.  Exits non-zero on any mismatch against the oracle."""
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)

from gcn10_b200 import capi, synth
from tests import lookups  # noqa: E402
from oracle import oracle as O  # noqa: E402

port = O.Port()
tables = port.load_tables(lookups.write_default_lookups(tempfile.mkdtemp()))
w, h = 4111, 300
gt, sgt, hsx, hsy = synth.block_geometry(-114.0, 42.0, w, h, hsg_origin_shift=(0.0004, 0.0009), margin_cells=1)
esa = synth.esa_tile(w, h, seed=2234, patch=64)
hsg = synth.hsg_tile(hsx, hsy, seed=3234, patch=3)
want = port.block_rows(esa, gt, hsg, sgt, tables)
with capi.Context(0) as ctx:
    ctx.set_luts(tables)
    got = ctx.block(esa, gt, hsg, sgt)
    assert np.array_equal(got, want), "block"
    band = ctx.block_rows(np.ascontiguousarray(esa[256:300]), h, 256, gt, hsg, sgt, plane_mask=capi.MASK_DRAINED)
    assert np.array_equal(band[:9], want[:9, 256:300]), "block_rows"
    res = ctx.block_deflate(esa, gt, hsg, sgt, plane_mask=0b101 | (1 << 17))
    for k, tiles in res["tiles"].items():
        for (tr, tx), z in tiles.items():
            raw = np.frombuffer(zlib.decompress(z), dtype=np.uint8).reshape(256, 256)
            y0, x0 = tr * 256, tx * 256
            hh, ww = min(256, h - y0), min(256, w - x0)
            assert np.array_equal(raw[:hh, :ww], want[k, y0:y0 + hh, x0:x0 + ww]), ("deflate", k, tr, tx)
print("sanitize case ok")
