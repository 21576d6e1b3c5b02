#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE'S OWN OBJECT CODE.

Run in the build container (needs /root/reference to build oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden.py

Every expected output below is produced by oracle/_ref/libgcn10_ref.so, i.e. by the reference's
src/cn.c (process_block, load_lookup_table, modify_hysogs_data, calculate_cn) and src/raster.c
(load_raster window arithmetic, save_raster) compiled unmodified over the RAM GDAL/OGR stand-ins in
oracle/refshim/.  The inputs are synthetic and seeded (gcn10_b200/synth.py) and are stored next to
the outputs, so the fixtures do not depend on the generator staying bit-stable.

Files written (tests/golden/):
    blocks.npz    block cases: full input rasters + dataset geotransforms + bbox -> 18 planes,
                  window size and clipped geotransform as the reference saved them
    windows.json  load_raster window arithmetic cases (incl. the real 4 320 000 x 1 728 000 VRT)
    luts.npz      effective uint8 lookup behaviour of the shipped CSVs and of a hostile custom CSV
                  set, read back through process_block on a probe raster of every (lc, hsg) pair
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gcn10_b200 import synth
from tests import lookups  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PX = 1.0 / 12000.0
PX_VRT = 8.3333333333330430e-05
HSG_PX = 1.0 / 480.0


def hostile_lookup_rows():
    """A user CSV set that pokes every branch of load_lookup_table (cn.c:50-83)."""
    rows = dict(lookups.DEFAULT_CN)
    rows["10_A"] = (255, 254, 256, 300, -1, -200, 0, 1000, 99)     # >=255 -> nodata; negatives wrap
    rows["20_E"] = (11, 12, 13, 14, 15, 16, 17, 18, 19)            # unknown letter -> D (overrides 20_D)
    rows["200_B"] = (1, 2, 3, 4, 5, 6, 7, 8, 9)                    # class outside the ESA legend
    rows["255_C"] = (21, 22, 23, 24, 25, 26, 27, 28, 29)
    rows["256_A"] = (31, 32, 33, 34, 35, 36, 37, 38, 39)           # lc out of range -> ignored
    rows["-5_A"] = (41, 42, 43, 44, 45, 46, 47, 48, 49)            # negative lc -> ignored
    rows["7_"] = (51, 52, 53, 54, 55, 56, 57, 58, 59)              # empty letter -> D
    rows["abc_A"] = (61, 62, 63, 64, 65, 66, 67, 68, 69)           # atoi -> 0 : class 0 gets a row
    return rows


def write_hostile(directory):
    lookups.write_default_lookups(directory, hostile_lookup_rows())
    # plus lines without underscore / without value / blank, appended to one file
    with open(os.path.join(directory, "default_lookup_g_ii.csv"), "ab") as f:
        f.write(b"nounderscore,5\r\n\r\n30_B\r\n40_C,\r\n50_D,77,extra\r\n")
    return directory


def block_cases():
    cases = {}

    def add(name, esa, esa_t, hsg, hsg_t, bbox):
        cases[name] = dict(esa=esa, esa_t=np.array(esa_t, dtype=np.float64), hsg=hsg,
                           hsg_t=np.array(hsg_t, dtype=np.float64), bbox=np.array(bbox, dtype=np.float64))

    # 1. exact 1/12000 pixel, 25:1, window strictly inside both rasters
    W, H = 420, 300
    add("basic", synth.esa_tile(W, H, 1, patch=40), (-114.0, PX, 0, 42.0, 0, -PX),
        synth.hsg_tile(30, 24, 2, patch=3), (-114.0 - 2 * HSG_PX, HSG_PX, 0, 42.0 + 3 * HSG_PX, 0, -HSG_PX),
        (-114.0 + 20 * PX, 42.0 - 280 * PX, -114.0 + 400 * PX, 42.0 - 10 * PX))
    # 2. real VRT pixel size, HSG grid shifted by a fraction of a cell, bbox hanging over the raster's
    #    north-west corner (negative-offset clamps, raster.c:134-141) and over its south-east corner
    W, H = 380, 260
    add("vrt_px_clipped", synth.esa_tile(W, H, 3, patch=32), (33.0, PX_VRT, 0, -57.0, 0, -PX_VRT),
        synth.hsg_tile(22, 18, 4, patch=2), (33.0 - 0.0013, HSG_PX, 0, -57.0 + 0.0007, 0, -HSG_PX),
        (33.0 - 15 * PX_VRT, -57.0 - 400 * PX_VRT, 33.0 + 500 * PX_VRT, -57.0 + 9 * PX_VRT))
    # 3. many exact-tie columns/rows (x = 12 mod 25) at an origin where FMA contraction flips ties
    W, H = 2600, 64
    add("ties_lon-3_lat3", synth.esa_tile(W, H, 5, patch=24), (-3.0, PX, 0, 3.0, 0, -PX),
        synth.hsg_tile(105, 4, 6, "random"), (-3.0, HSG_PX, 0, 3.0, 0, -HSG_PX),
        (-3.0, 3.0 - H * PX, -3.0 + W * PX, 3.0))
    # 4. coastal: nodata / water heavy land cover, dual-group and nodata heavy soils
    W, H = 333, 222
    add("coastal", synth.esa_tile(W, H, 7, "coastal", patch=28), (10.0, PX, 0, 5.0, 0, -PX),
        synth.hsg_tile(16, 11, 8, "coastal", patch=2), (10.0, HSG_PX, 0, 5.0, 0, -HSG_PX),
        (10.0, 5.0 - H * PX, 10.0 + W * PX, 5.0))
    # 5. every byte value in both rasters
    rng = np.random.default_rng(9)
    W, H = 257, 190
    add("all_bytes", rng.integers(0, 256, (H, W), dtype=np.uint8), (-60.0, PX, 0, -10.0, 0, -PX),
        rng.integers(0, 256, (9, 12), dtype=np.uint8), (-60.0, HSG_PX, 0, -10.0, 0, -HSG_PX),
        (-60.0, -10.0 - H * PX, -60.0 + W * PX, -10.0))
    # 6. HSG window far too small: every index clamps (cn.c:228-229)
    W, H = 300, 200
    add("hsg_clamped", synth.esa_tile(W, H, 10, patch=30), (100.0, PX, 0, 20.0, 0, -PX),
        synth.hsg_tile(3, 2, 11, "random"), (100.0, HSG_PX, 0, 20.0, 0, -HSG_PX),
        (100.0, 20.0 - H * PX, 100.0 + W * PX, 20.0))
    # 7. non-25 ratio with a fractional cell size
    W, H = 310, 170
    add("ratio_7p3", synth.esa_tile(W, H, 12, patch=20), (0.0, PX, 0, 0.0, 0, -PX),
        synth.hsg_tile(50, 30, 13, patch=3), (-0.0004, PX * 7.3, 0, 0.0003, 0, -PX * 7.3),
        (0.0, -H * PX, W * PX, 0.0))
    return cases


def lut_probe_raster():
    """One pixel per (land cover 0..255) x (HSG code from a list): row = hsg code, col = lc."""
    hsg_codes = list(range(0, 16)) + [100, 254, 255]
    esa = np.tile(np.arange(256, dtype=np.uint8), (len(hsg_codes), 1))
    # one HSG cell per raster row, 1024 units wide (ci = 0 for every x).  With the HSG grid half a
    # cell above the block, cn.c:224,226 gives cj = y + 1, so a dummy row sits on top; the bbox is
    # one row taller than the land-cover raster so the HSG window keeps its last row.
    hsg = np.array([255] + hsg_codes, dtype=np.uint8)[:, None]
    esa_t = (0.0, 1.0, 0, 0.0, 0, -1.0)
    hsg_t = (0.0, 1024.0, 0, 0.5, 0, -1.0)
    bbox = (0.0, -float(len(hsg_codes) + 1), 256.0, 0.0)
    return esa, esa_t, hsg, hsg_t, bbox, hsg_codes


def window_cases():
    vrt_t = (-180.0, PX_VRT, 0.0, 84.0, 0.0, -PX_VRT)
    VW, VH = 4320000, 1728000
    hs_t = (-180.0, HSG_PX, 0.0, 84.0, 0.0, -HSG_PX)
    HW, HH = 172800, 69120
    cases = []

    def add(rw, rh, t, bbox):
        cases.append(dict(rw=rw, rh=rh, t=list(t), bbox=list(bbox)))

    for lon, lat in [(-114, 42), (-180, 84), (177, 84), (-180, -57), (177, -57), (0, 3), (-3, 3), (33, -57)]:
        add(VW, VH, vrt_t, (lon, lat - 3, lon + 3, lat))
        add(HW, HH, hs_t, (lon, lat - 3, lon + 3, lat))
    add(VW, VH, vrt_t, (-183, 81, -180, 84))          # entirely west of the raster -> negative count
    add(VW, VH, vrt_t, (-181.5, 82.5, -178.5, 85.5))  # overlaps the NW corner
    add(VW, VH, vrt_t, (178.5, -61.5, 181.5, -58.5))  # overlaps the SE corner
    add(VW, VH, vrt_t, (180, 0, 183, 3))              # starts at the east edge -> invalid
    add(1000, 800, (5.0, 0.01, 0, 50.0, 0, -0.01), (5.005, 49.001, 7.777, 49.999))
    add(1000, 800, (5.0, 0.01, 0, 50.0, 0, -0.01), (5.0, 42.0, 15.0, 50.0))
    add(1000, 800, (5.0, 0.01, 0, 50.0, 0, -0.01), (4.0, 41.0, 16.0, 51.0))
    add(10, 10, (0.0, 1.0, 0, 0.0, 0, -1.0), (2.5, -7.5, 2.6, -7.4))       # sub-pixel bbox
    add(10, 10, (0.0, 1.0, 0, 0.0, 0, -1.0), (3.0, -3.0, 3.0, -3.0))       # empty bbox -> invalid
    return cases


def main():
    O.build(ref=True)
    ref = O.Ref()
    tmp = tempfile.mkdtemp(prefix="gcn10_golden_")
    default_dir = lookups.write_default_lookups(os.path.join(tmp, "default"))
    hostile_dir = write_hostile(os.path.join(tmp, "hostile"))

    out = {}
    for name, c in block_cases().items():
        r = ref.run_block(c["esa"], c["esa_t"], c["hsg"], c["hsg_t"], c["bbox"], default_dir, block_id=2234)
        assert r["nplanes"] == 18, (name, r["log"])
        assert r["paths"] == [f"cn_rasters_{cd}/cn_{h}_{a}_2234.tif" for cd in O.CONDS for h in O.HCS for a in O.ARCS]
        for k, v in c.items():
            out[f"{name}/{k}"] = v
        out[f"{name}/planes"] = r["planes"]
        out[f"{name}/gt"] = np.array(r["gt"], dtype=np.float64)
        print(f"{name:18s} window {r['w']}x{r['h']}  nodata {float((r['planes'] == 255).mean()):.3f}")
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **out)

    wins = []
    for c in window_cases():
        res = ref.window(c["rw"], c["rh"], c["t"], c["bbox"])
        c["expect"] = None if res is None else dict(xoff=res[0], yoff=res[1], xsize=res[2], ysize=res[3], gt=list(res[4]))
        wins.append(c)
    with open(os.path.join(HERE, "windows.json"), "w") as f:
        json.dump(wins, f, indent=1)
    print(f"{len(wins)} window cases; sizes on the real VRT:",
          sorted({(w['expect']['xsize'], w['expect']['ysize']) for w in wins if w['expect'] and w['rw'] == 4320000}))

    esa, esa_t, hsg, hsg_t, bbox, codes = lut_probe_raster()
    luts = {"hsg_codes": np.array(codes, dtype=np.int32)}
    for label, d in (("default", default_dir), ("hostile", hostile_dir)):
        r = ref.run_block(esa, esa_t, hsg, hsg_t, bbox, d, block_id=1)
        assert r["nplanes"] == 18 and (r["w"], r["h"]) == (256, len(codes)), (label, r["w"], r["h"], r["log"])
        luts[label] = r["planes"]            # [18, len(codes), 256]: plane, hsg code, land cover
    np.savez_compressed(os.path.join(HERE, "luts.npz"), **luts)
    # every block extent of the shipped shapefile (all are integer-degree 3 x 3 squares)
    shp = "/root/reference/blocks/esa_extent_blocks.shp"
    if os.path.exists(shp):
        from gcn10_b200 import hostlib
        b = hostlib.Blocks(shp)
        rows = []
        for i in b.ids():
            x0, y0, x1, y1 = b.bbox(i)
            assert (x1 - x0, y1 - y0) == (3.0, 3.0) and x0 == int(x0) and y0 == int(y0)
            rows.append([i, int(x0), int(y1)])
        with open(os.path.join(HERE, "block_extents.json"), "w") as f:
            json.dump({"comment": "id, west edge (deg), north edge (deg) of every 3x3 degree block of "
                                  "/root/reference/blocks/esa_extent_blocks.shp (read with gcn10_b200/hostlib.Blocks); "
                                  "generated by tests/golden/make_golden.py", "blocks": rows}, f, separators=(",", ":"))
    for fn in ("blocks.npz", "windows.json", "luts.npz", "block_extents.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
