"""The reference-side binding, compiled and run: integration/cn_gpu.c is a replacement for the reference's src/cn.c
built against the reference's own src/global.h and its unmodified src/raster.c (over the RAM GDAL/OGR/MPI stand-ins of
oracle/refshim) and linked with libgcn10cuda.so -> oracle/_ref/libgcn10_gpu_ref.so.  oracle/_ref/libgcn10_ref.so is the
same thing with the reference's cn.c.  Both run process_block() on the same rasters; what reaches save_raster() --
18 buffers, their order, sizes, geotransform and file names -- must be identical."""
import os

import numpy as np
import pytest

from gcn10_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

GPU_REF_SO = os.path.join(os.path.dirname(O.REF_SO), "libgcn10_gpu_ref.so")
PX = 1.0 / 12000.0
PX_VRT = 8.3333333333330430e-05
HSG_PX = 1.0 / 480.0


@pytest.fixture(scope="module")
def gpu_binding():
    if not os.path.exists(GPU_REF_SO):
        pytest.skip("oracle/_ref/libgcn10_gpu_ref.so not present (make -C integration, where /root/reference exists)")
    return O.Ref(GPU_REF_SO)


CASES = [
    # name, esa raster (W, H, px, lon0, lat0), hsg raster (w, h, shift in cells), bbox in esa pixels (x0, y0, x1, y1)
    ("inside", (3000, 2200, PX, -114.0, 42.0), (130, 100, (1.3, 2.6)), (100, 200, 1400, 1500)),
    ("over_the_se_corner", (3000, 2200, PX, -114.0, 42.0), (130, 100, (1.3, 2.6)), (1700, 900, 3300, 2500)),
    ("vrt_pixel_ties", (4100, 1300, PX_VRT, -3.0, 3.0), (175, 60, (0.0, 0.0)), (0, 0, 4100, 1300)),
    ("ragged_tiny", (333, 77, PX, 33.0, -57.0), (20, 9, (0.4, 0.2)), (3, 5, 320, 70)),
]


@pytest.mark.parametrize("name,esa_spec,hsg_spec,box", CASES, ids=[c[0] for c in CASES])
def test_binding_hands_save_raster_what_the_reference_does(name, esa_spec, hsg_spec, box, ref, gpu_binding, lookup_dir):
    W, H, px, lon0, lat0 = esa_spec
    hw, hh, (sx, sy) = hsg_spec
    esa_t = (lon0, px, 0.0, lat0, 0.0, -px)
    hsg_t = (lon0 - sx * HSG_PX, HSG_PX, 0.0, lat0 + sy * HSG_PX, 0.0, -HSG_PX)
    esa = synth.esa_tile(W, H, 41 + len(name), patch=64)
    hsg = synth.hsg_tile(hw, hh, 42 + len(name), "coastal", patch=3)
    x0, y0, x1, y1 = box
    bbox = (lon0 + x0 * px, lat0 - y1 * px, lon0 + x1 * px, lat0 - y0 * px)
    want = ref.run_block(esa, esa_t, hsg, hsg_t, bbox, lookup_dir, block_id=2234)
    got = gpu_binding.run_block(esa, esa_t, hsg, hsg_t, bbox, lookup_dir, block_id=2234)
    assert want["nplanes"] == 18, want["log"]
    assert got["nplanes"] == 18, got["log"]
    assert (got["w"], got["h"], got["gt"]) == (want["w"], want["h"], want["gt"])
    assert got["paths"] == want["paths"]                        # save order and names (cn.c:236,258-259,308)
    assert got["options"] == want["options"]                    # raster.c:204-210: TILED=YES, COMPRESS=DEFLATE
    for k in range(18):
        assert np.array_equal(got["planes"][k], want["planes"][k]), f"{name}: plane {k} ({want['paths'][k]})"
    assert got["log"].count("completed condition for 2234") == want["log"].count("completed condition for 2234") == 18


def test_binding_keeps_the_skip_tier_and_the_underscore_rule(ref, gpu_binding, lookup_dir, tmp_path):
    esa_t = (-114.0, PX, 0.0, 42.0, 0.0, -PX)
    hsg_t = (-114.0, HSG_PX, 0.0, 42.0, 0.0, -HSG_PX)
    esa = synth.esa_tile(600, 500, 5, patch=40)
    hsg = synth.hsg_tile(30, 25, 6, patch=3)
    # a block outside the rasters: both skip it with the reference's message, nothing is saved (cn.c:188-192)
    far = (10.0, 10.0, 11.0, 11.0)
    a = ref.run_block(esa, esa_t, hsg, hsg_t, far, lookup_dir, block_id=7)
    b = gpu_binding.run_block(esa, esa_t, hsg, hsg_t, far, lookup_dir, block_id=7)
    assert a["nplanes"] == b["nplanes"] == 0
    assert "esa load failed for block 7" in a["log"] and "esa load failed for block 7" in b["log"]
    # no overwrite and an existing output: that raster (only) gets the trailing underscore (cn.c:320-360)
    bbox = (-114.0, 42.0 - 499.75 * PX, -114.0 + 599.75 * PX, 42.0)
    for lib in (ref, gpu_binding):
        d = tmp_path / ("ref" if lib is ref else "gpu")
        (d / "cn_rasters_undrained").mkdir(parents=True)
        (d / "cn_rasters_undrained" / "cn_f_ii_9.tif").write_bytes(b"x")
        r = lib.run_block(esa, esa_t, hsg, hsg_t, bbox, lookup_dir, block_id=9, overwrite=False, scratch_dir=str(d))
        assert r["nplanes"] == 18
        assert r["paths"][13] == "cn_rasters_undrained/cn_f_ii_9_.tif"
        assert [p for i, p in enumerate(r["paths"]) if i != 13 and p.endswith("_.tif")] == []
        if lib is ref:
            want = r
    assert r["paths"] == want["paths"] and np.array_equal(r["planes"], want["planes"])
