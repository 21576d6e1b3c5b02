"""The optional GDAL input backend of the host library (gcn10_b200/host/host_raster_gdal.c, make GDAL=1).

No GDAL is installed in this image, so the backend is compiled here against oracle/refshim -- the RAM stand-in for
GDAL the reference's own raster.c is compiled against for the oracle -- and driven through the same gh_raster_* calls
the host program makes.  What this pins: the GDAL calls and their order (raster.c:118-179), the window / pitch handling,
the error strings, and the routing rules of gh_raster_open (which inputs go to GDAL and which stay with the built-in
readers).  Reading real formats is GDAL's job and cannot be exercised here.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from gcn10_b200 import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gcn10_b200", "host")
SHIM = os.path.join(ROOT, "oracle", "refshim")


@pytest.fixture(scope="module")
def gdal_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gdalhost") / "libgcn10host_gdal.so")
    srcs = [os.path.join(HOST, f) for f in ("host_core.c", "host_tiff.c", "host_raster.c", "host_raster_gdal.c")]
    srcs.append(os.path.join(SHIM, "fake_gdal.c"))
    # the write side of the stand-in reports to ref_api.c, which is not part of this build
    stub = str(tmp_path_factory.mktemp("gdalstub") / "sink_stub.c")
    with open(stub, "w") as f:
        f.write('#include "refshim.h"\n'
                'void refshim_sink_deliver(refshim_sink *s, const char *p, const void *b, int w, int h)\n'
                '{ (void)s; (void)p; (void)b; (void)w; (void)h; }\n'
                'double refshim_now(void) { return 0.0; }\n')
    srcs.append(stub)
    cmd = ["gcc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-fPIC", "-ffp-contract=off", "-pthread", "-DGCN10_WITH_GDAL",
           "-I" + os.path.join(SHIM, "include"), "-I" + SHIM, "-shared", "-o", out, *srcs, "-lz", "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    L = hostlib.load(out)
    L.refshim_reset.restype = None
    L.refshim_add_raster.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return L


def _register(L, name, data, gt):
    assert L.refshim_add_raster(name.encode(), data.ctypes.data, data.shape[1], data.shape[0], (C.c_double * 6)(*gt)) == 0


def test_plain_build_has_no_gdal_and_says_so(monkeypatch, tmp_path):
    L = hostlib.load()
    assert L.gh_raster_have_gdal() == 0
    monkeypatch.setenv("GCN10_RASTER_BACKEND", "gdal")
    with pytest.raises(hostlib.HostError) as ei:
        hostlib.Raster(str(tmp_path / "x.tif"))
    assert "gdal open failed" in ei.value.msg and "make GDAL=1" in ei.value.msg


def test_gdal_backend_reads_windows_like_rasterio(gdal_lib, monkeypatch):
    monkeypatch.delenv("GCN10_RASTER_BACKEND", raising=False)
    L = gdal_lib
    assert L.gh_raster_have_gdal() == 1
    L.refshim_reset()
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, size=(300, 517), dtype=np.uint8)
    gt = (-3.0, 1.0 / 12000, 0.0, 4.0, 0.0, -1.0 / 12000)
    # a name no local file has: fopen fails, GDALOpen takes it (how /vsicurl/ inputs arrive)
    _register(L, "/vsicurl/https://example.invalid/lc.tif", data, gt)
    r = hostlib.Raster("/vsicurl/https://example.invalid/lc.tif", lib=L)
    assert r.backend == "gdal" and not r.is_mosaic and r.source_count == 1
    assert (r.width, r.height) == (517, 300) and r.gt == gt
    for (x, y, w, h) in [(0, 0, 517, 300), (5, 7, 256, 256), (516, 299, 1, 1), (100, 0, 417, 1)]:
        assert np.array_equal(r.read(x, y, w, h), data[y:y + h, x:x + w])
    # a row pitch wider than the window (the pipeline's band buffers): only the window's bytes are written
    assert np.array_equal(r.read(9, 11, 40, 13, pitch=64), data[11:24, 9:49])
    # GDAL hands out pixels, not compressed tiles: the caller decodes on the host
    rc, parts = r.window_parts(0, 0, 256, 256)
    assert rc == 1 and parts == []
    # outside the raster: RasterIO's failure, worded like raster.c:182
    for bad in [(-1, 0, 5, 5), (0, 0, 518, 1), (0, 299, 1, 2), (0, 0, 0, 1)]:
        with pytest.raises(hostlib.HostError) as ei:
            r.read(*bad)
        assert "gdalrasterio error" in ei.value.msg
    r.close()
    with pytest.raises(hostlib.HostError) as ei:
        hostlib.Raster("/vsicurl/https://example.invalid/other.tif", lib=L)
    assert ei.value.msg == "gdal open failed: /vsicurl/https://example.invalid/other.tif"     # raster.c:121


def test_routing_between_builtin_readers_and_gdal(gdal_lib, monkeypatch, tmp_path):
    L = gdal_lib
    L.refshim_reset()
    rng = np.random.default_rng(6)
    data = rng.integers(0, 256, size=(64, 96), dtype=np.uint8)
    gt = (10.0, 0.5, 0.0, 20.0, 0.0, -0.5)
    tif = str(tmp_path / "lc.tif")
    hostlib.tiff_write(tif, data, gt, threads=2)
    other = (255 - data).copy()
    _register(L, tif, other, gt)                 # GDAL would see different pixels: tells the two readers apart
    monkeypatch.delenv("GCN10_RASTER_BACKEND", raising=False)
    r = hostlib.Raster(tif, lib=L)               # a DEFLATE GeoTIFF stays with the built-in reader (compressed tiles)
    assert r.backend == "geotiff" and np.array_equal(r.read(0, 0, 96, 64), data)
    rc, parts = r.window_parts(0, 0, 96, 64)
    assert rc == 0 and len(parts) == 1
    r.close()
    monkeypatch.setenv("GCN10_RASTER_BACKEND", "gdal")
    r = hostlib.Raster(tif, lib=L)               # forced
    assert r.backend == "gdal" and np.array_equal(r.read(0, 0, 96, 64), other)
    r.close()
    monkeypatch.delenv("GCN10_RASTER_BACKEND")
    # a local file the built-in readers reject (not a TIFF, not a VRT) falls through to GDAL
    img = str(tmp_path / "lc.img")
    with open(img, "wb") as f:
        f.write(b"EHFA_HEADER_TAG" + bytes(64))
    _register(L, img, data, gt)
    r = hostlib.Raster(img, lib=L)
    assert r.backend == "gdal" and np.array_equal(r.read(3, 4, 50, 20), data[4:24, 3:53])
    r.close()
    # so does a VRT outside the 1:1 mosaic subset (here: a scaled source)
    vrt = str(tmp_path / "scaled.vrt")
    with open(vrt, "w") as f:
        f.write('<VRTDataset rasterXSize="48" rasterYSize="32"><GeoTransform>10,1,0,20,0,-1</GeoTransform>'
                '<VRTRasterBand dataType="Byte" band="1"><SimpleSource>'
                '<SourceFilename relativeToVRT="1">lc.tif</SourceFilename><SourceBand>1</SourceBand>'
                '<SrcRect xOff="0" yOff="0" xSize="96" ySize="64"/><DstRect xOff="0" yOff="0" xSize="48" ySize="32"/>'
                '</SimpleSource></VRTRasterBand></VRTDataset>')
    small = data[::2, ::2].copy()
    _register(L, vrt, small, (10.0, 1.0, 0.0, 20.0, 0.0, -1.0))
    r = hostlib.Raster(vrt, lib=L)
    assert r.backend == "gdal" and (r.width, r.height) == (48, 32) and np.array_equal(r.read(0, 0, 48, 32), small)
    r.close()
    # the plain build reports the VRT's own error for the same file
    with pytest.raises(hostlib.HostError):
        hostlib.Raster(vrt)
