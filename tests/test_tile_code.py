"""CPU checks of the tuned Huffman code of the GPU tile encoder (gcn10_b200/csrc/tile_code.h).

save_raster() of the reference produces zlib streams (/root/reference/src/raster.c:204-219); whatever our encoder
writes must be inflatable by zlib, which is strict about dynamic-block headers (complete literal/length code,
valid code-length code).  tests/harness/tile_code_host.cpp builds the code for a set of byte values and encodes
tiles with the kernel's token rules in scalar code; zlib must give the tile back.  The GPU kernel itself is
checked in tests/test_gpu_deflate.py."""
from __future__ import annotations

import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

from gcn10_b200 import synth
from tests import lookups

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "harness", "tile_code_host.cpp")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tile_code") / "libtile_code_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", out, SRC])
    lib = ctypes.CDLL(out)
    lib.gcn10_test_tile_code.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 4
    lib.gcn10_test_tile_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    lib.gcn10_test_tile_encode.restype = ctypes.c_uint32
    return lib


def _present(values):
    p = np.zeros(256, dtype=np.uint8)
    p[list(values)] = 1
    return p


def _encode(lib, present, tile):
    out = np.zeros(200000, dtype=np.uint8)
    n = lib.gcn10_test_tile_encode(present.ctypes.data, np.ascontiguousarray(tile).ctypes.data, out.ctypes.data, out.size)
    assert n > 0
    return out[:n].tobytes()


def _cn_values():
    """every byte the shipped lookup tables can put into a plane, plus nodata and the tile padding"""
    import tempfile
    from oracle import oracle as O
    t = O.Port().load_tables(lookups.write_default_lookups(tempfile.mkdtemp()))
    vals = {int(v) for v in np.unique(t) if 0 <= v < 255}
    return sorted(vals | {0, 255})


def test_code_for_the_shipped_tables(harness):
    vals = _cn_values()
    p = _present(vals)
    lb, hb, l284, eob = (ctypes.c_int() for _ in range(4))
    assert harness.gcn10_test_tile_code(p.ctypes.data, lb, hb, l284, eob) == 0
    assert lb.value == 8, "64 values: one leaf of the length tree, two bits deep, split into 64 code words"
    assert l284.value <= 4, "the repeated-row length bucket must be cheap"
    assert hb.value < 19 + 8 * 100, "header under 100 bytes"


@pytest.mark.parametrize("nvals", [1, 2, 3, 31, 62, 63, 64, 65, 96, 97, 127, 128, 129, 192, 200, 224, 240])
def test_streams_inflate_with_zlib(harness, nvals):
    rng = np.random.default_rng(nvals)
    vals = sorted(rng.choice(256, size=nvals, replace=False).tolist())
    p = _present(vals)
    # piecewise constant tile (soil cells x land-cover patches) mapped onto the value set, plus a noisy band
    base = synth.esa_tile(256, 256, nvals, patch=20).astype(np.int64)
    tile = np.asarray(vals, dtype=np.uint8)[base % nvals]
    tile[100:110] = np.asarray(vals, dtype=np.uint8)[rng.integers(0, nvals, size=(10, 256))]
    z = _encode(harness, p, tile)
    assert zlib.decompress(z) == tile.tobytes()
    if nvals <= 64:
        assert len(z) < len(zlib.compress(tile.tobytes(), 1)) * 2


def test_more_than_240_values_is_refused(harness):
    p = _present(range(241))
    out = np.zeros(1000, dtype=np.uint8)
    tile = np.zeros((256, 256), dtype=np.uint8)
    assert harness.gcn10_test_tile_encode(p.ctypes.data, tile.ctypes.data, out.ctypes.data, out.size) == 0


def test_smaller_than_the_fixed_code(harness):
    """On a Curve-Number-like tile the tuned code beats RFC 1951's fixed code (what the encoder used before)."""
    vals = _cn_values()
    p = _present(vals)
    base = synth.esa_tile(256, 256, 3, patch=48).astype(np.int64)
    cells = (np.arange(256)[:, None] // 25 * 11 + np.arange(256)[None, :] // 25)
    tile = np.asarray(vals, dtype=np.uint8)[(base + cells) % len(vals)]
    z = _encode(harness, p, tile)
    assert zlib.decompress(z) == tile.tobytes()
    c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    fixed = c.compress(tile.tobytes()) + c.flush()
    assert len(z) < len(fixed), (len(z), len(fixed))
    print("tuned", len(z), "zlib fixed", len(fixed), "zlib 6", len(zlib.compress(tile.tobytes(), 6)))


def _tokens(lib, tile):
    hist = np.zeros(34, dtype=np.uint64)
    lib.gcn10_test_tile_tokens.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.gcn10_test_tile_tokens(np.ascontiguousarray(tile).ctypes.data, hist.ctypes.data)
    return hist.astype(np.int64)


def test_token_statistics_of_known_tiles(harness):
    """The parse model behind the code's length table (tools/token_stats.py), on tiles whose tokens can be counted by
    hand: the kernel's row-run rule (cn_deflate_fused.cuh: fused_row_run) and the greedy row parse."""
    # one value: a literal and a run of 255 in row 0, then 255 rows that repeat the row above = 65280 bytes at distance
    # 256 = 253 matches of 258 and one of 6
    h = _tokens(harness, np.full((256, 256), 7, np.uint8))
    assert h[0] == 1 and h[1] == 1
    assert h[2 + 27] == 1 and h[2 + 28] == 253 and h[2 + 3] == 1 and h[2:31].sum() == 255
    assert h[32] == 254 and h[33] == 1
    # vertical stripes 25 pixels wide (soil cells): row 0 = 11 literals + 10 runs of 24 + one of 5, the rest repeats
    t = np.tile((np.arange(256) // 25).astype(np.uint8), (256, 1))
    h = _tokens(harness, t)
    assert h[0] == 11 and h[33] == 11 and h[32] == 254
    # every row differs from the row above in one cell: two matches above around a literal + run per row, no row-runs;
    # row 0 has nothing above it: three literals, three runs
    t = np.zeros((256, 256), np.uint8)
    for r in range(256):
        t[r, 100:125] = r % 2 + 1
    h = _tokens(harness, t)
    assert h[2 + 28] == 0, "no two rows in sequence repeat the row above"
    assert h[0] == 3 + 255 and h[32] == 2 * 255 and h[33] == 3 + 255


def test_length_table_follows_the_measured_model(harness):
    """Symbols that the measured model makes frequent are short: the row-run token (258), the single repeated row
    (227..257) and length 5 outrank the rest; nothing is longer than 12 bits (tuned_match_code packs a token into 27)."""
    p = _present(_cn_values())
    bits = (ctypes.c_int * 29)()
    eob, lit, hdr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert harness.gcn10_test_tile_code_lengths(p.ctypes.data_as(ctypes.c_void_p), bits, ctypes.byref(eob),
                                                ctypes.byref(lit), ctypes.byref(hdr)) == 0
    b = list(bits)
    assert b[28] == min(b) and b[28] <= 3 and b[27] <= 4 and b[2] <= 4
    assert max(b + [eob.value]) <= 12
    assert max(b[8:27]) <= 8, "lengths 11..226 are a third of all tokens"
    assert lit.value == 8 and hdr.value <= 19 + 8 * 48
