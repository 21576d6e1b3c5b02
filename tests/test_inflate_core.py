"""CPU checks of the decode-lane half of the GPU tile inflater (gcn10_b200/csrc/inflate_core.h).

The land-cover window the reference obtains from GDAL (load_raster, /root/reference/src/raster.c:106-189)
is, for the ESA WorldCover GeoTIFFs, a set of zlib-compressed TIFF tiles; zlib (the codec GDAL's GTiff
driver links) is therefore the checker here.  tests/harness/inflate_host.cpp compiles the very header the
CUDA kernel includes and emulates the warp loop in scalar code; these tests compare its output with
zlib.decompress over every block type, strategy and level, at all 16 byte alignments, and feed it damaged
streams.  The GPU kernel itself is compared with zlib in tests/test_gpu_inflate.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

from gcn10_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "harness", "inflate_host.cpp")
HDR = os.path.join(ROOT, "gcn10_b200", "csrc", "inflate_core.h")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("inflate_harness") / "libinflate_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", out, SRC])
    lib = ctypes.CDLL(out)
    lib.gcn10_test_inflate.restype = ctypes.c_int
    lib.gcn10_test_inflate.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                       ctypes.c_uint32, ctypes.c_void_p]
    return lib


def run(lib, stream: bytes, out_len: int, misalign: int = 0):
    out = np.full(out_len + 64, 0xEE, dtype=np.uint8)
    stats = np.zeros(4, dtype=np.uint64)
    rc = lib.gcn10_test_inflate(stream, len(stream), misalign, out.ctypes.data, out_len, stats.ctypes.data)
    assert (out[out_len:] == 0xEE).all(), "harness wrote past the tile"
    return rc, out[:out_len], stats


def deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def corpus():
    rng = np.random.default_rng(7)
    tile = synth.esa_tile(1024, 1024, 2234).tobytes()
    noisy = synth.esa_tile(512, 512, 99, profile="random").tobytes()
    small = synth.esa_tile(256, 256, 5, patch=24).tobytes()
    period = rng.integers(0, 256, 30011, dtype=np.uint8).tobytes()
    far = (period * 12)[:300000]                       # matches at distance 30011: the batch-closing rule
    runs = (b"\x0a" * 70000 + b"\x14" * 3 + b"\x1e" * 259 + b"\x28" * 258) * 2
    text = bytes(rng.integers(0, 256, 200000, dtype=np.uint8))     # incompressible -> stored blocks inside level 6
    skew = rng.choice(np.arange(256, dtype=np.uint8), size=400000,
                      p=np.r_[[0.6], np.full(255, 0.4 / 255)]).tobytes()    # long Huffman codes (> 10 bits)
    return {"tile1024": tile, "noisy512": noisy, "small256": small, "far": far, "runs": runs, "random": text,
            "skew": skew, "one": b"\x07", "empty": b""}


CORPUS = corpus()


@pytest.mark.parametrize("name", sorted(CORPUS))
@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE),
                                            (6, zlib.Z_FILTERED)])
def test_matches_zlib(harness, name, level, strategy):
    data = CORPUS[name]
    stream = deflate(data, level, strategy)
    assert zlib.decompress(stream) == data
    rc, out, stats = run(harness, stream, len(data))
    assert rc == 0, f"inflate error {rc}"
    assert out.tobytes() == data


@pytest.mark.parametrize("misalign", range(16))
def test_every_stream_alignment(harness, misalign):
    data = CORPUS["small256"]
    rc, out, _ = run(harness, deflate(data, 6), len(data), misalign)
    assert rc == 0 and out.tobytes() == data


def test_small_windows_and_memlevels(harness):
    # wbits 9..15 changes CMF and the reachable distances; memLevel 1 gives many short dynamic blocks
    data = CORPUS["tile1024"][:300000]
    for wbits in (9, 12, 15):
        for memlevel in (1, 9):
            stream = deflate(data, 6, wbits=wbits, memlevel=memlevel)
            rc, out, stats = run(harness, stream, len(data))
            assert rc == 0 and out.tobytes() == data
    assert stats[3] >= 1


def test_block_mix_with_full_flushes(harness):
    # stored, fixed and dynamic blocks interleaved in one stream (Z_FULL_FLUSH emits empty stored blocks)
    rng = np.random.default_rng(3)
    c = zlib.compressobj(6)
    parts, raw = [], []
    for k in range(12):
        chunk = (rng.integers(0, 256, 5000, dtype=np.uint8).tobytes() if k % 3 == 0
                 else synth.esa_tile(300, 40, k, patch=16).tobytes())
        raw.append(chunk)
        parts.append(c.compress(chunk))
        parts.append(c.flush(zlib.Z_FULL_FLUSH if k % 2 else zlib.Z_SYNC_FLUSH))
    parts.append(c.flush())
    data, stream = b"".join(raw), b"".join(parts)
    rc, out, stats = run(harness, stream, len(data))
    assert rc == 0 and out.tobytes() == data
    assert stats[3] > 12


def test_gpu_encoder_style_stream(harness):
    # what deflate_tiles_kernel emits: one fixed-Huffman block whose matches use distances 1 and 256 only
    data = synth.esa_tile(256, 256, 11, patch=40).tobytes()
    stream = deflate(data, 6, zlib.Z_FIXED)
    rc, out, _ = run(harness, stream, len(data))
    assert rc == 0 and out.tobytes() == data


def test_size_mismatch_is_reported(harness):
    data = CORPUS["small256"]
    stream = deflate(data, 6)
    rc, _, _ = run(harness, stream, len(data) - 1)
    assert rc == 7          # kErrOverflow: more bytes than the tile holds
    rc, _, _ = run(harness, stream, len(data) + 1)
    assert rc == 9          # kErrShort


def test_bad_headers(harness):
    data = CORPUS["small256"]
    stream = bytearray(deflate(data, 6))
    bad = bytearray(stream)
    bad[0] = 0x79                      # method 9
    assert run(harness, bytes(bad), len(data))[0] == 1
    bad = bytearray(stream)
    bad[1] ^= 0x01                     # FCHECK
    assert run(harness, bytes(bad), len(data))[0] == 1
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(data) + raw.flush()
    assert run(harness, b"\x78\x9c" + bytes([body[0] | 0x06]) + body[1:], len(data))[0] == 2    # BTYPE 3


def test_damaged_streams_terminate_with_an_error_or_wrong_bytes_never_overrun(harness):
    rng = np.random.default_rng(11)
    data = CORPUS["tile1024"][:200000]
    stream = deflate(data, 6)
    detected = 0
    for _ in range(300):
        bad = bytearray(stream)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(2, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        rc, out, _ = run(harness, bytes(bad), len(data))      # run() asserts the guard band
        try:
            ok = zlib.decompress(bytes(bad)) == data
        except zlib.error:
            ok = False
        if rc != 0:
            detected += 1
        else:
            # accepted = structure intact AND the Adler-32 trailer matches what was decoded: the bytes are right
            # (zlib additionally rejects incomplete code sets, so `ok` may be False here)
            assert out.tobytes() == data
        if ok:
            assert rc == 0
    assert detected > 250
    for cut in (3, 10, len(stream) // 2, len(stream) - 5):
        rc, _, _ = run(harness, stream[:cut], len(data))
        assert rc != 0


def test_adler32_trailer_is_verified(harness):
    """Payload damage that leaves the Huffman structure valid is caught by the checksum, as zlib does for GDAL
    (a read error at /root/reference/src/raster.c:182-186): status kErrChecksum = 10."""
    data = CORPUS["small256"]
    stored = bytearray(deflate(data, 0))                    # stored blocks: every payload byte is a literal
    assert run(harness, bytes(stored), len(data))[0] == 0
    bad = bytearray(stored)
    bad[7 + 1000] ^= 0x10
    rc, out, _ = run(harness, bytes(bad), len(data))
    assert rc == 10 and out.tobytes() != data
    for stream in (deflate(data, 6), deflate(data, 6, zlib.Z_FIXED), bytes(stored)):
        for k in range(1, 5):                                # each byte of the trailer
            bad = bytearray(stream)
            bad[-k] ^= 0x01
            assert run(harness, bytes(bad), len(data))[0] == 10
        assert run(harness, stream[:-2], len(data))[0] == 8  # kErrInput: the stream ends inside the trailer
    # a literal swapped for another literal of the same code length (Huffman-only stream, one flipped bit)
    hs = deflate(data, 6, zlib.Z_HUFFMAN_ONLY)
    hits = 0
    for pos in range(len(hs) // 2, len(hs) // 2 + 40):
        bad = bytearray(hs)
        bad[pos] ^= 0x04
        rc, out, _ = run(harness, bytes(bad), len(data))
        assert rc != 0 or out.tobytes() == data
        hits += rc == 10
    assert hits > 0


def test_symbol_statistics_are_exported(harness):
    data = CORPUS["tile1024"]
    rc, out, stats = run(harness, deflate(data, 6), len(data))
    assert rc == 0
    nsym, nmatch, nsteps, nblocks = (int(v) for v in stats)
    assert 0 < nmatch <= nsym and nsteps >= nsym // 32 and nblocks >= 1
