#!/usr/bin/env python
"""Where the model table of build_tile_code() (gcn10_b200/csrc/tile_code.h) comes from, and what it costs.

    python tools/token_stats.py [--rows 1024] [--profiles worldcover:2234 worldcover:2015 coastal:2301]

CPU only.  For each synthetic block (the generators of bench.py / BASELINE.json) the CPU oracle computes the nine
drained planes of the first `rows` rows, the planes are numbered into value records the way the library does
(gcn10_cuda.cu: build_fused_tables), and the scalar model of the fused kernel's parse in
tests/harness/tile_code_host.cpp counts the tokens of every 256 x 256 tile.  Printed: the tokens per 1000 (the table to
paste into tile_code.h), and bytes per tile stream with the code the library builds today against a code fitted to
that block alone -- save_raster() of the reference lets zlib fit one per tile (/root/reference/src/raster.c:204-219).
"""
from __future__ import annotations

import argparse
import ctypes
import heapq
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gcn10_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import lookups  # noqa: E402


def harness():
    out = os.path.join(tempfile.mkdtemp(), "libtile_code_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "harness", "tile_code_host.cpp")])
    return ctypes.CDLL(out)


def huffman_depths(freq):
    h = [(f, i, None, None) for i, f in enumerate(freq) if f > 0]
    heapq.heapify(h)
    n = len(freq)
    depth = [0] * n
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        n += 1
        heapq.heappush(h, (a[0] + b[0], n, a, b))

    def walk(node, d):
        if node[2] is None:
            depth[node[1]] = max(d, 1)
        else:
            walk(node[2], d + 1)
            walk(node[3], d + 1)
    walk(h[0], 0)
    return depth


def block_tokens(lib, port, tables, profile, seed, rows, w=36000, h=36000):
    gt, sgt, hsx, hsy = synth.block_geometry(-114.0, 42.0, w, h)
    esa = synth.esa_tile(w, rows, seed, profile)
    hsg = synth.hsg_tile(hsx, hsy, seed + 1000, profile)
    planes = port.block_rows(esa, gt, hsg, sgt, tables, 0, rows, h)[:9]
    key = np.zeros((rows, w), np.uint64)
    for k in range(9):
        key = key * np.uint64(257) + planes[k]
    _, ids = np.unique(key, return_inverse=True)
    ids = ids.reshape(rows, w).astype(np.uint8)
    values = np.unique(planes)
    hist = np.zeros(34, np.uint64)
    stream = np.zeros(34, np.uint64)                            # what-if: matches that run on across row ends
    for ty in range(rows // 256):
        for tx in range((w + 255) // 256):
            t = np.full((256, 256), 255, np.uint8)              # id of the padding right of the raster
            blk = ids[ty * 256:(ty + 1) * 256, tx * 256:(tx + 1) * 256]
            t[:, :blk.shape[1]] = blk
            lib.gcn10_test_tile_tokens(t.ctypes.data_as(ctypes.c_void_p), hist.ctypes.data_as(ctypes.c_void_p))
            lib.gcn10_test_tile_tokens_stream(t.ctypes.data_as(ctypes.c_void_p), stream.ctypes.data_as(ctypes.c_void_p), 1)
    return hist.astype(np.float64), values, stream.astype(np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--profiles", nargs="*", default=["worldcover:2234", "worldcover:2015", "coastal:2301"])
    a = ap.parse_args()
    lib = harness()
    port = O.Port()
    tables = port.load_tables(lookups.write_default_lookups(tempfile.mkdtemp()))
    present = np.zeros(256, np.uint8)
    present[[int(v) for v in np.unique(tables) if 0 <= v < 255] + [0, 255]] = 1
    len_bits = (ctypes.c_int * 29)()
    eob, lit, hdr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.gcn10_test_tile_code_lengths(present.ctypes.data_as(ctypes.c_void_p), len_bits, ctypes.byref(eob),
                                            ctypes.byref(lit), ctypes.byref(hdr)) == 0
    today = [eob.value] + list(len_bits)
    k = int(np.ceil(np.log2(max(int(present.sum()), 1))))          # the literals: one leaf split into 2^k code words
    print(f"code today: literals {lit.value} bits, header {hdr.value} bits, end-of-block + length symbols {today}")
    blend = np.zeros(31)
    for spec in a.profiles:
        profile, seed = spec.split(":")
        hist, _, stream = block_tokens(lib, port, tables, profile, int(seed), a.rows)
        tiles = hist[1]
        sym = np.concatenate(([hist[1]], hist[2:31]))
        tokens = hist[0] + hist[2:31].sum()
        fixed = hist[31] + 7 * hist[32] + hist[33]              # extra length bits, distance codes (+ 6 extra bits at 256)
        tail = (hdr.value + 7) // 8 + 4                         # header, Adler-32

        def size(bits, lit_bits=lit.value):
            return ((hist[0] * lit_bits + (sym * np.asarray(bits)).sum() + fixed) / tiles) / 8 + tail
        joint = huffman_depths(list(sym) + [hist[0]])          # fitted to this block: the literals as one more leaf
        lit_fit, fitted = joint[-1] + k, np.asarray(joint[:-1])
        print(f"{spec}: {tokens / tiles:.0f} tokens per tile ({hist[0] / tiles:.0f} literals, {hist[32] / tiles:.0f} "
              f"above, {hist[33] / tiles:.0f} runs); bytes per stream: today {size(today):.1f}, fitted to this block "
              f"{size(fitted, lit_fit):.1f}")
        ssym = np.concatenate(([stream[1]], stream[2:31]))
        sbytes = ((stream[0] * lit.value + (ssym * np.asarray(today)).sum() + stream[31] + 7 * stream[32] + stream[33])
                  / tiles) / 8 + tail
        print(f"    what-if, same code, greedy parse over the tile as one stream (matches run on across row ends, runs win "
              f"ties): {(stream[0] + stream[2:31].sum()) / tiles:.0f} tokens, {sbytes:.1f} bytes per stream")
        blend += np.concatenate(([hist[0]], sym)) / tokens * 1000 / len(a.profiles)
    print("per 1000 tokens: literals %d; end-of-block and length symbols 257..285:" % round(blend[0]))
    print("   ", [max(2, int(round(x))) for x in blend[1:]])


if __name__ == "__main__":
    main()
