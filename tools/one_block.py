#!/usr/bin/env python
"""One 36000 x 36000 block through gcn10_cuda_block_tiles_deflate (compressed land-cover tiles in -> GPU inflate ->
fused Curve Number + tile DEFLATE -> compressed tiles out), `--reps` times: the short command the ncu captures of
inflate_tiles_kernel and cn_deflate_fused_kernel are taken on (profiles/).  Prints the library's own kernel times."""
import argparse
import json
import os
import sys
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench as B  # noqa: E402
from gcn10_b200 import capi, synth  # noqa: E402
from tests import lookups  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=36000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--planes", type=int, default=9, choices=[9, 18])
    ap.add_argument("--streams", type=int, default=0, help="strip slots (0 = library default); 1 = strictly serial strips")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    tables = B.load_tables_host(lookups.write_default_lookups(tempfile.mkdtemp()))
    ctx = capi.Context(0)
    ctx.set_luts(tables)
    w = h = a.tile
    gt, sgt, hsx, hsy = synth.block_geometry(-114.0, 42.0, w, h)
    esa = synth.esa_tile(w, h, 2234, device=dev).cpu().numpy()
    hsg = synth.hsg_tile(hsx, hsy, 3234)
    T = 1024
    tx_n, ty_n = (w + T - 1) // T, (h + T - 1) // T

    def one(i):
        ty, tx = divmod(i, tx_n)
        t = np.zeros((T, T), dtype=np.uint8)
        part = esa[ty * T:(ty + 1) * T, tx * T:(tx + 1) * T]
        t[:part.shape[0], :part.shape[1]] = part
        return zlib.compress(t.tobytes(), 6)

    with ThreadPoolExecutor(16) as ex:
        streams = list(ex.map(one, range(tx_n * ty_n)))
    sizes = np.array([len(z) for z in streams], dtype=np.uint32)
    offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
    total = int(sizes.sum())
    pin = capi.PinnedArray(ctx.lib, (total,))
    pin.array[:] = np.frombuffer(b"".join(streams), dtype=np.uint8)
    src = capi.TileSource(T, T, tx_n, ty_n, 0, 0, pin.array, offsets, sizes)
    nb = [0]

    def on_strip(st):
        nb[0] += st.blob_bytes
        return 0

    mask = capi.MASK_DRAINED if a.planes == 9 else capi.MASK_ALL
    if a.streams:
        ctx.set_option("streams", a.streams)
    out = []
    for _ in range(a.reps):
        nb[0] = 0
        t0 = time.perf_counter()
        ctx.block_tiles_deflate(src, w, h, gt, hsg, sgt, plane_mask=mask, on_strip=on_strip)
        wall = (time.perf_counter() - t0) * 1e3
        out.append({"block_ms": round(wall, 3), "inflate_ms": ctx.last_inflate_ms(),
                    "fused_ms_sum_over_strips": ctx.last_kernel_ms(), "out_bytes": nb[0]})
    print(json.dumps(out))
    ctx.close()
    pin.free()


if __name__ == "__main__":
    main()
