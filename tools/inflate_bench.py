#!/usr/bin/env python
"""GPU tile-inflate timing on the benchmark tile (development tool; bench.py is the judged benchmark).

    python tools/inflate_bench.py --tile 36000 --tiles 1024,256 --profile worldcover,random

For every (profile, input tile size): compresses the synthetic land-cover tile with zlib level 6 on the host
(thread pool; zlib releases the GIL), then times
  * the inflate kernel alone (CUDA events inside the library, gcn10_cuda_last_inflate_ms),
  * gcn10_cuda_inflate_tiles end to end (H2D of the compressed tiles + kernel + D2H of the raster),
  * gcn10_cuda_block_tiles_deflate (compressed in -> compressed out) against gcn10_cuda_block_deflate
    (raw raster in -> compressed out),
and checks the inflated raster against the source bytes.
"""
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

from gcn10_b200 import capi, synth
from tests import lookups  # noqa: E402
import bench as B  # noqa: E402


def compress_tiles(raster, tw, th, level, pinned_alloc):
    h, w = raster.shape
    tiles_x, tiles_y = (w + tw - 1) // tw, (h + th - 1) // th

    def one(i):
        ty, tx = divmod(i, tiles_x)
        t = np.zeros((th, tw), dtype=np.uint8)
        part = raster[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
        t[:part.shape[0], :part.shape[1]] = part
        return zlib.compress(t.tobytes(), level)

    with ThreadPoolExecutor(max_workers=min(64, os.cpu_count() or 8)) as ex:
        streams = list(ex.map(one, range(tiles_x * tiles_y)))
    sizes = np.array([len(z) for z in streams], dtype=np.uint32)
    offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
    total = int(sizes.sum())
    blob = pinned_alloc(total)
    pos = 0
    for z in streams:
        blob[pos:pos + len(z)] = np.frombuffer(z, dtype=np.uint8)
        pos += len(z)
    return capi.TileSource(tw, th, tiles_x, tiles_y, 0, 0, blob, offsets, sizes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=36000)
    ap.add_argument("--tiles", default="1024,256")
    ap.add_argument("--profile", default="worldcover")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--patch", type=int, default=192)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--strip-rows", default="2048")
    ap.add_argument("--streams", default="4")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    tables = B.load_tables_host(lookups.write_default_lookups(tempfile.mkdtemp()))
    ctx = capi.Context(0)
    ctx.set_luts(tables)
    w = h = a.tile
    gt, sgt, hsx, hsy = synth.block_geometry(-114.0, 42.0, w, h)
    pins = []

    def pinned(n):
        p = capi.PinnedArray(ctx.lib, (max(n, 1),))
        pins.append(p)
        return p.array

    for prof in a.profile.split(","):
        esa = pinned(w * h).reshape(h, w)
        esa[:] = synth.esa_tile(w, h, 2234, prof, device=dev, patch=a.patch).cpu().numpy()
        hsg = synth.hsg_tile(hsx, hsy, 3234, prof)
        out = pinned(w * h).reshape(h, w)
        for tsz in [int(x) for x in a.tiles.split(",")]:
            t0 = time.perf_counter()
            src = compress_tiles(esa, tsz, tsz, a.level, pinned)
            t_comp = time.perf_counter() - t0
            ms_k, ms_e2e = [], []
            for _ in range(a.reps):
                t0 = time.perf_counter()
                ctx.inflate_tiles(src, w, h, out=out)
                ms_e2e.append((time.perf_counter() - t0) * 1e3)
                ms_k.append(ctx.last_inflate_ms())
            ok = bool(np.array_equal(out, esa))
            # decoder warp alone (the writer warp drops every batch): which half of the pair sets the pace?
            ctx.set_option("inflate_probe", 1)
            ms_p = []
            for _ in range(3):
                ctx.inflate_tiles(src, w, h, out=out, want_status=True)
                ms_p.append(ctx.last_inflate_ms())
            ctx.set_option("inflate_probe", 2)
            ms_q = []
            for _ in range(3):
                ctx.inflate_tiles(src, w, h, out=out, want_status=True)
                ms_q.append(ctx.last_inflate_ms())
            ctx.set_option("inflate_probe", 0)
            rec = {"profile": prof, "tile": a.tile, "in_tile": tsz, "ntiles": int(src.sizes.size),
                   "compressed_mb": round(src.blob.size / 1e6, 2), "ratio": round(w * h / max(src.blob.size, 1), 2),
                   "host_compress_s": round(t_comp, 2), "inflate_kernel_ms": round(float(np.median(ms_k)), 3),
                   "decoder_only_ms": round(float(np.median(ms_p)), 3),
                   "no_flush_ms": round(float(np.median(ms_q)), 3),
                   "inflate_out_gbs": round(w * h / 1e6 / float(np.median(ms_k)), 1),
                   "inflate_to_host_ms": round(float(np.median(ms_e2e)), 2), "matches_source": ok}
            if not a.skip_e2e:
                nbytes = [0]

                def sink(st):
                    nbytes[0] += st.blob_bytes
                    return 0
                for sr in [int(x) for x in a.strip_rows.split(",")]:
                    for ns in [int(x) for x in a.streams.split(",")]:
                        ctx.set_option("strip_rows", sr)
                        ctx.set_option("streams", ns)
                        def piped():
                            ctx.tiles_prefetch(src, w, h)           # the next block's tiles, while this block runs
                            ctx.block_tiles_deflate(src, w, h, gt, hsg, sgt, capi.MASK_DRAINED, on_strip=sink)
                        ctx.tiles_prefetch(src, w, h)
                        for name, fn in (("tiles_in_prefetch", piped),
                                         ("tiles_in", lambda: ctx.block_tiles_deflate(src, w, h, gt, hsg, sgt, capi.MASK_DRAINED, on_strip=sink)),
                                         ("raster_in", lambda: ctx.block_deflate(esa, gt, hsg, sgt, capi.MASK_DRAINED, on_strip=sink))):
                            fn()
                            ts = []
                            for _ in range(a.reps):
                                nbytes[0] = 0
                                t0 = time.perf_counter()
                                fn()
                                ts.append((time.perf_counter() - t0) * 1e3)
                            tag = f"{name}_s{sr}_n{ns}"
                            rec[f"e2e_{tag}_ms"] = round(float(np.median(ts)), 2)
                            rec[f"e2e_{tag}_gpx_s"] = round(w * h / 1e6 / float(np.median(ts)), 2)
                            rec[f"e2e_{tag}_kernels_ms"] = round(ctx.last_kernel_ms(), 2)
                            rec[f"e2e_{tag}_d2h_mb"] = round(nbytes[0] / 1e6, 1)
            print(json.dumps(rec), flush=True)
    for p in pins:
        p.free()
    ctx.close()


if __name__ == "__main__":
    main()
