#!/usr/bin/env python
"""BASELINE configs[4]: the full synthetic global sweep over the extents of blocks/esa_extent_blocks.shp.

    python tools/global_sweep.py [--limit N]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/global_sweep.py                                                  # the per-GPU block queue on 8 GPUs

Every one of the 2651 block extents of the reference's shapefile (pinned in tests/golden/block_extents.json) goes
through the product's compressed chain with the REAL geometry of the reference's shipped configuration:

* the land cover is the mosaic of landcover/esa_worldcover_2021.vrt: 4 320 000 x 1 728 000 px at
  8.3333333333330430e-05 deg, one 36000 x 36000 GeoTIFF (1024 x 1024 DEFLATE tiles) per block extent, nothing where
  the shapefile has no block (ocean).  With that pixel size load_raster()'s window of a 3-degree block
  (/root/reference/src/raster.c:126-162, evaluated here by the host library's gh_raster_window) is 36001 x 36001
  and spills one pixel into the files to the east and south, so a block is assembled from up to FOUR sources:
  gcn10_cuda_block_parts_deflate with 1296 + 36 + 36 + 1 compressed tiles, window pixels of absent neighbours
  reading as the VRT's NoDataValue 0;
* the soils are the window of a global 1/480 deg HYSOGs grid (172 800 x 69 120 cells);
* 3.4 TB of land cover cannot be stored, so the tile FILES are synthesised: one 36000 x 36000 WorldCover-like raster
  is compressed once (zlib level 6, 1024 x 1024 tiles) and file k holds its tile grid rotated by a k-dependent number
  of whole tiles -- every block gets different land cover without compressing 2651 x 1.3 GB on the host; eight
  1442 x 1442 soil windows are used in turn.

Blocks are claimed from the dynamic queue (dist.BlockQueue: atomic fetch-and-add on the job's store, the
cross-process twin of the gcn10 executable's atomic counter; the reference's static round-robin is main.c:171), the
claimed next block is prefetched (upload + GPU inflate) beside the current one, and the compressed Curve Number
tiles of all 18 rasters arrive in page-locked host memory -- what the executable's sink appends to its GeoTIFFs.
A sampled tile row of the first block of every rank is checked against the CPU oracle.  One JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench as B  # noqa: E402
from gcn10_b200 import capi, hostlib, synth  # noqa: E402
from gcn10_b200 import dist as gdist  # noqa: E402
from tests import lookups  # noqa: E402

VRT_W, VRT_H = 4320000, 1728000
VRT_PX = 8.3333333333330430e-05
VRT_GT = (-180.0, VRT_PX, 0.0, 84.0, 0.0, -VRT_PX)
HSG_PX = 1.0 / 480.0
HSG_W, HSG_H = 360 * 480, 144 * 480
HSG_GT = (-180.0, HSG_PX, 0.0, 84.0, 0.0, -HSG_PX)
FILE_PX = 36000
T = 1024
NT = (FILE_PX + T - 1) // T            # 36 tiles per file side


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--limit", type=int, default=0, help="only the first N extents (0 = all 2651)")
    ap.add_argument("--planes", default="all", choices=["all", "drained"])
    ap.add_argument("--variants", type=int, default=8)
    ap.add_argument("--check", type=int, default=1, help="oracle check of one tile row of each rank's first block")
    a = ap.parse_args()
    # stdout carries exactly one JSON line: NCCL prints its version banner to stdout on the first collective
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank, local_rank, world = gdist.env_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = gdist.Group("nccl", device=dev)
    import tempfile
    lookup_dir = lookups.write_default_lookups(tempfile.mkdtemp(prefix="gcn10_lookups_"))
    tables = B.load_tables_host(lookup_dir)
    ctx = capi.Context(local_rank)
    ctx.set_luts(tables)
    lib = ctx.lib
    lib.gcn10_cuda_bind_host_thread(local_rank)
    mask = capi.MASK_ALL if a.planes == "all" else capi.MASK_DRAINED

    with open(os.path.join(ROOT, "tests", "golden", "block_extents.json")) as f:
        extents = [(int(i), float(w_), float(n_)) for i, w_, n_ in json.load(f)["blocks"]]
    if a.limit:
        extents = extents[:a.limit]
    present = {(w_, n_): i for i, w_, n_ in extents}          # which 3-degree cells have a file

    # ---- the one compressed file all tile files are rotations of
    t0 = time.time()
    base = synth.esa_tile(FILE_PX, FILE_PX, 2234, device=dev).cpu().numpy()

    def one(i):
        ty, tx = divmod(i, NT)
        t = np.zeros((T, T), dtype=np.uint8)
        part = base[ty * T:(ty + 1) * T, tx * T:(tx + 1) * T]
        t[:part.shape[0], :part.shape[1]] = part
        return zlib.compress(t.tobytes(), 6)

    with ThreadPoolExecutor(max_workers=max(4, (os.cpu_count() or 8) // world)) as ex:
        streams = list(ex.map(one, range(NT * NT)))
    sizes = np.array([len(z) for z in streams], dtype=np.uint32).reshape(NT, NT)
    offsets = np.concatenate([[0], np.cumsum(sizes.reshape(-1)[:-1], dtype=np.uint64)]).astype(np.uint64).reshape(NT, NT)
    total = int(sizes.sum())
    blob_pin = capi.PinnedArray(lib, (total,))
    pos = 0
    for z in streams:
        blob_pin.array[pos:pos + len(z)] = np.frombuffer(z, dtype=np.uint8)
        pos += len(z)
    del streams
    blob = blob_pin.array
    t_setup = time.time() - t0

    def file_shift(bid):
        v = bid % a.variants
        return (5 * v) % NT, (11 * v) % NT                    # (dy, dx) in whole tiles

    def file_tables(bid, ty0, ty1, tx0, tx1):
        """offsets / sizes of tile rows [ty0, ty1) x columns [tx0, tx1) of block bid's file."""
        dy, dx = file_shift(bid)
        rows = (np.arange(ty0, ty1) - dy) % NT
        cols = (np.arange(tx0, tx1) - dx) % NT
        return (np.ascontiguousarray(offsets[np.ix_(rows, cols)]).reshape(-1),
                np.ascontiguousarray(sizes[np.ix_(rows, cols)]).reshape(-1))

    def file_pixels(bid, y0, y1, x0, x1):
        """decoded pixels [y0, y1) x [x0, x1) of block bid's file (for the oracle check)."""
        dy, dx = file_shift(bid)
        out = np.zeros((y1 - y0, x1 - x0), dtype=np.uint8)
        for ty in range(y0 // T, (y1 - 1) // T + 1):
            for tx in range(x0 // T, (x1 - 1) // T + 1):
                sy, sx = (ty - dy) % NT, (tx - dx) % NT
                tile = np.zeros((T, T), dtype=np.uint8)
                part = base[sy * T:(sy + 1) * T, sx * T:(sx + 1) * T]
                tile[:part.shape[0], :part.shape[1]] = part
                ya, yb = max(y0, ty * T), min(y1, (ty + 1) * T)
                xa, xb = max(x0, tx * T), min(x1, (tx + 1) * T)
                out[ya - y0:yb - y0, xa - x0:xb - x0] = tile[ya - ty * T:yb - ty * T, xa - tx * T:xb - tx * T]
        return out

    hsgs = [np.ascontiguousarray(synth.hsg_tile(1442, 1442, 7000 + v)) for v in range(a.variants)]

    keep = []                                                  # ctypes objects of blocks in flight

    def prepare(k):
        """geometry + part list of extent k, as load_raster() over the VRT and the HYSOGs grid would give them"""
        bid, west, north = extents[k]
        bbox = (west, north - 3.0, west + 3.0, north)
        xo, yo, w, h, gt = hostlib.raster_window(VRT_W, VRT_H, VRT_GT, bbox)
        hxo, hyo, hsx, hsy, sgt = hostlib.raster_window(HSG_W, HSG_H, HSG_GT, bbox)
        parts = []
        # the files are placed on the 36000-pixel grid of the VRT (DstRect of source (cx, cy) = (36000 cx, 36000 cy))
        for cy in range(yo // FILE_PX, (yo + h - 1) // FILE_PX + 1):
            for cx in range(xo // FILE_PX, (xo + w - 1) // FILE_PX + 1):
                fwest, fnorth = -180.0 + 3.0 * cx, 84.0 - 3.0 * cy
                fid = present.get((fwest, fnorth))
                if fid is None:
                    continue                                   # no file: NoDataValue
                x0, x1 = max(xo, cx * FILE_PX), min(xo + w, (cx + 1) * FILE_PX)
                y0, y1 = max(yo, cy * FILE_PX), min(yo + h, (cy + 1) * FILE_PX)
                sx0, sy0 = x0 - cx * FILE_PX, y0 - cy * FILE_PX                  # inside the file
                tx0, tx1 = sx0 // T, (sx0 + (x1 - x0) - 1) // T + 1
                ty0, ty1 = sy0 // T, (sy0 + (y1 - y0) - 1) // T + 1
                o, z = file_tables(fid, ty0, ty1, tx0, tx1)
                src = capi.TileSource(T, T, tx1 - tx0, ty1 - ty0, sx0 - tx0 * T, sy0 - ty0 * T, blob, o, z)
                parts.append((src, x0 - xo, y0 - yo, x1 - x0, y1 - y0, fid, sx0, sy0))
        arr = capi.parts_array([p[:5] for p in parts])
        hs = hsgs[(k // a.variants) % a.variants]
        hs = np.ascontiguousarray(hs[:hsy, :hsx])
        return dict(k=k, bid=bid, w=w, h=h, gt=(C.c_double * 6)(*gt), sgt=(C.c_double * 6)(*sgt), hsx=hsx, hsy=hsy,
                    hsg=hs, parts=parts, arr=arr, n=len(parts), gt_t=gt, sgt_t=sgt)

    nbytes = [0, 0]
    captured = {}

    def _sink(_user, sp):
        st = sp.contents
        nbytes[0] += st.blob_bytes
        nbytes[1] += st.n_planes * st.n_tile_rows * st.tiles_x * 12 + 16
        if captured.get("want") is not None and st.tile_row0 <= captured["want"] < st.tile_row0 + st.n_tile_rows:
            tr = captured["want"] - st.tile_row0
            blob_ = C.string_at(st.blob, st.blob_bytes)
            rows = {}
            for j in range(st.n_planes):
                i0 = (j * st.n_tile_rows + tr) * st.tiles_x
                rows[st.plane_ids[j]] = [blob_[st.offsets[i0 + tx]: st.offsets[i0 + tx] + st.sizes[i0 + tx]]
                                         for tx in range(st.tiles_x)]
            captured["tiles"] = rows
        return 0

    cb = capi.TILE_SINK(_sink)

    def prefetch(b):
        if b["n"] and lib.gcn10_cuda_parts_prefetch(ctx.h, b["arr"], b["n"], 0, b["w"], b["h"]):
            raise RuntimeError(lib.gcn10_cuda_last_error().decode())

    def process(b):
        if not b["n"]:
            return False                                        # (cannot happen: every extent has its own file)
        rc = lib.gcn10_cuda_block_parts_deflate(ctx.h, b["arr"], b["n"], 0, b["w"], b["h"], b["gt"], b["hsg"].ctypes.data,
                                                b["hsx"], b["hsy"], b["hsx"], b["sgt"], mask, cb, None)
        if rc:
            raise RuntimeError(lib.gcn10_cuda_last_error().decode())
        return True

    # warm-up (allocations, first-launch costs) outside the timed sweep
    wq = gdist.BlockQueue(group, min(len(extents), world), "warm")
    i = wq.claim()
    if i is not None:
        process(prepare(i))
    group.barrier()

    q = gdist.BlockQueue(group, len(extents), "sweep")
    nbytes[0] = nbytes[1] = 0
    shapes = {}
    nparts_hist = {}
    px_done = 0
    checked = None
    t0 = time.perf_counter()
    cur = q.claim()
    cur_b = prepare(cur) if cur is not None else None
    if cur_b is not None:
        prefetch(cur_b)
    n_mine = 0
    while cur_b is not None:
        nxt = q.claim()
        nxt_b = prepare(nxt) if nxt is not None else None
        if nxt_b is not None:
            prefetch(nxt_b)
        do_check = a.check and n_mine == 0
        if do_check:
            captured["want"] = (cur_b["h"] - 1) // 256 // 2     # a tile row in the middle of the block
        process(cur_b)
        if do_check:
            checked = (cur_b, captured.pop("tiles", None), captured.pop("want"))
        shapes[(cur_b["w"], cur_b["h"])] = shapes.get((cur_b["w"], cur_b["h"]), 0) + 1
        nparts_hist[cur_b["n"]] = nparts_hist.get(cur_b["n"], 0) + 1
        px_done += cur_b["w"] * cur_b["h"]
        n_mine += 1
        cur_b = nxt_b
    group.barrier()
    dt = group.max(time.perf_counter() - t0)

    # ---- the oracle check (outside the timed region)
    check_msg = None
    if checked and checked[1]:
        from oracle import oracle as O
        b, rows, tr = checked
        port = O.Port()
        y0, y1 = tr * 256, min(b["h"], tr * 256 + 256)
        esa = np.zeros((y1 - y0, b["w"]), dtype=np.uint8)
        for src, dx, dy, pw, ph, fid, sx0, sy0 in b["parts"]:
            ya, yb = max(y0, dy), min(y1, dy + ph)
            if ya < yb:
                esa[ya - y0:yb - y0, dx:dx + pw] = file_pixels(fid, sy0 + ya - dy, sy0 + yb - dy, sx0, sx0 + pw)
        want = port.block_rows(esa, b["gt_t"], b["hsg"], b["sgt_t"], tables, y0=y0, y1=y1, h=b["h"])
        bad = 0
        for k_, tiles in rows.items():
            raw = b"".join(zlib.decompress(z) for z in tiles)
            band = np.frombuffer(raw, dtype=np.uint8).reshape(len(tiles), 256, 256).transpose(1, 0, 2).reshape(256, -1)
            bad += int((band[:y1 - y0, :b["w"]] != want[k_]).sum())
        check_msg = f"rank {rank}: block {b['bid']} tile row {tr}: {len(rows)} planes x {b['w']} px, {bad} bytes differ"
        if bad:
            raise SystemExit("global_sweep: " + check_msg)

    counts = group.gather_ints(n_mine)
    total_px = group.sum(px_done)
    d2h = group.sum(nbytes[0] + nbytes[1])
    shape_list = sorted((f"{w}x{h}", n) for (w, h), n in shapes.items())
    if rank == 0:
        print(json.dumps({
            "workload": "BASELINE configs[4]: synthetic global sweep over the extents of esa_extent_blocks.shp, VRT geometry "
                        "(36001 x 36001 windows assembled from up to four tile files), compressed tiles in and out",
            "blocks": len(extents), "gpus": world, "blocks_per_rank": counts, "seconds": dt,
            "value": total_px / dt / 1e9, "unit": "Gpixel/s", "planes": 18 if mask == capi.MASK_ALL else 9,
            "ms_per_block": dt / max(1, len(extents)) * 1e3 * world, "pixels": total_px,
            "d2h_bytes_total": int(d2h), "h2d_bytes_per_block": int(total + 2 * 36 * 22000),
            "window_shapes_rank0": shape_list, "parts_per_block_rank0": sorted(nparts_hist.items()),
            "setup_seconds": t_setup, "oracle_check": check_msg,
        }), file=real_stdout, flush=True)
    ctx.close()
    blob_pin.free()
    group.close()


if __name__ == "__main__":
    main()
