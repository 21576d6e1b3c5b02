#!/usr/bin/env python
"""Throughput of gcn10_cuda_block_tiles_deflate with several blocks in flight on ONE GPU (development tool).

Each in-flight block has its own gcn10_ctx (its own streams and device buffers) and its own host thread, the
way the host program runs several workers per GPU; the kernels of different contexts overlap on the device
(the inflate kernel and the fused Curve Number + DEFLATE kernel are both latency bound, not issue bound).

    python tools/inflight_bench.py --inflight 1,2,3 --blocks 12
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gcn10_b200 import capi, synth
from tests import lookups  # noqa: E402
import bench as B  # noqa: E402
from tools.inflate_bench import compress_tiles  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=36000)
    ap.add_argument("--in-tile", type=int, default=1024)
    ap.add_argument("--inflight", default="1,2,3")
    ap.add_argument("--blocks", type=int, default=12)
    ap.add_argument("--profile", default="worldcover")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    tables = B.load_tables_host(lookups.write_default_lookups(tempfile.mkdtemp()))
    w = h = a.tile
    gt, sgt, hsx, hsy = synth.block_geometry(-114.0, 42.0, w, h)
    lib = capi.load()
    pins = []

    def pinned(n):
        p = capi.PinnedArray(lib, (max(n, 1),))
        pins.append(p)
        return p.array

    esa = synth.esa_tile(w, h, 2234, a.profile, device=dev).cpu().numpy()
    hsg = synth.hsg_tile(hsx, hsy, 3234, a.profile)
    boot = capi.Context(0, lib)
    src = compress_tiles(esa, a.in_tile, a.in_tile, 6, pinned)
    del esa
    for nfl in [int(x) for x in a.inflight.split(",")]:
        ctxs = [capi.Context(0, lib) for _ in range(nfl)]
        for c in ctxs:
            c.set_luts(tables)
            c.block_tiles_deflate(src, w, h, gt, hsg, sgt, capi.MASK_DRAINED, on_strip=lambda st: 0)    # warm-up
        nxt = [0]
        lock = threading.Lock()

        def worker(c):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= a.blocks:
                    return
                c.block_tiles_deflate(src, w, h, gt, hsg, sgt, capi.MASK_DRAINED, on_strip=lambda st: 0)

        th = [threading.Thread(target=worker, args=(c,)) for c in ctxs]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"inflight": nfl, "blocks": a.blocks, "ms_per_block": round(dt / a.blocks * 1e3, 2),
                          "gpx_s": round(a.blocks * w * h / dt / 1e9, 1)}), flush=True)
        for c in ctxs:
            c.close()
    boot.close()
    for p in pins:
        p.free()


if __name__ == "__main__":
    main()
