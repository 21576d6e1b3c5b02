#!/usr/bin/env python
"""Kernel-only sweep of the fused Curve Number kernel (device-resident buffers, CUDA events).

    python tools/kbench.py --rows-per-cta 32,64,128,256 --profile worldcover,random --planes 9,18

Prints one line per configuration: ms per launch, Gpixel/s, algorithmic GB/s and fraction of the
measured HBM copy peak.  Development tool; bench.py is the judged benchmark.
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from gcn10_b200 import capi, synth
from tests import lookups  # noqa: E402
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=36000)
    ap.add_argument("--rows-per-cta", default="128")
    ap.add_argument("--profile", default="worldcover")
    ap.add_argument("--planes", default="9")
    ap.add_argument("--tma", default="1")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--once", action="store_true", help="one launch per config, no timing loop (for ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    tables = B.load_tables_host(lookups.write_default_lookups(tempfile.mkdtemp()))
    ctx = capi.Context(0)
    ctx.set_luts(tables)
    w = h = a.tile
    gt, sgt, hsx, hsy = synth.block_geometry(-114.0, 42.0, w, h)
    peak, _ = B.measured_peak()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    d_out = torch.empty((18, h, w), dtype=torch.uint8, device=dev)
    for prof in a.profile.split(","):
        d_esa = synth.esa_tile(w, h, 2234, prof, device=dev)
        d_hsg = torch.from_numpy(synth.hsg_tile(hsx, hsy, 3234, prof)).to(dev)
        for planes in [int(x) for x in a.planes.split(",")]:
            mask = {1: 1 << 7, 9: capi.MASK_DRAINED, 18: capi.MASK_ALL}[planes]
            ptrs = [d_out[k].data_ptr() for k in range(18)]
            for tma in [int(x) for x in a.tma.split(",")]:
                ctx.set_option("tma", tma)
                for rpc in [int(x) for x in a.rows_per_cta.split(",")]:
                    ctx.set_option("rows_per_cta", rpc)

                    def step():
                        ctx.block_device(d_esa.data_ptr(), w, h, w, gt, d_hsg.data_ptr(), hsx, hsy, hsx, sgt,
                                         mask, ptrs, w, stream=stream.cuda_stream)
                    if a.once:
                        step()
                        torch.cuda.synchronize()
                        continue
                    for _ in range(3):
                        step()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(a.reps):
                        e0 = torch.cuda.Event(enable_timing=True)
                        e1 = torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        step()
                        e1.record(stream)
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ts.sort()
                    ms = ts[len(ts) // 2]
                    by = float(w) * h * (1 + planes) + hsx * hsy
                    print(json.dumps({"profile": prof, "planes": planes, "tma": tma, "rows_per_cta": rpc,
                                      "ms_med": round(ms, 4), "ms_min": round(ts[0], 4),
                                      "gpx_s": round(w * h / ms / 1e6, 1), "gb_s": round(by / ms / 1e6, 1),
                                      "frac_peak": round(by / ms / 1e6 / peak, 4)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
