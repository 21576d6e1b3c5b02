#!/usr/bin/env python
"""Mutation fuzzing of everything the host program parses from disk (development tool, CPU only).

    python tools/fuzz_host_inputs.py --seeds 1-12 --count 150 [--no-asan]

The input files of a run are not trusted: GeoTIFFs (own tiled DEFLATE files, libtiff LZW and raw strips), VRT mosaics,
the block shapefile (.shp / .dbf), the configuration file, the lookup CSVs and the block list.  Every mutant (random
bytes, 32-bit boundary values, truncations, insertions; biased towards headers and directories) is opened and read
through gcn10_b200/hostlib in a child process, against a build of the host library with AddressSanitizer and UBSan
(built here with gcc into the work directory), so that out-of-bounds reads that do not happen to crash are seen too.
A child that dies or times out is a finding: the mutant is kept next to the log line.  The findings of round 2
(zero tile sizes, directory and table counts sizing allocations, dBASE header / record lengths reaching past the file)
are fixed and pinned by tests/test_host.py::test_*_survives_damaged_*.
"""
import argparse
import os
import random
import shutil
import struct
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HOST = os.path.join(ROOT, "gcn10_b200", "host")


def child(path, kind):
    import ctypes as C
    from gcn10_b200 import hostlib
    if os.environ.get("FUZZ_HOST_LIB"):
        hostlib._lib = hostlib.load(os.environ["FUZZ_HOST_LIB"])
    L = hostlib.load()
    try:
        if kind in ("tif", "vrt"):
            r = hostlib.Raster(path)
            if 0 < r.width <= 5000 and 0 < r.height <= 5000:
                r.read(0, 0, r.width, r.height)
                r.window_parts(0, 0, r.width, r.height)
            r.close()
        elif kind == "shp":
            b = hostlib.Blocks(path)
            b.ids()
            b.bbox(2234)
            b.close()
        elif kind == "cfg":
            cfg = hostlib.Config()
            e = C.create_string_buffer(512)
            if L.gh_config_parse(path.encode(), C.byref(cfg), e, 512) == 0:
                L.gh_config_free(C.byref(cfg))
        elif kind == "csv":
            hostlib.load_lookup_tables(os.path.dirname(path))
        elif kind == "lst":
            hostlib.read_block_list(path)
    except hostlib.HostError:
        pass


def mutate(rng, data):
    data = bytearray(data)
    mode = rng.random()
    for _ in range(rng.randint(1, 8)):
        n = len(data)
        if n < 8:
            break
        if mode < 0.5:
            pick = rng.random()
            pos = rng.randrange(n) if pick < 0.5 else rng.randrange(min(n, 400)) if pick < 0.75 else n - 1 - rng.randrange(min(n, 600))
            data[pos] = rng.randrange(256)
        elif mode < 0.7:
            pos = rng.randrange(n - 4)
            data[pos:pos + 4] = struct.pack("<I", rng.choice([0, 1, 0xFFFFFFFF, 0x7FFFFFFF, 0x80000000, rng.randrange(1 << 32)]))
        elif mode < 0.85:
            data = data[:max(1, rng.randrange(n))]
        else:
            pos = rng.randrange(n)
            data[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 64)))
    return bytes(data)


def run_seed(seed, count, work, env):
    import numpy as np
    from PIL import Image
    from gcn10_b200 import hostlib
    from tests import fixtures, lookups
    rng = random.Random(seed)
    base = os.path.join(work, f"base_{seed}")
    os.makedirs(base, exist_ok=True)
    a = (np.arange(300 * 520).reshape(300, 520) % 7 * 10).astype(np.uint8)
    gt = (-114.0, 1 / 12000, 0.0, 42.0, 0.0, -1 / 12000)
    hostlib.tiff_write(f"{base}/base.tif", a, gt, threads=2)
    Image.fromarray(a).save(f"{base}/lzw.tif", compression="tiff_lzw")
    Image.fromarray(a).save(f"{base}/raw.tif")
    fixtures.write_vrt(f"{base}/base.vrt", 520, 300, gt, [("base.tif", 0, 0, 0, 0, 520, 300)])
    fixtures.write_block_shapefile(f"{base}/base.shp", [(2234, -114.0, 39.0, -111.0, 42.0), (7, 0.0, 0.0, 3.0, 3.0)])
    lookups.write_default_lookups(f"{base}/lk")
    fixtures.write_config(f"{base}/base.cfg", "a.tif", "b.tif", "c.shp", "lk", "logs")
    with open(f"{base}/base.lst", "w") as f:
        f.write("11\n12\n 13 \n\n99\n")

    def rd(name):
        with open(f"{base}/{name}", "rb") as f:
            return f.read()

    src = {"tif": [rd("base.tif"), rd("lzw.tif"), rd("raw.tif")], "vrt": [rd("base.vrt")], "shp": [rd("base.shp")],
           "dbf": [rd("base.dbf")], "cfg": [rd("base.cfg")], "csv": [rd("lk/default_lookup_g_ii.csv")], "lst": [rd("base.lst")]}
    findings = []
    d = os.path.join(work, f"case_{seed}")
    for i in range(count):
        kind = rng.choice(["tif", "tif", "vrt", "vrt", "shp", "dbf", "cfg", "csv", "lst"])
        data = mutate(rng, rng.choice(src[kind]))
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        ck = kind
        if kind == "csv":
            shutil.copytree(f"{base}/lk", f"{d}/lk")
            p = f"{d}/lk/default_lookup_g_ii.csv"
        elif kind in ("shp", "dbf"):
            p, ck = f"{d}/f.shp", "shp"
            for ext in ("shp", "shx", "dbf"):
                shutil.copy(f"{base}/base.{ext}", f"{d}/f.{ext}")
        else:
            p = f"{d}/f.{kind}"
            if kind == "vrt":
                shutil.copy(f"{base}/base.tif", f"{d}/base.tif")
        with open(f"{d}/f.dbf" if kind == "dbf" else p, "wb") as f:
            f.write(data)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", p, ck], capture_output=True,
                               timeout=120, env=env)
            why = None if r.returncode == 0 else f"rc={r.returncode} " + r.stderr.decode(errors="replace")[-1500:]
        except subprocess.TimeoutExpired:
            why = "timeout"
        if why:
            keep = os.path.join(work, f"finding_{seed}_{i}.{kind}")
            with open(keep, "wb") as f:
                f.write(data)
            findings.append((keep, why))
    return seed, findings


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", nargs=2, metavar=("PATH", "KIND"))
    ap.add_argument("--seeds", default="1-8")
    ap.add_argument("--count", type=int, default=100)
    ap.add_argument("--no-asan", action="store_true")
    ap.add_argument("--work", default=None)
    a = ap.parse_args()
    if a.child:
        child(*a.child)
        return 0
    work = a.work or tempfile.mkdtemp(prefix="gcn10_fuzz_")
    os.makedirs(work, exist_ok=True)
    env = dict(os.environ)
    if not a.no_asan:
        lib = os.path.join(work, "libgcn10host_asan.so")
        srcs = [os.path.join(HOST, f) for f in ("host_core.c", "host_tiff.c", "host_raster.c", "host_raster_gdal.c")]
        subprocess.run(["gcc", "-std=c11", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-pthread", "-fsanitize=address,undefined",
                        "-fno-sanitize-recover=undefined", "-shared", "-o", lib, *srcs, "-lz", "-lm"], check=True)
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
        env.update(FUZZ_HOST_LIB=lib, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", PYTHONMALLOC="malloc")
    lo, _, hi = a.seeds.partition("-")
    seeds = list(range(int(lo), int(hi or lo) + 1))
    total = 0
    with ThreadPoolExecutor(max_workers=min(len(seeds), os.cpu_count() or 4)) as ex:
        for seed, findings in ex.map(lambda s: run_seed(s, a.count, work, env), seeds):
            for keep, why in findings:
                print(f"FINDING seed {seed}: {keep}\n    " + why.replace("\n", "\n    ")[-600:])
            total += len(findings)
    print(f"{len(seeds)} seeds x {a.count} mutants, {total} findings; work directory {work}")
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
