"""Readers for the committed golden fixtures (tests/golden/, made by make_golden.py from the
reference's own object code)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def block_cases():
    z = np.load(os.path.join(GOLDEN, "blocks.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names}


def window_cases():
    with open(os.path.join(GOLDEN, "windows.json")) as f:
        return json.load(f)


def luts():
    z = np.load(os.path.join(GOLDEN, "luts.npz"))
    return {k: z[k] for k in z.files}


def block_extents():
    """[(id, west, north)] of the 2651 blocks of the reference's shapefile."""
    with open(os.path.join(GOLDEN, "block_extents.json")) as f:
        return [tuple(r) for r in json.load(f)["blocks"]]
