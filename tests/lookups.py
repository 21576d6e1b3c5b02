"""Default Curve Number lookup data and CSV emission.

GCN10 reads nine CSV files ``default_lookup_<hc>_<arc>.csv`` (hc in p,f,g = poor/fair/good
hydrologic condition; arc in i,ii,iii = antecedent runoff condition) from the configured
``lookup_table_path`` (/root/reference/src/cn.c:21-22).  Each has a header line
``grid_code,cn`` and 44 rows ``<ESA class>_<A|B|C|D>,<cn>`` (/root/reference/lookups/).

The reference tree is not available where the tests and the benchmark run, so the default
tables are kept here as one matrix (values as shipped in /root/reference/lookups/*.csv,
column order = the reference's loop order p_i, p_ii, p_iii, f_i, ... g_iii, cn.c:146-147)
and written back out as CSV files on request, with the same byte-level quirks as the shipped
files (seven of nine carry a UTF-8 BOM and CRLF line ends; p_i and p_ii are plain LF) so that
the host parser is exercised on exactly what a user of the reference has on disk.
"""
from __future__ import annotations

import os

HCS = ("p", "f", "g")
ARCS = ("i", "ii", "iii")
VARIANTS = tuple(f"{h}_{a}" for h in HCS for a in ARCS)
ESA_CLASSES = (10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 100)
HSG_LETTERS = ("A", "B", "C", "D")

# variant files that are plain LF without BOM in the reference tree; the others are BOM + CRLF
_PLAIN_LF = ("p_i", "p_ii")

DEFAULT_CN = {
    "10_A": (45, 26, 65, 19, 36, 56, 30, 15, 50),
    "10_B": (66, 46, 82, 40, 60, 78, 55, 35, 74),
    "10_C": (77, 59, 89, 54, 73, 87, 70, 51, 85),
    "10_D": (83, 67, 93, 62, 79, 91, 77, 59, 89),
    "20_A": (63, 43, 80, 35, 55, 74, 49, 30, 69),
    "20_B": (77, 59, 89, 53, 72, 86, 68, 48, 84),
    "20_C": (85, 70, 94, 64, 81, 92, 79, 62, 91),
    "20_D": (88, 75, 95, 72, 86, 94, 84, 68, 93),
    "30_A": (68, 48, 84, 30, 49, 69, 39, 21, 59),
    "30_B": (79, 62, 91, 50, 69, 84, 61, 41, 78),
    "30_C": (86, 72, 94, 62, 79, 91, 74, 55, 88),
    "30_D": (89, 76, 96, 68, 84, 93, 80, 63, 91),
    "40_A": (72, 53, 86, 70, 50, 85, 67, 47, 83),
    "40_B": (81, 64, 92, 80, 62, 91, 78, 60, 90),
    "40_C": (88, 75, 95, 87, 73, 95, 85, 70, 94),
    "40_D": (91, 80, 97, 90, 78, 97, 89, 76, 96),
    "50_A": (89, 76, 96, 76, 89, 96, 89, 76, 96),
    "50_B": (92, 81, 97, 81, 92, 97, 92, 81, 97),
    "50_C": (94, 85, 98, 85, 94, 98, 94, 85, 98),
    "50_D": (95, 87, 98, 87, 95, 98, 95, 87, 98),
    "60_A": (65, 45, 82, 45, 65, 82, 65, 45, 82),
    "60_B": (79, 62, 91, 62, 79, 91, 79, 62, 91),
    "60_C": (87, 73, 95, 73, 87, 95, 87, 73, 95),
    "60_D": (90, 78, 96, 78, 90, 96, 90, 78, 96),
    "70_A": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "70_B": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "70_C": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "70_D": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "80_A": (100, 100, 100, 100, 100, 100, 100, 100, 100),
    "80_B": (100, 100, 100, 100, 100, 100, 100, 100, 100),
    "80_C": (100, 100, 100, 100, 100, 100, 100, 100, 100),
    "80_D": (100, 100, 100, 100, 100, 100, 100, 100, 100),
    "90_A": (80, 63, 91, 63, 80, 91, 80, 63, 91),
    "90_B": (80, 63, 91, 63, 80, 91, 80, 63, 91),
    "90_C": (80, 63, 91, 63, 80, 91, 80, 63, 91),
    "90_D": (80, 63, 91, 63, 80, 91, 80, 63, 91),
    "95_A": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "95_B": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "95_C": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "95_D": (0, 0, 0, 0, 0, 0, 0, 0, 0),
    "100_A": (74, 55, 88, 55, 74, 88, 74, 55, 88),
    "100_B": (77, 59, 89, 59, 77, 89, 77, 59, 89),
    "100_C": (78, 60, 90, 60, 78, 90, 78, 60, 90),
    "100_D": (79, 62, 91, 62, 79, 91, 79, 62, 91),
}


def lookup_csv_bytes(variant: str, rows=None) -> bytes:
    """Bytes of ``default_lookup_<variant>.csv`` for the default matrix (or custom rows)."""
    col = VARIANTS.index(variant)
    rows = DEFAULT_CN if rows is None else rows
    lines = ["grid_code,cn"] + [f"{code},{vals[col]}" for code, vals in rows.items()]
    if variant in _PLAIN_LF:
        return ("\n".join(lines) + "\n").encode("ascii")
    return b"\xef\xbb\xbf" + ("\r\n".join(lines) + "\r\n").encode("ascii")


def write_default_lookups(directory: str, rows=None) -> str:
    """Write the nine default lookup CSVs into ``directory`` and return it."""
    os.makedirs(directory, exist_ok=True)
    for v in VARIANTS:
        with open(os.path.join(directory, f"default_lookup_{v}.csv"), "wb") as f:
            f.write(lookup_csv_bytes(v, rows))
    return directory
