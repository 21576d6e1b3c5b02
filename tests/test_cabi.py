"""The C-ABI shared library: loads without a GPU, exports exactly what include/gcn10_cuda.h declares,
fails loudly (no CPU fallback) when no device is usable, and validates arguments."""
import ctypes as C
import os
import re
import subprocess

import pytest

from gcn10_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gcn10_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcn10_cuda_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    for name in _declared():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gcn10_cuda_\w+)", out))
    assert exported == set(_declared()), exported ^ set(_declared())


def test_version_string():
    lib = capi.load()
    assert b"sm_100a" in lib.gcn10_cuda_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.load()
    assert lib.gcn10_cuda_device_count() < 0
    with pytest.raises(capi.Gcn10Error) as e:
        capi.Context(0, lib)
    assert e.value.code in (-5, -2)
    assert lib.gcn10_cuda_last_error() != b""


def test_null_arguments_are_rejected():
    lib = capi.load()
    assert lib.gcn10_cuda_create(0, None) == -1
    assert lib.gcn10_cuda_set_luts(None, None) == -1
    assert lib.gcn10_cuda_synchronize(None) == -1
    assert lib.gcn10_cuda_set_option(None, b"tma", 1) == -1
    assert lib.gcn10_cuda_host_register(None, 0) == -1
    lib.gcn10_cuda_destroy(None)        # must be a no-op


def test_missing_library_raises(tmp_path):
    with pytest.raises(FileNotFoundError):
        capi.load(str(tmp_path / "libgcn10cuda.so"))


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under gcn10_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gcn10_b200")):
        for fn in files:
            if fn.endswith((".py", ".c", ".h", ".cu", ".cuh", ".cpp")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                if re.search(r"\boracle\b", txt) and "oracle" in txt and ("import" in txt or "#include" in txt):
                    for ln in txt.splitlines():
                        if re.search(r"(from|import)\s+oracle|#include\s+[\"<].*oracle|oracle/", ln):
                            bad.append((fn, ln.strip()))
    assert not bad, bad
