"""The C-ABI library loads and exports every symbol include/ntlink_b200.h declares (no compute: CPU only)."""
import ctypes
import os
import re

import pytest

import util
from ntlink_b200 import _lib, build

HEADER = os.path.join(util.REPO, "include", "ntlink_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ntl_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signatures out of sync with the header"


def test_no_cpu_fallback():
    "without a CUDA device the product refuses to initialise instead of computing on the host"
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.ntl_init(0, ctypes.byref(h)) == -1
    assert not h.value
    from ntlink_b200 import Context, NtlError
    with pytest.raises(NtlError):
        Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(util.REPO, "ntlink_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "pair_oracle" not in text and "indexlr_oracle" not in text and "oracle/" not in text, f
