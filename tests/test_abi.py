"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/rcd.h declares, struct layouts match, and the product never reaches into oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "realtime-collision-detection_b200")


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rcd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcd_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rcd_b200.host import _native as N
    lib = N.load()
    declared = _header_symbols()
    assert len(declared) >= 15
    assert sorted(N.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rcd_version() == 100


def test_struct_layouts_match_header():
    from rcd_b200.host import _native as N
    assert ctypes.sizeof(N.RcdConfig) == 4 + 4 + 8 + 8 + 12 + 12
    assert ctypes.sizeof(N.RcdCounts) == 8 * 13
    assert N.PAIR_DTYPE.itemsize == 48
    assert N.PAIR_DTYPE.fields["priority"][1] == 44 and N.PAIR_DTYPE.fields["d_closest"][1] == 40


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rcd_b200.host import _native as N
    from rcd_b200.host.engine import FrameEngine
    with pytest.raises(N.NativeError) as e:
        FrameEngine(1000)
    assert e.value.code == N.RCD_ENODEVICE
    assert "no CPU fallback" in str(e.value)


def test_null_and_bad_arguments_return_codes_not_crashes():
    from rcd_b200.host import _native as N
    lib = N.load()
    assert lib.rcd_create(None, None) == N.RCD_EINVAL
    h = ctypes.c_void_p()
    cfg = N.RcdConfig()
    cfg.max_objects = 0
    assert lib.rcd_create(ctypes.byref(cfg), ctypes.byref(h)) == N.RCD_EINVAL
    assert b"max_objects" in lib.rcd_last_error(None)
    assert lib.rcd_destroy(None) == N.RCD_OK
    assert lib.rcd_step(None, 0, 100.0, 10.0) == N.RCD_EINVAL
    assert lib.rcd_sync(None) == N.RCD_EINVAL


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package, nor the alias, may use it."""
    offenders = []
    for base in (PKG, os.path.join(ROOT, "rcd_b200")):
        for dirpath, _dirs, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/", text, flags=re.M):
                        offenders.append(os.path.join(dirpath, f))
    assert offenders == []
