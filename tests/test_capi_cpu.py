"""The C-ABI library loads without a GPU, exports every symbol include/dspmap_c.h declares, and refuses to run without
a device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from semantic_dsp_map_b200 import BASELINE, _capi, tiny

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_header_symbols_all_exported():
    header = open(os.path.join(ROOT, "include", "dspmap_c.h")).read()
    declared = set(re.findall(r"\b(dspmap_[a-z_]+)\s*\(", header))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = _capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dspmap_version() >= 100


def test_struct_sizes_match_header_layout():
    # natural alignment, as a C compiler lays the header's structs out
    assert C.sizeof(_capi.Config) == 22 * 4
    assert C.sizeof(_capi.Params) == 15 * 4
    assert C.sizeof(_capi.Mask) == 8 + 4 * 4 + 6 * 8 + 2 * 4
    assert C.sizeof(_capi.ObjectMove) == 4 + 16 * 4
    assert C.sizeof(_capi.FrameStats) == 12 * 4 + 8 * 4


def test_config_validation_and_no_cpu_fallback():
    lib = _capi.load()
    h = C.c_void_p()
    bad = tiny(BASELINE["B"], voxel_bits=(10, 10, 10), slot_bits=4).c_config()   # 34 index bits (operations.h:55-58)
    assert lib.dspmap_create(C.byref(bad), C.byref(h)) == -2
    assert not h
    assert lib.dspmap_create(None, C.byref(h)) == -1
    if not _has_gpu():
        ok = tiny(BASELINE["B"]).c_config()
        assert lib.dspmap_create(C.byref(ok), C.byref(h)) == -6   # DSPMAP_ERR_NO_DEVICE
        assert not h


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "semantic_dsp_map_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "libdsp_oracle" not in text and "orc_" not in text, f
