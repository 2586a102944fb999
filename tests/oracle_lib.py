"""Loader for the CPU oracle (oracle/libdsp_oracle.so).  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU
baseline legs may import this module; the product package never does."""
import ctypes as C
import os
import subprocess

import numpy as np

from semantic_dsp_map_b200 import _capi
from semantic_dsp_map_b200.map import Engine, SemanticDSPMap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libdsp_oracle.so")

_ORC_SYMBOLS = {k: v for k, v in _capi.SYMBOLS.items()
                if k in ("dspmap_clear", "dspmap_set_params", "dspmap_set_label_luts", "dspmap_set_random_table", "dspmap_set_colors", "dspmap_update",
                         "dspmap_get_result", "dspmap_get_stats", "dspmap_dump")}
_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            subprocess.check_call(["make", "-C", ORACLE_DIR])
        lib = C.CDLL(ORACLE_LIB)
        _capi.bind(lib, _ORC_SYMBOLS, "dspmap_", "orc_")
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(_capi.Config), C.c_int]
        lib.orc_destroy.restype = None
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_get_extra.restype = C.c_uint64
        lib.orc_get_extra.argtypes = [C.c_void_p, C.c_int]
        lib.orc_set_diagnostics.restype = None
        lib.orc_set_diagnostics.argtypes = [C.c_void_p, C.c_int]
        lib.orc_set_undefined_label_carry.restype = None
        lib.orc_set_undefined_label_carry.argtypes = [C.c_void_p, C.c_int]
        for fn in (lib.orc_rgb_to_hsv, lib.orc_hsv_to_rgb):
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orc_set_invalid_sigma.restype = None
        lib.orc_set_invalid_sigma.argtypes = [C.c_void_p, C.c_float]
        lib.orc_stage_seconds.restype = C.c_double
        lib.orc_stage_seconds.argtypes = [C.c_void_p, C.c_int]
        _lib = lib
    return _lib


class OracleEngine(Engine):
    """Same surface as the CUDA Engine, backed by the oracle.  order: 1 canonical (what the engine matches), 0 reference BFS."""

    def __init__(self, cfg, order=1):
        self.order = order
        super().__init__(load_oracle(), "orc_", cfg)

    def _create(self, cfg, *args):
        cc = cfg.c_config()
        h = self.lib.orc_create(C.byref(cc), self.order)
        if not h:
            raise _capi.DspMapError(_capi.ERRORS and -2, "orc_create")
        self._h = C.c_void_p(h)

    def check(self, rc, where):
        if rc != _capi.DSPMAP_OK:
            raise _capi.DspMapError(rc, where)

    def set_diagnostics(self, on):
        self.lib.orc_set_diagnostics(self._h, int(on))

    def set_invalid_sigma(self, value):
        """The sigma an invalid pixel carries (uninitialised in the reference, pointcloud_tools.h:224-232)."""
        self.lib.orc_set_invalid_sigma(self._h, float(value))

    def set_undefined_label_carry(self, on):
        """Label / track of an occupied voxel whose histogram is empty: 0 (default) or what the previous voxel left, as a
        compiled reference does (semantic_dsp_map.h:1245-1246)."""
        self.lib.orc_set_undefined_label_carry(self._h, int(on))

    def extra(self, which):
        return int(self.lib.orc_get_extra(self._h, which))

    def stage_seconds(self):
        return [float(self.lib.orc_stage_seconds(self._h, k)) for k in range(8)]


def oracle_map(cfg, labels=None, order=1):
    return SemanticDSPMap(cfg, labels, _engine=OracleEngine(cfg, order))
