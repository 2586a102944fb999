"""Loader for oracle/_ref/libdsp_ref_s<SETTING>.so: the REFERENCE's own hot-path code, compiled unmodified from
/root/reference/include against the stand-in third-party headers of oracle/ref_stubs (see oracle/ref_harness.cpp).

TEST INFRASTRUCTURE: used only to pin the CPU oracle (tests/test_reference_pin_cpu.py, tests/golden/make_ref_golden.py).
The reference keeps its map in process-wide globals and function statics, so one process can run ONE scenario: use
``run_scenario_subprocess`` (fresh interpreter per scenario)."""
import ctypes as C
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from semantic_dsp_map_b200 import _capi  # noqa: E402
from semantic_dsp_map_b200.labels import ZED2_ROWS  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
# colour columns of cfg/object_info_zed2.csv are not needed for labels; tests that check colours pass their own table
MAX_OBSERVATION_LOST_TIME = 20   # cfg/options_zed2.yaml


class RefMask(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("row_stride", C.c_int32), ("track_id", C.c_int32),
                ("label", C.c_char_p), ("kpts", C.c_void_p), ("n_kpts", C.c_int32)]


def ref_lib_path(setting=3):
    return os.path.join(REF_DIR, f"libdsp_ref_s{setting}.so")


def have_ref(setting=3):
    return os.path.exists(ref_lib_path(setting))


def write_label_csv(path, rows=ZED2_ROWS, colours=None):
    """The CSV format ObjectInfoHandler::readObjectInfo parses (utils/object_info_handler.h:104-150; docs/custom_files.md)."""
    with open(path, "w") as f:
        f.write("label_id,label,instance_id(-1 means not fixed),color_b,color_g,color_r\n")
        for label_id, name, inst in rows:
            b, g, r = (colours or {}).get(label_id, ((37 * label_id) % 256, (91 * label_id + 13) % 256, (53 * label_id + 101) % 256))
            f.write(f"{label_id},{name},{inst},{b},{g},{r}\n")


class RefMap:
    """One per process.  Same update() arguments as semantic_dsp_map_b200.SemanticDSPMap."""

    def __init__(self, cfg, csv_path, setting=3, eval_format=False, max_lost=MAX_OBSERVATION_LOST_TIME):
        lib = C.CDLL(ref_lib_path(setting))
        lib.ref_describe.argtypes = [C.POINTER(_capi.Config)]
        lib.ref_read_object_info.argtypes = [C.c_char_p]
        lib.ref_set_params.argtypes = [C.POINTER(_capi.Params), C.c_int, C.c_int]
        lib.ref_set_random_table.argtypes = [C.c_void_p, C.c_uint32]
        lib.ref_set_color_map.argtypes = [C.c_void_p]
        lib.ref_update.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(RefMask), C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(_capi.ObjectMove), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32]
        lib.ref_last_seconds.restype = C.c_double
        lib.ref_time_stamp.restype = C.c_uint32
        lib.ref_get_cloud.restype = C.c_uint32
        lib.ref_get_cloud.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32]
        lib.ref_voxel_positions.restype = C.c_uint32
        lib.ref_voxel_positions.argtypes = [C.c_void_p, C.c_uint32]
        lib.ref_move_order.argtypes = [C.POINTER(C.c_int32), C.c_int32]
        lib.ref_dump.argtypes = [C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        self.lib, self.cfg = lib, cfg
        built = _capi.Config()
        lib.ref_describe(C.byref(built))
        want = cfg.c_config()
        for name, _ in _capi.Config._fields_:
            if name in ("shard_rank", "shard_count", "device", "max_visible", "reserved0", "window_half"):
                continue
            a, b = getattr(built, name), getattr(want, name)
            if (abs(a - b) > 1e-6 * max(1.0, abs(b))) if isinstance(a, float) else (a != b):
                raise ValueError(f"the reference was compiled with {name}={a}, the test config has {b}")
        self.max_movable = lib.ref_read_object_info(csv_path.encode())
        p = cfg.c_params()
        lib.ref_set_params(C.byref(p), max_lost, int(eval_format))

    def set_random_table(self, table):
        t = np.ascontiguousarray(table, np.float32)
        if self.lib.ref_set_random_table(t.ctypes.data, t.size) != 0:
            raise ValueError("the reference's random table has exactly 1,000,000 entries (utils/basic_algorithms.h:377)")

    def set_color_map(self, perm256):
        a = np.ascontiguousarray(perm256, np.int32)
        assert a.size == 256
        self.lib.ref_set_color_map(a.ctypes.data)

    def update(self, depth, masks, position, orientation, if_get_freespace=False, moves=(), remove_tracks=(), fetch=True):
        depth = np.ascontiguousarray(depth, np.float32)
        keep = []
        marr = (RefMask * max(len(masks), 1))()
        for i, m in enumerate(masks):
            mk = np.ascontiguousarray(m.mask, np.uint8)
            kp = np.ascontiguousarray(np.asarray(m.kpts_current, np.float64).reshape(-1, 3))
            keep += [mk, kp]
            marr[i].data, marr[i].rows, marr[i].cols, marr[i].row_stride = mk.ctypes.data, mk.shape[0], mk.shape[1], mk.strides[0]
            marr[i].track_id, marr[i].label = int(m.track_id), m.label.encode()
            marr[i].kpts, marr[i].n_kpts = kp.ctypes.data, kp.shape[0]
        mv = (_capi.ObjectMove * max(len(moves), 1))()
        for i, (track, t) in enumerate(moves):
            mv[i].track_id = int(track)
            flat = np.asarray(t, np.float32).reshape(16)
            for k in range(16):
                mv[i].t[k] = float(flat[k])
        rm = (C.c_int32 * max(len(remove_tracks), 1))(*[int(t) for t in remove_tracks])
        pos = (C.c_double * 3)(*[float(x) for x in position])
        quat = (C.c_double * 4)(*[float(x) for x in orientation])
        rc = self.lib.ref_update(depth.ctypes.data, depth.shape[0], depth.shape[1], depth.strides[0] // 4, marr, len(masks), pos, quat, mv, len(moves),
                                 rm, len(remove_tracks), int(bool(if_get_freespace)))
        assert rc == 0
        return self.clouds() if fetch else None

    def clouds(self):
        out = {}
        for which, name in ((0, "occ"), (1, "free")):
            n = self.lib.ref_get_cloud(which, None, None, 0)
            xyz, rgb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.uint8)
            if n:
                self.lib.ref_get_cloud(which, xyz.ctypes.data, rgb.ctypes.data, n)
            out[name + "_xyz"], out[name + "_rgb"] = xyz, rgb
        return out

    def voxel_positions(self):
        n = self.cfg.n_voxels
        xyz = np.zeros((n, 3), np.float32)
        assert self.lib.ref_voxel_positions(xyz.ctypes.data, n) == n
        return xyz

    def move_order(self):
        buf = (C.c_int32 * 256)()
        n = self.lib.ref_move_order(buf, 256)
        return [buf[i] for i in range(n)]

    def dump(self, what, dtype=np.uint8):
        n = C.c_size_t()
        assert self.lib.ref_dump(what, None, 0, C.byref(n)) == 0
        buf = np.zeros(n.value, np.uint8)
        if n.value:
            assert self.lib.ref_dump(what, buf.ctypes.data, n.value, C.byref(n)) == 0
        return buf.view(dtype)

    def seconds(self):
        return float(self.lib.ref_last_seconds())


def voxel_index_of(points_xyz, all_voxel_xyz):
    """Emission order is ascending storage index and every voxel has a distinct corner: recover the indices by a merge."""
    if len(points_xyz) == 0:
        return np.zeros(0, np.uint32)
    key = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)  # noqa: E731
    ka, kp = key(all_voxel_xyz), key(points_xyz)
    mix = lambda k: (k[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (k[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) ^ (k[:, 2] * np.uint64(0x165667B19E3779F9))  # noqa: E731
    with np.errstate(over="ignore"):
        ha, hp = mix(ka), mix(kp)
    order = np.argsort(ha, kind="stable")
    pos = np.searchsorted(ha[order], hp)
    idx = order[np.clip(pos, 0, len(order) - 1)]
    assert np.array_equal(all_voxel_xyz[idx].view(np.uint32), np.ascontiguousarray(points_xyz, np.float32).view(np.uint32)), "point is not a voxel corner"
    return idx.astype(np.uint32)


def run_scenario_subprocess(module, func, *args, timeout=1200):
    """Run ``module.func(*args)`` in a fresh interpreter (one reference map per process) and return its pickled result."""
    with tempfile.TemporaryDirectory() as td:
        inp, outp = os.path.join(td, "in.pkl"), os.path.join(td, "out.pkl")
        with open(inp, "wb") as f:
            pickle.dump(args, f)
        code = (f"import sys, pickle; sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'tests')!r}, {os.path.join(ROOT, 'tests', 'golden')!r}]\n"
                f"import {module} as M\nargs = pickle.load(open({inp!r}, 'rb'))\nres = getattr(M, {func!r})(*args)\n"
                f"pickle.dump(res, open({outp!r}, 'wb'))\n")
        subprocess.run([sys.executable, "-c", code], check=True, timeout=timeout, stdout=subprocess.DEVNULL)
        with open(outp, "rb") as f:
            return pickle.load(f)
