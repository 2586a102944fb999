"""ctypes view of include/dspmap_c.h.

The structures here are the C-ABI's own (dspmap_config, dspmap_params, ...).  The product
library is ``libdspmap_b200.so`` next to this file; loading it fails loudly when it has not
been built — there is no CPU fallback and nothing here ever touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdspmap_b200.so")

DSPMAP_OK = 0
DSPMAP_ERR_CAPACITY = -5
ERRORS = {
    0: "OK", -1: "INVALID_ARG", -2: "CONFIG", -3: "IMAGE_SIZE", -4: "CUDA", -5: "CAPACITY",
    -6: "NO_DEVICE", -7: "UNSUPPORTED", -8: "COMM",
}
COMM_HANDLE_BYTES = 128

DUMP_POINTS, DUMP_PARTICLES, DUMP_PIXEL_OFFSETS, DUMP_PIXEL_LISTS, DUMP_CK, DUMP_RING, DUMP_VOXEL_STATE, DUMP_PARTICLES_HASH = 1, 2, 3, 4, 5, 6, 7, 8


class Config(C.Structure):
    _fields_ = [
        ("voxel_bits_x", C.c_int32), ("voxel_bits_y", C.c_int32), ("voxel_bits_z", C.c_int32),
        ("slot_bits", C.c_int32), ("voxel_size", C.c_float),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("image_width", C.c_int32), ("image_height", C.c_int32), ("boost_mode", C.c_int32),
        ("depth_min", C.c_float), ("depth_max", C.c_float),
        ("consider_instance", C.c_int32), ("zed2_filters", C.c_int32), ("window_half", C.c_int32),
        ("shard_rank", C.c_int32), ("shard_count", C.c_int32), ("device", C.c_int32),
        ("max_visible", C.c_uint32), ("reserved0", C.c_uint32),
    ]


class Params(C.Structure):
    _fields_ = [
        ("detection_probability", C.c_float), ("noise_number", C.c_float),
        ("nb_ptc_num_per_point", C.c_int32), ("occupancy_threshold", C.c_float),
        ("forgetting_rate", C.c_float), ("max_forget_count", C.c_int32),
        ("id_transition_probability", C.c_float),
        ("depth_noise_first_order", C.c_float), ("depth_noise_zero_order", C.c_float),
        ("consider_depth_noise", C.c_int32), ("use_independent_filter", C.c_int32),
        ("zero_center", C.c_int32),
        ("uninit_sigma_mode", C.c_int32), ("uninit_sigma_value", C.c_float),
        ("reference_iteration_order", C.c_int32),
    ]


class Mask(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("row_stride", C.c_int32), ("is_static", C.c_int32),
        ("track_id", C.c_int32), ("label_id", C.c_int32),
        ("bbox_min", C.c_double * 3), ("bbox_max", C.c_double * 3),
        ("rows", C.c_int32), ("cols", C.c_int32),
    ]


class ObjectMove(C.Structure):
    _fields_ = [("track_id", C.c_int32), ("t", C.c_float * 16)]


class ObjectObservation(C.Structure):
    _fields_ = [("track_id", C.c_int32), ("label_known", C.c_int32), ("label_id", C.c_int32), ("n_kpts", C.c_int32),
                ("kpts_current", C.c_void_p), ("kpts_previous", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [
        ("count", C.c_uint32), ("free_count", C.c_uint32),
        ("voxel_index", C.c_void_p), ("xyz", C.c_void_p), ("weight_sum", C.c_void_p),
        ("track_id", C.c_void_p), ("label_id", C.c_void_p), ("state", C.c_void_p),
        ("free_voxel_index", C.c_void_p), ("free_xyz", C.c_void_p), ("rgb", C.c_void_p),
    ]


class Colors(C.Structure):
    _fields_ = [
        ("enabled", C.c_int32), ("eval_format", C.c_int32), ("scheme", C.c_int32), ("background_label", C.c_int32),
        ("label_bgr", (C.c_uint8 * 3) * 256), ("instance_map", C.c_uint8 * 256),
    ]


class FrameStats(C.Structure):
    _fields_ = [
        ("time_stamp", C.c_uint32), ("n_swept_voxels", C.c_uint32), ("n_swept_nonempty", C.c_uint32),
        ("n_visible", C.c_uint32), ("n_valid_points", C.c_uint32), ("n_born", C.c_uint32),
        ("n_birth_voxels", C.c_uint32), ("n_resampled", C.c_uint32), ("n_occupied", C.c_uint32),
        ("n_nonempty_voxels", C.c_uint32), ("n_moved", C.c_uint32), ("n_removed", C.c_uint32),
        ("ms_stage", C.c_float * 8),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "ms_stage"}
        d["ms_stage"] = list(self.ms_stage)
        return d


STAGE_NAMES = ("project", "occ_headers_kernel", "predict", "visibility", "update", "birth", "occupancy", "total")

# name -> (restype, argtypes); every symbol include/dspmap_c.h declares
SYMBOLS = {
    "dspmap_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "dspmap_destroy": (None, [C.c_void_p]),
    "dspmap_clear": (C.c_int, [C.c_void_p]),
    "dspmap_last_error": (C.c_char_p, [C.c_void_p]),
    "dspmap_version": (C.c_int, []),
    "dspmap_set_params": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
    "dspmap_set_label_luts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint16, C.c_uint16]),
    "dspmap_set_random_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "dspmap_set_colors": (C.c_int, [C.c_void_p, C.POINTER(Colors)]),
    "dspmap_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Mask), C.c_int32,
                                C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ObjectMove), C.c_int32,
                                C.POINTER(C.c_int32), C.c_int32, C.c_int32]),
    "dspmap_update_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Mask), C.c_int32,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ObjectMove), C.c_int32,
                                       C.POINTER(C.c_int32), C.c_int32, C.c_int32]),
    "dspmap_update_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Mask), C.c_int32,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ObjectMove), C.c_int32,
                                      C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_uint32)]),
    "dspmap_update_device_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Mask), C.c_int32,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ObjectMove), C.c_int32,
                                             C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_uint32)]),
    "dspmap_wait_result": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(Result)]),
    "dspmap_get_result": (C.c_int, [C.c_void_p, C.POINTER(Result)]),
    "dspmap_get_result_device": (C.c_int, [C.c_void_p, C.POINTER(Result)]),
    "dspmap_get_stats": (C.c_int, [C.c_void_p, C.POINTER(FrameStats)]),
    "dspmap_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "dspmap_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dspmap_get_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "dspmap_object_particle_count": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_uint32)]),
    "dspmap_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dspmap_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dspmap_selftest_division": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_uint64)]),
    "dspmap_selftest_colour": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dspmap_dump": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    # host object layer (no GPU needed)
    "dspmap_objects_create": (C.c_int, [C.POINTER(Config), C.c_int32, C.POINTER(C.c_void_p)]),
    "dspmap_objects_destroy": (None, [C.c_void_p]),
    "dspmap_objects_clear": (C.c_int, [C.c_void_p]),
    "dspmap_objects_set_movement_filter": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "dspmap_objects_set_seed": (C.c_int, [C.c_void_p, C.c_uint32]),
    "dspmap_objects_step": (C.c_int, [C.c_void_p, C.POINTER(ObjectObservation), C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                      C.c_int32, C.POINTER(ObjectMove), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32,
                                      C.POINTER(C.c_int32)]),
    "dspmap_objects_last_plan": (C.c_int, [C.c_void_p, C.POINTER(ObjectMove), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "dspmap_objects_query": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]),
    "dspmap_objects_forget": (C.c_int, [C.c_void_p, C.c_int32]),
}

_lib = None


def bind(lib, symbols=SYMBOLS, prefix_from="dspmap_", prefix_to="dspmap_"):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in symbols.items():
        fn = getattr(lib, name.replace(prefix_from, prefix_to, 1))
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Load the CUDA engine.  Raises if the shared library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). The engine has no CPU fallback.")
        _lib = bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL))
    return _lib


class DspMapError(RuntimeError):
    def __init__(self, rc, where, detail=""):
        self.rc = rc
        super().__init__(f"{where}: {ERRORS.get(rc, rc)} ({rc}) {detail}".strip())
