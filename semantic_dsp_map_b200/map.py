"""Host-side mirror of the reference's ``SemanticDSPMap`` interface over the C ABI.

The reference exposes one class, ``SemanticDSPMap`` (include/semantic_dsp_map.h:23-251), that the ROS node drives with
five setters and ``update()`` (src/mapping.cpp:135-140, 273).  This module keeps the same method names, argument
meaning and error behaviour for the hot path, and evaluates on the host only what the reference evaluates with string
keys and tiny per-object loops (label look-ups, bbox from key points, the track-id wrap) before handing plain arrays
to ``libdspmap_b200.so``.  The C++ adapter a ROS build would use is shown in INTEGRATION.md; this is its Python twin,
used by the parity tests and the bench.  There is no CPU fallback: without the CUDA library or a GPU every entry point
raises.
"""
from __future__ import annotations

import ctypes as C
import sys
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi
from .configs import MapConfig
from .labels import LabelTables, ZED2_ROWS

_DBL_MAX = sys.float_info.max
_DBL_MIN = sys.float_info.min  # std::numeric_limits<double>::min(): smallest POSITIVE double (pointcloud_tools.h:178-180)


@dataclass
class MaskKpts:
    """include/utils/tracking_result_handler.h:15-26."""
    track_id: int
    label: str
    mask: np.ndarray                                   # uint8, input resolution
    bbox: tuple = (0, 0, 0, 0)
    kpts_current: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    kpts_previous: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))


@dataclass
class OccupancyResult:
    """What getOccupancyResult emits (semantic_dsp_map.h:1239-1383), before colouring."""
    voxel_index: np.ndarray
    xyz: np.ndarray
    weight_sum: np.ndarray
    track_id: np.ndarray
    label_id: np.ndarray
    state: np.ndarray
    free_voxel_index: np.ndarray
    free_xyz: np.ndarray
    rgb: np.ndarray = None     # (n, 3) uint8 r, g, b of the occupied points (zeros until set_colors enabled them)


_memview = C.pythonapi.PyMemoryView_FromMemory
_memview.restype = C.py_object
_memview.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int]


def _as_array(ptr, count, dtype, shape=None, copy=True):
    """count entries of `dtype` (rows of shape[1] entries when `shape` is given) at address `ptr`, as a numpy view or copy."""
    width = shape[1] if shape else 1
    if count == 0 or not ptr:
        return np.zeros((0, width) if shape else (0,), dtype)
    n = count * width
    a = np.frombuffer(_memview(ptr, n * _ITEMSIZE[dtype], 0x100), dtype, n)   # 0x100 = PyBUF_READ
    if copy:
        a = a.copy()
    return a.reshape(count, width) if shape else a


_ITEMSIZE = {np.uint32: 4, np.float32: 4, np.uint16: 2, np.uint8: 1, np.int8: 1}


class Engine:
    """Thin owner of one ``dspmap_t`` handle; used for both the CUDA library and (in tests) the oracle library."""

    def __init__(self, lib, prefix: str, cfg: MapConfig, create_args=()):
        self.lib, self.prefix, self.cfg = lib, prefix, cfg
        self._h = C.c_void_p()
        self._keep = []
        self._create(cfg, *create_args)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _create(self, cfg, *args):
        cc = cfg.c_config(*args)
        rc = self._fn("create")(C.byref(cc), C.byref(self._h))
        if rc != _capi.DSPMAP_OK:
            raise _capi.DspMapError(rc, "dspmap_create")

    def check(self, rc, where):
        if rc != _capi.DSPMAP_OK:
            detail = self._fn("last_error")(self._h).decode() if hasattr(self.lib, self.prefix + "last_error") else ""
            raise _capi.DspMapError(rc, where, detail)

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setters -------------------------------------------------------------------------------------------
    def set_params(self, p: _capi.Params):
        self.check(self._fn("set_params")(self._h, C.byref(p)), "dspmap_set_params")

    def set_label_luts(self, tables: LabelTables):
        v2t = np.ascontiguousarray(tables.static_value_to_track(), np.uint16)
        tracks, labels = tables.static_track_labels()
        tracks = np.ascontiguousarray(tracks, np.uint16)
        labels = np.ascontiguousarray(labels, np.uint8)
        self.check(self._fn("set_label_luts")(self._h, v2t.ctypes.data, tracks.ctypes.data, labels.ctypes.data, len(tracks),
                                              tables.max_movable_instance_id & 0xFFFF, tables.sky_track() & 0xFFFF),
                   "dspmap_set_label_luts")

    def set_colors(self, c: _capi.Colors):
        self.check(self._fn("set_colors")(self._h, C.byref(c)), "dspmap_set_colors")

    def set_random_table(self, table: np.ndarray):
        t = np.ascontiguousarray(table, np.float32)
        self.check(self._fn("set_random_table")(self._h, t.ctypes.data, t.size), "dspmap_set_random_table")

    def clear(self):
        self.check(self._fn("clear")(self._h), "dspmap_clear")

    # ---- sharded maps ------------------------------------------------------------------------------------------
    def comm_export(self) -> bytes:
        buf = (C.c_uint8 * _capi.COMM_HANDLE_BYTES)()
        self.check(self._fn("comm_export")(self._h, buf), "dspmap_comm_export")
        return bytes(buf)

    def comm_connect(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == _capi.COMM_HANDLE_BYTES * len(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.check(self._fn("comm_connect")(self._h, buf), "dspmap_comm_connect")

    # ---- frame ---------------------------------------------------------------------------------------------------
    # numpy mirrors of dspmap_mask / dspmap_object_move (same size and field offsets as the ctypes structures): descriptor
    # arrays are filled with a handful of vectorised assignments instead of one Python attribute store per field
    MASK_DTYPE = np.dtype([("data", "<u8"), ("row_stride", "<i4"), ("is_static", "<i4"), ("track_id", "<i4"), ("label_id", "<i4"),
                           ("bbox_min", "<f8", (3,)), ("bbox_max", "<f8", (3,)), ("rows", "<i4"), ("cols", "<i4")])
    MOVE_DTYPE = np.dtype([("track_id", "<i4"), ("t", "<f4", (16,))])

    @classmethod
    def pack_masks(cls, masks) -> np.ndarray:
        """list of mask descriptors (dicts, see SemanticDSPMap.lower_masks) -> array of dspmap_mask."""
        if isinstance(masks, np.ndarray):
            return masks
        arr = np.zeros(len(masks), cls.MASK_DTYPE)
        for i, m in enumerate(masks):
            arr[i] = (m["data"], m["row_stride"], int(m["is_static"]), int(m["track_id"]), int(m["label_id"]), m["bbox_min"], m["bbox_max"],
                      int(m.get("rows", 0)), int(m.get("cols", 0)))
        return arr

    @classmethod
    def pack_moves(cls, moves) -> np.ndarray:
        if isinstance(moves, np.ndarray):
            return moves
        arr = np.zeros(len(moves), cls.MOVE_DTYPE)
        for i, (track, t) in enumerate(moves):
            arr[i] = (int(track), np.asarray(t, np.float32).reshape(16))
        return arr

    def update_raw(self, fn_name, depth_ptr, rows, cols, stride, masks, position, orientation, moves=(), remove_tracks=(),
                   want_freespace=False):
        n_masks, n_moves = len(masks), len(moves)
        marr, mvarr = self.pack_masks(masks), self.pack_moves(moves)
        rm = np.asarray(remove_tracks if len(remove_tracks) else (0,), np.int32)
        pose = np.empty(7, np.float64)
        pose[:3], pose[3:] = position, orientation
        args = (self._h, depth_ptr, rows, cols, stride, C.cast(marr.ctypes.data, C.POINTER(_capi.Mask)), n_masks,
                C.cast(pose.ctypes.data, C.POINTER(C.c_double)), C.cast(pose.ctypes.data + 24, C.POINTER(C.c_double)),
                C.cast(mvarr.ctypes.data, C.POINTER(_capi.ObjectMove)), n_moves,
                C.cast(rm.ctypes.data, C.POINTER(C.c_int32)), len(remove_tracks), int(bool(want_freespace)))
        if fn_name.endswith("_async"):      # pipelined frame (dspmap_c.h): returns the frame's ticket
            ticket = C.c_uint32()
            self.pending_ticket = 0
            rc = self._fn(fn_name)(*args, C.byref(ticket))
            if rc != _capi.DSPMAP_OK and ticket.value:
                # this frame IS enqueued; the error belongs to the frame before it (found when its counters were read)
                self.pending_ticket = ticket.value
            self.check(rc, "dspmap_" + fn_name)
            return ticket.value
        self.check(self._fn(fn_name)(*args), "dspmap_" + fn_name)

    def result(self, device=False) -> _capi.Result:
        r = _capi.Result()
        self.check(self._fn("get_result_device" if device else "get_result")(self._h, C.byref(r)), "dspmap_get_result")
        return r

    def wait_result(self, ticket, copy=True) -> OccupancyResult:
        """Result of the pipelined frame `ticket` (dspmap_wait_result); copy=False: views valid until two more frames."""
        r = _capi.Result()
        self.check(self._fn("wait_result")(self._h, ticket, C.byref(r)), "dspmap_wait_result")
        return self.result_arrays(copy, r)

    def result_arrays(self, copy=True, r=None) -> OccupancyResult:
        """copy=False returns views of the handle's pinned host buffers, valid until the next update (dspmap_c.h)."""
        r = r or self.result()
        n, nf = r.count, r.free_count
        return OccupancyResult(
            _as_array(r.voxel_index, n, np.uint32, None, copy), _as_array(r.xyz, n, np.float32, (n, 3), copy),
            _as_array(r.weight_sum, n, np.float32, None, copy), _as_array(r.track_id, n, np.uint16, None, copy),
            _as_array(r.label_id, n, np.uint8, None, copy), _as_array(r.state, n, np.int8, None, copy),
            _as_array(r.free_voxel_index, nf, np.uint32, None, copy), _as_array(r.free_xyz, nf, np.float32, (nf, 3), copy),
            _as_array(r.rgb, n, np.uint8, (n, 3), copy))

    def stats(self) -> dict:
        s = _capi.FrameStats()
        self.check(self._fn("get_stats")(self._h, C.byref(s)), "dspmap_get_stats")
        return s.as_dict()

    def dump(self, what: int, dtype=np.uint8) -> np.ndarray:
        n = C.c_size_t()
        self.check(self._fn("dump")(self._h, what, None, 0, C.byref(n)), "dspmap_dump")
        buf = np.zeros(n.value, np.uint8)
        if n.value:
            self.check(self._fn("dump")(self._h, what, buf.ctypes.data, n.value, C.byref(n)), "dspmap_dump")
        return buf.view(dtype)


def connect_shards(dsp_map: "SemanticDSPMap", all_gather_object):
    """Exchange the exchange-arena handles of a sharded map.  ``all_gather_object(obj) -> list`` is the caller's plumbing
    (e.g. a wrapper of torch.distributed.all_gather_object over NCCL or gloo): it must return every rank's object in rank
    order."""
    handles = all_gather_object(dsp_map.engine.comm_export())
    if len(handles) != dsp_map.shard_count:
        raise ValueError(f"expected {dsp_map.shard_count} handles, got {len(handles)}")
    dsp_map.engine.comm_connect(handles)
    return handles


class SemanticDSPMap:
    """The reference's public class, same method names (include/semantic_dsp_map.h:23-251)."""

    def __init__(self, config: MapConfig, labels: Optional[LabelTables] = None, device: int = 0, _engine: Optional[Engine] = None,
                 shard_rank: int = 0, shard_count: int = 1):
        """shard_count > 1: this process holds storage-z slabs [rank*NZ/G, (rank+1)*NZ/G) of one map on `device`; call
        connect_shards() on every rank before the first update (include/dspmap_c.h, "sharded maps")."""
        self.config = config
        self.labels = labels or LabelTables.from_rows(ZED2_ROWS)
        self.engine = _engine or Engine(_capi.load(), "dspmap_", config, (shard_rank, shard_count, device, 0))
        self.shard_rank, self.shard_count = shard_rank, shard_count
        self._params = config.c_params()
        self.engine.set_params(self._params)
        self.engine.set_label_luts(self.labels)
        self._movement = (0.1, 0.69, 0.1, 0.15)          # constructor defaults, semantic_dsp_map.h:36-39
        self.max_obersevation_lost_time = 5              # :30
        self.objects = None                              # host object layer, created at the first update() that needs it
        self._keep = []
        self._eval_format = False
        self._colors = None
        # mirror of ObjectSet's two tables as far as the hot path reads them (object_layer.h:345-425): which movable tracks
        # are tracked objects, and which tracks can own particles without being one ("floating", semantic_dsp_map.h:712-737)
        self._known_objects = set()
        self._born_tracks = set()
        self._met_labels = set()
        self._refresh_lut_tracks()

    # -- setters, semantic_dsp_map.h:101-166 ----------------------------------------------------------------------------
    def setMapParameters(self, detection_probability, noise_number, nb_ptc_num_per_point, occupancy_threshold, max_obersevation_lost_time,
                         forgetting_rate=1.0, max_forget_count=5, match_score_threshold=0.5, id_transition_probability=0.1):
        p = self._params
        p.detection_probability, p.noise_number = detection_probability, noise_number
        p.nb_ptc_num_per_point, p.occupancy_threshold = nb_ptc_num_per_point, occupancy_threshold
        p.forgetting_rate, p.max_forget_count = forgetting_rate, max_forget_count
        p.id_transition_probability = id_transition_probability
        self.max_obersevation_lost_time = max_obersevation_lost_time
        self.match_score_threshold = match_score_threshold
        self.engine.set_params(p)

    def setMapOptions(self, if_consider_depth_noise, if_use_independent_filter):
        self._params.consider_depth_noise = int(bool(if_consider_depth_noise))
        self._params.use_independent_filter = int(bool(if_use_independent_filter))
        self.engine.set_params(self._params)

    def setVisualizeOptions(self, visualize_with_zero_center, if_out_evaluation_format=False):
        self._params.zero_center = int(bool(visualize_with_zero_center))
        self._eval_format = bool(if_out_evaluation_format)
        self.engine.set_params(self._params)
        if self._colors is not None:
            self._colors.eval_format = int(self._eval_format)
            self.engine.set_colors(self._colors)

    def enable_colors(self, label_bgr=None, instance_map=None, scheme=None):
        """Turn on the colouring of the occupied list (getOccupancyResult, semantic_dsp_map.h:1274-1351).

        label_bgr: {label_id: (b, g, r)} = g_label_color_map_default (default: the colour columns of the object-info CSV the
        label tables were read from, or the built-in colours);
        instance_map: the 256-entry permutation color_map_int_256_ (a std::random_device shuffle in the reference, so the
        caller supplies it; default: a fixed seeded permutation); scheme: which ``#if SETTING`` colouring block applies
        (default: 3 when the ZED2 filters are on, 0 for a map without instances, else 1)."""
        c = _capi.Colors()
        c.enabled, c.eval_format = 1, int(self._eval_format)
        cfg = self.config
        c.scheme = scheme if scheme is not None else (3 if cfg.zed2_filters else (1 if cfg.consider_instance else 0))
        c.background_label = self.labels.label_id_of("Background")              # :1276, default-insert semantics
        for label_id, bgr in (self.labels.label_colors if label_bgr is None else label_bgr).items():
            for k in range(3):
                c.label_bgr[int(label_id) & 0xFF][k] = int(bgr[k]) & 0xFF
        perm = np.random.default_rng(256).permutation(256) if instance_map is None else np.asarray(instance_map)
        for k in range(256):
            c.instance_map[k] = int(perm[k]) & 0xFF
        self._colors = c
        self.engine.set_colors(c)

    def setBeyesianMovementParameters(self, distance_threshold, probability_threshold, increment, decrement):   # semantic_dsp_map.h:142-148
        self._movement = (distance_threshold, probability_threshold, increment, decrement)
        if self.objects is not None:
            self.objects.set_movement_filter(*self._movement)

    def _object_layer(self):
        if self.objects is None:
            from .objects import ObjectLayer
            self.objects = ObjectLayer(self.config)
            self.objects.set_movement_filter(*self._movement)
        return self.objects

    def _run_object_layer(self, ins_seg_result, camera_position, camera_orientation, time_stamp):
        """objectLevelUpdate on this frame's tracking result, then the move / drop decisions (semantic_dsp_map.h:188-190,
        306-566, 588-704) - on the host, through the C ABI's object layer."""
        max_movable = self.labels.max_movable_instance_id
        obs = []
        for m in ins_seg_result:
            if m.label == "static" or m.track_id > max_movable:                      # :317
                continue
            known = m.label in self.labels.label_id_map or m.label in self._met_labels
            obs.append((m.track_id, known, self.labels.label_id_of(m.label), m.kpts_current, m.kpts_previous))
        return self._object_layer().step(obs, camera_position, camera_orientation, time_stamp, self.max_obersevation_lost_time)

    def setDepthNoiseModelParameters(self, first_order, zero_order):
        self._params.depth_noise_first_order, self._params.depth_noise_zero_order = first_order, zero_order
        self.engine.set_params(self._params)

    def followCompiledReferenceSigma(self, on: bool = True, value: float = 0.0):
        """Not in the reference's API: where the reference reads an uninitialised LabeledPoint::sigma (invalid pixels and
        pixels its ZED2 bounding-box gate turned into background, pointcloud_tools.h:224-232, 256-272) use ``value`` - a
        compiled reference reads 0.0 there - instead of the depth-noise model (dspmap_params.uninit_sigma_mode)."""
        self._params.uninit_sigma_mode, self._params.uninit_sigma_value = int(bool(on)), float(value)
        self.engine.set_params(self._params)

    def setReferenceIterationOrder(self, on: bool = True):
        """Not in the reference's API: build the pixel lists in the order the reference's breadth-first search visits the voxels
        in (dspmap_params.reference_iteration_order) instead of ascending particle index.  A scene without moving objects then
        reproduces the reference bit for bit in every float; costs a few milliseconds per frame."""
        self._params.reference_iteration_order = int(bool(on))
        self.engine.set_params(self._params)

    def setOccupancyThreshold(self, thr):
        self._params.occupancy_threshold = thr
        self.engine.set_params(self._params)

    def clear(self):
        self.engine.clear()
        self._known_objects.clear()          # object_set_.clear(), semantic_dsp_map.h:74-81
        self._born_tracks.clear()
        self._met_labels.clear()
        if self.objects is not None:
            self.objects.clear()

    def _refresh_lut_tracks(self):
        max_movable = self.labels.max_movable_instance_id
        self._lut_movable_tracks = {int(t) for t in np.unique(self.labels.static_value_to_track()) if int(t) <= max_movable}

    def _floating_tracks(self, ins_seg_result, moves, remove_tracks, camera_position=(0.0, 0.0, 0.0), builtin=False):
        """semantic_dsp_map.h:712-737 ("floating objects"): after the moves and the removals of a frame, every track that owns
        particles in obj_ptc_hash_map but has no entry in object_tracking_hash_map is removed with all its particles.  Particles
        enter obj_ptc_hash_map at birth when their track id is <= g_max_movable_object_instance_id (:1160-1162); objects enter
        object_tracking_hash_map when objectLevelUpdate sees their mask (:346-364), which only runs with g_consider_instance,
        and only when the label is known (:330-334; pointcloud_tools.h:168 default-inserts an unknown label, so it is known
        from the next frame on) and some key point lies within 1.2 x the map's half size of the camera (:348-360: an object
        reported WITHOUT key points is never registered).
        Hence e.g. in static mode the pixels of a movable CLASS of the static mask (instance id -1 -> track 0,
        pointcloud_tools.h:137-138) live for exactly one frame.  Returns the tracks to append to this frame's removals."""
        max_movable = self.labels.max_movable_instance_id
        if builtin:   # the object layer itself says which tracks are objects
            if self.config.consider_instance:
                self._born_tracks.update(int(m.track_id) for m in ins_seg_result if m.label != "static" and m.track_id <= max_movable)
                self._met_labels.update(m.label for m in ins_seg_result if m.label != "static")
            self._born_tracks -= {int(t) for t in remove_tracks}
            floating = sorted(t for t in (self._lut_movable_tracks | self._born_tracks) if not self.objects.is_tracked(t) and t not in remove_tracks)
            for t in floating:
                self.objects.forget(t)
            self._born_tracks -= set(floating)
            return floating
        if self.config.consider_instance:
            half = float(np.float32(self.config.voxel_size) * np.float32(1 << (max(self.config.voxel_bits) - 1))) * 1.2
            cam = np.asarray(camera_position, np.float64).reshape(3)
            for m in ins_seg_result:
                if m.label != "static" and m.track_id <= max_movable:
                    self._born_tracks.add(int(m.track_id))
                    if int(m.track_id) in self._known_objects or not (m.label in self.labels.label_id_map or m.label in self._met_labels):
                        continue
                    k = np.asarray(m.kpts_current, np.float64).reshape(-1, 3)
                    if len(k) and np.abs(k - cam).max(axis=1).min() <= half:
                        self._known_objects.add(int(m.track_id))
            self._met_labels.update(m.label for m in ins_seg_result if m.label != "static")
        for track, _ in moves:
            self._known_objects.add(int(track))
        for track in remove_tracks:
            self._known_objects.discard(int(track))
        floating = sorted((self._lut_movable_tracks | self._born_tracks) - self._known_objects - {int(t) for t in remove_tracks})
        self._born_tracks -= set(floating) | {int(t) for t in remove_tracks}
        return floating

    # -- host evaluation of the string-keyed parts of generateLabeledPointCloud ---------------------------------------------
    def lower_masks(self, ins_seg_result: Sequence[MaskKpts]):
        """MaskKpts -> C-ABI mask descriptors (pointcloud_tools.h:116-215; semantic_dsp_map.h:179-186)."""
        out = []
        max_movable = self.labels.max_movable_instance_id
        for m in ins_seg_result:
            if m.label != "static" and m.track_id > max_movable:        # semantic_dsp_map.h:179-186
                m.track_id = m.track_id % max_movable
            mask = np.ascontiguousarray(m.mask, np.uint8)
            if mask.ndim != 2:
                raise TypeError("a mask must be a 2-D uint8 array (CV_8UC1)")
            self._keep.append(mask)
            k = np.asarray(m.kpts_current, np.float64).reshape(-1, 3)                    # pointcloud_tools.h:178-199
            mn = np.minimum(_DBL_MAX, k.min(0)) if len(k) else np.full(3, _DBL_MAX)
            mx = np.maximum(_DBL_MIN, k.max(0)) if len(k) else np.full(3, _DBL_MIN)
            out.append(dict(data=mask.ctypes.data, row_stride=mask.strides[0], is_static=(m.label == "static"),
                            track_id=m.track_id, label_id=self.labels.label_id_of(m.label),
                            bbox_min=(mn - 1.0).tolist(), bbox_max=(mx + 1.0).tolist(), rows=mask.shape[0], cols=mask.shape[1]))
        return out

    def _prepare_frame(self, depth_value_mat, ins_seg_result, camera_position, camera_orientation, time_stamp_double, moves, remove_tracks):
        """The host part of update() before the engine call: argument checks, mask descriptors, the object layer, the
        floating-object rule (semantic_dsp_map.h:176-190, 306-566, 712-737)."""
        depth = np.asarray(depth_value_mat)
        if depth.dtype != np.float32 or depth.ndim != 2:
            raise TypeError("depth must be a 2-D float32 array (CV_32FC1, metres)")
        if depth.strides[1] != 4:
            depth = np.ascontiguousarray(depth)
        self._keep = [depth]
        masks = self.lower_masks(ins_seg_result)
        builtin = moves is None
        if builtin:
            moves, dropped = self._run_object_layer(ins_seg_result, camera_position, camera_orientation, time_stamp_double) if self.config.consider_instance else ([], [])
            self._object_layer()
            remove_tracks = list(remove_tracks) + dropped
        saved = (set(self._born_tracks), set(self._known_objects), set(self._met_labels))
        remove_tracks = list(remove_tracks) + self._floating_tracks(ins_seg_result, moves, remove_tracks, camera_position, builtin)
        return depth, masks, moves, remove_tracks, saved

    # -- semantic_dsp_map.h:170-251 -----------------------------------------------------------------------------------------
    def update(self, depth_value_mat: np.ndarray, ins_seg_result: Sequence[MaskKpts], camera_position, camera_orientation,
               if_get_freespace: bool = False, time_stamp_double: float = 0.0, moves=None, remove_tracks=(),
               copy_result: bool = True) -> OccupancyResult:
        """One frame.  ``camera_orientation`` is (w, x, y, z).  With ``moves=None`` (default) the object layer of the
        reference (objectLevelUpdate) runs on the host from the masks' key points and ``time_stamp_double``, like in the
        reference.  A caller with an object layer of their own passes ``moves`` = [(track_id, 4x4)] and ``remove_tracks`` =
        [track_id] instead (an empty list = "no object moves this frame")."""
        depth, masks, moves, remove_tracks, saved = self._prepare_frame(depth_value_mat, ins_seg_result, camera_position, camera_orientation,
                                                                        time_stamp_double, moves, remove_tracks)
        try:
            self.engine.update_raw("update", depth.ctypes.data, depth.shape[0], depth.shape[1], depth.strides[0] // 4, masks,
                                   camera_position, camera_orientation, moves, remove_tracks, if_get_freespace)
        except _capi.DspMapError:
            # a refused frame changed nothing in the engine (arguments are validated before any state is touched): the
            # bookkeeping of which tracks are due for removal goes back to where it was, they go with the next frame
            self._born_tracks, self._known_objects, self._met_labels = saved
            raise
        return self.engine.result_arrays(copy_result)

    def update_async(self, depth_value_mat: np.ndarray, ins_seg_result: Sequence[MaskKpts], camera_position, camera_orientation,
                     if_get_freespace: bool = False, time_stamp_double: float = 0.0, moves=None, remove_tracks=()) -> int:
        """Not in the reference's API: update() with a deferred hand-over (dspmap_update_async, include/dspmap_c.h).  Stages
        and enqueues this frame and returns its ticket while it runs; ``wait_result(ticket)`` gives what ``update()`` would
        have returned.  Calling it for frame t+1 before ``wait_result(t)`` lets frame t's device->host transfer overlap
        frame t+1's kernels.  Results, order of operations on the map and the host object layer are those of update()."""
        depth, masks, moves, remove_tracks, saved = self._prepare_frame(depth_value_mat, ins_seg_result, camera_position, camera_orientation,
                                                                        time_stamp_double, moves, remove_tracks)
        try:
            return self.engine.update_raw("update_async", depth.ctypes.data, depth.shape[0], depth.shape[1], depth.strides[0] // 4, masks,
                                          camera_position, camera_orientation, moves, remove_tracks, if_get_freespace)
        except _capi.DspMapError:
            if not getattr(self.engine, "pending_ticket", 0):      # the frame itself was refused (not: an earlier frame failed)
                self._born_tracks, self._known_objects, self._met_labels = saved
            raise

    def wait_result(self, ticket: int, copy_result: bool = True) -> OccupancyResult:
        return self.engine.wait_result(ticket, copy_result)
