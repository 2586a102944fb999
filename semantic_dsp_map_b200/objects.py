"""Host object layer over the C ABI (include/dspmap_c.h, "object layer"; implementation adapter/dsp_object_tracker.h).

What the reference's objectLevelUpdate (include/semantic_dsp_map.h:306-566) and the head of subObjectLevelUpdate (:588-704)
decide per frame: which objects are tracked, which are moving and by which rigid matrix, which are dropped.  Host only - works
without a GPU; `SemanticDSPMap.update` runs it when the caller does not bring moves of their own.
"""
import ctypes as C
from typing import Sequence

import numpy as np

from . import _capi

STATIC, ALWAYS_MOVING, MATCHED_KEYPOINTS, BOX_KEYPOINTS = 0, 1, 2, 3


class ObjectLayer:
    def __init__(self, config, mode=None):
        """mode: which "#if SETTING" flavour of the reference to follow; default = box key points for the ZED2 preset
        (SETTING 3), matched key points otherwise (SETTING 2 rule)."""
        self._lib = _capi.load()
        self.mode = (BOX_KEYPOINTS if config.zed2_filters else MATCHED_KEYPOINTS) if mode is None else int(mode)
        self._h = C.c_void_p()
        cfg = config.c_config()
        rc = self._lib.dspmap_objects_create(C.byref(cfg), self.mode, C.byref(self._h))
        if rc != 0:
            raise _capi.DspMapError(rc, "dspmap_objects_create")
        self._moves = (_capi.ObjectMove * 64)()
        self._dropped = (C.c_int32 * 64)()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dspmap_objects_destroy(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the library may already be gone
            pass

    def _check(self, rc, where):
        if rc != 0:
            raise _capi.DspMapError(rc, where)

    def clear(self):
        self._check(self._lib.dspmap_objects_clear(self._h), "dspmap_objects_clear")

    def set_movement_filter(self, distance_threshold, probability_threshold, increment, decrement):
        self._check(self._lib.dspmap_objects_set_movement_filter(self._h, distance_threshold, probability_threshold, increment, decrement),
                    "dspmap_objects_set_movement_filter")

    def set_seed(self, seed: int):
        self._check(self._lib.dspmap_objects_set_seed(self._h, seed), "dspmap_objects_set_seed")

    def step(self, observations: Sequence[tuple], position, orientation_wxyz, time_stamp: float, max_lost: int):
        """observations: (track_id, label_known, label_id, kpts_current (n, 3), kpts_previous (n, 3) or None).
        Returns (moves, dropped): [(track_id, float32 4x4)], [track_id]."""
        n = len(observations)
        arr = (_capi.ObjectObservation * max(n, 1))()
        keep = []
        for i, (track, known, label_id, cur, prev) in enumerate(observations):
            cur = np.ascontiguousarray(np.asarray(cur, np.float64).reshape(-1, 3))
            keep.append(cur)
            arr[i].track_id, arr[i].label_known, arr[i].label_id, arr[i].n_kpts = int(track), int(bool(known)), int(label_id), len(cur)
            arr[i].kpts_current = cur.ctypes.data if len(cur) else None
            arr[i].kpts_previous = None
            if prev is not None and len(prev) == len(cur) and len(cur):
                prev = np.ascontiguousarray(np.asarray(prev, np.float64).reshape(-1, 3))
                keep.append(prev)
                arr[i].kpts_previous = prev.ctypes.data
        pose = np.empty(7, np.float64)
        pose[:3], pose[3:] = position, orientation_wxyz
        n_moves, n_dropped = C.c_int32(), C.c_int32()
        if n > len(self._moves):
            self._moves, self._dropped = (_capi.ObjectMove * (2 * n))(), (C.c_int32 * (2 * n))()
        rc = self._lib.dspmap_objects_step(self._h, arr, n, C.cast(pose.ctypes.data, C.POINTER(C.c_double)), C.cast(pose.ctypes.data + 24, C.POINTER(C.c_double)),
                                           float(time_stamp), int(max_lost), self._moves, len(self._moves), C.byref(n_moves), self._dropped, len(self._dropped),
                                           C.byref(n_dropped))
        if rc == _capi.DSPMAP_ERR_CAPACITY:      # the tracker has stepped; its plan waits in the handle: fetch it into larger arrays
            cap = 2 * max(n_moves.value, n_dropped.value, 1)
            self._moves, self._dropped = (_capi.ObjectMove * cap)(), (C.c_int32 * cap)()
            rc = self._lib.dspmap_objects_last_plan(self._h, self._moves, cap, C.byref(n_moves), self._dropped, cap, C.byref(n_dropped))
        self._check(rc, "dspmap_objects_step")
        moves = [(int(self._moves[i].track_id), np.asarray(self._moves[i].t, np.float32).reshape(4, 4).copy()) for i in range(n_moves.value)]
        return moves, [int(self._dropped[i]) for i in range(n_dropped.value)]

    def query(self, track_id: int) -> dict:
        tracked, has_flag, moving, p = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        t = np.full(16, np.nan)
        self._check(self._lib.dspmap_objects_query(self._h, int(track_id), C.byref(tracked), C.byref(has_flag), C.byref(moving), C.byref(p),
                                                   C.cast(t.ctypes.data, C.POINTER(C.c_double))), "dspmap_objects_query")
        return dict(tracked=bool(tracked.value), has_flag=bool(has_flag.value), moving=bool(moving.value), moved_probability=p.value,
                    matrix=t.reshape(4, 4) if has_flag.value else None)

    def is_tracked(self, track_id: int) -> bool:
        tracked = C.c_int32()
        self._lib.dspmap_objects_query(self._h, int(track_id), C.byref(tracked), None, None, None, None)
        return bool(tracked.value)

    def forget(self, track_id: int):
        self._check(self._lib.dspmap_objects_forget(self._h, int(track_id)), "dspmap_objects_forget")
