"""Pipelined frames (dspmap_update_async / dspmap_wait_result): frame t+1 is staged and enqueued while frame t runs and frame
t-1's result travels to the host.  Nothing about the arithmetic changes, so everything - every result field of every frame,
the particle store, the counters - must equal the synchronous calls bit for bit, and through them the oracle."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, _capi, make_random_table, tiny

from oracle_lib import oracle_map
from parity import COUNT_STATS, PARTICLE, assert_same

pytestmark = pytest.mark.gpu

FIELDS = ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state", "free_voxel_index", "free_xyz", "rgb")


def _same_result(a, b, frame):
    for name in FIELDS:
        assert_same("result." + name, getattr(a, name), getattr(b, name), frame)


def _sync_run(m, scene, n, want_free=False):
    out = []
    for t in range(n):
        fr = scene.frame(t)
        out.append(m.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=want_free, moves=fr.moves, remove_tracks=fr.remove_tracks))
    return out


def _pipelined_run(m, scene, n, want_free=False, copy=True):
    """The loop of include/dspmap_c.h: submit frame t, then collect frame t-1."""
    out, prev = [], None
    for t in range(n):
        fr = scene.frame(t)
        k = m.update_async(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=want_free, moves=fr.moves, remove_tracks=fr.remove_tracks)
        assert k == t + 1                                  # the ticket is the frame's time stamp
        if prev is not None:
            out.append(m.wait_result(prev, copy_result=copy))
        prev = k
    out.append(m.wait_result(prev, copy_result=copy))
    return out


@pytest.mark.parametrize("name,want_free", [("B", False), ("C", True), ("A", False)])
def test_pipelined_equals_synchronous_and_oracle(name, want_free):
    cfg = tiny(BASELINE[name])
    table = make_random_table(cfg.seed, 50_000)
    maps = [SemanticDSPMap(cfg), SemanticDSPMap(cfg), oracle_map(cfg)]
    for m in maps:
        m.enable_colors()
        m.engine.set_random_table(table)
    n = 14
    sync = _sync_run(maps[0], Scene(cfg), n, want_free)
    pipe = _pipelined_run(maps[1], Scene(cfg), n, want_free)
    orc = _sync_run(maps[2], Scene(cfg), n, want_free)
    assert any(len(r.voxel_index) for r in sync)
    for t in range(n):
        _same_result(pipe[t], sync[t], t)
        _same_result(pipe[t], orc[t], t)
    assert_same("particles", maps[1].engine.dump(_capi.DUMP_PARTICLES).view(PARTICLE), maps[2].engine.dump(_capi.DUMP_PARTICLES).view(PARTICLE), n)
    sp, ss = maps[1].engine.stats(), maps[0].engine.stats()
    for k in COUNT_STATS:
        assert sp[k] == ss[k], (k, sp[k], ss[k])


def test_views_stay_valid_for_two_more_frames():
    cfg = tiny(BASELINE["B"])
    table = make_random_table(cfg.seed, 50_000)
    a, b = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    for m in (a, b):
        m.enable_colors()
        m.engine.set_random_table(table)
    n = 10
    sync = _sync_run(a, Scene(cfg), n)
    scene, held = Scene(cfg), {}
    tickets = []
    for t in range(n):
        fr = scene.frame(t)
        tickets.append(b.update_async(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks))
        if t >= 1:
            held[t - 1] = b.wait_result(tickets[t - 1], copy_result=False)       # views of the pinned buffers
        if t >= 2:
            # frame t-2's views: two more frames (t-1, t) have been submitted and wait_result(t) has not been called yet
            _same_result(held[t - 2], sync[t - 2], t - 2)
    with pytest.raises(_capi.DspMapError):
        b.wait_result(tickets[n - 3])                                            # no longer held
    _same_result(b.wait_result(tickets[n - 1]), sync[n - 1], n - 1)


def test_mixing_synchronous_calls_drains_the_pipeline():
    cfg = tiny(BASELINE["B"])
    table = make_random_table(cfg.seed, 50_000)
    a, b = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    for m in (a, b):
        m.engine.set_random_table(table)
    sync = _sync_run(a, Scene(cfg), 9)
    scene = Scene(cfg)
    for t in range(9):
        fr = scene.frame(t)
        if t % 3 == 2:      # a synchronous frame in between: waits for the frames in flight, then behaves as always
            r = b.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks)
            _same_result(r, sync[t], t)
        else:
            k = b.update_async(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks)
            if t % 3 == 1:  # dump / stats / get_result while a frame is in flight
                assert b.engine.stats()["time_stamp"] == k
                assert_same("particles", b.engine.dump(_capi.DUMP_PARTICLES).view(PARTICLE), _replay_store(cfg, table, t + 1), t)
                _same_result(b.engine.result_arrays(), sync[t], t)
    b.clear()
    a.clear()
    for m in (a, b):
        m.engine.set_random_table(table)
    s2, p2 = _sync_run(a, Scene(cfg), 3), _pipelined_run(b, Scene(cfg), 3)
    for t in range(3):
        _same_result(p2[t], s2[t], t)


def _replay_store(cfg, table, n):
    m = SemanticDSPMap(cfg)
    m.engine.set_random_table(table)
    _sync_run(m, Scene(cfg), n)
    return m.engine.dump(_capi.DUMP_PARTICLES).view(PARTICLE)


def test_device_resident_pipelined_frames():
    """dspmap_update_device_async: images already in device memory, result left on the device."""
    import torch
    cfg = tiny(BASELINE["B"])
    table = make_random_table(cfg.seed, 50_000)
    a, b = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    for m in (a, b):
        m.enable_colors()
        m.engine.set_random_table(table)
    n = 8
    sync = _sync_run(a, Scene(cfg), n)
    scene, eng, keep = Scene(cfg), b.engine, []
    tickets = []
    for t in range(n):
        fr = scene.frame(t)
        lowered = b.lower_masks(fr.masks)
        d_depth = torch.from_numpy(fr.depth).cuda()
        d_masks = [torch.from_numpy(np.ascontiguousarray(mk.mask)).cuda() for mk in fr.masks]
        for lm, dm in zip(lowered, d_masks):
            lm["data"], lm["row_stride"] = dm.data_ptr(), dm.stride(0)
        keep.append((d_depth, d_masks))
        torch.cuda.synchronize()
        tickets.append(eng.update_raw("update_device_async", d_depth.data_ptr(), fr.depth.shape[0], fr.depth.shape[1], d_depth.stride(0), lowered,
                                      fr.position, fr.orientation, fr.moves, fr.remove_tracks, False))
        if t >= 1:
            r = _capi.Result()
            eng.check(eng.lib.dspmap_wait_result(eng._h, tickets[t - 1], C.byref(r)), "dspmap_wait_result")
            assert r.count == len(sync[t - 1].voxel_index)
            if r.count:
                src = torch.as_tensor(_DevArray(r.voxel_index, r.count), device="cuda")
                assert np.array_equal(src.cpu().numpy().view(np.uint32), sync[t - 1].voxel_index)
    _same_result(eng.result_arrays(), sync[n - 1], n - 1)       # synchronous read of the last frame drains the pipeline


class _DevArray:
    """__cuda_array_interface__ over a raw device pointer."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
