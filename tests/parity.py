"""Stage-by-stage comparison of the CUDA engine with the CPU oracle on identical seeded frames."""
import numpy as np

from semantic_dsp_map_b200 import _capi

PARTICLE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("ts", "<u2"), ("track", "<u2"), ("label", "<u2"),
                     ("status", "u1"), ("forget", "u1")])
POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("sigma", "<f4"), ("track", "<u2"), ("label", "u1"), ("valid", "u1")])
COUNT_STATS = ("time_stamp", "n_swept_voxels", "n_swept_nonempty", "n_visible", "n_valid_points", "n_born", "n_birth_voxels",
               "n_resampled", "n_occupied", "n_nonempty_voxels", "n_moved", "n_removed")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_same(name, a, b, frame):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"frame {frame}: {name} shape {a.shape} != {b.shape}"
    if a.dtype.fields:
        for fld in a.dtype.names:
            assert_same(f"{name}.{fld}", a[fld], b[fld], frame)
        return
    neq = bits(a) != bits(b)
    if neq.any():
        idx = np.flatnonzero(neq.reshape(-1))[:5]
        raise AssertionError(f"frame {frame}: {name} differs at {int(neq.sum())} of {a.size} entries; first {idx.tolist()}: "
                             f"engine {a.reshape(-1)[idx].tolist()} oracle {b.reshape(-1)[idx].tolist()}")


def compare_frame(eng_map, orc_map, res_e, res_o, frame, stages=True, particles=True):
    """eng_map / orc_map are SemanticDSPMap objects over the CUDA engine and the oracle; res_* their update() results."""
    e, o = eng_map.engine, orc_map.engine
    if stages:
        assert_same("ring", e.dump(_capi.DUMP_RING, np.uint32), o.dump(_capi.DUMP_RING, np.uint32), frame)
        assert_same("points", e.dump(_capi.DUMP_POINTS).view(POINT), o.dump(_capi.DUMP_POINTS).view(POINT), frame)
        assert_same("pixel_offsets", e.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32), o.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32), frame)
        assert_same("pixel_lists", e.dump(_capi.DUMP_PIXEL_LISTS, np.uint32), o.dump(_capi.DUMP_PIXEL_LISTS, np.uint32), frame)
        assert_same("ck", e.dump(_capi.DUMP_CK, np.float32), o.dump(_capi.DUMP_CK, np.float32), frame)
    if particles:
        pe, po = e.dump(_capi.DUMP_PARTICLES).view(PARTICLE), o.dump(_capi.DUMP_PARTICLES).view(PARTICLE)
        assert_same("particles", pe, po, frame)
    assert_same("voxel_state", e.dump(_capi.DUMP_VOXEL_STATE, np.int8), o.dump(_capi.DUMP_VOXEL_STATE, np.int8), frame)
    for name in ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state", "free_voxel_index", "free_xyz", "rgb"):
        assert_same("result." + name, getattr(res_e, name), getattr(res_o, name), frame)
    se, so = e.stats(), o.stats()
    for k in COUNT_STATS:
        assert se[k] == so[k], f"frame {frame}: stats.{k} engine {se[k]} oracle {so[k]}"


def run_pair(eng_map, orc_map, scene, n_frames, table, check_frames=None, want_free=False, stages=True, particles=True):
    for m in (eng_map, orc_map):
        m.engine.set_random_table(table)
    for t in range(n_frames):
        fr = scene.frame(t)
        re = eng_map.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=want_free, moves=fr.moves,
                            remove_tracks=fr.remove_tracks)
        fr2 = scene.frame(t)  # update() may rewrite track ids in place, like the reference; give the oracle a fresh copy
        ro = orc_map.update(fr2.depth, fr2.masks, fr2.position, fr2.orientation, if_get_freespace=want_free, moves=fr2.moves,
                            remove_tracks=fr2.remove_tracks)
        if check_frames is None or t in check_frames:
            compare_frame(eng_map, orc_map, re, ro, t, stages, particles)
