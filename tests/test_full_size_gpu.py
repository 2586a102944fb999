"""BASELINE.json's configurations at FULL size, in the steady state: the CUDA engine against the CPU oracle on the seeded
sequence of SURVEY.md section 8(d), every frame's result and counters, the stage dumps and the particle store at the frames
the survey names (1, 2, 5, 30; the sequence is cut at 30 so that the suite stays within minutes - bench.py carries the
comparison to frame 130, `parity_vs_reference`).

  B  256 x 64 x 256 voxels x 8 slots, 1280 x 720 (boost), ZED2 parameters, moving boxes      31 frames
  C  same grid, VKITTI2 camera 1242 x 375, 11 x 11 window, voxel 0.2 m                        31 frames
  D  512 x 128 x 512 voxels x 16 slots (12.9 GB store), 2 m per frame                         11 frames, store compared by digest

Bit-exact for every integer AND every float (tests/parity.py)."""
import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, _capi, make_random_table

from oracle_lib import oracle_map
from parity import COUNT_STATS, PARTICLE, POINT, assert_same

pytestmark = pytest.mark.gpu

RESULT_FIELDS = ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state", "free_voxel_index", "free_xyz", "rgb")


def run_full(name, n_frames, stage_frames, store_frames, store_digest_only=False, free_frames=()):
    cfg = BASELINE[name]
    eng, orc = SemanticDSPMap(cfg), oracle_map(cfg)
    table = make_random_table(cfg.seed)
    for m in (eng, orc):
        m.enable_colors()
        m.engine.set_random_table(table)
    scene = Scene(cfg)
    n_occ = n_vis = 0
    for t in range(n_frames):
        want_free = t in free_frames
        fr = scene.frame(t)
        re = eng.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=want_free, moves=fr.moves, remove_tracks=fr.remove_tracks)
        fr = scene.frame(t)
        ro = orc.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=want_free, moves=fr.moves, remove_tracks=fr.remove_tracks)
        for f in RESULT_FIELDS:
            assert_same("result." + f, getattr(re, f), getattr(ro, f), t)
        se, so = eng.engine.stats(), orc.engine.stats()
        for k in COUNT_STATS:
            assert se[k] == so[k], f"frame {t}: stats.{k} engine {se[k]} oracle {so[k]}"
        e, o = eng.engine, orc.engine
        if t in stage_frames:
            assert_same("ring", e.dump(_capi.DUMP_RING, np.uint32), o.dump(_capi.DUMP_RING, np.uint32), t)
            assert_same("points", e.dump(_capi.DUMP_POINTS).view(POINT), o.dump(_capi.DUMP_POINTS).view(POINT), t)
            assert_same("pixel_offsets", e.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32), o.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32), t)
            assert_same("pixel_lists", e.dump(_capi.DUMP_PIXEL_LISTS, np.uint32), o.dump(_capi.DUMP_PIXEL_LISTS, np.uint32), t)
            assert_same("ck", e.dump(_capi.DUMP_CK, np.float32), o.dump(_capi.DUMP_CK, np.float32), t)
            assert_same("voxel_state", e.dump(_capi.DUMP_VOXEL_STATE, np.int8), o.dump(_capi.DUMP_VOXEL_STATE, np.int8), t)
        if t in store_frames:
            he, ho = e.dump(_capi.DUMP_PARTICLES_HASH, np.uint64), o.dump(_capi.DUMP_PARTICLES_HASH, np.uint64)
            assert he[0] == ho[0], f"frame {t}: particle store digest engine {he[0]:#x} oracle {ho[0]:#x}"
            if not store_digest_only:
                assert_same("particles", e.dump(_capi.DUMP_PARTICLES).view(PARTICLE), o.dump(_capi.DUMP_PARTICLES).view(PARTICLE), t)
        n_occ, n_vis = se["n_occupied"], se["n_visible"]
    return n_occ, n_vis, eng.engine.stats()


def test_config_B_full_size_31_frames():
    n_occ, n_vis, st = run_full("B", 31, stage_frames=(1, 2, 5, 30), store_frames=(5, 30), free_frames=(30,))
    assert n_occ > 50_000 and n_vis > 100_000 and st["n_resampled"] > 0 and st["n_moved"] > 0     # steady state: resampling and object moves are live


def test_config_C_full_size_31_frames():
    n_occ, n_vis, st = run_full("C", 31, stage_frames=(1, 2, 5, 30), store_frames=(30,), store_digest_only=True)
    assert n_occ > 20_000 and n_vis > 50_000 and st["n_resampled"] > 0


def test_config_D_full_size_11_frames_16_slots():
    # 16 slots per voxel, 2 m per frame (ring steps of 13 voxels), 33.5 M voxels: the multi-chunk occupancy path at its real size
    n_occ, n_vis, st = run_full("D", 11, stage_frames=(1, 5, 10), store_frames=(5, 10), store_digest_only=True, free_frames=(10,))
    assert n_occ > 20_000 and n_vis > 50_000
