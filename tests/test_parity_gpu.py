"""CUDA engine vs CPU oracle, through the C ABI, stage by stage.  Bit-exact for everything (indices, labels, AND floats:
the engine evaluates every float expression in the oracle's order without FMA contraction)."""
import dataclasses

import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, make_random_table, tiny

from oracle_lib import oracle_map
from parity import run_pair

pytestmark = pytest.mark.gpu


def _pair(cfg):
    return SemanticDSPMap(cfg), oracle_map(cfg)


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_tiny_config_every_stage(name):
    cfg = tiny(BASELINE[name])
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 12, make_random_table(cfg.seed, 50_000))


def test_freespace_and_no_zero_center():
    cfg = dataclasses.replace(tiny(BASELINE["B"]), zero_center=False)
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 6, make_random_table(cfg.seed, 50_000), want_free=True)


def test_no_depth_noise_birth_path():
    # mapping.cpp:109-118: if_consider_depth_noise false -> addNewbornParticleAndResample, sigma 0.1
    cfg = dataclasses.replace(tiny(BASELINE["B"]), consider_depth_noise=False, detection_probability=1.0, noise_number=0.001,
                              occupancy_threshold=0.1, nb_ptc_num_per_point=3)
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 8, make_random_table(cfg.seed, 50_000))


def test_fast_motion_ring_wrap():
    # 2 m/frame on a 9.6 m map: the ring wraps several times and large steps are sub-stepped (operations.h:71-90)
    cfg = dataclasses.replace(tiny(BASELINE["B"]), speed=2.0)
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 12, make_random_table(cfg.seed, 50_000))


def test_hoisted_reciprocal_division_is_ieee_exact():
    # the stencil kernels divide by a per-thread sigma with a hoisted reciprocal (InvDiv); it must equal `/` bit for bit
    import ctypes as C
    cfg = tiny(BASELINE["B"])
    eng = SemanticDSPMap(cfg).engine
    for seed, (b_lo, b_hi, a_max) in enumerate([(0.05, 1.5, 64.0), (0.3, 0.6, 16.0), (0.1, 0.1000001, 4.0), (1e-3, 50.0, 1e4)]):
        bad = C.c_uint64(123)
        eng.check(eng.lib.dspmap_selftest_division(eng._h, 1 << 30, 1234 + seed, b_lo, b_hi, a_max, C.byref(bad)), "selftest")
        assert bad.value == 0, (b_lo, b_hi, a_max, bad.value)


@pytest.mark.parametrize("name", ["A_static", "B_zed2", "C_vkitti2", "B_fast_ring_wrap", "B_no_depth_noise"])
def test_engine_reproduces_committed_golden_vectors(name):
    # no oracle in the loop: the CUDA engine against tests/golden/*.npz (results + SHA-256 of the full particle store)
    from test_oracle_golden_cpu import check_against_golden
    check_against_golden(name, lambda cfg: SemanticDSPMap(cfg))


def test_error_codes_through_the_c_abi():
    from semantic_dsp_map_b200 import _capi
    cfg = tiny(BASELINE["B"])
    m = SemanticDSPMap(cfg)
    with pytest.raises(_capi.DspMapError) as e:
        m.update(np.ones((cfg.image_height, cfg.image_width + 2), np.float32), [], np.zeros(3), np.array([1.0, 0, 0, 0]))
    assert e.value.rc == -3    # DSPMAP_ERR_IMAGE_SIZE instead of the reference's silent out-of-bounds read
    with pytest.raises(TypeError):
        m.update(np.ones((cfg.image_height, cfg.image_width), np.float64), [], np.zeros(3), np.array([1.0, 0, 0, 0]))
    # an empty frame list is fine: a frame with no masks at all
    r = m.update(np.full((cfg.image_height, cfg.image_width), 2.0, np.float32), [], np.zeros(3), np.array([1.0, 0, 0, 0]))
    assert r.voxel_index.size >= 0


def test_clear_then_continue_matches_oracle():
    # SemanticDSPMap::clear (semantic_dsp_map.h:74-81): the store, the slab stamps and the time stamp reset; the ring
    # offsets stay where the last pose left them
    cfg = tiny(BASELINE["B"])
    eng, orc = _pair(cfg)
    table = make_random_table(cfg.seed, 50_000)
    scene = Scene(cfg)
    run_pair(eng, orc, scene, 4, table, check_frames=(3,))
    eng.clear(); orc.clear()
    from parity import compare_frame
    for t in range(4, 8):
        fr = scene.frame(t)
        re = eng.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
        fr = scene.frame(t)
        ro = orc.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
        compare_frame(eng, orc, re, ro, t)


def test_track_removed_and_moved_in_one_frame():
    # the reference removes a track AFTER moving it (semantic_dsp_map.h:588-737); plus a plain removal of another track
    cfg = tiny(BASELINE["B"])
    eng, orc = _pair(cfg)
    table = make_random_table(cfg.seed, 50_000)
    scene = Scene(cfg)
    from parity import compare_frame
    for m in (eng, orc):
        m.engine.set_random_table(table)
    for t in range(7):
        outs = []
        for m in (eng, orc):
            fr = scene.frame(t)
            rm = []
            if t == 4:
                rm = [1]                      # track 1 is moving in this scene: moved and removed in the same frame
            if t == 5:
                rm = [2]                      # plain removal
            outs.append(m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=rm))
        compare_frame(eng, orc, outs[0], outs[1], t)
