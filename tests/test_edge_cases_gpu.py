"""Edge cases through the C ABI: degenerate frames, strided inputs, every slot-count template, capacity / unsupported errors,
the device-resident entry point, and full-size (config B) parity + size-independent properties."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, MaskKpts, Scene, SemanticDSPMap, _capi, make_random_table, tiny

from oracle_lib import oracle_map
from parity import PARTICLE, compare_frame, run_pair

pytestmark = pytest.mark.gpu


def _pair(cfg):
    return SemanticDSPMap(cfg), oracle_map(cfg)


@pytest.mark.parametrize("slot_bits", [1, 2, 4])
def test_every_slot_count_template(slot_bits):
    # P = 2 (one usable slot), 4 and 16 slots per voxel: the kernels are templated on the slot bits
    cfg = dataclasses.replace(tiny(BASELINE["B"]), slot_bits=slot_bits)
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 8, make_random_table(cfg.seed, 50_000))


def test_custom_window_and_independent_filter():
    cfg = dataclasses.replace(tiny(BASELINE["B"]), window_half=2, use_independent_filter=True, forgetting_rate=0.5, max_forget_count=3)
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 8, make_random_table(cfg.seed, 50_000))


def test_degenerate_frames():
    cfg = tiny(BASELINE["B"])
    eng, orc = _pair(cfg)
    H, W = cfg.image_height, cfg.image_width
    static = np.zeros((H, W), np.uint8)
    frames = [
        np.zeros((H, W), np.float32),                                   # every pixel below the depth range
        np.full((H, W), cfg.depth_max + 5.0, np.float32),               # every pixel "too far": frees what it sees
        np.full((H, W), np.float32(2.0)),                               # a wall
        np.full((H, W), cfg.depth_max + 5.0, np.float32),               # ... and the wall is gone again
        np.where(np.arange(W)[None, :] % 2 == 0, np.float32(1.0), np.float32(0.0)) * np.ones((H, 1), np.float32),
    ]
    for t, depth in enumerate(frames):
        depth = np.ascontiguousarray(depth, np.float32)
        out = []
        for m in (eng, orc):
            masks = [] if t == 2 else [MaskKpts(65535, "static", static)]    # one frame without any mask at all
            out.append(m.update(depth, masks, [0.0, 0.0, 0.05 * t], [1.0, 0.0, 0.0, 0.0], if_get_freespace=True))
        compare_frame(eng, orc, out[0], out[1], t)
    assert len(out[0].voxel_index) >= 0


def test_strided_host_images():
    # cv::Mat rows may be padded: row_stride > cols for the depth image and the masks
    cfg = tiny(BASELINE["B"])
    eng, orc = _pair(cfg)
    scene = Scene(cfg)
    table = make_random_table(cfg.seed, 50_000)
    eng.engine.set_random_table(table); orc.engine.set_random_table(table)
    for t in range(4):
        fr = scene.frame(t)
        big = np.full((cfg.image_height, cfg.image_width + 24), np.float32(77.0))
        big[:, :cfg.image_width] = fr.depth
        strided_depth = big[:, :cfg.image_width]
        assert not strided_depth.flags.c_contiguous
        masks = []
        for mk in fr.masks:
            mb = np.full((cfg.image_height, cfg.image_width + 8), 255, np.uint8)
            mb[:, :cfg.image_width] = mk.mask
            masks.append(dataclasses.replace(mk, mask=mb[:, :cfg.image_width]))
        # lower_masks makes masks contiguous, so hand the strided views straight to the engine for the depth image
        re = eng.update(strided_depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
        fr2 = scene.frame(t)
        ro = orc.update(fr2.depth, fr2.masks, fr2.position, fr2.orientation, moves=fr2.moves)
        compare_frame(eng, orc, re, ro, t)


def test_device_resident_entry_point_matches_host_entry_point():
    import torch
    cfg = tiny(BASELINE["B"])
    host, dev = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    table = make_random_table(cfg.seed, 50_000)
    host.engine.set_random_table(table); dev.engine.set_random_table(table)
    scene = Scene(cfg)
    for t in range(5):
        fr = scene.frame(t)
        rh = host.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
        fr = scene.frame(t)
        lowered = dev.lower_masks(fr.masks)
        d_depth = torch.from_numpy(fr.depth).cuda()
        d_masks = [torch.from_numpy(np.ascontiguousarray(mk.mask)).cuda() for mk in fr.masks]
        for lm, dm in zip(lowered, d_masks):
            lm["data"], lm["row_stride"] = dm.data_ptr(), dm.stride(0)
        torch.cuda.synchronize()
        dev.engine.update_raw("update_device", d_depth.data_ptr(), fr.depth.shape[0], fr.depth.shape[1], d_depth.stride(0), lowered,
                              fr.position, fr.orientation, fr.moves, fr.remove_tracks, False)
        rd = dev.engine.result_arrays()
        for name in ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state"):
            np.testing.assert_array_equal(getattr(rh, name).view(np.uint8), getattr(rd, name).view(np.uint8), err_msg=f"frame {t} {name}")


def test_capacity_and_unsupported_errors():
    cfg = tiny(BASELINE["B"])
    lib = _capi.load()
    # a visible-particle list that is far too small must be reported, not overrun
    h = C.c_void_p()
    cc = cfg.c_config(0, 1, 0, 64)
    assert lib.dspmap_create(C.byref(cc), C.byref(h)) == 0
    from semantic_dsp_map_b200.map import Engine
    eng = Engine.__new__(Engine)
    eng.lib, eng.prefix, eng.cfg, eng._h, eng._keep = lib, "dspmap_", cfg, h, []
    m = SemanticDSPMap(cfg, _engine=eng)
    scene = Scene(cfg)
    with pytest.raises(_capi.DspMapError) as e:
        for t in range(6):
            fr = scene.frame(t)
            m.update(fr.depth, fr.masks, fr.position, fr.orientation)
    assert e.value.rc == -5       # DSPMAP_ERR_CAPACITY
    # an absurd number of newborn particles per point exceeds the 2^24 candidate ranks: reported as unsupported
    m2 = SemanticDSPMap(dataclasses.replace(cfg, nb_ptc_num_per_point=6000))
    fr = scene.frame(0)
    with pytest.raises(_capi.DspMapError) as e:
        m2.update(fr.depth, fr.masks, fr.position, fr.orientation)
    assert e.value.rc == -7       # DSPMAP_ERR_UNSUPPORTED


@pytest.mark.parametrize("nb", [2, 3])
def test_several_noisy_newborn_particles_per_point(nb):
    # nb_ptc_num_per_point > 1 with the depth-noise model: each copy is perturbed by sigma * table draws, consumed in birth
    # order (semantic_dsp_map.h:1177-1230); the table is short here so the cursor wraps many times
    cfg = dataclasses.replace(tiny(BASELINE["B"]), nb_ptc_num_per_point=nb)
    eng, orc = _pair(cfg)
    run_pair(eng, orc, Scene(cfg), 8, make_random_table(cfg.seed, 7_919))


def test_object_particle_count_matches_the_store():
    cfg = tiny(BASELINE["B"])
    m = SemanticDSPMap(cfg)
    m.engine.set_random_table(make_random_table(cfg.seed, 50_000))
    scene = Scene(cfg)
    for t in range(4):
        fr = scene.frame(t)
        m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
    store = m.engine.dump(_capi.DUMP_PARTICLES).view(PARTICLE)
    live = store["status"] != 0
    live[:: 1 << cfg.slot_bits] = False          # slot 0 is the time particle
    for track in (1, 2, 65534):
        n = C.c_uint32()
        m.engine.check(m.engine.lib.dspmap_object_particle_count(m.engine._h, track, C.byref(n)), "count")
        assert n.value == int((live & (store["track"] == track)).sum()), track


def test_full_size_config_b_parity_and_properties():
    """BASELINE.json's metric configuration at full size: 4 frames against the oracle (result + counters), then
    size-independent properties of the output."""
    cfg = BASELINE["B"]
    eng, orc = _pair(cfg)
    table = make_random_table(cfg.seed)
    scene = Scene(cfg)
    run_pair(eng, orc, scene, 4, table, stages=True, particles=False)
    fr = scene.frame(4)
    r = eng.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=True, moves=fr.moves)
    st = eng.engine.stats()
    v = r.voxel_index.astype(np.int64)
    assert np.all(np.diff(v) > 0), "occupied voxels must come out in ascending storage index"
    assert np.all(np.diff(r.free_voxel_index.astype(np.int64)) > 0)
    assert not np.intersect1d(r.voxel_index, r.free_voxel_index).size
    assert set(np.unique(r.state).tolist()) <= {1, 2}
    assert np.all(r.weight_sum[r.state == 1] > cfg.occupancy_threshold)
    assert st["n_occupied"] == len(v) and st["n_visible"] > 0 and st["n_born"] > 0
    vs = eng.engine.dump(_capi.DUMP_VOXEL_STATE, np.int8)
    assert (vs > 0).sum() == len(v) and (vs == 0).sum() == len(r.free_voxel_index)
    # the emitted position is the voxel's min corner in the (zero-centred) map frame: a lattice point
    cam = np.asarray(fr.position, np.float32)
    lattice = (r.xyz + cam) / np.float32(cfg.voxel_size)
    assert np.abs(lattice - np.round(lattice)).max() < 2e-2


def test_same_inputs_same_bits():
    cfg = tiny(BASELINE["C"])
    a, b = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    table = make_random_table(cfg.seed, 50_000)
    a.engine.set_random_table(table); b.engine.set_random_table(table)
    scene = Scene(cfg)
    for t in range(6):
        fr = scene.frame(t)
        ra = a.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
        fr = scene.frame(t)
        rb = b.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
    np.testing.assert_array_equal(a.engine.dump(_capi.DUMP_PARTICLES), b.engine.dump(_capi.DUMP_PARTICLES))
    np.testing.assert_array_equal(ra.weight_sum.view(np.uint32), rb.weight_sum.view(np.uint32))


class _BackwardScene:
    """The seeded scene with the camera backing away diagonally while rolling and pitching: negative ring steps on every axis,
    wrap-around, a general quaternion and pixels the ZED2 bounding-box gate turns into background (same trajectory as the
    reference-pin scenario s3_backward_diagonal_rotating)."""

    def __init__(self, cfg):
        self.scene = Scene(cfg)

    def frame(self, t):
        fr = self.scene.frame(t)
        fr.position = np.array([-0.9 * t, 0.23 * t, -1.7 * t + 0.4], np.float64)
        q = np.array([np.cos(0.07 * t), 0.3 * np.sin(0.07 * t), np.sin(0.11 * t) * 0.2 + 0.05, -0.25 * np.sin(0.05 * t)], np.float64)
        fr.orientation = q / np.linalg.norm(q)
        return fr


@pytest.mark.parametrize("base", ["B", "C"])
def test_backward_diagonal_rotating_trajectory(base):
    cfg = tiny(BASELINE[base])
    eng, orc = SemanticDSPMap(cfg), oracle_map(cfg)
    run_pair(eng, orc, _BackwardScene(cfg), 8, make_random_table(cfg.seed, 50_000), want_free=True)
    assert eng.engine.stats()["n_occupied"] > 0


def test_multi_chunk_occupancy_pass_matches_oracle():
    """512 x 64 x 512 voxels = 16384 occupancy tiles = four scan chunks: the tile-count scan hands over between blocks and the
    emit pass runs from the non-empty-tile list (the path large maps take); free space on, so most tiles emit."""
    cfg = dataclasses.replace(tiny(BASELINE["B"]), voxel_bits=(9, 6, 9))
    eng, orc = SemanticDSPMap(cfg), oracle_map(cfg)
    for m in (eng, orc):
        m.enable_colors()
    run_pair(eng, orc, Scene(cfg), 4, make_random_table(cfg.seed, 50_000), want_free=True, stages=False, particles=False)
    assert eng.engine.stats()["n_occupied"] > 0


class _QuirkScene:
    """The seeded scene with the two label quirks of the ZED2 table in the static mask (the patch of
    tests/golden/make_ref_golden.py's s3_fast_remove_quirks): value 10 = "sky" (the spelling the Sky filter misses or hits through the
    LUT, pointcloud_tools.h:236) and value 11 = "person", a MOVABLE class in the static mask (its pixels carry track 0 and live one
    frame: the floating-object rule, semantic_dsp_map.h:712-737)."""

    def __init__(self, cfg):
        self.scene = Scene(cfg)

    def frame(self, t):
        fr = self.scene.frame(t)
        cls = fr.masks[0].mask
        h, w = cls.shape
        cls[: h // 6, : w // 3] = 10
        cls[h // 2:, w // 2: w // 2 + w // 10] = 11
        return fr


@pytest.mark.parametrize("follow_compiled_sigma", [False, True])
def test_sky_and_person_values_in_the_static_mask(follow_compiled_sigma):
    # both label quirks on the GPU, every stage, with the ZED2 filters on: "sky" pixels miss the "Sky" filter (pointcloud_tools.h:236), whose
    # key - absent from the table, hence 0 - hits the "person" pixels (instance -1 -> track 0) instead.  The second run
    # also follows a COMPILED reference at its uninitialised-sigma sites (dspmap_params.uninit_sigma_mode: sigma = 0 for invalid
    # pixels and bbox-gated background pixels, which makes the stencil passes divide by zero the IEEE way)
    cfg = tiny(BASELINE["B"])
    eng, orc = _pair(cfg)
    if follow_compiled_sigma:
        for m in (eng, orc):
            m.followCompiledReferenceSigma(True, 0.0)
    run_pair(eng, orc, _QuirkScene(cfg), 8, make_random_table(cfg.seed, 50_000), want_free=True)
    pts = eng.engine.dump(_capi.DUMP_POINTS).view(np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("sigma", "<f4"), ("track", "<u2"), ("label", "u1"), ("valid", "u1")]))
    pts = pts.reshape(cfg.internal_size)
    h, w = pts.shape
    sky = pts[: h // 6 - 1, : w // 3 - 1]
    sky = sky[sky["valid"] == 1]
    assert len(sky) and (sky["track"] == 65524).all(), "'sky' is not the 'Sky' the ZED2 filter looks for: ordinary observations of track 65524"
    person = pts[h // 2 + 1:, w // 2 + 1: w // 2 + w // 10 - 1]
    assert (person["valid"] == 0).all(), "the table has no 'Sky', so the filter's key is instance 0 (operator[] default) = the 'person' pixels"
    if follow_compiled_sigma:
        assert np.all(pts["sigma"][pts["valid"] == 0] == 0.0)


@pytest.mark.parametrize("name", ["A", "C"])
def test_movable_class_in_the_static_mask_lives_one_frame(name):
    # without the ZED2 filters (static mode A, VKITTI2 camera C) the "person" pixels of the static mask give birth to particles
    # of track 0, which no object owns: the floating-object rule (semantic_dsp_map.h:712-737) removes them in the next frame;
    # "sky" pixels are ordinary observations here
    cfg = tiny(BASELINE[name])
    eng, orc = _pair(cfg)
    run_pair(eng, orc, _QuirkScene(cfg), 5, make_random_table(cfg.seed, 50_000), want_free=True)
    assert eng.engine.stats()["n_removed"] > 0, "the track-0 particles of the 'person' pixels must be removed again"


def test_more_removals_than_one_batch_and_too_many_moves():
    # the reference has no limit on removals per frame (every unregistered detection is one): 300 go in batches; more moves
    # than the engine's table holds are refused BEFORE anything changes
    cfg = tiny(BASELINE["B"])
    eng, orc = _pair(cfg)
    table = make_random_table(cfg.seed, 50_000)
    for m in (eng, orc):
        m.engine.set_random_table(table)
    scene = Scene(cfg)
    for t in range(5):
        outs = []
        for m in (eng, orc):
            fr = scene.frame(t)
            rm = list(range(3, 303)) + [2] if t == 3 else []
            outs.append(m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=rm))
        compare_frame(eng, orc, outs[0], outs[1], t)
    fr = scene.frame(5)
    before = eng.engine.stats()["time_stamp"]
    ring = eng.engine.dump(_capi.DUMP_RING, np.uint32).copy()
    with pytest.raises(_capi.DspMapError) as e:
        eng.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=[(k + 1, np.eye(4, dtype=np.float32)) for k in range(400)])
    assert e.value.rc == -5
    np.testing.assert_array_equal(eng.engine.dump(_capi.DUMP_RING, np.uint32), ring)      # refused before the ring moved
    fr = scene.frame(5)
    re = eng.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
    fr = scene.frame(5)
    ro = orc.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
    compare_frame(eng, orc, re, ro, 5)
    assert eng.engine.stats()["time_stamp"] == before + 1


def test_mask_of_another_size_is_refused():
    cfg = tiny(BASELINE["B"])
    m = SemanticDSPMap(cfg)
    depth = np.full((cfg.image_height, cfg.image_width), 2.0, np.float32)
    short = np.zeros((cfg.image_height - 8, cfg.image_width), np.uint8)
    with pytest.raises(_capi.DspMapError) as e:
        m.update(depth, [MaskKpts(65535, "static", short)], np.zeros(3), np.array([1.0, 0, 0, 0]))
    assert e.value.rc == -3
