"""Host-side logic: label tables (with the reference's default-insert quirks), mask lowering, the synthetic scene."""
import sys

import numpy as np

from semantic_dsp_map_b200 import BASELINE, SETTINGS, LabelTables, MaskKpts, Scene, SemanticDSPMap, ZED2_ROWS, make_random_table, tiny
from oracle_lib import oracle_map


def test_zed2_label_luts_reproduce_reference_quirks():
    t = LabelTables.from_rows(ZED2_ROWS)
    v2t = t.static_value_to_track()
    # static-mask pixel value v means label id v+1 (pointcloud_tools.h:137)
    assert v2t[0] == 65534 and v2t[9] == 65525          # road, terrain
    # movable classes have no fixed instance id: operator[] default-inserts 0 (pointcloud_tools.h:137-138)
    assert v2t[11] == 0 and v2t[13] == 0                # person, Vehicle
    assert v2t[200] == 0                                # unknown class value
    # the ZED2 sky filter looks up "Sky" but the CSV spells it "sky": the key is absent -> 0 (pointcloud_tools.h:236)
    assert t.sky_track() == 0
    assert t.max_movable_instance_id == 65516           # min static instance id - 1 (object_info_handler.h:65-69)
    tracks, labels = t.static_track_labels()
    assert dict(zip(tracks.tolist(), labels.tolist()))[65534] == 1


def test_builtin_label_tables():
    t = LabelTables.builtin()
    assert t.sky_track() == 65533 and t.max_movable_instance_id == 65523
    assert t.label_id_of("Car") == 14 and t.label_id_of("nonexistent") == 0


def test_mask_lowering_bbox_and_track_wrap():
    cfg = tiny(BASELINE["B"])
    m = oracle_map(cfg)
    mask = np.zeros((cfg.image_height, cfg.image_width), np.uint8)
    kp = np.array([[-3.0, -2.0, -5.0], [-1.0, -0.5, -4.0]])
    big = m.labels.max_movable_instance_id + 7
    lowered = m.lower_masks([MaskKpts(65535, "static", mask), MaskKpts(big, "Vehicle", mask, kpts_current=kp)])
    assert lowered[0]["is_static"] and not lowered[1]["is_static"]
    assert lowered[1]["track_id"] == big % m.labels.max_movable_instance_id        # semantic_dsp_map.h:179-186
    assert lowered[1]["label_id"] == 14
    # min starts at DBL_MAX, max at the smallest POSITIVE double (pointcloud_tools.h:178-180): all-negative key points
    # leave max at ~0, so the +1 m margin gives ~1.0 instead of kpt_max + 1
    assert lowered[1]["bbox_min"] == [-4.0, -3.0, -6.0]
    assert lowered[1]["bbox_max"] == [sys.float_info.min + 1.0] * 3
    # no key points: min stays DBL_MAX - 1
    lowered = m.lower_masks([MaskKpts(3, "Vehicle", mask)])
    assert lowered[0]["bbox_min"][0] > 1e300


def test_scene_is_deterministic_and_shaped():
    cfg = tiny(BASELINE["B"])
    a, b = Scene(cfg).frame(3), Scene(cfg).frame(3)
    assert a.depth.dtype == np.float32 and a.depth.shape == (cfg.image_height, cfg.image_width)
    np.testing.assert_array_equal(a.depth, b.depth)
    assert [m.track_id for m in a.masks] == [m.track_id for m in b.masks]
    assert a.masks[0].label == "static" and a.masks[0].mask.dtype == np.uint8
    assert not np.isnan(a.depth).any()
    far = (a.depth > cfg.depth_max).mean()
    assert 0.01 < far < 0.5
    np.testing.assert_array_equal(make_random_table(7, 100), make_random_table(7, 100))


def test_internal_size_and_presets():
    assert BASELINE["B"].internal_size == (360, 640) and BASELINE["C"].internal_size == (375, 1242)
    assert BASELINE["B"].n_slots == 33_554_432 and BASELINE["E"].n_slots == 2 ** 31
    assert SETTINGS[3].voxel_bits == (7, 5, 7) and SETTINGS[3].boost_mode


def test_public_interface_mirrors_the_reference_class():
    # include/semantic_dsp_map.h:74-251 public section
    for name in ("clear", "setMapParameters", "setMapOptions", "setVisualizeOptions", "setBeyesianMovementParameters",
                 "setDepthNoiseModelParameters", "setOccupancyThreshold", "update"):
        assert callable(getattr(SemanticDSPMap, name))


def test_oracle_rejects_wrong_image_size():
    from semantic_dsp_map_b200 import _capi
    cfg = tiny(BASELINE["B"])
    m = oracle_map(cfg)
    depth = np.ones((cfg.image_height + 1, cfg.image_width), np.float32)
    try:
        m.update(depth, [], np.zeros(3), np.array([1.0, 0, 0, 0]))
    except _capi.DspMapError as e:
        assert e.rc == -3
    else:
        raise AssertionError("expected DSPMAP_ERR_IMAGE_SIZE")


def test_label_tables_keep_the_colour_columns(tmp_path):
    """g_label_color_map_default: built-in colours (utils/data_base.h:216-232) and the colour columns of an object-info CSV
    (object_info_handler.h:56), as (b, g, r); enable_colors() sends them to the engine by default."""
    from semantic_dsp_map_b200 import LabelTables
    b = LabelTables.builtin()
    assert b.label_colors[7] == (100, 60, 100) and b.label_colors[14] == (80, 127, 255) and len(b.label_colors) == 15
    p = tmp_path / "object_info.csv"
    p.write_text("label_id,label,instance_id(-1 means not fixed),color_b,color_g,color_r\n0,background,65535,0,0,0\n1,road,65534,128,64,128\n14,Vehicle,-1,142,0,0\n")
    t = LabelTables.read_object_info(str(p))
    assert t.label_colors == {0: (0, 0, 0), 1: (128, 64, 128), 14: (142, 0, 0)}
    assert t.max_movable_instance_id == 65533 and t.label_id_map["Vehicle"] == 14
    assert LabelTables.from_rows([(1, "road", 65534)]).label_colors == {}          # rows without colour columns: no colours
