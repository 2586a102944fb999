"""Colouring of the occupied list on the GPU (dspmap_set_colors; getOccupancyResult, semantic_dsp_map.h:1274-1351):
the device's two colour-space conversions against the oracle's (themselves pinned to cv2 in tests/test_colour_cpu.py) on
all 2^24 inputs, and the RGB bytes of every occupied voxel against the oracle in display and evaluation format."""
import dataclasses

import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, make_random_table, tiny
from oracle_lib import load_oracle, oracle_map
from parity import run_pair

pytestmark = pytest.mark.gpu

LABEL_BGR = {k: ((37 * k) % 256, (91 * k + 13) % 256, (53 * k + 101) % 256) for k in range(20)}


def test_device_colour_conversions_equal_the_oracles_on_all_inputs():
    m = SemanticDSPMap(tiny(BASELINE["B"]))
    n = 1 << 24
    a, b = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    m.engine.check(m.engine.lib.dspmap_selftest_colour(m.engine._h, a.ctypes.data, b.ctypes.data), "dspmap_selftest_colour")
    i = np.arange(n, dtype=np.uint32)
    src = np.ascontiguousarray(np.stack([i >> 16, (i >> 8) & 255, i & 255], -1).astype(np.uint8))
    lib = load_oracle()
    want = np.empty_like(src)
    unpack = lambda w: np.stack([w & 255, (w >> 8) & 255, (w >> 16) & 255], -1).astype(np.uint8)  # noqa: E731
    lib.orc_rgb_to_hsv(src.ctypes.data, want.ctypes.data, n)
    np.testing.assert_array_equal(unpack(a), want)
    lib.orc_hsv_to_rgb(src.ctypes.data, want.ctypes.data, n)
    ok = src[:, 0] < 180
    np.testing.assert_array_equal(unpack(b)[ok], want[ok])


@pytest.mark.parametrize("base,eval_format,zero_center", [("B", False, True), ("B", True, False), ("C", False, False), ("A", False, False)])
def test_occupied_colours_match_the_oracle(base, eval_format, zero_center):
    cfg = dataclasses.replace(tiny(BASELINE[base]), zero_center=zero_center)
    eng, orc = SemanticDSPMap(cfg), oracle_map(cfg)
    perm = np.random.default_rng(3).permutation(256)
    for m in (eng, orc):
        m.setVisualizeOptions(zero_center, eval_format)
        m.enable_colors(label_bgr=LABEL_BGR, instance_map=perm)
    run_pair(eng, orc, Scene(cfg), 6, make_random_table(cfg.seed, 50_000), want_free=True)
    fr = Scene(cfg).frame(6)
    r = eng.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
    assert len(np.unique(r.rgb, axis=0)) > 3 and r.rgb.any()


def test_colours_off_by_default_and_bad_scheme_rejected():
    cfg = tiny(BASELINE["B"])
    eng = SemanticDSPMap(cfg)
    fr = Scene(cfg).frame(0)
    eng.update(fr.depth, fr.masks, fr.position, fr.orientation)
    fr = Scene(cfg).frame(1)
    r = eng.update(fr.depth, fr.masks, fr.position, fr.orientation)
    assert len(r.rgb) == len(r.voxel_index) and not r.rgb.any()
    with pytest.raises(Exception):
        eng.enable_colors(scheme=2)
