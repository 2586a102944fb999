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
