"""dspmap_params.reference_iteration_order: the pixel lists in the order the reference's breadth-first search visits the voxels
in (mc_ring/operations.h:1326-1456), replayed level by level on the device.  Held against the oracle in ITS reference order
(order 0: a real FIFO search) - which tests/test_reference_pin_cpu.py pins bit for bit to the reference's own code - on scenes
without object moves (the one other order the reference depends on, std::unordered_set iteration over an object's particles,
is not reproducible even between two runs of the reference: its jitter table is seeded from std::random_device).
Every stage dump - the pixel lists themselves among them -, every particle record and every result field, bit for bit."""
import dataclasses

import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, _capi, make_random_table, tiny

from oracle_lib import oracle_map
from parity import compare_frame

pytestmark = pytest.mark.gpu


def _run(cfg, n_frames, stages=True, particles=True, want_free=True):
    eng, orc = SemanticDSPMap(cfg), oracle_map(cfg, order=0)
    eng.setReferenceIterationOrder(True)
    table = make_random_table(cfg.seed, 50_000)
    for m in (eng, orc):
        m.enable_colors()
        m.engine.set_random_table(table)
    scene = Scene(cfg)
    differs_from_canonical = False
    for t in range(n_frames):
        out = []
        for m in (eng, orc):
            fr = scene.frame(t)
            out.append(m.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=want_free, moves=[], remove_tracks=[2] if t == 5 else []))
        compare_frame(eng, orc, out[0], out[1], t, stages, particles)
        if stages:
            lists = eng.engine.dump(_capi.DUMP_PIXEL_LISTS, np.uint32)
            off = eng.engine.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32)
            asc = np.concatenate([np.sort(lists[a:b]) for a, b in zip(off[:-1], off[1:]) if b - a > 1] or [np.zeros(0, np.uint32)])
            cur = np.concatenate([lists[a:b] for a, b in zip(off[:-1], off[1:]) if b - a > 1] or [np.zeros(0, np.uint32)])
            differs_from_canonical |= bool((asc != cur).any())
    return eng, differs_from_canonical


@pytest.mark.parametrize("name", ["B", "A", "C"])
def test_reference_order_matches_the_oracles_search_bit_for_bit(name):
    eng, differs = _run(tiny(BASELINE[name]), 8)
    assert eng.engine.stats()["n_visible"] > 0
    assert differs, "the scene never produced a pixel list whose reference order differs from ascending particle index: the test shows nothing"


def test_reference_order_with_sixteen_slots_and_a_yawing_camera():
    cfg = dataclasses.replace(tiny(BASELINE["B"]), slot_bits=4)
    _run(cfg, 6)


def test_reference_order_at_full_size_config_b():
    # the benchmark configuration at full size: 6 frames, results + counters + particle store digest
    cfg = BASELINE["B"]
    eng, orc = SemanticDSPMap(cfg), oracle_map(cfg, order=0)
    eng.setReferenceIterationOrder(True)
    table = make_random_table(cfg.seed)
    for m in (eng, orc):
        m.enable_colors()
        m.engine.set_random_table(table)
    scene = Scene(cfg)
    for t in range(6):
        out = []
        for m in (eng, orc):
            fr = scene.frame(t)
            out.append(m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=[], remove_tracks=[]))
        compare_frame(eng, orc, out[0], out[1], t, stages=(t in (1, 5)), particles=False)
        he, ho = eng.engine.dump(_capi.DUMP_PARTICLES_HASH, np.uint64), orc.engine.dump(_capi.DUMP_PARTICLES_HASH, np.uint64)
        assert he[0] == ho[0], f"frame {t}: particle store digest"
    assert eng.engine.stats()["n_visible"] > 50_000
