"""The whole Python host stack on the CPU against the reference's node: the recorded replay of the "moving" scene
(tests/golden/make_node_golden.py) goes through wire.py (ROS 1 message decoding, syncCallback's rules), SemanticDSPMap.update
(label look-ups, track-id wrap, floating-object rule) and the host object layer (dspmap_objects_*: registration, Kabsch on the
key points, moving-or-not filter, coasting, drop) - with the CPU oracle standing in for the CUDA engine below the C ABI - and is
compared with the clouds the reference's own node published for the same replay (tests/golden/node_reference_moving.npz).
Same statistical comparison as the GPU test (the reference jitters moved particles from a random table)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_node_golden as G  # noqa: E402
from oracle_lib import oracle_map  # noqa: E402
from test_adapter import _compare_moving  # noqa: E402

from semantic_dsp_map_b200 import SETTINGS, LabelTables, make_random_table, wire  # noqa: E402


def test_python_host_stack_on_the_oracle_replays_the_moving_scene_like_the_reference_node():
    with tempfile.TemporaryDirectory() as td:
        rec = G.write_package(td, moving=True)
        labels = LabelTables.read_object_info(os.path.join(td, "cfg", "object_info.csv"))
        m = oracle_map(SETTINGS[3], labels, order=0)                                # the reference's own iteration order
        m.setMapParameters(0.8, 0.2, 1, 0.15, 3, 1.0, 5, 0.6, 0.5)                 # cfg/options.yaml of the replay package
        m.setMapOptions(True, False)
        m.setVisualizeOptions(False, True)
        m.setBeyesianMovementParameters(0.2, 0.3, 0.2, 0.05)
        m.setDepthNoiseModelParameters(0.02, 0.3)
        m.enable_colors()                                                          # evaluation format: rgb = (label, track id)
        m.engine.set_random_table(make_random_table(7))
        clouds, tracked = [], []
        with open(rec, "rb") as f:
            for _, r in wire.replay(m, wire.read_records(f)):
                c = np.zeros(len(r.voxel_index), G.REC)
                c["xyz"], c["rgb"] = r.xyz, r.rgb
                clouds.append(c)
                tracked.append(m.objects.is_tracked(3))
    # registered at the first frame, dropped three steps after the detector lost it (frames 7 and 8 coast)
    assert tracked == [True] * 9 + [False] * 2, tracked
    _compare_moving(clouds, np.load(os.path.join(ROOT, "tests", "golden", "node_reference_moving.npz")), G)


def test_static_replay_on_the_oracle_equals_the_reference_node_exactly():
    """No moved particles, hence no random numbers: with the oracle in the reference's iteration order the host stack must
    publish EXACTLY the clouds of the reference node - same counts in every frame, same voxels, same colour on every voxel."""
    with tempfile.TemporaryDirectory() as td:
        rec = G.write_package(td, moving=False)
        labels = LabelTables.read_object_info(os.path.join(td, "cfg", "object_info.csv"))
        m = oracle_map(SETTINGS[3], labels, order=0)
        m.engine.set_undefined_label_carry(True)                                   # a compiled reference's behaviour at its one UB site
        m.setMapParameters(0.8, 0.2, 1, 0.15, 20, 1.0, 5, 0.6, 0.5)
        m.setMapOptions(True, False)
        m.setVisualizeOptions(False, True)
        m.setBeyesianMovementParameters(0.2, 0.3, 0.2, 0.05)
        m.setDepthNoiseModelParameters(0.02, 0.3)
        m.enable_colors()
        m.engine.set_random_table(make_random_table(7))
        with open(rec, "rb") as f:
            results = [r for _, r in wire.replay(m, wire.read_records(f))]
    want = np.load(os.path.join(ROOT, "tests", "golden", "node_reference.npz"))
    assert [len(r.voxel_index) for r in results] == want["counts"].tolist()
    assert want["counts"][-1] > 1000
    last = results[-1]
    got = np.concatenate([G.lattice(last.xyz).astype(np.int32), last.rgb.astype(np.int32)], axis=1)
    ref = np.concatenate([want["last_lattice"].astype(np.int32), want["last_rgb"].astype(np.int32)], axis=1)
    assert np.array_equal(got, ref)                                                 # same order too: ascending storage index


def test_quirks_replay_on_the_oracle_equals_the_reference_node_exactly():
    """A yawing, climbing, side-stepping camera over a static world with the label quirks of the path in the static mask ("sky"
    spelt unlike the filter's "Sky", a movable class, a value outside the table), far and too-close pixels, a detection with two
    key points (registered, never moved) and one without key points (never registered: floating, its particles live one
    frame).  Nothing moves, so the reference draws no random numbers: EXACT equality with the reference node's clouds."""
    with tempfile.TemporaryDirectory() as td:
        rec = G.write_quirk_package(td)
        labels = LabelTables.read_object_info(os.path.join(td, "cfg", "object_info.csv"))
        m = oracle_map(SETTINGS[3], labels, order=0)
        m.engine.set_undefined_label_carry(True)
        # the two places where the reference reads uninitialised memory: sigma of an invalid pixel and of a pixel the
        # bounding-box gate turned into background - a compiled reference reads 0 there when the depth-noise model is on
        m.engine.set_invalid_sigma(0.0)
        m.setMapParameters(0.8, 0.2, 1, 0.15, 20, 1.0, 5, 0.6, 0.5)
        m.setMapOptions(True, False)
        m.setVisualizeOptions(False, True)
        m.setBeyesianMovementParameters(0.2, 0.3, 0.2, 0.05)
        m.setDepthNoiseModelParameters(0.02, 0.3)
        m.enable_colors()
        m.engine.set_random_table(make_random_table(7))
        with open(rec, "rb") as f:
            results = [r for _, r in wire.replay(m, wire.read_records(f))]
        assert m.objects.is_tracked(5) and not m.objects.is_tracked(6)
    want = np.load(os.path.join(ROOT, "tests", "golden", "node_reference_quirks.npz"))
    assert [len(r.voxel_index) for r in results] == want["counts"].tolist()
    assert want["counts"][-1] > 5000
    for t in (1, G.QUIRK_FRAMES - 1):
        got = np.concatenate([G.lattice(results[t].xyz).astype(np.int32), results[t].rgb.astype(np.int32)], axis=1)
        ref = np.concatenate([want[f"lattice_{t}"].astype(np.int32), want[f"rgb_{t}"].astype(np.int32)], axis=1)
        assert np.array_equal(got, ref), t
