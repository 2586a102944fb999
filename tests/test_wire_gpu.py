"""Recorded wire traffic replayed through the engine gives what the direct calls give (semantic_dsp_map_b200/wire.py)."""
import io

import numpy as np
import pytest

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, make_random_table, tiny, wire

pytestmark = pytest.mark.gpu


def test_replay_of_recorded_messages_equals_direct_updates():
    cfg = tiny(BASELINE["B"])
    scene = Scene(cfg)
    f = io.BytesIO()
    frames = [scene.frame(t) for t in range(6)]
    for t, fr in enumerate(frames):
        h = wire.Header(seq=t, secs=100 + t, nsecs=0, frame_id="zed2")          # seq 0 and 1 are dropped by the node
        objs = [wire.WireObject(m.label, m.track_id, wire.image_from_array(m.mask, "mono8", h), (0, 0, 0), (0, 0, 0),
                                np.asarray(m.kpts_current, np.float32).reshape(-1, 3), np.asarray(m.kpts_current, np.float32).reshape(-1, 3)) for m in fr.masks]
        wire.write_record(f, wire.encode_image(wire.image_from_array(fr.depth, "32FC1", h)), wire.encode_pose_stamped(h, fr.position, fr.orientation),
                          wire.encode_mask_group(wire.MaskGroup(h, objs)))
    f.seek(0)
    a, b = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    table = make_random_table(cfg.seed, 50_000)
    a.engine.set_random_table(table)
    b.engine.set_random_table(table)
    moves = {t: fr.moves for t, fr in enumerate(frames)}
    got = list(wire.replay(a, wire.read_records(f), moves_for=lambda fr: (moves[fr.seq], ())))
    assert [seq for seq, _ in got] == [2, 3, 4, 5]
    for seq, r in got:
        fr = scene.frame(seq)
        # key points travel as float32 on the wire (mask_kpts_msgs/Keypoint): give the direct call the same values
        for m in fr.masks:
            m.kpts_current = np.asarray(m.kpts_current, np.float32).astype(np.float64)
        d = b.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves)
        for name in ("voxel_index", "track_id", "label_id", "state"):
            np.testing.assert_array_equal(getattr(r, name), getattr(d, name), err_msg=f"frame {seq} {name}")
        np.testing.assert_array_equal(r.weight_sum.view(np.uint32), d.weight_sum.view(np.uint32))
    assert len(got[-1][1].voxel_index) > 0
