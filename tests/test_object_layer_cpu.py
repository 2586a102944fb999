"""The adapter's host object layer (adapter/dsp_object_tracker.h: what objectLevelUpdate, ObjectSet and Kabsch / RANSAC of the
reference decide per frame - which objects are tracked, which are moving, by which rigid matrix, which are dropped) against
the reference's OWN object layer on the scenarios of tests/object_layer/scenarios.h (tests/golden/object_layer_reference.txt,
made by tests/golden/make_object_layer_golden.py from the unmodified headers).  Discrete decisions and the moved
probability must be identical; the rigid matrices agree to 1e-9 (two different SVD routines; the reference's RANSAC sample
is random, the motions are rigid)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "object_layer")
EXE = os.path.join(ROOT, "tests", "_built", "tracker_harness")
LIB_DIR = os.path.join(ROOT, "semantic_dsp_map_b200")
T_TOL = 1e-9


def golden():
    out, name = {}, None
    for ln in open(os.path.join(ROOT, "tests", "golden", "object_layer_reference.txt")).read().splitlines():
        if ln.startswith("SCENARIO "):
            name = ln.split()[1]
            out[name] = []
        elif ln:
            out[name].append(ln)
    return out


@pytest.fixture(scope="module", params=["c_abi", "direct_header"])
def harness(request):
    """c_abi: the dspmap_objects_* entry points of libdspmap_b200.so (host only, no GPU needed); direct_header: the C++ header
    the way adapter/semantic_dsp_map.h uses it."""
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    exe = EXE + "_" + request.param
    extra = ["-DDIRECT_HEADER"] if request.param == "direct_header" else ["-L" + LIB_DIR, "-l:libdspmap_b200.so", "-Wl,-rpath," + LIB_DIR]
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", os.path.join(SRC, "tracker_harness.cpp"), "-o", exe] + extra)
    return exe


GOLDEN = golden()


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_object_layer_follows_the_reference(harness, name):
    r = subprocess.run([harness, name], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
    got = [ln for ln in r.stderr.splitlines() if ln.startswith("OBJ ")]
    want = GOLDEN[name]
    assert len(got) == len(want) > 0
    moving = dropped = 0
    for g, w in zip(got, want):
        gs, ws = g.split(), w.split()
        assert gs[:7] == ws[:7], (g, w)            # frame, track, tracked/absent, has_flag, moving, moved probability (printed with 17 digits)
        assert len(gs) == len(ws), (g, w)
        for a, b in zip(gs[7:], ws[7:]):
            assert abs(float(a) - float(b)) <= T_TOL, (g, w)
        moving += "moving=1" in w
        dropped += "absent" in w
    assert moving > 0                              # every scenario has a moving object: the comparison is not vacuous
    if harness.endswith("c_abi"):                  # the moves handed to dspmap_update are the float casts of the moving objects' matrices
        moves = [ln.split() for ln in r.stderr.splitlines() if ln.startswith("MOVE ")]
        drops = [ln.split() for ln in r.stderr.splitlines() if ln.startswith("DROP ")]
        state = {(w.split()[1], w.split()[2]): w.split() for w in want}
        assert len(moves) == sum("moving=1" in w for w in want)
        for _, t, track, tx, ty, tz in moves:
            ws = state[(t, track)]
            assert ws[5] == "moving=1"
            assert [float(tx), float(ty), float(tz)] == pytest.approx([float(ws[10]), float(ws[14]), float(ws[18])], abs=1e-6)
        for _, t, track in drops:                  # a dropped object is absent from that frame on
            assert state[(t, track)][3] == "absent" and state[(str(int(t) - 1), track)][5] == "moving=1"


def test_scenarios_cover_the_branches():
    text = "\n".join(sum(GOLDEN.values(), []))
    assert "absent" in text and "has_flag=0" in text and "moving=0 p=" in text and "moving=1" in text
    # a recovered rotation: some matrix has an off-diagonal rotation term well away from zero
    assert any(abs(float(ln.split()[9])) > 0.05 for ln in text.splitlines() if len(ln.split()) > 9)


def test_python_wrapper_of_the_object_layer():
    """semantic_dsp_map_b200/objects.py over dspmap_objects_* (no GPU): a box translating 0.3 m a frame is registered, found
    moving with the right matrix, coasts when lost and is dropped after max_lost steps; a far object is never registered."""
    import numpy as np
    from semantic_dsp_map_b200 import SETTINGS
    from semantic_dsp_map_b200.objects import BOX_KEYPOINTS, ObjectLayer
    ol = ObjectLayer(SETTINGS[3])
    assert ol.mode == BOX_KEYPOINTS
    ol.set_movement_filter(0.2, 0.3, 0.2, 0.05)
    base = np.array([[-0.9, 1.0, 5.9], [0.9, 1.0, 5.9], [-0.9, 1.0, 10.1], [-0.9, -0.5, 5.9]])
    far = base + [0.0, 0.0, 100.0]
    pos, quat = [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]
    for t in range(5):
        k = base + [0.3 * t, 0.0, 0.0]
        moves, dropped = ol.step([(4, True, 14, k, None), (9, True, 14, far, None)], pos, quat, 0.1 * t, 3)
        assert dropped == [] and not ol.is_tracked(9)
        if t == 0:
            assert moves == [] and ol.query(4) == dict(tracked=True, has_flag=False, moving=False, moved_probability=0.5, matrix=None)
        else:
            (track, T), = moves
            assert track == 4 and T.dtype == np.float32
            assert np.allclose(T, [[1, 0, 0, 0.3], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], atol=1e-6)
    q = ol.query(4)
    assert q["moving"] and np.allclose(q["matrix"][:3, 3], [0.3, 0, 0], atol=1e-12)
    seen_drop = False
    for t in range(5, 9):                                   # lost: coasts on the estimated velocity, then is dropped
        moves, dropped = ol.step([], pos, quat, 0.1 * t, 3)
        if dropped:
            assert dropped == [4] and moves == [] and not ol.is_tracked(4)
            seen_drop = True
            break
        (track, T), = moves
        assert track == 4 and np.allclose(T[:3, :3], np.eye(3)) and T[0, 3] > 0.25   # pure translation (object_layer.h:190)
    assert seen_drop
    ol.clear()
    assert not ol.is_tracked(4)
