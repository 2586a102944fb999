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


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", os.path.join(SRC, "tracker_harness.cpp"), "-o", EXE])
    return EXE


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


def test_scenarios_cover_the_branches():
    text = "\n".join(sum(GOLDEN.values(), []))
    assert "absent" in text and "has_flag=0" in text and "moving=0 p=" in text and "moving=1" in text
    # a recovered rotation: some matrix has an off-diagonal rotation term well away from zero
    assert any(abs(float(ln.split()[9])) > 0.05 for ln in text.splitlines() if len(ln.split()) > 9)
