"""The CUDA engine against the REFERENCE'S OWN CODE, directly and bit for bit.

tests/golden/ref_*static*.npz are outputs of oracle/_ref/libdsp_ref_s<SETTING>.so - the reference's hot path compiled unmodified
from /root/reference/include (oracle/build_ref.sh) - on seeded scenes in which no object is moved (tests/golden/make_ref_golden.py).
The engine, with the pixel lists in the reference's iteration order (dspmap_params.reference_iteration_order: its breadth-first
search replayed on the device) and its uninitialised-sigma sites set to what that build reads (dspmap_params.uninit_sigma_mode),
must reproduce every frame of them: SHA-256 of the ring state, of the labelled points, of the pixel offsets, of the per-pixel
particle lists IN ORDER, of the complete particle store as 24-byte reference records, of the occupied and free voxel indices and
positions and of the RGB bytes of every emitted point.  No oracle in between, no tolerance."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_ref_golden as G  # noqa: E402

pytestmark = pytest.mark.gpu

STATIC = sorted(n for n, s in G.SCENARIOS.items() if s.get("static"))
DIGESTS = ("ring", "points", "pixel_offsets", "pixel_lists", "particles", "occ_index", "occ_xyz", "occ_rgb", "free_index", "free_xyz")


@pytest.mark.parametrize("name", STATIC)
def test_engine_reproduces_the_reference_build_bit_for_bit(name):
    spec = G.SCENARIOS[name]
    want = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_" + name + ".npz"))
    got = G.run_oracle(name, None, engine=True)
    skip = {"occ_index", "free_index"} if G.scenario_config(spec).zero_center else set()      # not recoverable from camera-relative points
    for t in range(spec["n_frames"]):
        np.testing.assert_array_equal(got[f"f{t}_counts"], want[f"f{t}_counts"], err_msg=f"{name}: frame {t} counts (occupied, free, listed)")
        for k, what in enumerate(DIGESTS):
            if what not in skip:
                assert got[f"f{t}_digests"][k] == want[f"f{t}_digests"][k], f"{name}: frame {t}: {what} differs from the reference's own output"
    last = spec["n_frames"] - 1
    np.testing.assert_array_equal(got[f"f{last}_occ_rgb"], want[f"f{last}_occ_rgb"])
    assert int(want[f"f{last}_counts"][0]) > 5000 and int(want[f"f{last}_counts"][2]) > 1000      # not vacuous: occupied voxels, listed particles
