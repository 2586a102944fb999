"""Generate tests/golden/*.npz from the CPU oracle (canonical order) on the seeded synthetic scene.

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference ships no tests/fixtures for this path and cannot be built here, so
these vectors come from oracle/dsp_oracle.cpp, our line-by-line restatement of the reference.  They pin the oracle
against drift and give the GPU tests a /root/reference-free ground truth on the GPU box.
Per case: the full result of selected frames (voxel index, position, weight sum, track, label, state), the per-frame
counters, and a SHA-256 of every frame's complete particle store / labelled point image / ck image.
"""
import dataclasses
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from semantic_dsp_map_b200 import BASELINE, Scene, _capi, make_random_table, tiny  # noqa: E402
from oracle_lib import oracle_map  # noqa: E402
from parity import COUNT_STATS  # noqa: E402

CASES = {
    "A_static": dict(base="A", n_frames=10, keep=(0, 4, 9), overrides={}),
    "B_zed2": dict(base="B", n_frames=10, keep=(0, 4, 9), overrides={}),
    "C_vkitti2": dict(base="C", n_frames=8, keep=(0, 3, 7), overrides={}),
    "B_fast_ring_wrap": dict(base="B", n_frames=10, keep=(4, 9), overrides=dict(speed=2.0)),
    "B_no_depth_noise": dict(base="B", n_frames=6, keep=(5,), overrides=dict(consider_depth_noise=False, detection_probability=1.0,
                                                                             noise_number=0.001, occupancy_threshold=0.1, nb_ptc_num_per_point=3)),
}


def case_config(spec):
    return dataclasses.replace(tiny(BASELINE[spec["base"]]), **spec["overrides"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(spec, make_map):
    cfg = case_config(spec)
    m = make_map(cfg)
    m.engine.set_random_table(make_random_table(cfg.seed, 50_000))
    scene = Scene(cfg)
    out = {}
    stats, digests = [], []
    for t in range(spec["n_frames"]):
        fr = scene.frame(t)
        r = m.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=True, moves=fr.moves, remove_tracks=fr.remove_tracks)
        st = m.engine.stats()
        stats.append([st[k] for k in COUNT_STATS])
        digests.append([sha(m.engine.dump(_capi.DUMP_PARTICLES)), sha(m.engine.dump(_capi.DUMP_POINTS)), sha(m.engine.dump(_capi.DUMP_CK)),
                        sha(r.free_voxel_index)])
        if t in spec["keep"]:
            for name in ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state"):
                out[f"f{t}_{name}"] = getattr(r, name)
    out["stats"] = np.asarray(stats, np.int64)
    out["digests"] = np.asarray(digests)
    return out


def main():
    for name, spec in CASES.items():
        out = run_case(spec, oracle_map)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes", out["stats"][-1].tolist())


if __name__ == "__main__":
    main()
