"""How far apart are the canonical order (ascending particle index inside a pixel's list: the engine's default) and the reference's
iteration order (its breadth-first search, dspmap_params.reference_iteration_order) on a BASELINE configuration?  Two engines on the
same seeded frames (object moves included: both take a moving object's particles in ascending index order), per frame: occupied
voxels present in only one of the two results, voxels whose fused (label, track) differs, largest relative difference of the weight
sums of common voxels.  usage: python tools/order_flips.py [config] [frames]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, make_random_table  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "B"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    cfg = BASELINE[name]
    a, b = SemanticDSPMap(cfg), SemanticDSPMap(cfg)
    b.setReferenceIterationOrder(True)
    table = make_random_table(cfg.seed)
    for m in (a, b):
        m.enable_colors()
        m.engine.set_random_table(table)
    scene = Scene(cfg)
    rows = []
    for t in range(n):
        res = []
        for m in (a, b):
            fr = scene.frame(t)
            res.append(m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks))
        ra, rb = res
        ia, ib = ra.voxel_index, rb.voxel_index
        only = len(np.setxor1d(ia, ib))
        common, pa, pb = np.intersect1d(ia, ib, return_indices=True)
        lab = int(((ra.label_id[pa] != rb.label_id[pb]) | (ra.track_id[pa] != rb.track_id[pb])).sum())
        wa, wb = ra.weight_sum[pa], rb.weight_sum[pb]
        rel = float(np.max(np.abs(wa - wb) / np.maximum(np.abs(wb), 1e-30))) if len(common) else 0.0
        rows.append(dict(frame=t, occupied=int(len(ib)), only_in_one=int(only), label_or_track_differs=lab, max_rel_weight_diff=rel))
    tot = sum(r["occupied"] for r in rows)
    out = dict(config=cfg.name, frames=n, occupied_voxel_frames=tot, only_in_one=sum(r["only_in_one"] for r in rows),
               label_or_track_differs=sum(r["label_or_track_differs"] for r in rows), max_rel_weight_diff=max(r["max_rel_weight_diff"] for r in rows),
               worst_frame=max(rows, key=lambda r: r["only_in_one"]), last_frame=rows[-1])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
