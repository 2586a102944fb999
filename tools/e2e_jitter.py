"""Per-frame wall time of the end-to-end call (host buffers in, host result views out): percentiles over a run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, make_random_table
cfg = BASELINE["B"]; sc = Scene(cfg); frames = [sc.frame(t) for t in range(130)]
m = SemanticDSPMap(cfg); m.engine.set_random_table(make_random_table(cfg.seed))
ts = []
for fr in frames:
    t0 = time.perf_counter()
    m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks, copy_result=False)
    ts.append(time.perf_counter() - t0)
a = 1e3 * np.asarray(ts[30:])
print("ms/frame: mean %.3f median %.3f p10 %.3f p90 %.3f p99 %.3f max %.3f  -> %.0f frames/s" % (a.mean(), np.median(a), np.percentile(a, 10), np.percentile(a, 90), np.percentile(a, 99), a.max(), 1e3 / a.mean()))
