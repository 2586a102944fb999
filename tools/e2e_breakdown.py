"""Where the end-to-end (host buffers in, host result out) frame time goes.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, make_random_table
cfg = BASELINE["B"]
sc = Scene(cfg)
frames = [sc.frame(t) for t in range(40)]
m = SemanticDSPMap(cfg)
m.engine.set_random_table(make_random_table(cfg.seed))
acc = np.zeros(4)
for k, fr in enumerate(frames):
    t0 = time.perf_counter()
    depth = fr.depth
    m._keep = [depth]
    masks = m.lower_masks(fr.masks)
    t1 = time.perf_counter()
    m.engine.update_raw("update", depth.ctypes.data, depth.shape[0], depth.shape[1], depth.strides[0] // 4, masks, fr.position, fr.orientation, fr.moves, fr.remove_tracks, False)
    t2 = time.perf_counter()
    r = m.engine.result_arrays()
    t3 = time.perf_counter()
    if k >= 20:
        acc += [t1 - t0, t2 - t1, t3 - t2, t3 - t0]
print("ms per frame: lower_masks %.3f  dspmap_update %.3f  result copy %.3f  total %.3f" % tuple(1e3 * acc / 20))
print(m.engine.stats())
