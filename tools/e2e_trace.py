"""Where an end-to-end frame spends its host time: Python preparation, the C call that stages + enqueues, the wait for the
result, building the views.  usage: python tools/e2e_trace.py [frames] [sync|async]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semantic_dsp_map_b200 import BASELINE, SemanticDSPMap, make_random_table  # noqa: E402
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    mode = sys.argv[2] if len(sys.argv) > 2 else "async"
    cfg = BASELINE["B"]
    frames = bench.generate_frames("B", n)
    m = SemanticDSPMap(cfg)
    m.enable_colors()
    m.engine.set_random_table(make_random_table(cfg.seed))
    import gc
    gc.disable()
    T = np.zeros((n, 4))
    prev = 0
    t_all = time.perf_counter()
    for k, fr in enumerate(frames):
        t0 = time.perf_counter()
        if mode == "async":
            depth, masks, moves, rm, _ = m._prepare_frame(fr.depth, fr.masks, fr.position, fr.orientation, 0.0, fr.moves, fr.remove_tracks)
            t1 = time.perf_counter()
            tk = m.engine.update_raw("update_async", depth.ctypes.data, depth.shape[0], depth.shape[1], depth.strides[0] // 4, masks, fr.position, fr.orientation, moves, rm, False)
            t2 = time.perf_counter()
            if prev:
                m.wait_result(prev, copy_result=False)
            t3 = time.perf_counter()
            prev = tk
        else:
            depth, masks, moves, rm, _ = m._prepare_frame(fr.depth, fr.masks, fr.position, fr.orientation, 0.0, fr.moves, fr.remove_tracks)
            t1 = time.perf_counter()
            m.engine.update_raw("update", depth.ctypes.data, depth.shape[0], depth.shape[1], depth.strides[0] // 4, masks, fr.position, fr.orientation, moves, rm, False)
            t2 = time.perf_counter()
            m.engine.result_arrays(False)
            t3 = time.perf_counter()
        T[k] = (t1 - t0, t2 - t1, t3 - t2, t3 - t0)
    total = time.perf_counter() - t_all
    w = T[n // 2:] * 1e3
    print(f"{mode}: {n / total:.0f} frames/s overall; steady half, ms median: python prepare {np.median(w[:,0]):.3f}  C update {np.median(w[:,1]):.3f}  "
          f"wait/result {np.median(w[:,2]):.3f}  iteration {np.median(w[:,3]):.3f}")


if __name__ == "__main__":
    main()
