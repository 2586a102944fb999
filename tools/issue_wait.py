"""Host time to ISSUE a frame's launches vs time spent WAITING for the device (DSPMAP_TRACE=1), device-resident and end to end."""
import os, sys
os.environ["DSPMAP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from semantic_dsp_map_b200 import BASELINE, Engine, Scene, SemanticDSPMap, make_random_table
cfg = BASELINE["B"]; sc = Scene(cfg); frames = [sc.frame(t) for t in range(64)]
m = SemanticDSPMap(cfg); m.engine.set_random_table(make_random_table(cfg.seed))
dev = torch.device("cuda", 0)
print("== device-resident", file=sys.stderr)
for fr in frames:
    lowered = m.lower_masks(fr.masks)
    d = torch.from_numpy(fr.depth).to(dev); dm = [torch.from_numpy(np.ascontiguousarray(k.mask)).to(dev) for k in fr.masks]
    for lm, t in zip(lowered, dm): lm["data"], lm["row_stride"] = t.data_ptr(), t.stride(0)
    m.engine.update_raw("update_device", d.data_ptr(), fr.depth.shape[0], fr.depth.shape[1], d.stride(0), Engine.pack_masks(lowered), fr.position, fr.orientation, Engine.pack_moves(fr.moves), [], False)
m.clear()
print("== end to end", file=sys.stderr)
for fr in frames:
    m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, copy_result=False)
