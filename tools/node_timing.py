"""Node-level timing with the reference's OWN instrumentation: src/mapping.cpp measures every update() and reports the running
average through ROS_INFO_THROTTLE ("Average time cost: %f s", :279-280).  The same unmodified node source, built against the
adapter (tests/_built/mapping_node_b200, GPU) and against the reference's header (tests/_built/mapping_node_ref, CPU, built by
__graft_entry__.build() where /root/reference exists), replays the same recorded file (SETTING 3 as shipped: 128 x 32 x 128
voxels, 1280 x 720 frames) and prints that line.

    python tools/node_timing.py [frames]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_node_golden as G  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    out = {"frames": frames, "workload": "SETTING 3 as shipped (128x32x128 voxels x 4 slots), 1280x720 depth + static mask + one instance mask"}
    with tempfile.TemporaryDirectory() as td:
        rec = G.write_package(td, moving=False, frames=frames, cam_step=0.05)
        for name in ("mapping_node_b200", "mapping_node_ref"):
            exe = os.path.join(ROOT, "tests", "_built", name)
            if not os.path.exists(exe):
                out[name] = "not built"
                continue
            t0 = time.time()
            try:
                clouds, log = G.run_node(exe, td, rec=rec, want_log=True)
            except subprocess.CalledProcessError as e:       # e.g. the engine build on a box without a GPU
                out[name] = "failed: " + (e.stderr or "").strip().splitlines()[-1]
                continue
            wall = time.time() - t0
            if os.environ.get("NODE_TIMING_LOG"):            # e.g. with DSPMAP_TRACE=1: the engine's issue / wait times per frame
                with open(os.environ["NODE_TIMING_LOG"] + "." + name, "w") as f:
                    f.write(log)
            m = re.search(r"Average time cost: ([0-9.eE+-]+) s", log)
            out[name] = {"average_update_s": float(m.group(1)) if m else None, "process_wall_s": round(wall, 3), "clouds": len(clouds),
                         "last_cloud_points": int(len(clouds[-1])) if clouds else 0}
    a, b = out.get("mapping_node_b200"), out.get("mapping_node_ref")
    if isinstance(a, dict) and isinstance(b, dict) and a["average_update_s"] and b["average_update_s"]:
        out["update_speedup"] = round(b["average_update_s"] / a["average_update_s"], 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
