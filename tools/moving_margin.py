"""How far the adapter's run of the "moving" drop-in scene sits from the reference fixture, per kept frame (the quantities
tests/test_adapter.py::_compare_moving bounds).  GPU box:  python tools/moving_margin.py [seed]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_dropin_golden as G  # noqa: E402


def main():
    seed = sys.argv[1] if len(sys.argv) > 1 else "20240911"
    lib_dir = os.path.join(ROOT, "semantic_dsp_map_b200")
    exe, out = "/tmp/dropin_adapter_moving", "/tmp/dropin_adapter_moving.bin"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I" + os.path.join(ROOT, "oracle", "ref_stubs"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "dropin_driver.cpp"), "-o", exe, "-L" + lib_dir, "-l:libdspmap_b200.so", "-Wl,-rpath," + lib_dir])
    subprocess.run([exe, str(G.MOVING_FRAMES), out, "moving"], check=True, capture_output=True, env=dict(os.environ, DSPMAP_ADAPTER_SEED=seed))
    clouds = G.read_clouds(out)
    want = np.load(os.path.join(ROOT, "tests", "golden", "dropin_reference_moving.npz"))
    print("counts", [len(c) for c in clouds], "reference", want["counts"].tolist())
    for t in G.MOVING_KEPT:
        lat, rgb = G.lattice(clouds[t]["xyz"]), clouds[t]["rgb"]
        rl, rr = want[f"lattice_{t}"], want[f"rgb_{t}"]
        v, rv = lat[rgb[:, 0] == 14], rl[rr[:, 0] == 14]
        a, b = {tuple(x) for x in lat[rgb[:, 0] != 14].tolist()}, {tuple(x) for x in rl[rr[:, 0] != 14].tolist()}
        line = f"frame {t}: vehicle voxels {len(v)} vs {len(rv)}; static differing {len(a ^ b)} of {len(b)} ({100.0 * len(a ^ b) / len(b):.2f} %)"
        if len(v) and len(rv):
            line += f"; centroid diff {np.abs(v.mean(0) - rv.mean(0)).round(2).tolist()}; min diff {np.abs(v.min(0) - rv.min(0)).tolist()}; max diff {np.abs(v.max(0) - rv.max(0)).tolist()}"
        print(line)


if __name__ == "__main__":
    main()
