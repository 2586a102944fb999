"""tests/golden/dropin_reference.npz: the occupied clouds that tests/dropin_driver.cpp produces when it is compiled against the
REFERENCE'S OWN include/semantic_dsp_map.h (-DUSE_REFERENCE; third-party headers from oracle/ref_stubs) - the whole
SemanticDSPMap::update() of the reference, object layer included, driven like src/mapping.cpp drives it.  Build container only:

    python tests/golden/make_dropin_golden.py
"""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FRAMES = 4
MOVING_FRAMES = 11                 # scene "moving" of tests/dropin_driver.cpp: the whole object layer runs (Kabsch, filter, coasting, drop)
MOVING_KEPT = (2, 6, 8, 9, 10)     # frames whose clouds are kept in dropin_reference_moving.npz
REC = np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3)])


def read_clouds(path):
    b = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(b):
        (n,) = struct.unpack_from("<I", b, pos)
        out.append(np.frombuffer(b, REC, n, pos + 4).copy())
        pos += 4 + n * REC.itemsize
    return out


def lattice(xyz, voxel=0.15):
    """Emitted points are voxel corners: global coordinates are integer multiples of the voxel size."""
    return np.round(np.asarray(xyz, np.float64) / voxel).astype(np.int16)


def main():
    with tempfile.TemporaryDirectory() as td:
        exe, out = os.path.join(td, "dropin_ref"), os.path.join(td, "out.bin")
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-w", "-DUSE_REFERENCE",
                               "-I" + os.path.join(ROOT, "oracle", "ref_stubs"), "-I/root/reference/include", os.path.join(ROOT, "tests", "dropin_driver.cpp"), "-o", exe])
        subprocess.check_call([exe, str(FRAMES), out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        clouds = read_clouds(out)
        subprocess.check_call([exe, str(MOVING_FRAMES), out, "moving"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        moving = read_clouds(out)
    kept = {}
    for t in MOVING_KEPT:
        kept[f"lattice_{t}"], kept[f"rgb_{t}"] = lattice(moving[t]["xyz"]), moving[t]["rgb"]
    np.savez_compressed(os.path.join(HERE, "dropin_reference_moving.npz"), counts=np.asarray([len(c) for c in moving], np.int64), **kept)
    print("moving:", [len(c) for c in moving], os.path.getsize(os.path.join(HERE, "dropin_reference_moving.npz")), "bytes")
    last = clouds[-1]
    np.savez_compressed(os.path.join(HERE, "dropin_reference.npz"), counts=np.asarray([len(c) for c in clouds], np.int64),
                        last_lattice=lattice(last["xyz"]), last_rgb=last["rgb"])
    print([len(c) for c in clouds], os.path.getsize(os.path.join(HERE, "dropin_reference.npz")), "bytes")


if __name__ == "__main__":
    main()
