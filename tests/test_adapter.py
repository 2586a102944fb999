"""adapter/semantic_dsp_map.h (the drop-in header for src/mapping.cpp) compiles, links and — on a GPU — produces the same
occupied cloud as the Python mirror on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

from semantic_dsp_map_b200 import LabelTables, MapConfig, MaskKpts, SemanticDSPMap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "adapter_driver")


def build_driver():
    lib_dir = os.path.join(ROOT, "semantic_dsp_map_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "oracle", "ref_stubs"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "adapter_driver.cpp"), "-o", EXE, "-L" + lib_dir, "-l:libdspmap_b200.so", "-Wl,-rpath," + lib_dir]
    subprocess.check_call(cmd)


def test_adapter_compiles_and_links():
    build_driver()
    assert os.path.exists(EXE)


def _fnv(xyz, r):
    h = 1469598103934665603
    bits = np.ascontiguousarray(xyz, np.float32).view(np.uint32)
    for row, rr in zip(bits.tolist(), r.tolist()):
        for w in row:
            h = ((h ^ w) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        h = ((h ^ rr) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_adapter_matches_python_mirror():
    build_driver()
    out = subprocess.run([EXE, "4"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    cfg = MapConfig(name="adapter", voxel_bits=(6, 5, 6), slot_bits=3, voxel_size=0.15, fx=66.0, fy=66.0, cx=79.5, cy=47.5,
                    image_width=160, image_height=96, boost_mode=False, depth_min=0.3, depth_max=15.0, consider_instance=True,
                    zed2_filters=False, detection_probability=0.8, noise_number=0.2, nb_ptc_num_per_point=1, occupancy_threshold=0.15,
                    forgetting_rate=1.0, max_forget_count=5, id_transition_probability=0.5, depth_noise_first_order=0.02,
                    depth_noise_zero_order=0.3, consider_depth_noise=True, use_independent_filter=False, zero_center=False)
    labels = LabelTables.from_rows([(0, "background", 65535), (1, "road", 65534), (14, "Vehicle", -1)])
    m = SemanticDSPMap(cfg, labels)
    m.setVisualizeOptions(False, True)          # like the driver: evaluation format
    m.enable_colors()
    i = np.arange(96, dtype=np.float32)[:, None] * np.ones((1, 160), np.float32)
    with np.errstate(divide="ignore"):
        depth = np.where(i > 52, np.float32(1.5) * np.float32(66.0) / (i - np.float32(47.0)), np.float32(3.5)).astype(np.float32)
    car = np.zeros((96, 160), np.uint8)
    car[30:50, 60:90] = 255
    for t, line in enumerate(out):
        r = m.update(depth, [MaskKpts(65535, "static", np.zeros((96, 160), np.uint8)), MaskKpts(3, "Vehicle", car)],
                     [0.0, 0.0, 0.1 * t], [1.0, 0.0, 0.0, 0.0])
        red = r.rgb[:, 0].astype(np.int64)
        f, n, h = line.split()
        assert int(f) == t and int(n) == len(r.voxel_index), (line, len(r.voxel_index))
        assert int(h, 16) == _fnv(r.xyz, red), line
    assert int(out[-1].split()[1]) > 0
