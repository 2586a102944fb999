"""Golden vectors from the REFERENCE ITSELF: oracle/_ref/libdsp_ref_s<SETTING>.so is the reference's own hot-path code
(include/semantic_dsp_map.h and everything below it), compiled unmodified from /root/reference/include by
oracle/build_ref.sh against the stand-in Eigen / OpenCV / PCL / Boost headers of oracle/ref_stubs (what that does and does
not anchor is stated in oracle/ref_harness.cpp and oracle/ref_stubs/Eigen/Dense).

    python tests/golden/make_ref_golden.py            # needs /root/reference (build container only)

Each scenario runs in a fresh process (the reference keeps its map in globals) on the seeded synthetic scene and stores,
per frame: SHA-256 of the complete particle store (24-byte reference records), of the labelled-point image, of the
per-pixel particle lists (BFS discovery order) and of the ring state; the occupied / free voxel indices (recovered from
the emitted voxel-corner positions); the emitted positions and RGB bytes; the order in which the reference visited
moving objects.  tests/test_reference_pin_cpu.py replays the same scenario through the CPU oracle (reference iteration
order) and demands identical bits — on the GPU box, where /root/reference does not exist, against these files.
"""
import dataclasses
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from semantic_dsp_map_b200 import BASELINE, SETTINGS, Scene, _capi, make_random_table  # noqa: E402
from semantic_dsp_map_b200.labels import ZED2_ROWS  # noqa: E402
from parity import PARTICLE, POINT  # noqa: E402

# id of the static-mask pixel values used by the quirk patch (cfg/object_info_zed2.csv: value v -> label v+1)
V_SKY, V_PERSON = 10, 11

SCENARIOS = {
    # the configuration the reference ships (settings.h:22, cfg/options_zed2.yaml), two boxes moving, free space on
    "s3_zed2_default": dict(setting=3, n_frames=6, overrides=dict(zero_center=False), eval_format=True),
    # output positions relative to the camera (the shipped visualisation option) and the display colour path
    "s3_zero_center_colours": dict(setting=3, n_frames=4, overrides=dict(zero_center=True), eval_format=False),
    # 2 m per frame: sub-stepped ring shift + wrap-around (operations.h:71-90); a track dropped by the object layer;
    # static-mask classes whose label keys miss in the string tables (sky, person: SURVEY 7.3-e)
    "s3_fast_remove_quirks": dict(setting=3, n_frames=6, overrides=dict(zero_center=False, speed=2.0), eval_format=True,
                                  remove_at={4: [1]}, quirk_patch=True),
    # src/mapping.cpp:109-118 fallback when if_consider_depth_noise is false
    "s3_no_depth_noise": dict(setting=3, n_frames=4, eval_format=True,
                              overrides=dict(zero_center=False, consider_depth_noise=False, detection_probability=1.0, noise_number=0.001,
                                             occupancy_threshold=0.1, nb_ptc_num_per_point=3)),
    # several noisy newborn particles per point (semantic_dsp_map.h:1177-1230) and independent per-instance filters
    "s3_noisy_births_independent": dict(setting=3, n_frames=4, eval_format=True,
                                        overrides=dict(zero_center=False, nb_ptc_num_per_point=3, use_independent_filter=True)),
    # SETTING 2 (Virtual KITTI 2 preset, settings.h:79-98): no BOOST_MODE, 11x11 update window, 8 slots, no ZED2 filters
    "s2_vkitti2": dict(setting=2, n_frames=3, overrides=dict(zero_center=False), eval_format=True),
    # the camera backs away diagonally (-x, +y, -z: every axis scrolls the other way and wraps) while rolling and pitching:
    # negative ring steps (operations.h:1131-1181, 1196-1230) and a general quaternion; the depth images no longer match the
    # pose, which parity does not need
    "s3_backward_diagonal_rotating": dict(setting=3, n_frames=6, overrides=dict(zero_center=True), eval_format=True, trajectory="backward"),
    # SETTING 0 (KITTI-360 preset, settings.h:32-52): static mode (g_consider_instance false), 256^3 voxels x 8 slots (a 3.2 GB
    # particle array), independent filters, display colours of the "#if SETTING == 0" block (label colour for every class)
    "s0_kitti360_static": dict(setting=0, n_frames=2, overrides=dict(zero_center=False), eval_format=False),
    # BASELINE.json's benchmark configuration (the ZED2 preset with 256 x 64 x 256 voxels x 8 slots) -
    # the build bench.py times as the CPU baseline
    "sB_benchmark_config": dict(setting="B", n_frames=2, overrides=dict(), eval_format=False),
    # ---- scenes WITHOUT object moves: nothing in them depends on std::unordered_set iteration, so the CUDA engine in the reference's
    # iteration order (dspmap_params.reference_iteration_order) is held to these files DIRECTLY (tests/test_reference_build_gpu.py) ----
    # the shipped configuration, instance masks present but no object ever moved; a track dropped at frame 4
    "s3_static_scene": dict(setting=3, n_frames=6, overrides=dict(zero_center=False), eval_format=True, static=True, remove_at={4: [1]}),
    # display colours, camera-relative output, the label quirks in the static mask, 2 m per frame (sub-stepped ring shift + wrap)
    "s3_static_fast_quirks": dict(setting=3, n_frames=5, overrides=dict(zero_center=True, speed=2.0), eval_format=False, static=True, quirk_patch=True),
    # SETTING 2: no BOOST_MODE, 11 x 11 window, 8 slots
    "s2_static_vkitti2": dict(setting=2, n_frames=3, overrides=dict(zero_center=False), eval_format=True, static=True),
    # the benchmark configuration (256 x 64 x 256 voxels x 8 slots) over 4 frames
    "sB_static_benchmark": dict(setting="B", n_frames=4, overrides=dict(), eval_format=False, static=True),
}


def scenario_config(spec):
    base = BASELINE["B"] if spec["setting"] == "B" else SETTINGS[spec["setting"]]
    return dataclasses.replace(base, **spec["overrides"])


def scenario_frame(scene, spec, t):
    fr = scene.frame(t)
    if spec.get("quirk_patch"):
        cls = fr.masks[0].mask
        h, w = cls.shape
        cls[: h // 6, : w // 3] = V_SKY
        cls[h // 2:, w // 2: w // 2 + w // 10] = V_PERSON
    if spec.get("trajectory") == "backward":
        fr.position = np.array([-0.9 * t, 0.23 * t, -1.7 * t + 0.4], np.float64)
        q = np.array([np.cos(0.07 * t), 0.3 * np.sin(0.07 * t), np.sin(0.11 * t) * 0.2 + 0.05, -0.25 * np.sin(0.05 * t)], np.float64)
        fr.orientation = q / np.linalg.norm(q)
    fr.remove_tracks = list(spec.get("remove_at", {}).get(t, []))
    if fr.remove_tracks:
        fr.moves = [(k, T) for k, T in fr.moves if k not in fr.remove_tracks]
    if spec.get("static"):
        fr.moves = []
    return fr


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def masked_points(raw):
    """The reference leaves invalid pixels uninitialised (pointcloud_tools.h:224-232): compare validity everywhere and
    the payload only where valid."""
    p = raw.view(POINT).copy()
    inv = p["valid"] == 0
    for f in ("x", "y", "z", "sigma", "track", "label"):
        p[f][inv] = 0
    return p


def frame_record(t, ring, points, offsets, lists, particles, occ_idx, occ_xyz, occ_rgb, free_idx, free_xyz, move_order, keep_arrays):
    """Digest order: ring, points, pixel offsets, pixel lists, particles, occupied index / xyz / rgb, free index / xyz."""
    occ_idx, free_idx = np.asarray(occ_idx, np.uint32), np.asarray(free_idx, np.uint32)
    occ_xyz, free_xyz, occ_rgb = np.asarray(occ_xyz, np.float32), np.asarray(free_xyz, np.float32), np.asarray(occ_rgb, np.uint8)
    rec = {
        f"f{t}_digests": np.asarray([sha(ring), sha(masked_points(points)), sha(offsets), sha(lists), sha(particles), sha(occ_idx), sha(occ_xyz),
                                     sha(occ_rgb), sha(free_idx), sha(free_xyz)]),
        f"f{t}_counts": np.asarray([len(occ_xyz), len(free_xyz), len(lists)], np.int64),
        f"f{t}_move_order": np.asarray(move_order, np.int32),
    }
    if keep_arrays:   # full arrays only for the last frame (positions and the free list stay digests): the fixtures stay small
        rec.update({f"f{t}_occ_index": occ_idx, f"f{t}_occ_rgb": occ_rgb})
    return rec


def run_reference(name):
    """Runs in a fresh process: the real reference code."""
    from ref_lib import RefMap, voxel_index_of, write_label_csv
    import tempfile
    spec = SCENARIOS[name]
    cfg = scenario_config(spec)
    with tempfile.TemporaryDirectory() as td:
        csv = os.path.join(td, "object_info.csv")
        write_label_csv(csv, ZED2_ROWS)
        ref = RefMap(cfg, csv, setting=spec["setting"], eval_format=spec["eval_format"])
    ref.set_random_table(make_random_table(cfg.seed))
    ref.set_color_map(colour_permutation())
    scene = Scene(cfg)
    out, seconds = {}, []
    for t in range(spec["n_frames"]):
        fr = scenario_frame(scene, spec, t)
        clouds = ref.update(fr.depth, fr.masks, fr.position, fr.orientation, True, fr.moves, fr.remove_tracks)
        seconds.append(ref.seconds())
        vox = ref.voxel_positions()
        if cfg.zero_center:   # points are emitted relative to the camera: indices are not recoverable from them, and need not be
            occ_idx, free_idx = np.zeros(0, np.uint32), np.zeros(0, np.uint32)
        else:
            occ_idx, free_idx = voxel_index_of(clouds["occ_xyz"], vox), voxel_index_of(clouds["free_xyz"], vox)
        out.update(frame_record(t, ref.dump(_capi.DUMP_RING, np.uint32), ref.dump(_capi.DUMP_POINTS), ref.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32),
                                ref.dump(_capi.DUMP_PIXEL_LISTS, np.uint32), ref.dump(_capi.DUMP_PARTICLES), occ_idx, clouds["occ_xyz"],
                                clouds["occ_rgb"], free_idx, clouds["free_xyz"], ref.move_order(), t == spec["n_frames"] - 1))
    out["seconds"] = np.asarray(seconds)
    return out


def csv_colours():
    """{label_id: (b, g, r)}: the colour columns ref_lib.write_label_csv writes by default."""
    return {label_id: ((37 * label_id) % 256, (91 * label_id + 13) % 256, (53 * label_id + 101) % 256) for label_id, _, _ in ZED2_ROWS}


def colour_permutation():
    """Stands in for the std::random_device shuffle of color_map_int_256_ (semantic_dsp_map.h:45-48)."""
    return np.random.default_rng(7).permutation(256).astype(np.int32)


def run_oracle(name, move_orders=None, order=0, engine=False):
    """The same scenario through the CPU oracle.  move_orders[t] = the order in which the reference visited the moving
    objects of frame t (std::unordered_map iteration, a host-container artefact the oracle takes as given).
    engine=True: through the CUDA engine instead, in the reference's iteration order (static scenarios)."""
    spec = SCENARIOS[name]
    cfg = scenario_config(spec)
    # The reference copies an uninitialised LabeledPoint for invalid pixels (pointcloud_tools.h:224-232) and later reads its
    # sigma (semantic_dsp_map.h:1047).  In oracle/_ref's build that stack slot reads 0 when the depth-noise model is on and
    # 0.1 when it is off; the oracle takes the value as a parameter so that everything else is compared bit for bit (the engine
    # has the same switch in dspmap_params).
    if engine:
        from semantic_dsp_map_b200 import SemanticDSPMap
        orc = SemanticDSPMap(cfg)
        orc.setReferenceIterationOrder(True)
        orc.followCompiledReferenceSigma(True, 0.0 if cfg.consider_depth_noise else 0.1)
    else:
        from oracle_lib import oracle_map
        orc = oracle_map(cfg, order=order)
        orc.engine.set_invalid_sigma(0.0 if cfg.consider_depth_noise else 0.1)
    orc.engine.set_random_table(make_random_table(cfg.seed))
    orc.setVisualizeOptions(cfg.zero_center, spec["eval_format"])
    orc.enable_colors(label_bgr=csv_colours(), instance_map=colour_permutation())
    scene = Scene(cfg)
    out, seconds = {}, []
    import time
    for t in range(spec["n_frames"]):
        fr = scenario_frame(scene, spec, t)
        moves = fr.moves
        if move_orders is not None:
            by_track = dict(fr.moves)
            moves = [(int(k), by_track[int(k)]) for k in move_orders[t]]
        t0 = time.perf_counter()
        r = orc.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=True, moves=moves, remove_tracks=fr.remove_tracks)
        seconds.append(time.perf_counter() - t0)
        e = orc.engine
        zc = cfg.zero_center
        out.update(frame_record(t, e.dump(_capi.DUMP_RING, np.uint32), e.dump(_capi.DUMP_POINTS), e.dump(_capi.DUMP_PIXEL_OFFSETS, np.uint32),
                                e.dump(_capi.DUMP_PIXEL_LISTS, np.uint32), e.dump(_capi.DUMP_PARTICLES),
                                np.zeros(0, np.uint32) if zc else r.voxel_index, r.xyz, r.rgb,
                                np.zeros(0, np.uint32) if zc else r.free_voxel_index, r.free_xyz, [k for k, _ in moves], True))
        out[f"f{t}_label"], out[f"f{t}_track"], out[f"f{t}_state"] = r.label_id, r.track_id, r.state
    out["seconds"] = np.asarray(seconds)
    return out


def main():
    from ref_lib import have_ref, run_scenario_subprocess
    only = sys.argv[1:]
    for name, spec in SCENARIOS.items():
        if only and name not in only:
            continue
        if not have_ref(spec["setting"]):
            print(name, "skipped: oracle/_ref/libdsp_ref_s%s.so not built" % spec["setting"])
            continue
        out = run_scenario_subprocess("make_ref_golden", "run_reference", name)
        path = os.path.join(HERE, "ref_" + name + ".npz")
        seconds = out.pop("seconds")                       # timing is not part of the fixture (keeps the files reproducible)
        np.savez_compressed(path, **out)
        out["seconds"] = seconds
        last = spec["n_frames"] - 1
        print(name, os.path.getsize(path), "bytes; last frame:", int(out[f"f{last}_counts"][0]), "occupied,", "%.2f s/frame" % out["seconds"].mean())


if __name__ == "__main__":
    main()
