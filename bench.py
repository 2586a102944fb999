#!/usr/bin/env python
"""bench.py — depth frames/sec of the particle-map update path (BASELINE.json metric) on config B
(ZED2 mode, 256x64x256 voxels = "256x256x64 grid", 8 slots/voxel, 1280x720 depth + masks).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--config B]

A step = one SemanticDSPMap::update() over one synthetic frame (SURVEY.md §8d scene; seeded).
  value : frames/s with the frame's depth + masks already resident in HBM, through dspmap_update_device_async +
          dspmap_wait_result (frame t+1 is enqueued while frame t runs; every frame's result is collected inside the timed
          region), timed with CUDA events on the stream the kernels run on, max over ranks.  `synchronous`: the same frames
          through dspmap_update_device, which returns after its frame has finished (round 1's `value`).
  e2e   : frames/s through the reference-facing call with HOST buffers (SemanticDSPMap.update_async + wait_result =
          dspmap_update_async + dspmap_wait_result): host->pinned->device copies of depth and masks and the device->host
          read of every frame's occupied list are inside the timed region; frame t's result transfer overlaps frame t+1.
          `synchronous_value`: the same through update() (= dspmap_update + dspmap_get_result), nothing overlapped.
  roofline    : predict+update stages against SURVEY 8(d)'s dense numerator; roofline_stencil: the two dominant kernels in pair
                evaluations/s; roofline_occ_headers: the HBM-streaming kernel (k_occ_headers) - all vs MEASURED_PEAKS.json.
  cpu_baseline: the reference's own code (oracle/_ref, unmodified headers; else the oracle port) timed on this box's host
                cores on a bounded sample of the same frames, with the reference's own per-stage timers beside it.
  parity_vs_reference: default list order vs the reference's breadth-first order on the first 40 benchmark frames (N = 1).
N > 1: one process per GPU (torchrun); each rank updates its own map replica from its own frame stream (weak scaling,
no data-path collective), and `sharded_D` adds ONE config-D map (512x128x512 voxels x 16 slots) sharded over the same GPUs
(strong scaling, NVLink peer-memory exchange).  --impl reference times the reference's CPU code instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "depth frames/sec on 256x256x64 grid, 8 particles/voxel"
UNIT = "frames/s"


def config_strings(cfg):
    """The workload description both arms print (the driver compares them)."""
    nx, ny, nz = (1 << b for b in cfg.voxel_bits)
    slots = cfg.n_slots
    hi, wi = cfg.internal_size
    return {"workload": cfg.name,
            "grid": f"{nx}x{ny}x{nz} voxels x {1 << cfg.slot_bits} slots ({slots / 1e6:.1f}M slots, {slots * 24 / 1e6:.0f} MB as 24 B records)",
            "image": f"{cfg.image_width}x{cfg.image_height} depth + static mask + instance masks -> {wi}x{hi} internal"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def _gen_frame(args):
    name, t, overrides = args
    from semantic_dsp_map_b200 import BASELINE, Scene
    import dataclasses
    cfg = dataclasses.replace(BASELINE[name], **overrides)
    return Scene(cfg).frame(t)


def generate_frames(name, n, overrides=None):
    """Frames are a pure function of (config, t); fan the ray casting out over the host cores (before CUDA is touched)."""
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(name, t, overrides or {}) for t in range(n)]
    world = int(os.environ.get("WORLD_SIZE", "1"))      # every rank generates the same frames: share the host cores
    workers = max(1, min((os.cpu_count() or 1) // world, 16, n))
    if workers == 1:
        return [_gen_frame(j) for j in jobs]
    with ProcessPoolExecutor(workers) as ex:
        return list(ex.map(_gen_frame, jobs, chunksize=max(1, n // (workers * 4))))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(self.samples)}


class _StdoutToStderr:
    """The reference chats on std::cout (setters, warnings): keep this process's stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference(cfg_name):
    """The CPU implementation the baselines time, as (step(frame) -> seconds, kind, description).

    kind "reference": the reference's OWN hot-path code, compiled from /root/reference/include by oracle/build_ref.sh with the
    benchmark's grid constants (oracle/_ref/libdsp_ref_sB.so; third-party headers replaced by the stand-ins of oracle/ref_stubs,
    see oracle/ref_harness.cpp) - available for config B.  kind "port": the oracle (oracle/dsp_oracle.cpp, pinned bit for bit to
    that build) in the reference's iteration order.  Both are single-threaded like the shipped code
    (semantic_dsp_map.h:749 calls the non-MT BFS)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from semantic_dsp_map_b200 import BASELINE, make_random_table
    cfg = BASELINE[cfg_name]
    import ref_lib
    if cfg_name == "B" and ref_lib.have_ref("B"):
        import tempfile
        with tempfile.TemporaryDirectory() as td, _StdoutToStderr():
            csv = os.path.join(td, "object_info.csv")
            ref_lib.write_label_csv(csv)
            ref = ref_lib.RefMap(cfg, csv, setting="B", eval_format=False)
            ref.set_random_table(make_random_table(cfg.seed))
            ref.set_color_map(np.random.default_rng(256).permutation(256))

        def step(fr):
            with _StdoutToStderr():
                ref.update(fr.depth, fr.masks, fr.position, fr.orientation, False, fr.moves, fr.remove_tracks, fetch=False)
            return ref.seconds()          # time inside the reference call (input Mats copied like cv_bridge::toCvCopy, then update())
        return step, "reference", "the reference's own code (oracle/_ref/libdsp_ref_sB.so: unmodified headers, benchmark grid constants, stand-in Eigen/OpenCV/PCL)"
    from oracle_lib import oracle_map
    m = oracle_map(cfg, order=0)
    m.enable_colors()
    m.engine.set_random_table(make_random_table(cfg.seed))

    def step(fr):
        t0 = time.perf_counter()
        m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks)
        return time.perf_counter() - t0
    return step, "port", "oracle/dsp_oracle.cpp in the reference's iteration order"


def run_reference(args, cfg_name):
    """The reference's own CPU implementation of the path, single thread like the shipped code."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from semantic_dsp_map_b200 import BASELINE
    cfg = BASELINE[cfg_name]
    # bounded: the CPU path runs at about 1-2 frames/s on this config
    steps, warmup = min(args.steps, 20), min(args.warmup, 5)
    frames = generate_frames(cfg_name, warmup + steps)
    step, kind, what = cpu_reference(cfg_name)
    for fr in frames[:warmup]:
        step(fr)
    dt = sum(step(fr) for fr in frames[warmup:])
    fps = steps / dt
    sample = f"frames {warmup}..{warmup + steps - 1} of the seeded sequence after {warmup} untimed frames (requested steps={args.steps}, warmup={args.warmup}; bounded); {what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(config_strings(cfg), l2="inputs larger than L2 (the particle store does not fit the CPU's caches either)", parallelism=("replicas: %d independent maps" % args.gpus) if args.gpus > 1 else "single"),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def measure_engine(m, frames, warmup, steps, stream, dev, barrier, n_prof, torch, device_arm_only=False):
    """All engine-arm measurements of one map over one frame sequence.  Returns a dict of raw numbers (per rank)."""
    from semantic_dsp_map_b200 import Engine, _capi
    eng = m.engine
    n_frames = warmup + steps

    def launches():
        n = C.c_uint64()
        eng.lib.dspmap_get_launch_count(eng._h, C.byref(n))
        return n.value

    # ---------- device-resident arm ("value") -----------------------------------------------------------------
    d_frames = []
    for fr in frames:
        lowered = m.lower_masks(fr.masks)
        d_depth = torch.from_numpy(fr.depth).to(dev)
        d_masks = [torch.from_numpy(np.ascontiguousarray(mk.mask)).to(dev) for mk in fr.masks]
        for lm, dm in zip(lowered, d_masks):
            lm["data"], lm["row_stride"] = dm.data_ptr(), dm.stride(0)
        # the descriptor arrays (dspmap_mask / dspmap_object_move: a few hundred bytes of host structs) are packed once here
        d_frames.append((d_depth, d_masks, Engine.pack_masks(lowered), Engine.pack_moves(fr.moves)))
    torch.cuda.synchronize(dev)
    res = _capi.Result()

    def step_device(k, fn="update_device"):
        fr, (d_depth, _, lowered, moves) = frames[k], d_frames[k]
        return eng.update_raw(fn, d_depth.data_ptr(), fr.depth.shape[0], fr.depth.shape[1], d_depth.stride(0), lowered,
                              fr.position, fr.orientation, moves, fr.remove_tracks, False)

    def timed(run):
        """`run` (the K timed frames) between two events on the engine's stream, and by the wall clock."""
        barrier()
        gc.collect()
        gc.disable()      # a cyclic-GC pass inside a 50 ms timed region would be a tenth of it; re-enabled after the last timed step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        run()
        e1.record(stream)
        barrier()
        wall = 1e3 * (time.perf_counter() - t0)
        gc.enable()
        return max(e0.elapsed_time(e1), 0.0), wall

    out = {}
    # pipelined (the headline): frame t+1 is enqueued while frame t runs; every frame's result is collected inside the timed region
    for k in range(warmup):      # the warm-up frames go through the same call as the timed ones (its first use allocates the second result set)
        eng.check(eng.lib.dspmap_wait_result(eng._h, step_device(k, "update_device_async"), C.byref(res)), "dspmap_wait_result")
    sampler = ClockSampler(dev.index)
    sampler.start()
    l0 = launches()

    def run_device_pipelined():
        prev = 0
        for k in range(warmup, n_frames):
            t = step_device(k, "update_device_async")
            if prev:
                eng.check(eng.lib.dspmap_wait_result(eng._h, prev, C.byref(res)), "dspmap_wait_result")
            prev = t
        eng.check(eng.lib.dspmap_wait_result(eng._h, prev, C.byref(res)), "dspmap_wait_result")

    ms, wall = timed(run_device_pipelined)
    out["ms_device"] = ms
    out["gpu_launches"] = launches() - l0
    out["clocks"] = sampler.summary()
    # the frames that follow, with the engine's per-stage CUDA events on: where a frame's time goes, and the counters of the roofline numerators
    stage_ms = np.zeros(8)
    acc = {"n_listed": 0, "n_occ": 0, "n_vis": 0, "n_swept": 0}
    eng.check(eng.lib.dspmap_set_profiling(eng._h, 1), "dspmap_set_profiling")
    for k in range(n_frames, n_frames + n_prof):
        step_device(k)
        st = eng.stats()
        stage_ms += np.asarray(st["ms_stage"])
        acc["n_occ"] += st["n_occupied"]; acc["n_vis"] += st["n_visible"]; acc["n_swept"] += st["n_swept_voxels"]
        acc["n_listed"] += st["n_nonempty_voxels"]
    eng.check(eng.lib.dspmap_set_profiling(eng._h, 0), "dspmap_set_profiling")
    out["stage_ms"], out["acc"] = stage_ms, acc
    if device_arm_only:
        out.update(ms_device_sync=out["ms_device"], ms_e2e=out["ms_device"], ms_e2e_sync=out["ms_device"], h2d=0, d2h=0)
        return out
    # synchronous calls on the same frames (every update returns after its frame has finished): what round 1 reported as `value`
    eng.clear()
    for k in range(warmup):
        step_device(k)
    ms, _ = timed(lambda: [step_device(k) for k in range(warmup, n_frames)])
    out["ms_device_sync"] = ms

    # ---------- end-to-end arm: host buffers in, host result out -------------------------------------------------
    h2d = d2h = 0
    fields = ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state", "rgb")
    eng.clear()
    for k in range(warmup):
        fr = frames[k]
        m.wait_result(m.update_async(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks), copy_result=False)

    def run_e2e_pipelined():
        # results are read where dspmap_wait_result leaves them: the handle's pinned host buffers (valid for two more frames)
        nonlocal h2d, d2h
        prev = 0
        for k in range(warmup, n_frames + 1):
            t = 0
            if k < n_frames:
                fr = frames[k]
                t = m.update_async(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks)
                h2d += fr.depth.nbytes + sum(mk.mask.nbytes for mk in fr.masks)
            if prev:
                r = m.wait_result(prev, copy_result=False)
                d2h += sum(getattr(r, a).nbytes for a in fields) + 80
            prev = t

    ms, wall = timed(run_e2e_pipelined)
    out["ms_e2e"] = max(ms, wall)
    out["h2d"], out["d2h"] = h2d // steps, d2h // steps
    eng.clear()
    for k in range(warmup):
        fr = frames[k]
        m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks)

    def run_e2e_sync():
        for k in range(warmup, n_frames):
            fr = frames[k]
            m.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks, copy_result=False)

    ms, wall = timed(run_e2e_sync)
    out["ms_e2e_sync"] = max(ms, wall)
    del d_frames
    return out


def order_parity(cfg, frames, n, device):
    """The engine's default (canonical) order of a pixel's particle list against the REFERENCE's iteration order on the benchmark
    frames: two engines, the second one replaying the reference's breadth-first search (dspmap_params.reference_iteration_order -
    the mode that tests/test_reference_build_gpu.py holds to the reference's own build bit for bit).  Counts occupied voxels that
    are in only one result, voxels whose fused label or track differs, and the largest relative weight-sum difference."""
    from semantic_dsp_map_b200 import SemanticDSPMap, make_random_table
    a, b = SemanticDSPMap(cfg, device=device), SemanticDSPMap(cfg, device=device)
    b.setReferenceIterationOrder(True)
    table = make_random_table(cfg.seed)
    for mm in (a, b):
        mm.enable_colors()
        mm.engine.set_random_table(table)
    total = only = lab = 0
    rel = 0.0
    for fr in frames[:n]:
        ra, rb = (mm.update(fr.depth, fr.masks, fr.position, fr.orientation, moves=fr.moves, remove_tracks=fr.remove_tracks) for mm in (a, b))
        only += len(np.setxor1d(ra.voxel_index, rb.voxel_index))
        common, pa, pb = np.intersect1d(ra.voxel_index, rb.voxel_index, return_indices=True)
        lab += int(((ra.label_id[pa] != rb.label_id[pb]) | (ra.track_id[pa] != rb.track_id[pb])).sum())
        if len(common):
            rel = max(rel, float(np.max(np.abs(ra.weight_sum[pa] - rb.weight_sum[pb]) / np.maximum(np.abs(rb.weight_sum[pb]), 1e-30))))
        total += len(rb.voxel_index)
    a.engine.close(); b.engine.close()
    return {"flips": only + lab, "total": total, "only_in_one_result": only, "label_or_track_differs": lab, "max_rel_weight_diff": rel, "frames": n,
            "what": "default (canonical) list order vs the reference's breadth-first order, both on the engine; the latter equals the reference's own build bit for bit (tests/test_reference_build_gpu.py)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="B")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-arm-only", action="store_true", help="profiling runs (ncu): only the device-resident pipelined arm and the per-stage frames")
    ap.add_argument("--no-sharded-d", action="store_true", help="N > 1: skip the extra measurement of ONE config-D map sharded over the N GPUs")
    ap.add_argument("--parallelism", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: independent map replicas (weak scaling, default) or ONE map sharded by storage-z slabs over the GPUs (strong)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, args.config)
    args.warmup = max(args.warmup, 3)

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    # the engine's host staging helpers (pageable -> pinned copies of dspmap_update) poll between jobs: share the host cores
    # between the ranks of one box instead of letting every rank start its default of six
    os.environ.setdefault("DSPMAP_STAGE_THREADS", str(max(1, min(6, (os.cpu_count() or 1) // world - 2))))
    from semantic_dsp_map_b200 import BASELINE, SemanticDSPMap, _capi, connect_shards, make_random_table
    cfg = BASELINE[args.config]
    n_frames = args.warmup + args.steps
    n_prof = min(20, args.steps)                             # frames after the timed ones: per-stage CUDA events and counters
    frames = generate_frames(args.config, n_frames + n_prof)          # before CUDA / NCCL are touched (fork safety)
    sharded = world > 1 and args.parallelism == "sharded"
    # N > 1, default run: besides the N replicas of the metric's configuration, ONE config-D map (512 x 128 x 512 voxels x 16 slots,
    # 2 m per frame: BASELINE.json's multi-GPU configuration) sharded over the same N GPUs
    extra_d = world > 1 and not sharded and args.config == "B" and not args.no_sharded_d and not args.device_arm_only
    d_warm, d_steps = 10, 40
    frames_d = generate_frames("D", d_warm + d_steps + 8) if extra_d else None

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_map(c, shard):
        mm = SemanticDSPMap(c, device=local, shard_rank=rank, shard_count=world) if shard else SemanticDSPMap(c, device=local)
        if shard:
            connect_shards(mm, gather)
        mm.enable_colors()                   # the reference colours every emitted point (getOccupancyResult): so do both arms
        mm.engine.set_random_table(make_random_table(c.seed))
        st = torch.cuda.Stream(device=dev)
        mm.engine.check(mm.engine.lib.dspmap_set_stream(mm.engine._h, C.c_void_p(st.cuda_stream)), "dspmap_set_stream")
        return mm, st

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    sharded_d = None
    if extra_d:
        # never at the price of the metric's own line: a failure here (a peer missing a device-side barrier ends every rank's frame
        # with DSPMAP_ERR_COMM after its 2 s time-out) is reported in the field and the replicas are measured all the same
        try:
            cfg_d = BASELINE["D"]
            md, stream_d = make_map(cfg_d, True)
            rd = measure_engine(md, frames_d, d_warm, d_steps, stream_d, dev, barrier, 8, torch)
            v = reduce_max([rd["ms_device"], rd["ms_e2e"], rd["ms_device_sync"], rd["ms_e2e_sync"]])
            barrier()
            md.engine.close()
            del md
            sharded_d = {"config": dict(config_strings(cfg_d), parallelism="sharded: one map, storage-z slabs over %d GPUs, NVLink peer-memory exchange, no collective library in the data path" % world),
                         "scaling": "strong", "steps": d_steps, "warmup": d_warm, "value": d_steps / (v[0] * 1e-3), "unit": UNIT, "ms_per_step": v[0] / d_steps,
                         "e2e": {"value": d_steps / (v[1] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": rd["h2d"], "d2h_bytes_per_step": rd["d2h"],
                                 "synchronous_value": d_steps / (v[3] * 1e-3)},
                         "synchronous": {"value": d_steps / (v[2] * 1e-3), "ms_per_step": v[2] / d_steps},
                         "stage_ms_rank0": {n: float(x / 8) for n, x in zip(_capi.STAGE_NAMES, rd["stage_ms"])},
                         "note": "every rank submits the same frame (pipelined calls) and meets the others at device-side barriers inside the frame; "
                                 "the same map on one GPU: profiles/r02_summary.md"}

        except Exception as exc:
            sharded_d = {"error": str(exc)[:300]}
            gc.enable()
            try:
                barrier()
            except Exception:
                pass

    m, stream = make_map(cfg, sharded)
    units = 1 if sharded else world          # maps updated per step by the whole job
    r = measure_engine(m, frames, args.warmup, args.steps, stream, dev, barrier, n_prof, torch, args.device_arm_only)
    ms_device, ms_e2e, ms_device_sync, ms_e2e_sync = reduce_max([r["ms_device"], r["ms_e2e"], r["ms_device_sync"], r["ms_e2e_sync"]])
    barrier()
    m.engine.close()
    del m

    if rank == 0:
        peak, peak_src = load_peaks()
        stage_ms, acc = r["stage_ms"], r["acc"]
        # Rooflines, all against the measured HBM peak (MEASURED_PEAKS.json).
        #  * roofline (the headline): the north-star's "fused predict+update" - everything from the visibility sweep to the weight
        #    write-back (k_sweep_select .. k_gather_visible, k_ck, k_weight_update; SURVEY.md section 8 rows a12 + a13) - charged
        #    with section 8(d)'s layout-independent numerator: 24 B per slot of every swept voxel + 8 B per updated particle + 24 B
        #    per internal pixel, over the time of those two stages (CUDA events inside the engine, on its stream).  The engine does
        #    NOT move those bytes - the header plane lets it open only non-empty voxels (`traffic` = what ncu saw it move) - and the
        #    two stencil passes that dominate are bound by the shared-memory pipe (three table look-ups per particle-pixel pair),
        #    so this is the dense formulation's yardstick, reported because it is the one the target is stated in.
        #  * roofline_stencil: the two dominant kernels alone (k_ck + k_weight_update = the `update` stage) in the unit that bounds
        #    them: (particle, pixel) pair evaluations per second against an issue-slot ceiling.
        #  * roofline_occ_headers: the one kernel that streams a whole plane from HBM (4-byte header of EVERY voxel in, one state
        #    byte per voxel out, voxels that need their slot records listed): 4*V + 1*V + 4*n_listed bytes over its own launch.
        n_vox = cfg.n_voxels if not sharded else cfg.n_voxels // world      # a rank streams its own slabs
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(cfg.name, {})
        tr = lambda k: (traffic[k]["dram_read"] + traffic[k]["dram_write"]) if k in traffic else None      # one ncu --set full capture, per frame / launch
        pu_bytes = (24.0 * cfg.n_slots / cfg.n_voxels * acc["n_swept"] + 8.0 * acc["n_vis"]) / n_prof + 24.0 * cfg.internal_size[0] * cfg.internal_size[1]
        pu_ms = (stage_ms[3] + stage_ms[4]) / n_prof
        occ_bytes = 5.0 * n_vox + 4.0 * acc["n_listed"] / n_prof
        occ_ms = stage_ms[1] / n_prof
        win = 2 * (cfg.window_half or (3 if cfg.boost_mode else 5)) + 1
        pairs = 2.0 * win * win * acc["n_vis"] / n_prof           # both passes evaluate every (listed particle, window pixel) pair
        upd_ms = stage_ms[4] / n_prof
        sm_mhz = (r["clocks"].get("sm_mhz") or 1965.0)
        issue_peak = 148 * 4 * sm_mhz * 1e6 / 40.0 * 32.0         # 148 SMs x 4 schedulers, ~40 warp instructions per 32 pairs (VERDICT r01's floor)
        out = {
            "metric": METRIC, "value": units * args.steps / (ms_device * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_device / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(config_strings(cfg), l2="inputs larger than L2 (the particle store does not fit the CPU's caches either)",
                           parallelism= ("sharded: one map, storage-z slabs over %d GPUs, NVLink peer-memory exchange" % world) if sharded else ("replicas: %d independent maps" % world if world > 1 else "single")),
            "e2e": {"value": units * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "call": "dspmap_update_async + dspmap_wait_result (frame t's result transfer overlaps frame t+1; every result collected inside the timed region)",
                    "synchronous_value": units * args.steps / (ms_e2e_sync * 1e-3)},
            "call": "dspmap_update_device_async + dspmap_wait_result",
            "synchronous": {"value": units * args.steps / (ms_device_sync * 1e-3), "ms_per_step": ms_device_sync / args.steps, "call": "dspmap_update_device (returns after the frame has finished)"},
            "gpu_launches": int(r["gpu_launches"]),
            "clocks": r["clocks"],
            "roofline": {"kernel": "predict+update: k_sweep_select, k_sweep_slots, pixel-list build, k_ck, k_weight_update", "bound": "hbm",
                         "achieved": pu_bytes / (pu_ms * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": pu_bytes / (pu_ms * 1e-3) / 1e9 / peak, "traffic": tr("predict_update"), "algorithmic_bytes": pu_bytes, "ms": pu_ms,
                         "note": "SURVEY 8(d) numerator of the dense formulation over the visibility + update stages; the engine opens only non-empty "
                                 "voxels (traffic << algorithmic bytes) and the stencil passes are bound by the shared-memory pipe, not by HBM: "
                                 "DESIGN.md section 4, profiles/r02_summary.md"},
            "roofline_stencil": {"kernel": "k_ck + k_weight_update (the update stage: the two dominant kernels)", "bound": "issue slots / shared-memory pipe",
                                 "achieved": pairs / (upd_ms * 1e-3) / 1e9, "peak": issue_peak / 1e9, "unit": "G pair evaluations/s",
                                 "frac": pairs / (upd_ms * 1e-3) / issue_peak, "pairs_per_frame": pairs, "ms": upd_ms,
                                 "dram_frac": ((tr("k_ck") or 0) + (tr("k_weight_update") or 0)) / (upd_ms * 1e-3) / 1e9 / peak if tr("k_ck") else None,
                                 "note": "ceiling = 148 SMs x 4 issue slots x SM clock at 40 warp instructions per 32 pairs (three IEEE divisions, three table look-ups, "
                                         "five products per pair without FMA contraction)"},
            "roofline_occ_headers": {"kernel": "k_occ_headers", "bound": "hbm", "achieved": occ_bytes / (occ_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": occ_bytes / (occ_ms * 1e-3) / 1e9 / peak, "traffic": tr("k_occ_headers"), "algorithmic_bytes": occ_bytes, "ms": occ_ms},
            "stage_ms": {n: float(v / n_prof) for n, v in zip(_capi.STAGE_NAMES, stage_ms)},
        }
        if sharded_d is not None:
            out["sharded_D"] = sharded_d
        if world == 1 and not sharded and not args.device_arm_only:
            try:
                out["parity_vs_reference"] = order_parity(cfg, frames, min(40, n_frames), local)
            except Exception as exc:      # a measurement beside the bench line, never a reason to lose it
                out["parity_vs_reference"] = {"error": str(exc)[:200]}
        if args.device_arm_only:
            out["note"] = "--device-arm-only: e2e and synchronous fields repeat the device arm (profiling run, not a bench line)"
        if not args.no_cpu_baseline and world == 1 and not args.device_arm_only:      # the CPU baseline is a rank-0, N = 1 measurement
            out["cpu_baseline"] = cpu_baseline(args.config, frames)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(cfg_name, frames, warmup=3, steps=12):
    """The CPU implementation (see cpu_reference), one thread, on the first frames of the same sequence."""
    steps = min(steps, len(frames) - warmup)
    step, kind, what = cpu_reference(cfg_name)
    for fr in frames[:warmup]:
        step(fr)
    dt = sum(step(fr) for fr in frames[warmup:warmup + steps])
    out = {"value": steps / dt, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"frames {warmup}..{warmup + steps - 1} of the same seeded sequence (map still filling; the CPU path slows as it fills); {what}",
           "host_cores_available": os.cpu_count()}
    try:
        split = cpu_stage_split(cfg_name, frames, 6)
        if split:
            out["stage_ms"] = split
    except Exception as exc:
        out["stage_ms"] = {"error": str(exc)[:200]}
    return out


def cpu_stage_split(cfg_name, frames, n):
    """Per-stage CPU times with the reference's own stage names, from the reference's own timers: the same code compiled with
    VERBOSE_MODE 1 (oracle/_ref/libdsp_ref_sBV.so, oracle/build_ref.sh) prints "Average time used for <stage>: <s> s" every
    frame (include/semantic_dsp_map.h:912-953); the last set over frames 0..n-1 is returned in milliseconds.  ("projecttion" is
    the visibility sweep inside "update weight", as the reference's t2_1 - t2 / t3 - t2 define them.)  In a child process: the
    verbose build owns a second 805 MB static particle array and writes to std::cout."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    if cfg_name != "B" or not ref_lib.have_ref("BV"):
        return None
    code = (
        "import sys, os, tempfile\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, 'tests'))\n"
        "from semantic_dsp_map_b200 import BASELINE, Scene, make_random_table\n"
        "import ref_lib\n"
        "cfg = BASELINE['B']\n"
        "td = tempfile.mkdtemp(); csv = os.path.join(td, 'o.csv'); ref_lib.write_label_csv(csv)\n"
        "ref = ref_lib.RefMap(cfg, csv, setting='BV', eval_format=False)\n"
        "ref.set_random_table(make_random_table(cfg.seed))\n"
        "sc = Scene(cfg)\n"
        f"for t in range({n}):\n"
        "    fr = sc.frame(t)\n"
        "    ref.update(fr.depth, fr.masks, fr.position, fr.orientation, False, fr.moves, fr.remove_tracks, fetch=False)\n"
        "sys.stdout.flush()\n")
    text = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300).stdout
    import re
    found = {}
    for name, val in re.findall(r"Average time used for ([a-z ]+): ([0-9.eE+-]+) s", text):
        found[name.strip()] = float(val) * 1e3          # later frames overwrite earlier ones: the running average after frame n-1
    if not found:
        return None
    found["frames"] = n
    found["source"] = "the reference's VERBOSE_MODE stage timers (oracle/_ref/libdsp_ref_sBV.so), running average over the first frames"
    return found


if __name__ == "__main__":
    main()
