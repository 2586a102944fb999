"""Sharded map vs the CPU oracle, bit for bit.  One process per shard:

    torchrun --nproc-per-node G tools/sharded_check.py [--config B] [--frames N] [--same-gpu]

--same-gpu puts every shard on cuda:0 (two processes time-slicing one GPU: slow, but exercises the whole exchange on a
1-GPU box).  Rank 0 compares the sharded map's result with the oracle's canonical order on the same seeded frames."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from semantic_dsp_map_b200 import BASELINE, Scene, SemanticDSPMap, connect_shards, make_random_table, tiny


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="B")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--same-gpu", action="store_true")
    ap.add_argument("--full", action="store_true", help="full-size config instead of the tiny variant")
    ap.add_argument("--pipelined", action="store_true", help="frames through update_async / wait_result (frame t+1 enqueued while frame t runs)")
    ap.add_argument("--grid", default="", help="voxel bits x,y,z overriding the (tiny) config's grid, e.g. 9,6,9")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dist.init_process_group("gloo")
    device = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(device)
    cfg = BASELINE[args.config] if args.full else tiny(BASELINE[args.config])
    if args.grid:
        import dataclasses
        cfg = dataclasses.replace(cfg, voxel_bits=tuple(int(b) for b in args.grid.split(",")))
    m = SemanticDSPMap(cfg, device=device, shard_rank=rank, shard_count=world)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    if world > 1:
        connect_shards(m, gather)
    table = make_random_table(cfg.seed, 50_000)
    m.engine.set_random_table(table)
    orc = None
    if rank == 0:
        from oracle_lib import oracle_map
        orc = oracle_map(cfg)
        orc.engine.set_random_table(table)
    scene = Scene(cfg)
    ok = True
    def check(t, r):
        nonlocal ok
        if rank != 0:
            return
        fr = scene.frame(t)
        ro = orc.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=True, moves=fr.moves, remove_tracks=[2] if t == 5 else [])
        for name in ("voxel_index", "xyz", "weight_sum", "track_id", "label_id", "state", "free_voxel_index", "free_xyz"):
            a, b = getattr(r, name), getattr(ro, name)
            same = a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
            if not same:
                ok = False
                print(f"frame {t}: {name} differs: engine {a.shape} oracle {b.shape}", flush=True)
        print(f"frame {t}: occupied {len(r.voxel_index)} free {len(r.free_voxel_index)} {'OK' if ok else 'MISMATCH'}", flush=True)

    prev = None
    for t in range(args.frames):
        fr = scene.frame(t)
        rm = [2] if t == 5 else []
        if args.pipelined:      # every rank submits frame t, then collects frame t-1 (rank 0 holds the whole result)
            k = m.update_async(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=True, moves=fr.moves, remove_tracks=rm)
            if prev is not None:
                check(t - 1, m.wait_result(prev))
            prev = k
        else:
            check(t, m.update(fr.depth, fr.masks, fr.position, fr.orientation, if_get_freespace=True, moves=fr.moves, remove_tracks=rm))
    if args.pipelined:
        check(args.frames - 1, m.wait_result(prev))
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not bool(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("sharded_check: PASS", flush=True)


if __name__ == "__main__":
    main()
