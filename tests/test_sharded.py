"""Sharded maps.  CPU (gloo, world_size 2): the handle-exchange plumbing.  GPU: a 2-shard map against the oracle, through
tools/sharded_check.py (two GPUs when the box has them, otherwise both shards on cuda:0)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from semantic_dsp_map_b200 import BASELINE, tiny
    from semantic_dsp_map_b200.map import connect_shards
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeEngine:                      # stands in for the CUDA engine: records what the plumbing hands it
        def __init__(self):
            self.connected = None

        def comm_export(self):
            return bytes([rank]) * 128

        def comm_connect(self, handles):
            self.connected = list(handles)

    class FakeMap:
        engine = FakeEngine()
        shard_rank, shard_count = rank, world

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    m = FakeMap()
    handles = connect_shards(m, gather)
    assert [h[0] for h in handles] == list(range(world)), handles
    assert m.engine.connected == handles
    # the slab ranges the C side derives from (rank, count) tile the storage-z axis exactly
    cfg = tiny(BASELINE["B"])
    nz = 1 << cfg.voxel_bits[2]
    lo, hi = rank * nz // world, (rank + 1) * nz // world
    ranges = gather((lo, hi))
    assert ranges[0][0] == 0 and ranges[-1][1] == nz and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_shard_handle_exchange_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, 29500 + os.getpid() % 400
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_connect_shards_rejects_wrong_world():
    sys.path.insert(0, ROOT)
    from semantic_dsp_map_b200.map import connect_shards

    class E:
        def comm_export(self):
            return b"\\0" * 128

        def comm_connect(self, h):
            raise AssertionError("must not be reached")

    class M:
        engine, shard_rank, shard_count = E(), 0, 2

    with pytest.raises(ValueError):
        connect_shards(M(), lambda obj: [obj])


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["B", "A"])
def test_two_shards_match_oracle_bit_for_bit(config):
    import torch
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29900 + os.getpid() % 90), os.path.join(ROOT, "tools", "sharded_check.py"), "--config", config, "--frames", "6"]
    if torch.cuda.device_count() < 2:
        cmd.append("--same-gpu")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sharded_check: PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_two_shards_pipelined_match_oracle_bit_for_bit():
    """The same through update_async / wait_result: every rank enqueues frame t+1 while frame t runs; both result sets live in the
    ranks' arenas and the peers' emit passes wait (through rank 0's barrier) for rank 0's transfer out of the set they overwrite."""
    import torch
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 90), os.path.join(ROOT, "tools", "sharded_check.py"), "--config", "B", "--frames", "8", "--pipelined"]
    if torch.cuda.device_count() < 2:
        cmd.append("--same-gpu")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sharded_check: PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_two_shards_of_a_multi_chunk_map_match_oracle():
    """512 x 64 x 512 voxels: every shard has 8192 occupancy tiles = two scan chunks, so the tile-count scan hands over between
    blocks and the emit pass runs from the non-empty-tile list (the large-map path) on both ranks."""
    import torch
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29800 + os.getpid() % 90), os.path.join(ROOT, "tools", "sharded_check.py"), "--config", "B", "--frames", "4", "--grid", "9,6,9"]
    if torch.cuda.device_count() < 2:
        cmd.append("--same-gpu")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "sharded_check: PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [4, 8])
def test_four_and_eight_shards_match_oracle_bit_for_bit(world):
    """One map over 4 and 8 GPUs (one process per GPU): every slab boundary, every peer arena, the per-rank particle lists of the
    second stencil pass.  Needs that many GPUs - the device-side barriers of 8 processes time-slicing ONE GPU would run into their
    own timeout - so the 1-GPU box of the round-end run skips it; profiles/ holds the log of the 8-GPU run."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    for config, extra in (("B", []), ("B", ["--grid", "9,6,9", "--frames", "4"])):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29700 + os.getpid() % 90), os.path.join(ROOT, "tools", "sharded_check.py"), "--config", config, "--frames", "6"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "sharded_check: PASS" in r.stdout, f"{config} {extra}: " + r.stdout[-2000:] + r.stderr[-2000:]
