"""Multi-GPU path, CPU tier: bench.py shards the batched worlds across ranks with no data-path collective — every rank
owns worlds [rank * n, (rank + 1) * n) and only the timing is reduced (MAX over ranks).  Here two gloo ranks run that
sharding logic with the CPU emulator standing in for the device and must reproduce the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLDS_PER_RANK, STEPS = 3, 20


def _run_block(first_world, n):
    from emu import EmuBatch
    from resolve2d_b200 import scenes
    b = EmuBatch(n, 2.0, 4)
    for w in range(n):
        scenes.build_batch_world(b.world(w), first_world + w, nx=8, ny=4)
    for _ in range(STEPS):
        b.process(scenes.DT, 4, 4)
    return np.concatenate([b.world(w).read_bodies()["pos"] for w in range(n)])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pos = _run_block(rank * WORLDS_PER_RANK, WORLDS_PER_RANK)
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # stands in for the per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                        # bench.py: max over ranks
    n = torch.tensor([pos.shape[0]], dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)                        # bench.py: units processed by all ranks
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), pos)
    if rank == 0:
        np.save(os.path.join(out_dir, "meta.npy"), np.array([t.item(), n.item()]))
    dist.destroy_process_group()


def test_two_ranks_shard_worlds_without_collectives(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    want = _run_block(0, world * WORLDS_PER_RANK)
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))   # sharded == unsharded, bit for bit
    t_max, n_total = np.load(tmp_path / "meta.npy")
    assert t_max == float(world) and n_total == want.shape[0]
