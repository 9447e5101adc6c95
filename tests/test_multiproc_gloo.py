"""world_size-2 run of the multi-GPU host logic on CPU (gloo): contiguous sharding of the instance batch,
shard-independent Monte-Carlo draws, the final gather and the max-over-ranks timing reduction.  The compute
stand-in on CPU is the oracle (tests only); on the GPU box each rank calls libftsb.so instead (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ftspice_b200 import netlist as nl
    from ftspice_b200.sharding import gather_rows, max_over_ranks, shard_range
    from oracle import orc
    circ = nl.load(nl.PROJ1_NL)
    first, count = shard_range(total, rank, world)
    vals, fns = nl.monte_carlo(circ, count, first=first)
    st, it, x = orc.Oracle(circ).batch_op(vals, fns, threads=1)
    gx = gather_rows(torch.from_numpy(x), total)
    git = gather_rows(torch.from_numpy(it), total)
    gst = gather_rows(torch.from_numpy(st), total)
    tmax = max_over_ranks([float(rank + 1)])
    if rank == 0:
        q.put((gx.numpy(), git.numpy(), gst.numpy(), tmax, first, count))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_matches_single_process(oracle_mod):
    from ftspice_b200 import netlist as nl
    total, world = 37, 2  # ragged: 19 + 18
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    gx, git, gst, tmax, first, count = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    circ = nl.load(nl.PROJ1_NL)
    vals, fns = nl.monte_carlo(circ, total)
    st, it, x = oracle_mod.Oracle(circ).batch_op(vals, fns, threads=2)
    assert np.array_equal(gx, x) and np.array_equal(git, it) and np.array_equal(gst, st)
    assert tmax == [2.0] and (first, count) == (0, 19)


def test_shard_range_covers_everything():
    from ftspice_b200.sharding import shard_range
    for total in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                f, c = shard_range(total, r, world)
                got += list(range(f, f + c))
            assert got == list(range(total))
