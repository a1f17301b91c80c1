"""CPU tests of the N>1 path (gloo, world_size 2): contiguous frame ranges per rank, MAX-reduced step time, gathered
per-frame counts.  The data path itself has no collective (frames are independent), so this is all the distributed
logic there is; each rank here runs the CPU oracle on its own frame range as a stand-in for its GPU."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from iv_slam_b200 import sharding


def test_frame_ranges_partition_the_batch():
    for total in (0, 1, 7, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.frame_range(g, world, total) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iv_slam_b200 import synthetic as S
    from oracle import oracle_lib as O
    s, e = sharding.frame_range(rank, world, total)
    # every rank regenerates the same global sequence and takes its own slice (as bench.py does per rank)
    L, R = S.make_stereo_batch(320, 240, total, 500, distinct=total)
    params = dict(nfeatures=300, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)
    nL, nM = O.stereo_batch(params, L[s:e], R[s:e], 100.0, 400.0, 1)
    dist.barrier()
    t = sharding.reduce_max(0.5 + rank)
    full = sharding.gather_counts(nL, total, rank, world)
    if rank == 0:
        q.put((t, full.tolist()))
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    total, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t, full = q.get()
    assert t == 1.5                                     # max over ranks of (0.5, 1.5)
    from iv_slam_b200 import synthetic as S
    from oracle import oracle_lib as O
    L, R = S.make_stereo_batch(320, 240, total, 500, distinct=total)
    nL, _ = O.stereo_batch(dict(nfeatures=300, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7), L, R, 100.0, 400.0, 1)
    assert full == nL.tolist() and sum(full) > 0
