"""world_size-2 gloo tests (CPU) of the host-side plumbing of the z-slab decomposition."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vegas_rs_b200 import distributed as vd
    nz, zoff = vd.slab_extent(16, rank, world)
    lo, hi = vd.neighbours(rank, world)
    blob = bytes([rank]) * 256
    blobs = vd.exchange_blobs(blob, dist)
    e, m = vd.reduce_observables(np.array([1.0 + rank, 2.0]), np.full((2, 3), float(rank + 1)), dist, "cpu")
    q.put((rank, nz, zoff, lo, hi, [b[0] for b in blobs], e.tolist(), m.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_plumbing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, nz, zoff, lo, hi, blobs, e, m in res:
        assert nz == 8 and zoff == 8 * rank
        assert (lo, hi) == ((rank - 1) % 2, (rank + 1) % 2)
        assert blobs == [0, 1]
        assert e == [3.0, 4.0] and m == [[3.0] * 3] * 2


def test_slab_extent_validation():
    from vegas_rs_b200 import distributed as vd
    assert vd.slab_extent(1024, 3, 8) == (128, 384)
    with pytest.raises(ValueError):
        vd.slab_extent(10, 0, 4)
    with pytest.raises(ValueError):
        vd.slab_extent(6, 0, 2)
    assert vd.neighbours(0, 8) == (7, 1) and vd.neighbours(7, 8) == (6, 0)
