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


def test_cooldown_point_list_matches_reference_quirks():
    """src/program.rs:202-211 accumulates T in f64."""
    from vegas_rs_b200 import distributed as vd
    pts = vd.cooldown_temperatures(6.0, 1.0, 0.05)
    assert len(pts) == 101 and pts[0] == 6.0 and pts[-1] == 1.0000000000000133
    pts = vd.cooldown_temperatures(4.0, 0.1, 0.1)
    assert len(pts) == 39 and pts[-1] == 0.1999999999999976
    assert len(vd.cooldown_temperatures(3.0, 0.05, 0.1)) == 30
    with pytest.raises(ValueError):
        vd.cooldown_temperatures(1.0, 2.0, 0.1)
    with pytest.raises(ValueError):
        vd.cooldown_temperatures(2.0, 1.0, 0.0)
    shards = [vd.shard_points(pts, r, 8) for r in range(8)]
    assert sorted(t for s in shards for t in s) == sorted(pts) and max(map(len, shards)) - min(map(len, shards)) <= 1
    assert all(s[0] > 3.0 and s[-1] < 1.0 for s in shards)  # every rank spans the range


def _points_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vegas_rs_b200 import distributed as vd

    class FakeMachine:  # records the one-point CoolDown calls a rank issues
        def __init__(self): self.lines = []
        def cooldown(self, tmax, tmin, rate, relax, steps):
            assert tmax == tmin
            self.lines.append(f"{tmax:.16f} 0.0 {relax} {steps}")

    pts = vd.cooldown_temperatures(3.0, 2.0, 0.25)
    m = FakeMachine()
    vd.sharded_cooldown(m, vd.shard_points(pts, rank, world), 0.25, 10, 20)
    merged = vd.gather_lines(m.lines, dist)
    q.put((rank, len(m.lines), merged))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cooldown_points_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 150
    procs = [ctx.Process(target=_points_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    assert [r[1] for r in res] == [3, 2]
    temps = [float(ln.split()[0]) for ln in res[0][2]]
    assert temps == [3.0, 2.75, 2.5, 2.25, 2.0] and res[0][2] == res[1][2]


def _group_machine_worker(rank, world, port, q):
    """One Machine per rank over the scripted test double of the device (tests/mock), grouped with the real
    torch.distributed reduction the TOML front end uses (run.group_reduce) on gloo."""
    import ctypes as C
    import subprocess
    import types
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "tests", "mock", "libvegas_host_mock.so")
    if rank == 0 and not os.path.exists(so):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(root, "include"), "-o", so,
                        os.path.join(root, "tests", "mock", "mock_vegas_gpu.cpp"), os.path.join(root, "vegas_rs_b200", "csrc", "vegas_host.cpp")], check=True)
    dist.barrier()
    lib = C.CDLL(so)
    lib.mock_gpu_create.restype, lib.mock_gpu_create.argtypes = C.c_void_p, [C.c_uint64, C.c_int, C.c_uint64]
    lib.mock_gpu_destroy.restype, lib.mock_gpu_destroy.argtypes = None, [C.c_void_p]
    from vegas_rs_b200 import ISING, run
    from vegas_rs_b200.machine import Machine
    h = C.c_void_p(lib.mock_gpu_create(10, 0, 2**64 - 1))
    m = Machine(types.SimpleNamespace(_h=h, model=ISING), lib=lib)
    lines, batches = [], []
    m.add_stat_sensor(lambda line, row: lines.append(row))
    m.add_observable_sensor(lambda *a: batches.append(a))
    m.set_group(run.group_reduce(dist, 0), 10 * world)
    m.set_thermostat(2.0)
    m.measure_for(6)
    q.put((rank, lines[0][2], batches[0][2], batches[0][5].tolist()))
    m.close(); lib.mock_gpu_destroy(h)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_group_machine_world2():
    """Both ranks' sensors see the SUM of the two slabs' per-step energies and State::len = the global site count."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 150
    procs = [ctx.Process(target=_group_machine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs: p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    want = [2 * (-0.5 * k + 2.0) for k in range(1, 7)]      # the scripted device: E(step k) = -k/2 + T on each rank
    for rank, mean_e, n, e in res:
        assert n == 20 and e == want and abs(mean_e - np.mean(want)) < 1e-12
