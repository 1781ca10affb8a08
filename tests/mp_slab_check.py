"""Multi-process z-slab check (needs >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mp_slab_check.py [ising|heisenberg|bcc|fcc]
Every rank owns one slab on its own GPU, halos travel over CUDA IPC peer memory; rank 0 also runs the whole
lattice on one handle and the states must agree bit for bit (same Philox keys)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegas_rs_b200 as vg
from vegas_rs_b200 import distributed as vd


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "ising"
    model = vg.ISING if kind.startswith("i") else vg.HEISENBERG
    uc, nb = (vg.FCC, 4) if kind == "fcc" else ((vg.BCC, 2) if kind == "bcc" else (vg.SC, 1))
    rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    Lx, Ly, Lz = (256, 64, 16 * world) if model == vg.ISING else ((64, 32, 16 * world) if nb == 1 else (24, 20, 8 * world))
    nz, zoff = vd.slab_extent(Lz, rank, world)
    rng = np.random.default_rng(7)
    if model == vg.ISING:
        full = (2 * rng.integers(0, 2, Lx * Ly * Lz) - 1).astype(np.int8)
    else:
        v = rng.normal(size=(Lx * Ly * Lz * nb, 3)); full = v / np.linalg.norm(v, axis=1, keepdims=True)
    plane = Lx * Ly * nb
    g = vg.GpuMetropolis(model, unitcell=uc, size=(Lx, Ly, nz), nz_global=Lz, z_offset=zoff, seed=11, device=dev)
    g.upload(full[zoff * plane:(zoff + nz) * plane])
    g.set_thermostat(2.0 if model == vg.HEISENBERG else 4.0, (0, 0, 1.0), 0.25)
    for kv in filter(None, os.environ.get("VEGAS_TUNE", "").split(",")):   # e.g. heis_pipe=1: force a kernel on the SLAB handles only
        k, v = kv.split("=")
        g.set_tuning(k, int(v))
    vd.connect_slabs(g, dist)
    slab_kernel = g.step_kernel
    steps = 6
    g.step_async(steps, True)
    e, m = g.read_observables(steps)
    e, m = vd.reduce_observables(e, m, dist, f"cuda:{dev}")
    mine = g.download()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    ok = True
    if rank == 0:
        whole = vg.GpuMetropolis(model, unitcell=uc, size=(Lx, Ly, Lz), seed=11, device=dev)
        whole.upload(full)
        whole.set_thermostat(2.0 if model == vg.HEISENBERG else 4.0, (0, 0, 1.0), 0.25)
        e_ref, m_ref = whole.step(steps)
        ref = whole.download()
        got = np.concatenate(parts)
        ok = np.array_equal(ref, got)
        tol = 0 if model == vg.ISING else 1e-6 * abs(e_ref[-1]) + 1e-6
        ok = ok and abs(e[-1] - e_ref[-1]) <= tol
        print(f"mp_slab_check model={kind} family={whole.kernel_family} slab_kernel={slab_kernel} world={world} identical_state={np.array_equal(ref, got)} "
              f"E_slabs={e[-1]:.6f} E_single={e_ref[-1]:.6f} -> {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    g.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
