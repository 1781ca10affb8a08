"""Multi-process slab-GROUP Machine check (needs >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mp_machine_check.py [ising|heisenberg]
The TOML front end runs Relax + CoolDown (src/program.rs:97-115, :182-214) on a lattice cut into one z-slab per rank,
every rank through its own Machine with the per-step (E, M) partial sums all-reduced (vegas_machine_set_group); rank 0
then runs the SAME input on one GPU.  Same Philox keys => the StatSensor lines must be identical for Ising (integer
sums) and agree to 1e-6 relative for Heisenberg (per-step E and M are f64 sums of fp32 per-thread partials, grouped
differently by the slabs; the fp32 bar of the observables is 1e-5)."""
import io
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vegas_rs_b200 import run

TOML = """
model = "{model}"
algorithm = "Metropolis"
[sample]
unitcell = {{ name = "sc" }}
size = {{ x = 64, y = 32, z = {nz} }}
pbc = {{ x = true, y = true, z = true }}
[[stages]]
program = "Relax"
steps = 200
temperature = {t0}
[[stages]]
program = "CoolDown"
max_temperature = {tmax}
min_temperature = {tmin}
cool_rate = {rate}
relax = 100
steps = 500
"""


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "ising"
    rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    text = TOML.format(model="Ising" if kind == "ising" else "Heisenberg", nz=32 * world,
                       **(dict(t0=6.0, tmax=5.0, tmin=4.0, rate=0.5) if kind == "ising" else dict(t0=2.5, tmax=1.8, tmin=1.2, rate=0.3)))
    cfg = run.parse_input(text)
    out = io.StringIO()
    run.run_input(cfg, seed=2026, out=out, device=dev, dist=dist)
    ok = True
    if rank == 0:
        group_lines = out.getvalue().strip().split("\n")
        single = io.StringIO()
        run.run_input(cfg, seed=2026, out=single, device=dev)
        single_lines = single.getvalue().strip().split("\n")
        if kind == "ising":
            ok = group_lines == single_lines
        else:
            a = [[float(x) for x in ln.split()] for ln in group_lines]
            b = [[float(x) for x in ln.split()] for ln in single_lines]
            ok = len(a) == len(b) and all(abs(x - y) <= 1e-6 * max(1.0, abs(y)) for ra, rb in zip(a, b) for x, y in zip(ra, rb))
        print(f"mp_machine_check model={kind} world={world} lines={len(group_lines)} first={group_lines[0][:60]!r} "
              f"identical={group_lines == single_lines} -> {'OK' if ok else 'FAIL'}", flush=True)
        if not ok:
            for x, y in zip(group_lines, single_lines):
                print("  group :", x); print("  single:", y)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
