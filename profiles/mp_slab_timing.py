"""torchrun timing probe for the IPC slab path: python -m torch.distributed.run ... profiles/mp_slab_timing.py Lz_per_gpu"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegas_rs_b200 as vg
from vegas_rs_b200 import distributed as vd

rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
Lz = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(1024, 1024, Lz), nz_global=Lz * world, z_offset=Lz * rank, seed=5, device=dev)
g.randomize(); g.set_thermostat(4.5)
vd.connect_slabs(g, dist)
for rep in range(3):
    g.step_async(3, True); g.synchronize(); dist.barrier()
    for n in (1, 10):
        t0 = time.perf_counter()
        g.timer_start(); g.step_async(n, True); ms = g.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        print(f"rank {rank} rep {rep} steps {n}: device {ms / n:.3f} ms/step, host wall {wall / n:.3f} ms/step", flush=True)
        dist.barrier()
g.close(); dist.destroy_process_group()
