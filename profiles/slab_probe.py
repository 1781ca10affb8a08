"""Timing probe for the z-slab path (python profiles/slab_probe.py [Lz_per_gpu]); run on a 2-GPU box.
Single process, two handles: (a) both on GPU 0, (b) on GPU 0 and GPU 1 with peer access."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vegas_rs_b200 as vg

Lz = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = 1024
steps = 10

def run(devs, label):
    n = len(devs)
    hs = [vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, Lz), nz_global=Lz * n, z_offset=Lz * r, seed=5, device=d)
          for r, d in enumerate(devs)]
    for h in hs:
        h.randomize(); h.set_thermostat(4.5)
    for r, h in enumerate(hs):
        h.slab_connect_local(hs[(r - 1) % n], hs[(r + 1) % n])
    for h in hs:
        h.randomize()
    for _ in range(3):
        for h in hs: h.step_async(1, True)
    for h in hs: h.synchronize()
    t0 = time.perf_counter()
    for h in hs: h.timer_start()
    for _ in range(steps):
        for h in hs: h.step_async(1, True)
    ms = [h.timer_stop() for h in hs]
    wall = time.perf_counter() - t0
    print(label, "ms/step per handle", [m / steps for m in ms], "wall ms/step", wall / steps * 1e3, flush=True)
    for h in hs: h.close()

g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, Lz), seed=5, device=0)
g.randomize(); g.set_thermostat(4.5); g.step_async(3, True); g.synchronize()
g.timer_start(); g.step_async(steps, True); print("single handle, no slab: ms/step", g.timer_stop() / steps, flush=True); g.close()
run([0, 0], "2 slabs on GPU0")
import torch
if torch.cuda.device_count() > 1:
    run([0, 1], "2 slabs on GPU0+GPU1 (peer access)")
