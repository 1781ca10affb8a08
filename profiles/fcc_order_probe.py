"""Does the fcc 384^3 step time depend on what the process allocated before?  (bench.py runs it after three other
workloads and measured 4.29 ms/step; a fresh process measures 3.42 ms/step.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench


def fcc(tag):
    g, w = bench.make_handle("heis_fcc_384", 0, 1, 0)
    g.randomize(); g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
    g.step_async(5, False); g.synchronize()
    g.timer_start(); g.step_async(10, True); ms = g.timer_stop() / 10
    free, total = torch.cuda.mem_get_info()
    print(f"{tag}: fcc {ms:.3f} ms/step   (device memory free {free / 2**30:.1f} GiB)", flush=True)
    g.close()


torch.cuda.set_device(0)
fcc("fresh process")
fcc("second handle")
for name in ("ising3d_1024", "heis3d_512"):
    g, w = bench.make_handle(name, 0, 1, 0)
    g.randomize(); g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
    g.step_async(20, True); g.synchronize()
    g.close()
fcc("after ising3d_1024 + heis3d_512 handles (no host round trip)")
g, w = bench.make_handle("heis3d_512", 0, 1, 0)
g.randomize(); g.set_thermostat(1.0)
host = torch.empty((g.n_sites, 3), dtype=torch.float64, pin_memory=True)
arr = host.numpy(); g.download_into(arr); g.step_host(arr); g.close()
del host, arr
fcc("after a 3 GiB host round trip (stream-ordered pool keeps its blocks)")
