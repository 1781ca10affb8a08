"""Where the end-to-end Ising step spends its time: python profiles/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
g, w = bench.make_handle("ising3d_1024", 0, 1, 0)
g.randomize(); g.set_thermostat(4.5)
n = g.n_sites
host = torch.empty(n, dtype=torch.int8, pin_memory=True); arr = host.numpy(); arr[:] = g.download()
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("upload   %.1f ms" % t(lambda: g.upload(arr)))
print("download %.1f ms" % t(lambda: g._check(g._lib.vegas_gpu_download_ising(g._h, arr.ctypes.data, n))))
print("step(1)  %.2f ms" % t(lambda: g.step(1)))
print("step_host %.1f ms" % t(lambda: g.step_host(arr)))
dev = torch.empty(n, dtype=torch.int8, device="cuda")
print("torch H2D %.1f ms, D2H %.1f ms" % (t(lambda: dev.copy_(host)), t(lambda: host.copy_(dev))))
pag = np.empty(n, np.int8)
print("upload pageable %.1f ms" % t(lambda: g.upload(pag)))
