"""fcc: whole lattice vs two in-process slabs on ONE GPU (separates the cost of the SLAB kernel variant from NVLink)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegas_rs_b200 as vg
L = 256
def run(handles, steps=10):
    for g in handles: g.randomize(); g.set_thermostat(3.2)
    if len(handles) > 1:
        for r, g in enumerate(handles): g.slab_connect_local(handles[(r - 1) % len(handles)], handles[(r + 1) % len(handles)])
    for g in handles: g.step_async(2, True)
    for g in handles: g.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for g in handles: g.step_async(1, True)
    for g in handles: g.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
whole = [vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(L, L, L), seed=1)]
print("whole   %.3f ms/step" % run(whole)); whole[0].close()
slabs = [vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(L, L, L // 2), nz_global=L, z_offset=r * L // 2, seed=1) for r in range(2)]
print("2 slabs %.3f ms/step (same GPU, sequential kernels)" % run(slabs))
