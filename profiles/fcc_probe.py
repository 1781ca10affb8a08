"""ms per step of the fcc 384^3 heis_basis kernel (recorded / unrecorded steps); usage: python profiles/fcc_probe.py [tag]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegas_rs_b200 as vg

tag = sys.argv[1] if len(sys.argv) > 1 else ""
L = int(os.environ.get("FCC_L", "384"))
g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(L, L, L), seed=12345)
g.randomize(); g.set_thermostat(3.2)
g.step_async(3, False); g.synchronize()
out = []
for rec in (True, False):
    g.timer_start(); g.step_async(10, rec); ms = g.timer_stop() / 10
    out.append(f"{'recorded' if rec else 'unrecorded'} {ms:.3f} ms/step ({g.n_sites / ms / 1e6:.2f} G attempts/s)")
print(tag, g.step_kernel, " | ".join(out), flush=True)
g.close()
