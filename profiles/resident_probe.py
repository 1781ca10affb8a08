"""Time per Monte Carlo step of small sc lattices: shared-memory-resident batch kernel vs launch-per-colour path.
usage (on the GPU box): python profiles/resident_probe.py  -> one line per (model, L)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegas_rs_b200 as vg


def us_per_step(g, batch, reps):
    g.step_async(batch, True); g.synchronize()
    g.timer_start()
    for _ in range(reps):
        g.step_async(batch, True)
    return g.timer_stop() * 1e3 / (batch * reps)


for model, name in ((vg.ISING, "ising"), (vg.HEISENBERG, "heis_f32")):
    for L in (4, 10, 12, 16, 20, 24):
        row = []
        for resident in (True, False):
            g = vg.GpuMetropolis(model, unitcell=vg.SC, size=(L, L, L), seed=1, force_general=True)
            g.set_tuning("resident_max", 16384 if resident else 0)
            if resident and not g.step_kernel.endswith("resident"):
                g.close(); row.append(float("nan")); continue
            g.randomize(); g.set_thermostat(4.5 if model == vg.ISING else 1.4)
            row.append(us_per_step(g, 2048 if resident else 256, 4))
            g.close()
        n = L ** 3
        print(f"{name} L={L} n={n} resident {row[0]:.3f} us/step ({n / row[0] / 1e3:.2f} G attempts/s)  "
              f"colour passes {row[1]:.3f} us/step ({n / row[1] / 1e3:.2f} G attempts/s)  speed-up {row[1] / row[0]:.1f}x", flush=True)
