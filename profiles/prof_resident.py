"""Driver for ncu: one recorded batch of the config[0] lattice on the resident kernel (python profiles/prof_resident.py [steps])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegas_rs_b200 as vg

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(10, 10, 10), seed=12345)
g.randomize(); g.set_thermostat(4.5)
g.step_async(steps, True); g.synchronize()
g.step_async(steps, True); g.synchronize()
e, m = g.read_observables(steps)
print(g.step_kernel, e[-1] / g.n_sites)
