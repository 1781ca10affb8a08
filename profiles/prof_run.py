"""Small driver for ncu: a few steps of one bench workload (python profiles/prof_run.py <workload> [steps])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g, w = bench.make_handle(name, 0, 1, 0)  # honours VEGAS_TUNE
g.randomize()
g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
g.step_async(steps, True)
g.synchronize()
e, m = g.read_observables(steps)
print(name, g.kernel_family, e[-1] / g.n_sites)
