"""heis3d_512 ms/step of the persistent wave kernel by (steps per launch, planes per chunk, lag); usage: python profiles/wave_k_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(1, 4, 5), (2, 4, 5), (2, 2, 4), (2, 2, 6), (2, 1, 6), (3, 2, 4), (4, 2, 4), (4, 1, 4), (2, 4, 3), (3, 1, 5)]
for k, planes, lag in combos:
    g, w = bench.make_handle("heis3d_512", 0, 1, 0)
    try:
        g.set_tuning("heis_wave_steps", k)
    except Exception as e:  # a library built before the key existed
        print("(no heis_wave_steps key)", end=" ")
    g.set_tuning("heis_wave_planes", planes); g.set_tuning("heis_wave_lag", lag)
    g.randomize(); g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
    g.step_async(12, False); g.synchronize()
    g.timer_start(); g.step_async(48, True); ms = g.timer_stop() / 48
    g.synchronize()
    print(f"steps/launch {k} planes {planes} lag {lag}: {ms:.4f} ms/step  {g.n_sites / ms / 1e6:.1f} G attempts/s  "
          f"{g.n_sites / ms / 1e6 * 24 / 6553.3 * 100:.1f} % of 6553 GB/s", flush=True)
    g.close()
