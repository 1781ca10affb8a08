"""Per-batch step time of one workload over a few seconds, with NVML SM clock / power / throttle reasons beside it:
does a long run drift (power cap, thermal)?  usage: python profiles/drift_probe.py heis_fcc_384|heis3d_512|ising3d_1024 [seconds]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml as nv
import bench

name = sys.argv[1]
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
nv.nvmlInit()
dev = nv.nvmlDeviceGetHandleByIndex(0)
g, w = bench.make_handle(name, 0, 1, 0)
g.randomize(); g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
g.step_async(3, False); g.synchronize()
t_end = time.time() + seconds
k = 0
while time.time() < t_end:
    g.timer_start(); g.step_async(10, True); ms = g.timer_stop() / 10
    mhz = nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM)
    mem = nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_MEM)
    pw = nv.nvmlDeviceGetPowerUsage(dev) / 1e3
    tmp = nv.nvmlDeviceGetTemperature(dev, nv.NVML_TEMPERATURE_GPU)
    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(dev)
    if k % 8 == 0:
        print(f"{name} batch {k:4d} {ms:8.4f} ms/step  sm {mhz} MHz mem {mem} MHz  {pw:6.1f} W  {tmp} C  reasons 0x{rs:x}", flush=True)
    k += 1
g.close()
