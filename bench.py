#!/usr/bin/env python
"""bench.py -- spin-flip attempts per second of the B200 Metropolis sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" is one Monte Carlo step (N attempts = Integrator::step, src/integrator.rs:66-138) with the
per-step energy / magnetisation observers fused in (src/instrument.rs:133-141).  Workloads (BASELINE.json configs):
  ising3d_1024  cfg[2]  Ising sc 1024^3 pbc, T=4.5 (default: > L2, and the z-slab config; weak scaling: 1024^3 per GPU)
  ising2d_8192  cfg[1]  Ising sc 8192^2 pbc, T=2.269 (8 MiB bit-packed: L2 resident, reported in "also")
  heis3d_512    cfg[3]  Heisenberg sc 512^3 pbc, Exchange+Anisotropy+Zeeman, T=1.0, |H|=1, fp32
  heis_fcc_384  cfg[4]  Heisenberg fcc 384^3 cells (226 M sites, z = 12), T=3.2, fp32 (z-slabs of 384 cell planes per GPU)
  ising_sc10_cfg0 cfg[0] docs/metropolis.toml's 10^3 lattice on the shared-memory-resident kernel (inside "also"; main
                        workload of --impl reference only)
Prints ONE JSON line on rank 0: the bench contract's keys plus roofline, cpu_baseline, e2e (Integrator::step's own
signature: host State in / out every step), e2e_machine (Machine::measure_for with the State resident), clocks,
gpu_launches, and "also" = the other workloads with their own roofline / e2e / clocks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "ising3d_1024": dict(model="ising", size=(1024, 1024, 1024), pbc=(True, True, True), T=4.5, H=0.0, bytes_per_attempt=0.25,
                         dtype="u32 bit-packed (1 bit/spin)", cpu_L=(128, 128, 128)),
    "ising2d_8192": dict(model="ising", size=(8192, 8192, 1), pbc=(True, True, False), T=2.269185314213022, H=0.0,
                         bytes_per_attempt=0.25, dtype="u32 bit-packed (1 bit/spin)", cpu_L=(1024, 1024, 1)),
    "heis3d_512": dict(model="heisenberg", size=(512, 512, 512), pbc=(True, True, True), T=1.0, H=1.0, bytes_per_attempt=24.0,
                       dtype="f32", cpu_L=(128, 128, 128), anisotropy=((0.0, 0.0, 1.0), 0.1)),
    # cfg[4]: fcc (4 sites per cell, z = 12), basis 4-colouring, heis_basis kernel (basis-split SoA, compile-time
    # neighbour table); z-slabs of 384 cell planes per GPU at N > 1 (weak scaling)
    "heis_fcc_384": dict(model="heisenberg", unitcell="fcc", size=(384, 384, 384), pbc=(True, True, True), T=3.2, H=0.0,
                         bytes_per_attempt=24.0, dtype="f32", cpu_L=(48, 48, 48)),
}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the committed `ncu --set full`
# captures (cold cache): (bytes, source)
NCU_TRAFFIC = {
    "ising3d_1024": {"ising_msc": (134.37e6 + 30.64e6, "profiles/r02_ising_msc.metrics.txt (one colour pass)")},
    "heis3d_512": {"heis_pipe": (1.914e9 + 1.534e9, "profiles/r02_heis_pipe.metrics.txt (one step = both colours, one launch)"),
                   "heis_wave": (2.887e9 + 1.562e9, "profiles/r01u_heis_wave.metrics.txt (one step = both colours)"),
                   "heis_stencil": (2.42e9, "profiles/r01o_heis_stencil.metrics.txt (one colour pass)")},
    "heis_fcc_384": {"heis_basis": (2.829e9 + 0.686e9, "profiles/r01z_heis_basis_vec.metrics.txt (one colour pass)"),
                     "basis_pair": (2.989e9 + 1.357e9, "profiles/r02_basis_pair.metrics.txt (one pair launch = two colours)"),
                     "basis_pipe": (7.813e9 + 2.714e9, "profiles/r02_basis_pipe.metrics.txt (one step = four colours, one launch)"),
                     "basis_wave": (6.068e9 + 3.168e9, "profiles/r02_basis_wave.metrics.txt (one step = four colours, one launch)")},
}
# step kernels that do every colour of a step in ONE launch
ONE_LAUNCH_KERNELS = ("heis_pipe", "basis_pipe", "basis_wave", "heis_wave", "heis_fused")
# the CPU arm always runs this many single-threaded replicas (or every core of a smaller host), so that the driver's
# GPU / reference ratio means the same thing on every box
REFERENCE_REPLICAS = 16
# cfg[0] (docs/metropolis.toml, the reference's own CPU-runnable case): 1000 sites, latency bound; runs on the
# shared-memory-resident kernel (one launch per batch of steps), reported in "also" without a roofline
SMALL_WORKLOADS = {
    "ising_sc10_cfg0": dict(model="ising", size=(10, 10, 10), pbc=(True, True, True), T=4.5, H=0.0, cpu_L=(10, 10, 10),
                            dtype="int8 in shared memory"),
}
ALL_WORKLOADS = {**WORKLOADS, **SMALL_WORKLOADS}
POLL_S = float(os.environ.get("VEGAS_BENCH_POLL_MS", "1")) * 1e-3  # pause between NVML polls of the clock sampler (back-to-back
# polling cost the fcc workload 1.7 %: profiles/r01v_poll_probe.txt)
METRIC = "spin-flip attempts/sec"
UNIT = "attempts/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError, TypeError):
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Polls NVML (SM clock, max clock, power, throttle reasons) as fast as it can while the timed region runs;
    nvidia-smi's own polling (>= 100 ms) is too coarse for a region of a few milliseconds."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.ready = index, [], False, threading.Event()
        self.mark0 = self.mark1 = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self.stop_flag:
                t = time.perf_counter()
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append((t, mhz, [k for k, b in bits.items() if mask & b]))
                self.ready.set()
                if POLL_S > 0:
                    time.sleep(POLL_S)
        except Exception as e:  # no NVML: report nothing rather than a wrong number
            self.error = repr(e)
            self.ready.set()

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2.0)
        inside = [r for r in self.rows if self.mark0 is not None and self.mark0 <= r[0] <= self.mark1] or self.rows[-3:]
        sm = [r[1] for r in inside]
        reasons = sorted({x for r in inside for x in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "max_mhz", None),
                "reasons": reasons, "samples": len(sm), "source": "NVML polled during the timed region"}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_replica(args):
    """One single-threaded run of the oracle port of the reference Metropolis (machine.rs:91-101 loop
    with StatSensor + ObservableSensor as `vegas run` configures them, input.rs:324-345): the model is built once,
    then `samples` timed samples of `steps` Monte Carlo steps each.  Returns [(attempts, seconds)] per sample."""
    name, steps, samples, seed = args
    from oracle import binding as ob
    w = ALL_WORKLOADS[name]
    L = w["cpu_L"]
    model = ob.ISING if w["model"] == "ising" else ob.HEISENBERG
    lat = ob.Lattice(ob.FCC if w.get("unitcell") == "fcc" else ob.SC, *L, pbc=w["pbc"])
    csr = ob.Csr.from_lattice(lat, 1.0, False)
    terms = [ob.TERM_EXCHANGE, ob.TERM_ZEEMAN]
    kw = {}
    if "anisotropy" in w:
        terms.append(ob.TERM_ANISOTROPY); kw = dict(aniso_axis=w["anisotropy"][0], aniso_k=w["anisotropy"][1])
    H = ob.Hamiltonian(model, terms, csr, **kw)
    rng = ob.OracleRng(seed)
    n = cpu_sites(name)
    state = H.rand_state(rng, n)
    m = ob.Machine(H, ob.PROPOSE_FLIP if model == ob.ISING else ob.PROPOSE_RANDOM, rng, state, n_sensors=2)
    m.set_thermostat(H.thermostat(w["T"], (0, 0, 1.0), w["H"]))
    m.relax_for(1)  # warm caches
    out = []
    for _ in range(samples):
        t0 = time.perf_counter()
        m.measure_for(steps)
        out.append((n * steps, time.perf_counter() - t0))
    return out


def cpu_run(name: str, steps: int, replicas: int, samples: int = 1):
    """`replicas` independent single-threaded copies (the reference has no intra-run parallelism); per sample:
    (attempts of all replicas / slowest replica's time, that time)."""
    if replicas == 1:
        res = [cpu_replica((name, steps, samples, 12345))]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(replicas) as pool:
            res = pool.map(cpu_replica, [(name, steps, samples, 12345 + r) for r in range(replicas)])
    out = []
    for k in range(samples):
        attempts = sum(r[k][0] for r in res)
        wall = max(r[k][1] for r in res)
        out.append((attempts / wall, wall))
    return out


def cpu_sites(name: str) -> int:
    L = ALL_WORKLOADS[name]["cpu_L"]
    return L[0] * L[1] * L[2] * (4 if ALL_WORKLOADS[name].get("unitcell") == "fcc" else 1)


def cpu_steps_for(name: str, budget_s: float) -> int:
    n = cpu_sites(name)
    return max(1, int(budget_s * (6e6 if n <= 4096 else 1.5e6) / n))  # ~1.5e6 attempts/s/core out of cache with two sensors


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    cores = min(os.cpu_count() or 1, REFERENCE_REPLICAS)
    # every timed "step" is a bounded sample of the workload; the whole --steps/--warmup run stays within ~3 minutes
    budget = max(0.25, min(6.0, 150.0 / max(1, args.warmup + args.steps)))
    per_step = cpu_steps_for(name, budget)
    vals = cpu_run(name, per_step, cores, args.warmup + args.steps)[args.warmup:]
    value = float(np.mean([v for v, _ in vals]))
    w = ALL_WORKLOADS[name]
    sample = (f"{cores} independent single-threaded replicas (the reference has no intra-run parallelism) of the oracle port, "
              f"{'fcc' if w.get('unitcell') == 'fcc' else 'sc'} {w['cpu_L']} sub-lattice of the workload (the reference CSR "
              f"layout, 16 B/nnz, cannot hold the full size), "
              f"{per_step} MC steps per timed step, StatSensor+ObservableSensor per-step E/M as `vegas run`")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean([wl for _, wl in vals]) * 1e3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": name, "lattice": list(w["cpu_L"]), "cores": cores, "host_cores": os.cpu_count(),
                       "note": "CPU oracle port of vegas-rs 0.9.0 Metropolis (Rust toolchain absent); the reference is single-threaded: "
                               f"{cores} independent replicas (fixed at {REFERENCE_REPLICAS} so that the ratio does not depend on the host), "
                               "per-core rate = value / cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def make_handle(name: str, rank: int, world: int, device: int, seed: int = 12345):
    import vegas_rs_b200 as vg
    w = WORKLOADS[name]
    size = list(w["size"])
    kw = dict(unitcell=vg.FCC if w.get("unitcell") == "fcc" else vg.SC, pbc=w["pbc"], seed=seed, device=device)
    if w["model"] == "ising":
        model = vg.ISING
    else:
        model = vg.HEISENBERG
        kw.update(precision=vg.F32, anisotropy=w.get("anisotropy"))
    if world > 1 and size[2] > 1:
        # weak scaling: every rank owns a full-size slab of a lattice that is `world` times taller
        kw.update(nz_global=size[2] * world, z_offset=size[2] * rank)
    else:
        kw["seed"] = seed + rank  # independent replicas (2D: one temperature point per GPU)
    g = vg.GpuMetropolis(model, size=tuple(size), **kw)
    for kv in filter(None, os.environ.get("VEGAS_TUNE", "").split(",")):  # tuning experiments, e.g. heis_fused=0
        k, v = kv.split("=")
        g.set_tuning(k, int(v))
    return g, w


def run_workload(name: str, steps: int, warmup: int, rank: int, world: int, device: int, dist, torch, e2e_steps: int,
                 machine_e2e: bool = True):
    g, w = make_handle(name, rank, world, device)
    slab = world > 1 and w["size"][2] > 1
    g.randomize()
    g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
    if slab:
        from vegas_rs_b200 import distributed as vd
        vd.connect_slabs(g, dist)
    n_local = g.n_sites
    step_kernel = g.step_kernel
    # ---- device-resident throughput: K steps, fused E/M on, CUDA events on the sweep stream
    g.step_async(warmup, False)
    g.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(device) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.ready.wait(5.0)
    if world > 1:
        dist.barrier()  # every rank enters the timed region together (slab neighbours spin on each other's flags)
    torch.cuda.synchronize()
    l0 = g.launches
    if sampler:
        sampler.mark0 = time.perf_counter()
    g.timer_start()
    RING = 4096   # rows of the device observable ring: longer runs record batch after batch (the last batch is read back)
    for first in range(0, steps, RING):
        g.step_async(min(RING, steps - first), True)
    ms = g.timer_stop()
    if sampler:
        sampler.mark1 = time.perf_counter()
    launches = g.launches - l0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms], device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.finish() if sampler else None
    n_read = steps - (steps - 1) // 4096 * 4096   # steps of the last recorded batch
    e_series, m_series = g.read_observables(n_read)
    if world > 1 and slab:
        t = torch.tensor(np.concatenate([e_series, m_series.ravel()]), device=f"cuda:{device}")
        dist.all_reduce(t)  # per-step scalars: sum of the slab partials
        e_series = t[:n_read].cpu().numpy()
    value = n_local * world * steps / (ms * 1e-3)
    # ---- end to end: Integrator::step's own signature, host State in -> host State out, pinned buffers
    e2e = None
    if e2e_steps > 0:
        if w["model"] == "ising":
            host = torch.empty(n_local, dtype=torch.int8, pin_memory=True)
        else:
            host = torch.empty((n_local, 3), dtype=torch.float64, pin_memory=True)
        arr = host.numpy(); g.download_into(arr)

        def host_step():
            if not slab:
                g.step_host(arr)  # Integrator::step: host State in -> host State out
            else:  # a slab's upload pushes its boundary planes to the neighbours: all ranks must have uploaded
                g.upload(arr); dist.barrier(); g.step(1); g.download_into(arr); dist.barrier()

        host_step()  # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=f"cuda:{device}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        nbytes = int(g.state_transfer_bytes)   # what crosses PCIe: the host State itself, or its sign bitmap (host-packed Ising path)
        e2e = {"value": n_local * world * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes + 32, "steps": e2e_steps, "host_state_bytes": arr.nbytes,
               "transfer": ("sign bitmap, packed / unpacked by the host threads of the C ABI chunk by chunk, copies overlapped"
                            if nbytes != arr.nbytes else "the reference's host layout as it is"),
               "api": "vegas_gpu_step_host_* (host State in, host State out, E and M back)" if not slab else
                      "vegas_gpu_upload_* + vegas_gpu_step + vegas_gpu_download_* per slab"}
    # ---- the same steps through the host layer: Machine::measure_for with StatSensor + ObservableSensor fed from the
    # device-reduced per-step (E, M); the State stays in HBM (it only crosses PCIe for a StateSensor dump)
    e2e_machine = None
    if machine_e2e:
        from vegas_rs_b200.machine import Machine
        m = Machine(g)
        if slab:   # slab group: one Machine per rank, the per-step (E, M) partial sums all-reduced before the sensors see them
            from vegas_rs_b200 import run as vrun
            m.set_group(vrun.group_reduce(dist, device), n_local * world)
        m.add_stat_sensor(lambda line, row: None)
        m.add_observable_sensor(lambda *a: None)
        m.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
        m.measure_for(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.measure_for(steps)
        dt = time.perf_counter() - t0
        m.close()
        if world > 1:
            t = torch.tensor([dt], device=f"cuda:{device}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_machine = {"value": n_local * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64,
                       "steps": steps, "api": "vegas_machine_measure_for (Machine::measure_for, src/machine.rs:116-125) with "
                                              "StatSensor + ObservableSensor; host wall clock" +
                                              ("; slab group (vegas_machine_set_group): per-step partial sums all-reduced over the ranks" if slab else "")}
    peak, peak_src = peaks()
    # sweep launches per step: one per colour, or ONE for the persistent wave kernel (both colours); a connected slab
    # adds wait / signal / boundary launches, so count colours there
    passes = 1 if step_kernel in ONE_LAUNCH_KERNELS else (g.n_colours if slab else max(1, round(launches / steps)))
    per_launch_s = ms * 1e-3 / (passes * steps)
    alg_bytes_per_launch = w["bytes_per_attempt"] * n_local / passes
    achieved = alg_bytes_per_launch / per_launch_s / 1e9
    traffic, traffic_src = NCU_TRAFFIC.get(name, {}).get(step_kernel, (None, None))
    if traffic is not None and slab:
        traffic_src += "; single-GPU capture of the same kernel (halo planes add " + ("1/%d" % w["size"][2]) + " of it per neighbour)"
    res = {"value": value, "ms_per_step": ms / steps, "launches": launches, "clocks": clocks, "e2e": e2e, "e2e_machine": e2e_machine,
           "family": step_kernel, "n_local": n_local,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_source": traffic_src,
                        "algorithmic_bytes_per_launch": alg_bytes_per_launch, "peak_source": peak_src, "kernel": f"{step_kernel} " + ("step (all colours, one launch)" if passes == 1 else "colour pass"),
                        "algorithmic_bytes_per_attempt": w["bytes_per_attempt"]},
           "energy_per_site_last": float(e_series[-1] / (n_local * (world if slab else 1)))}
    g.close()
    return res


def run_small_workload(name: str, device: int, torch, with_cpu: bool):
    """cfg[0]-sized lattice: batches of 4096 recorded steps on the shared-memory-resident kernel (one launch per batch),
    then the same through Machine::measure_for, next to the oracle port on ONE host core on the SAME lattice."""
    import vegas_rs_b200 as vg
    from vegas_rs_b200.machine import Machine
    w = SMALL_WORKLOADS[name]
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=w["size"], pbc=w["pbc"], seed=12345, device=device)
    g.randomize()
    g.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
    batch, reps = 4096, 8
    g.step_async(batch, True)
    g.synchronize()
    l0 = g.launches
    g.timer_start()
    for _ in range(reps):
        g.step_async(batch, True)
    ms = g.timer_stop()
    launches = g.launches - l0
    family = g.step_kernel
    m = Machine(g)
    m.add_stat_sensor(lambda line, row: None)
    m.add_observable_sensor(lambda *a: None)
    m.set_thermostat(w["T"], (0.0, 0.0, 1.0), w["H"])
    m.measure_for(batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.measure_for(batch * reps)
    dt = time.perf_counter() - t0
    m.close()
    n = g.n_sites
    g.close()
    res = {"value": n * batch * reps / (ms * 1e-3), "unit": UNIT, "us_per_step": ms * 1e3 / (batch * reps), "gpu_launches": launches,
           "steps": batch * reps, "family": family, "roofline": None,
           "e2e": {"value": n * batch * reps / dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64,
                   "api": "vegas_machine_measure_for with StatSensor + ObservableSensor; host wall clock"},
           "note": "docs/metropolis.toml lattice (1000 sites): latency bound, one CTA, State resident in shared memory; no HBM roofline"}
    if with_cpu:
        k = cpu_steps_for(name, 3.0)
        v, wall = cpu_run(name, k, 1)[0]
        res["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"oracle port, the same sc {w['size']} lattice, {k} MC steps, two sensors, {wall:.1f} s"}
    # the WHOLE documented program (Relax 1000 + CoolDown 6 -> 1 in 101 points of 1000 + 20000 steps, all three sensors,
    # parquet outputs) through the TOML front end, host wall clock; next to it what the reference's own execution model
    # (one thread) needs for the same number of attempts at the rate just measured
    try:
        import io
        import tempfile
        from vegas_rs_b200 import run as vrun
        with open(os.path.join(ROOT, "tests", "golden", "cfg0_ising_sc10.toml")) as f:
            cfg = vrun.parse_input(f.read())
        with tempfile.TemporaryDirectory() as td:
            cfg["output"]["observables"] = os.path.join(td, "output.parquet")
            cfg["output"]["state"]["path"] = os.path.join(td, "state.parquet")
            out = io.StringIO()
            t0 = time.perf_counter()
            vrun.run_input(cfg, seed=12345, out=out, device=device)
            wall_gpu = time.perf_counter() - t0
        from vegas_rs_b200.distributed import cooldown_temperatures
        points = len([ln for ln in out.getvalue().splitlines() if ln.strip()])
        stages = cfg["stages"]
        steps_total = sum(st["steps"] for st in stages if st["program"] == "Relax") + \
            sum(len(cooldown_temperatures(st["max_temperature"], st["min_temperature"], st["cool_rate"])) * (st["relax"] + st["steps"])
                for st in stages if st["program"] == "CoolDown")
        res["program"] = {"input": "tests/golden/cfg0_ising_sc10.toml (= docs/metropolis.toml)", "stat_lines": points, "mc_steps": steps_total,
                          "wall_s": wall_gpu, "api": "python -m vegas_rs_b200.run (TOML -> Machine -> Relax + CoolDown, StatSensor + ObservableSensor + StateSensor, parquet sinks)"}
        if "cpu_baseline" in res:
            res["program"]["cpu_wall_s_at_sample_rate"] = steps_total * n / res["cpu_baseline"]["value"]
    except Exception as e:  # the program leg never costs the line
        res["program"] = {"error": repr(e)[:300]}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ising3d_1024", choices=sorted(ALL_WORKLOADS))
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload in SMALL_WORKLOADS:
        raise SystemExit(f"bench.py: {args.workload} is reported inside \"also\" (and by --impl reference); pick a BASELINE workload")
    import torch
    import torch.distributed as dist
    from vegas_rs_b200 import _lib
    _lib.load()  # fail loudly if the CUDA extension is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sweep has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{device}"))
    main_res = run_workload(args.workload, args.steps, args.warmup, rank, world, device, dist, torch, args.e2e_steps,
                            args.e2e_steps > 0)
    also = {}
    if not args.no_also:
        for other in WORKLOADS:
            if other != args.workload:
                heavy = "unitcell" in WORKLOADS[other]   # 226 M sites: fewer steps, no 5 GB host round trip
                try:
                    r = run_workload(other, min(args.steps, 10) if heavy else args.steps, args.warmup, rank, world, device, dist, torch,
                                     0 if heavy else args.e2e_steps, args.e2e_steps > 0)
                except Exception as e:  # a secondary workload must never cost the headline line (single process only:
                    if world > 1:       # with several ranks a one-sided failure would desynchronise the collectives)
                        raise
                    also[other] = {"error": repr(e)}
                    continue
                also[other] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "roofline": r["roofline"],
                               "e2e": r["e2e"], "e2e_machine": r["e2e_machine"], "family": r["family"], "clocks": r["clocks"],
                               "gpu_launches": r["launches"],
                               "note": {"ising2d_8192": "8 MiB state is L2 resident: not an HBM measurement" + ("; independent replicas, one per GPU" if world > 1 else ""),
                                        "heis3d_512": "z-slabs of 512 planes per GPU" if world > 1 else "",
                                        "heis_fcc_384": "heis_basis vector kernel (16-byte loads, one Philox call per site), 4 colour passes per step" +
                                                        ("; z-slabs of 384 cell planes per GPU" if world > 1 else "")}.get(other, "")}
        if world == 1 and args.e2e_steps > 0:
            for small in SMALL_WORKLOADS:
                try:
                    also[small] = run_small_workload(small, device, torch, not args.no_cpu)
                except Exception as e:
                    also[small] = {"error": repr(e)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        k = cpu_steps_for(args.workload, 12.0)
        v, wall = cpu_run(args.workload, k, 1)[0]
        w = WORKLOADS[args.workload]
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"oracle port of the reference Metropolis, sc {w['cpu_L']} sub-lattice, {k} MC steps, "
                         f"StatSensor+ObservableSensor per-step E/M, {wall:.1f} s"}
    if rank == 0:
        w = WORKLOADS[args.workload]
        size = list(w["size"])
        if world > 1 and size[2] > 1:
            size[2] *= world
        line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
                "config": {"workload": args.workload, "lattice": size, "pbc": list(w["pbc"]), "temperature": w["T"], "field": w["H"],
                           "decomposition": ("z-slabs, peer-written halos" if world > 1 and w["size"][2] > 1 else
                                             ("independent replicas" if world > 1 else "single GPU")),
                           "observers": "energy+magnetisation fused in the last colour pass, every step",
                           "l2": "state (2 x 64 MiB colour arrays) exceeds L2; no flush" if args.workload == "ising3d_1024" else "see note"},
                "roofline": main_res["roofline"], "cpu_baseline": cpu, "e2e": main_res["e2e"], "e2e_machine": main_res["e2e_machine"], "gpu_launches": main_res["launches"],
                "clocks": main_res["clocks"], "kernel_family": main_res["family"], "also": also}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
