"""CPU-only tests of the C++ host layer (vegas_rs_b200/csrc/vegas_host.cpp: Machine, StatSensor / ObservableSensor /
StateSensor, Relax / CoolDown / HysteresisLoop) against the reference's semantics (src/machine.rs:91-125,
src/instrument.rs:61-351, src/program.rs:97-336).  The host layer is compiled together with a SCRIPTED test double of the
device handle (tests/mock/mock_vegas_gpu.cpp: step k reports fixed formulas, every call is logged) -- no GPU, no Monte
Carlo, nothing from the product package is replaced."""
import ctypes as C
import os
import subprocess
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_DIR = os.path.join(ROOT, "tests", "mock")
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def mock_lib():
    so = os.path.join(MOCK_DIR, "libvegas_host_mock.so")
    srcs = [os.path.join(MOCK_DIR, "mock_vegas_gpu.cpp"), os.path.join(ROOT, "vegas_rs_b200", "csrc", "vegas_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "vegas_rs_b200", "csrc", "vegas_host.hpp"), os.path.join(ROOT, "include", "vegas_host.h"),
                   os.path.join(ROOT, "include", "vegas_gpu.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", so, *srcs], check=True)
    lib = C.CDLL(so)
    lib.mock_gpu_create.restype, lib.mock_gpu_create.argtypes = C.c_void_p, [C.c_uint64, C.c_int, C.c_uint64]
    lib.mock_gpu_destroy.restype, lib.mock_gpu_destroy.argtypes = None, [C.c_void_p]
    lib.mock_gpu_log.restype, lib.mock_gpu_log.argtypes = C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.mock_gpu_steps.restype, lib.mock_gpu_steps.argtypes = C.c_uint64, [C.c_void_p]
    return lib


class Rig:
    """A Machine of the real host layer over the scripted device, with the three sensors recording what they receive."""

    def __init__(self, lib, n=10, heisenberg=False, fail_at=2**64 - 1, state_frequency=None, sensors=True):
        from vegas_rs_b200.machine import Machine
        from vegas_rs_b200 import ISING, HEISENBERG
        self.lib, self.n, self.heis = lib, n, heisenberg
        self.h = C.c_void_p(lib.mock_gpu_create(n, int(heisenberg), fail_at))
        self.m = Machine(types.SimpleNamespace(_h=self.h, model=HEISENBERG if heisenberg else ISING), lib=lib)
        self.lines, self.batches, self.dumps = [], [], []
        if not sensors:
            return
        self.m.add_stat_sensor(lambda line, row: self.lines.append((line, row)))
        self.m.add_observable_sensor(lambda *a: self.batches.append(a))
        if state_frequency is not None:
            self.m.add_state_sensor(state_frequency, lambda *a: self.dumps.append(a))

    def log(self):
        k = self.lib.mock_gpu_log(self.h, None, 0)
        out = np.zeros(k)
        self.lib.mock_gpu_log(self.h, out.ctypes.data_as(C.c_void_p), k)
        return out.reshape(-1, 4)

    def close(self):
        self.m.close()
        self.lib.mock_gpu_destroy(self.h)


def energy(k, T):
    return -0.5 * k + T


def magnitude(k, heis):
    if heis:
        a = float(k % 5)
        mag = np.sqrt((3 * a) ** 2 + (4 * a) ** 2)
        return 0.0 if abs(mag) < EPS else abs(mag)            # HeisenbergSpin::from_projections, src/state.rs:150-160
    return abs(float((k * 37) % 101) - 50.0)                  # IsingSpin::from_projections, src/state.rs:86-92


def test_machine_default_thermostat_and_hook_order(mock_lib):
    r = Rig(mock_lib)
    assert r.m.thermostat() == (2.8, 0.0)                     # Thermostat::new(2.8, Field::zero()), src/input.rs:274
    r.m.relax(5, 6.0)
    r.m.set_thermostat(3.0, (0, 0, 1.0), 0.25)
    r.m.measure_for(7)
    assert r.m.steps_done == 12 and mock_lib.mock_gpu_steps(r.h) == 12
    # ObservableSensor: one batch per stage, relax flag, stage counter over both kinds of stage (instrument.rs:229-252)
    assert [(b[0], b[1], b[2], b[3], b[4], len(b[5])) for b in r.batches] == [(True, 0, 10, 6.0, 0.0, 5), (False, 1, 10, 3.0, 0.25, 7)]
    assert np.array_equal(r.batches[0][5], [energy(k, 6.0) for k in range(1, 6)])
    assert np.array_equal(r.batches[1][6], [magnitude(k, False) for k in range(6, 13)])
    # StatSensor: measure stages only; population variance / (N T^2), / (N T); Binder cumulant of |M| (instrument.rs:98-131)
    assert len(r.lines) == 1
    e = np.array([energy(k, 3.0) for k in range(6, 13)]); m = np.array([magnitude(k, False) for k in range(6, 13)])
    row = r.lines[0][1]
    want = (3.0, 0.25, e.mean(), e.var() / (10 * 9.0), m.mean(), m.var() / (10 * 3.0), 1 - (m**4).mean() / (3 * (m**2).mean() ** 2))
    assert np.allclose(row, want, rtol=1e-13, atol=1e-13)
    assert r.lines[0][0] == " ".join(f"{v:.16f}" for v in row)                      # "{:.16} ..." line, instrument.rs:118-128
    # the device saw: thermostats in order, one recorded batch per stage
    log = r.log()
    assert [tuple(x) for x in log[log[:, 0] == 1][:, 1:3]] == [(2.8, 0.0), (6.0, 0.0), (3.0, 0.25)]
    assert [tuple(x) for x in log[log[:, 0] == 2][:, 1:]] == [(5, 1, 0), (7, 1, 5)]
    r.close()


def test_machine_chunks_long_stages_without_reordering_steps(mock_lib):
    """Stages longer than the device's observable ring (4096 rows) run as several batches; the sensors still see every
    step exactly once and in order (src/machine.rs:91-101)."""
    r = Rig(mock_lib, heisenberg=True, n=24)
    r.m.set_thermostat(1.5)
    r.m.measure_for(10000)
    steps = r.log()[r.log()[:, 0] == 2]
    assert [int(x) for x in steps[:, 1]] == [4096, 4096, 1808] and [int(x) for x in steps[:, 3]] == [0, 4096, 8192]
    (relax, stage, n, T, field, e, mag), = r.batches
    assert (relax, stage, n, T, len(e)) == (False, 0, 24, 1.5, 10000)
    assert np.array_equal(e, [energy(k, 1.5) for k in range(1, 10001)])
    assert np.array_equal(mag, [magnitude(k, True) for k in range(1, 10001)])     # |M| = norm of the three projections
    assert mag.min() == 0.0 and mag.max() == 20.0
    r.close()


def test_hooks_of_a_batch_are_replayed_while_the_next_batch_sweeps(mock_lib):
    """Without a StateSensor the machine launches batch k + 1 before it replays the hooks of batch k (the device sweeps
    while the host accumulates); the device call order shows it, and the sensors' view is unchanged."""
    r = Rig(mock_lib, n=16)
    r.m.set_thermostat(2.0)
    r.m.measure_for(9000)
    calls = [int(c) for c in r.log()[:, 0] if c in (2, 4)]           # 2 = step_async, 4 = read_observables
    assert calls == [2, 4, 2, 4, 2, 4]                                # launches and reads alternate (one device ring)
    (relax, stage, n, T, field, e, mag), = r.batches
    assert len(e) == 9000 and np.array_equal(e, [energy(k, 2.0) for k in range(1, 9001)])
    (line, row), = r.lines
    assert row[2] == np.mean([energy(k, 2.0) for k in range(1, 9001)]) or abs(row[2] - np.mean(e)) < 1e-9
    r.close()


def test_slab_group_machine_sums_the_partials_before_the_instruments(mock_lib):
    """vegas_machine_set_group: the per-step (E, Mx, My, Mz) of a batch are summed over the ranks (here: a scripted
    reduction that adds a second rank's partials) before StatSensor / ObservableSensor see them, and State::len is the
    global site count (src/instrument.rs:98-131 divides the variances by it)."""
    r = Rig(mock_lib, n=10)
    seen = []

    def reduce_sum(values):
        seen.append(len(values))
        k = len(values) // 4
        values[:k] += 100.0                      # the other slab's energy partial
        values[k:] *= 2.0                        # ... and a mirror-image magnetisation
    r.m.set_group(reduce_sum, 20)
    r.m.set_thermostat(3.0)
    r.m.relax_for(3)                             # nothing records during relax here? the ObservableSensor does
    r.m.measure_for(5000)
    assert seen == [4 * 3, 4 * 4096, 4 * 904]
    (_, _, n0, _, _, e0, m0), (relax, stage, n, T, field, e, mag) = r.batches
    assert n0 == n == 20
    assert np.array_equal(e, [energy(k, 3.0) + 100.0 for k in range(4, 5004)])
    assert np.array_equal(mag, [2.0 * magnitude(k, False) for k in range(4, 5004)])
    (line, row), = r.lines
    assert abs(row[3] - np.var(e) / (20 * 9.0)) < 1e-9 * abs(row[3])       # Cv = Var(E) / (N_global T^2)
    # a failing reduction surfaces as an error, with the Python exception that caused it
    r2 = Rig(mock_lib, n=10)

    def broken(values):
        raise RuntimeError("rank 1 went away")
    r2.m.set_group(broken, 20)
    with pytest.raises(RuntimeError, match="rank 1 went away"):
        r2.m.measure_for(2)
    r.close(); r2.close()


def test_python_exceptions_inside_sensor_callbacks_are_not_swallowed(mock_lib):
    r = Rig(mock_lib, n=8, sensors=False)

    def on_batch(*a):
        raise OSError("disk full")
    r.m.add_observable_sensor(on_batch)
    with pytest.raises(OSError, match="disk full"):
        r.m.measure_for(3)
    r.close()


@pytest.mark.parametrize("heis", [False, True], ids=["ising", "heisenberg"])
def test_state_sensor_schedule_and_contents(mock_lib, heis):
    """StateSensor (src/instrument.rs:265-351): a dump when step.is_multiple_of(frequency), the counter restarts every
    stage, the stage index counts relax and measure stages, and the dumped State is the one AFTER that step -- the
    Machine therefore cuts its device batches so that a due step is the last of its batch."""
    r = Rig(mock_lib, n=6, heisenberg=heis, state_frequency=4)
    r.m.relax(10, 2.0)
    r.m.measure_for(6)
    got = [(d[0], d[1], d[2]) for d in r.dumps]
    assert got == [(True, 0, 0), (True, 0, 4), (True, 0, 8), (False, 1, 0), (False, 1, 4)]
    # the device step count when each dump was taken: step index s of a stage = the state after s + 1 steps of it
    device_steps = [1, 5, 9, 11, 15]
    for d, k in zip(r.dumps, device_steps):
        state = d[5]
        want = np.array([1 if (k + i) & 1 else -1 for i in range(6)])
        if heis:
            assert state.shape == (6, 3) and np.array_equal(state[:, 2], want) and not state[:, :2].any()
        else:
            assert state.shape == (6,) and np.array_equal(state, want)
    log = r.log()
    assert [int(x) for x in log[log[:, 0] == 3][:, 1]] == device_steps
    # batches end on the due steps: 1 | 4 | 4 | 1 (relax 10) and 1 | 4 | 1 (measure 6)
    assert [int(x) for x in log[log[:, 0] == 2][:, 1]] == [1, 4, 4, 1, 1, 4, 1]
    # every step still reaches the observable sensor once
    assert [len(b[5]) for b in r.batches] == [10, 6]
    assert np.array_equal(np.concatenate([b[5] for b in r.batches]), [energy(k, 2.0) for k in range(1, 17)])
    r.close()


def test_state_sensor_frequency_zero_dumps_only_step_zero(mock_lib):
    r = Rig(mock_lib, n=4, state_frequency=0)
    r.m.relax(5, 1.0)
    r.m.relax(3, 1.0)
    assert [(d[1], d[2]) for d in r.dumps] == [(0, 0), (1, 0)]
    r.close()


def test_cooldown_program_sequence(mock_lib):
    """CoolDown::run (src/program.rs:182-214): T -= cool_rate in f64 until T < min; relax then measure at every point."""
    r = Rig(mock_lib)
    r.m.cooldown(3.0, 2.0, 0.1, 3, 5)
    temps, t = [], 3.0
    while True:
        temps.append(t)
        t -= 0.1
        if t < 2.0:
            break
    assert len(temps) == 10 and temps[-1] != 2.0 + 0.1          # the 2.0 point is lost to rounding
    assert [l[1][0] for l in r.lines] == temps
    assert [(b[0], b[3], len(b[5])) for b in r.batches] == [x for T in temps for x in ((True, T, 3), (False, T, 5))]
    assert r.m.steps_done == 10 * 8
    log = r.log()
    assert list(log[log[:, 0] == 1][:, 1]) == [2.8] + temps
    r.close()


def test_hysteresis_program_sequence(mock_lib):
    """HysteresisLoop::run (src/program.rs:281-336): 0 -> +max (overshooting by one step), down to -max (overshooting),
    back up; the thermostat receives the signed magnitude, the sensors report Field::magnitude() = |H|."""
    r = Rig(mock_lib, heisenberg=True)
    r.m.hysteresis(4, 2, 1.25, 1.0, 0.5)
    signed = [0.0, 0.5, 1.0, 1.5, 1.0, 0.5, 0.0, -0.5, -1.0, -1.5, -1.0, -0.5, 0.0, 0.5, 1.0]
    log = r.log()
    th = log[log[:, 0] == 1]
    assert list(th[2:, 2]) == signed and set(th[2:, 1]) == {1.25} and set(th[2:, 3]) == {1.0}   # Field::new(S::up(), magnitude)
    assert [l[1][1] for l in r.lines] == [abs(x) for x in signed]
    assert [b[4] for b in r.batches] == [abs(x) for x in signed for _ in (0, 1)]
    assert r.m.steps_done == len(signed) * 6
    r.close()


def test_program_errors_and_device_failures_propagate(mock_lib):
    from vegas_rs_b200.machine import ProgramError
    from vegas_rs_b200 import VegasGpuError
    r = Rig(mock_lib)
    for call, name in ((lambda: r.m.relax(0, 1.0), "NoSteps"), (lambda: r.m.relax(5, 0.0), "ZeroTemperature"),
                       (lambda: r.m.cooldown(1.0, 2.0, 0.1, 1, 1), "TemperatureMaxLessThanMin"),
                       (lambda: r.m.cooldown(2.0, 1.0, 0.0, 1, 1), "ZeroCoolRate"), (lambda: r.m.cooldown(2.0, 0.0, 0.1, 1, 1), "ZeroTemperature"),
                       (lambda: r.m.hysteresis(1, 1, 1.0, 0.0, 0.1), "ZeroField"), (lambda: r.m.hysteresis(1, 1, 1.0, 1.0, 0.0), "ZeroFieldStep"),
                       (lambda: r.m.hysteresis(0, 1, 1.0, 1.0, 0.1), "NoSteps")):
        with pytest.raises(ProgramError) as ei:
            call()
        assert name in str(ei.value)                           # ProgramError variants, src/error.rs:31-46
    assert mock_lib.mock_gpu_steps(r.h) == 0 and not r.lines   # nothing ran
    r.close()
    bad = Rig(mock_lib, fail_at=9000)
    with pytest.raises(VegasGpuError) as ei:
        bad.m.relax(20000, 2.0)
    assert "scripted device failure" in str(ei.value) and ei.value.code == -2
    assert bad.m.steps_done == 8192                            # the two batches before the failing one were delivered
    bad.close()


@pytest.mark.parametrize("heis", [False, True], ids=["ising", "heisenberg"])
def test_toml_front_end_writes_reference_outputs(mock_lib, tmp_path, heis):
    """`vegas run` wiring (src/input.rs:264-345) over the scripted device: the stdout lines of StatSensor, the observables
    parquet (one row per step of every stage, relax flag, stage and step counters, n, T, |H|) and the state parquet (one
    row per site and dump) -- docs/metropolis.toml's shape, shortened."""
    import io
    import pyarrow.parquet as pq
    from vegas_rs_b200 import run
    text = open(os.path.join(ROOT, "tests", "golden", "cfg0_ising_sc10.toml")).read()
    text = text.replace("steps = 20000", "steps = 6").replace("relax = 1000", "relax = 3").replace("steps = 1000", "steps = 5")
    text = text.replace("cool_rate = 0.05", "cool_rate = 2.5").replace("frequency = 1000", "frequency = 4")
    text = text.replace("./output.parquet", str(tmp_path / "output.parquet")).replace("./state.parquet", str(tmp_path / "state.parquet"))
    if heis:
        text = text.replace('model = "Ising"', 'model = "Heisenberg"')
    cfg = run.parse_input(text)
    assert cfg["model"] == ("Heisenberg" if heis else "Ising") and cfg["size"] == (10, 10, 10)
    r = Rig(mock_lib, n=1000, heisenberg=heis, sensors=False)
    out = io.StringIO()
    run.run_stages(cfg, r.m, out)
    # stages: Relax(5 @ 6.0), CoolDown 6.0 -> 1.0 by 2.5 = points 6.0, 3.5, 1.0, each relax 3 + measure 6
    lines = out.getvalue().strip().split("\n")
    assert [l.split()[0] for l in lines] == ["6.0000000000000000", "3.5000000000000000", "1.0000000000000000"]
    obs = pq.read_table(tmp_path / "output.parquet").to_pydict()
    stages = [(True, 5, 6.0)] + [x for T in (6.0, 3.5, 1.0) for x in ((True, 3, T), (False, 6, T))]
    assert obs["relax"] == [rx for rx, k, _ in stages for _ in range(k)]
    assert obs["stage"] == [i for i, (_, k, _) in enumerate(stages) for _ in range(k)]
    assert obs["step"] == [j for _, k, _ in stages for j in range(k)]
    assert set(obs["n"]) == {1000} and set(obs["field"]) == {0.0}
    assert obs["temperature"] == [T for _, k, T in stages for _ in range(k)]
    device_step = np.arange(1, 33)
    assert np.array_equal(obs["energy"], [energy(k, T) for k, T in zip(device_step, obs["temperature"])])
    assert np.array_equal(obs["magnetization"], [magnitude(k, heis) for k in device_step])
    st = pq.read_table(tmp_path / "state.parquet").to_pydict()
    dumps = [(True, 0, 0, 6.0), (True, 0, 4, 6.0)] + [x for i, T in enumerate((6.0, 3.5, 1.0))
                                                     for x in ((True, 1 + 2 * i, 0, T), (False, 2 + 2 * i, 0, T), (False, 2 + 2 * i, 4, T))]
    assert len(st["id"]) == 1000 * len(dumps) and st["id"][:1000] == list(range(1000))
    assert [(st["relax"][i], st["stage"][i], st["step"][i], st["temperature"][i]) for i in range(0, len(st["id"]), 1000)] == dumps
    assert set(st["sz"]) == {1.0, -1.0} and set(st["sx"]) == {0.0} and set(st["sy"]) == {0.0}
    assert not os.path.exists(tmp_path / "output.parquet.tmp") and not os.path.exists(tmp_path / "state.parquet.tmp")
    r.close()


def test_run_input_needs_the_cuda_device(built):
    """`python -m vegas_rs_b200.run` has no CPU fallback: without a device the run fails with the library's message."""
    import torch
    from vegas_rs_b200 import run, VegasGpuError
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = run.parse_input(open(os.path.join(ROOT, "tests", "golden", "cfg0_ising_sc10.toml")).read())
    cfg["output"] = None
    with pytest.raises(VegasGpuError) as ei:
        run.run_input(cfg, seed=1)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
