"""Python binding of the C++ host layer (include/vegas_host.h): Machine, instruments and programs of the
reference (src/machine.rs, src/instrument.rs, src/program.rs) over a device-resident GpuMetropolis."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .gpu_metropolis import ISING, GpuMetropolis, VegasGpuError

STAT_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)
OBS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_double),
                     C.POINTER(C.c_double), C.c_uint64)
STATE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_void_p, C.c_uint64)
REDUCE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_uint64)

HOST_SYMBOLS = [
    ("vegas_machine_create", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("vegas_machine_destroy", None, [C.c_void_p]),
    ("vegas_machine_last_error", C.c_char_p, [C.c_void_p]),
    ("vegas_machine_add_stat_sensor", C.c_int, [C.c_void_p, STAT_CB, C.c_void_p]),
    ("vegas_machine_add_observable_sensor", C.c_int, [C.c_void_p, OBS_CB, C.c_void_p]),
    ("vegas_machine_add_state_sensor", C.c_int, [C.c_void_p, C.c_uint64, STATE_CB, C.c_void_p]),
    ("vegas_machine_set_thermostat", C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_double]),
    ("vegas_machine_thermostat", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("vegas_machine_relax_for", C.c_int, [C.c_void_p, C.c_uint64]),
    ("vegas_machine_measure_for", C.c_int, [C.c_void_p, C.c_uint64]),
    ("vegas_machine_steps_done", C.c_uint64, [C.c_void_p]),
    ("vegas_machine_set_group", C.c_int, [C.c_void_p, REDUCE_CB, C.c_void_p, C.c_uint64]),
    ("vegas_program_relax", C.c_int, [C.c_void_p, C.c_uint64, C.c_double]),
    ("vegas_program_cooldown", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64]),
    ("vegas_program_hysteresis", C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double]),
]

# ProgramError (src/error.rs:31-46) as returned through the C ABI
PROGRAM_ERRORS = {-10: "NoSteps", -11: "ZeroTemperature", -12: "TemperatureMaxLessThanMin", -13: "ZeroCoolRate",
                  -14: "ZeroField", -15: "ZeroFieldStep"}


class ProgramError(VegasGpuError):
    pass


def bind_host_symbols(lib):
    """ctypes signatures of include/vegas_host.h on `lib` (the CUDA library, or the host layer linked against the scripted
    test double of tests/mock/ in the CPU-only tests)."""
    if not getattr(lib, "_host_bound", False):
        for name, res, args in HOST_SYMBOLS:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib._host_bound = True
    return lib


def _load():
    return bind_host_symbols(_lib.load())


class Machine:
    """Machine::new(Thermostat::new(2.8, Field::zero()), hamiltonian, integrator, instruments, state)
    (src/input.rs:273-279) with the GPU handle standing in for hamiltonian + integrator + state."""

    def __init__(self, gpu: GpuMetropolis, lib=None):
        self._lib = _load() if lib is None else bind_host_symbols(lib)   # lib: tests only (host layer over a test double)
        self.gpu = gpu
        self._m = C.c_void_p()
        rc = self._lib.vegas_machine_create(gpu._h, C.byref(self._m))
        if rc:
            raise VegasGpuError(rc, "vegas_machine_create failed")
        self._keep = []
        self._cb_error = None

    def _guard(self, fn):
        """ctypes swallows exceptions raised inside a callback: keep the first one and re-raise it when the C call returns"""
        def wrapped(*a):
            if self._cb_error is None:
                try:
                    return fn(*a)
                except Exception as e:
                    self._cb_error = e
        return wrapped

    def _check(self, rc):
        err, self._cb_error = self._cb_error, None
        if err is not None:
            raise err
        if rc:
            msg = (self._lib.vegas_machine_last_error(self._m) or b"").decode()
            if rc in PROGRAM_ERRORS:
                raise ProgramError(rc, PROGRAM_ERRORS[rc] + ": " + msg)
            raise VegasGpuError(rc, msg)

    def close(self):
        if self._m and self._m.value:
            self._lib.vegas_machine_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- instruments (called in the order added, src/machine.rs:96-98)
    def add_stat_sensor(self, on_line):
        """on_line(line: str, row: tuple of the 7 numbers) -- StatSensor, src/instrument.rs:61-142"""
        cb = STAT_CB(self._guard(lambda u, line, *row: on_line(line.decode(), row)))
        self._keep.append(cb)
        self._check(self._lib.vegas_machine_add_stat_sensor(self._m, cb, None))

    def add_observable_sensor(self, on_batch):
        """on_batch(relax, stage, n, T, field, energy[], magnetization[]) -- ObservableSensor, src/instrument.rs:145-263"""
        def tramp(u, relax, stage, n, T, field, e, m, ln):
            ea = np.ctypeslib.as_array(e, (ln,)).copy() if ln else np.zeros(0)
            ma = np.ctypeslib.as_array(m, (ln,)).copy() if ln else np.zeros(0)
            on_batch(bool(relax), stage, n, T, field, ea, ma)
        cb = OBS_CB(self._guard(tramp))
        self._keep.append(cb)
        self._check(self._lib.vegas_machine_add_observable_sensor(self._m, cb, None))

    def add_state_sensor(self, frequency: int, on_state):
        """on_state(relax, stage, step, T, field, state) -- StateSensor, src/instrument.rs:265-351"""
        ising = self.gpu.model == ISING
        def tramp(u, relax, stage, step, T, field, ptr, n):
            if ising:
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int8)), (n,)).copy()
            else:
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), (n, 3)).copy()
            on_state(bool(relax), stage, step, T, field, a)
        cb = STATE_CB(self._guard(tramp))
        self._keep.append(cb)
        self._check(self._lib.vegas_machine_add_state_sensor(self._m, frequency, cb, None))

    # ---- slab group (multi-GPU): one Machine per rank over its own z-slab of one lattice
    def set_group(self, reduce_sum, n_sites_global: int):
        """reduce_sum(values: np.ndarray) sums the array IN PLACE over the ranks (e.g. an all-reduce); every rank must run
        the same program.  The instruments then see the whole lattice's E, |M| and State::len = n_sites_global."""
        def tramp(u, ptr, ln):
            try:
                reduce_sum(np.ctypeslib.as_array(ptr, (ln,)))
                return 0
            except Exception as e:  # ctypes would swallow it: report a failed reduction instead
                self._cb_error = e
                return 1
        cb = REDUCE_CB(tramp)
        self._keep.append(cb)
        self._check(self._lib.vegas_machine_set_group(self._m, cb, None, n_sites_global))

    # ---- machine (src/machine.rs:104-125)
    def set_thermostat(self, temperature, field_dir=(0.0, 0.0, 1.0), field_mag=0.0):
        d = np.asarray(field_dir, np.float64)
        self._check(self._lib.vegas_machine_set_thermostat(self._m, temperature, d.ctypes.data_as(C.c_void_p), field_mag))

    def thermostat(self):
        t, f = C.c_double(), C.c_double()
        self._check(self._lib.vegas_machine_thermostat(self._m, C.byref(t), C.byref(f)))
        return t.value, f.value

    def relax_for(self, steps): self._check(self._lib.vegas_machine_relax_for(self._m, steps))
    def measure_for(self, steps): self._check(self._lib.vegas_machine_measure_for(self._m, steps))
    @property
    def steps_done(self): return self._lib.vegas_machine_steps_done(self._m)

    # ---- programs (src/program.rs)
    def relax(self, steps, temperature): self._check(self._lib.vegas_program_relax(self._m, steps, temperature))

    def cooldown(self, max_temperature, min_temperature, cool_rate, relax, steps):
        self._check(self._lib.vegas_program_cooldown(self._m, max_temperature, min_temperature, cool_rate, relax, steps))

    def hysteresis(self, steps, relax, temperature, max_field, field_step):
        self._check(self._lib.vegas_program_hysteresis(self._m, steps, relax, temperature, max_field, field_step))
