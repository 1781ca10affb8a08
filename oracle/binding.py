"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under vegas_rs_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ISING, HEISENBERG = 0, 1
SC, BCC, FCC = 0, 1, 2
TERM_GAUGE, TERM_ANISOTROPY, TERM_ZEEMAN, TERM_EXCHANGE = 0, 1, 2, 3
PROPOSE_FLIP, PROPOSE_RANDOM = 0, 1

u64, f64 = C.c_uint64, C.c_double
p_u64, p_f64 = C.POINTER(C.c_uint64), C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("vegas_oracle.c", "vegas_replay.c", "vegas_oracle.h", "Makefile")]
    stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class Rng(C.Structure):
    _fields_ = [("state", C.c_uint64 * 2), ("inc", C.c_uint64 * 2)]
    _pack_ = 16


class LatticeS(C.Structure):
    _fields_ = [("n_sites", u64), ("n_edges", u64), ("src", p_u64), ("dst", p_u64)]


class CsrS(C.Structure):
    _fields_ = [("n", u64), ("row_ptr", p_u64), ("col_idx", p_u64), ("values", p_f64)]


class HamS(C.Structure):
    _fields_ = [("model", C.c_int), ("n_terms", C.c_int), ("terms", C.c_int * 4), ("gauge", f64), ("aniso_k", f64),
                ("aniso_axis", f64 * 3), ("exchange", C.POINTER(CsrS))]


class ThermoS(C.Structure):
    _fields_ = [("temperature", f64), ("field_dir", f64 * 3), ("field_mag", f64)]


class StatRow(C.Structure):
    _fields_ = [("temperature", f64), ("field", f64), ("mean_e", f64), ("cv", f64), ("mean_m", f64), ("chi", f64),
                ("binder", f64)]


class MachineS(C.Structure):
    _fields_ = [("h", C.POINTER(HamS)), ("th", ThermoS), ("proposal", C.c_int), ("rng", C.c_void_p),
                ("state", C.c_void_p), ("n", u64), ("n_sensors", C.c_int), ("rows", C.POINTER(StatRow)),
                ("rows_cap", u64), ("rows_len", u64), ("obs_energy", p_f64), ("obs_mag", p_f64), ("obs_cap", u64),
                ("obs_len", u64), ("attempts", u64), ("scratch", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.vo_rng_seed.argtypes = [vp, u64]
    L.vo_rng_u64.argtypes = [vp]; L.vo_rng_u64.restype = u64
    L.vo_rng_f64.argtypes = [vp]; L.vo_rng_f64.restype = f64
    L.vo_rng_below.argtypes = [vp, u64]; L.vo_rng_below.restype = u64
    L.vo_lattice_build.argtypes = [C.c_int, u64, u64, u64, C.c_int, C.c_int, C.c_int, C.POINTER(LatticeS)]
    L.vo_lattice_free.argtypes = [C.POINTER(LatticeS)]
    L.vo_csr_from_lattice.argtypes = [C.POINTER(LatticeS), f64, C.c_int, C.POINTER(CsrS)]
    L.vo_csr_from_triplets.argtypes = [u64, u64, vp, vp, vp, C.POINTER(CsrS)]
    L.vo_csr_free.argtypes = [C.POINTER(CsrS)]
    L.vo_thermostat.argtypes = [f64, vp, f64]; L.vo_thermostat.restype = ThermoS
    L.vo_energy.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), vp, u64, u64]; L.vo_energy.restype = f64
    L.vo_total_energy.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), vp, u64]; L.vo_total_energy.restype = f64
    L.vo_site_energies.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), vp, u64, vp]
    L.vo_delta_energies.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), vp, u64, vp, vp]
    L.vo_magnetization.argtypes = [C.c_int, vp, u64, vp]; L.vo_magnetization.restype = f64
    L.vo_marsaglia.argtypes = [vp, vp]
    L.vo_state_rand.argtypes = [C.c_int, vp, vp, u64]
    L.vo_metropolis_step.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), C.c_int, vp, vp, u64]
    L.vo_metropolis_step.restype = u64
    for f in ("vo_acc_mean", "vo_acc_variance", "vo_acc_binder"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = f64
    L.vo_acc_reset.argtypes = [vp]; L.vo_acc_collect.argtypes = [vp, f64]
    L.vo_machine_init.argtypes = [C.POINTER(MachineS), C.POINTER(HamS), C.c_int, vp, vp, u64, C.c_int]
    L.vo_machine_free.argtypes = [C.POINTER(MachineS)]
    L.vo_relax_for.argtypes = [C.POINTER(MachineS), u64]
    L.vo_measure_for.argtypes = [C.POINTER(MachineS), u64]
    L.vo_program_relax.argtypes = [C.POINTER(MachineS), u64, f64]
    L.vo_program_cooldown.argtypes = [C.POINTER(MachineS), f64, f64, f64, u64, u64]
    L.vo_program_hysteresis.argtypes = [C.POINTER(MachineS), u64, u64, f64, f64, f64]
    L.vo_cooldown_points.argtypes = [f64, f64, f64, vp, u64]; L.vo_cooldown_points.restype = u64
    L.vo_hysteresis_points.argtypes = [f64, f64, vp, u64]; L.vo_hysteresis_points.restype = u64
    L.vo_stat_line.argtypes = [C.POINTER(StatRow), C.c_char_p, C.c_size_t]
    L.vo_philox4x32_10.argtypes = [vp, vp, vp]
    L.vo_philox4x32.argtypes = [vp, vp, C.c_int, vp]
    L.vo_replay_ising_msc.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), C.c_int, u64, u64, u64, u64, u64, vp]
    L.vo_replay_ising_msc.restype = u64
    L.vo_replay_ising_sites.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), C.c_int, u64, u64, u64, vp, C.c_int, vp]
    L.vo_replay_ising_sites.restype = u64
    L.vo_replay_heisenberg.argtypes = [C.POINTER(HamS), C.POINTER(ThermoS), C.c_int, C.c_int, u64, u64, u64, vp,
                                       C.c_int, vp]
    L.vo_replay_heisenberg.restype = u64
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleRng:
    def __init__(self, seed: int):
        self.buf = (C.c_uint8 * 64)()
        addr = C.addressof(self.buf)
        self.p = C.c_void_p((addr + 15) & ~15)
        lib().vo_rng_seed(self.p, seed)

    def u64(self) -> int: return lib().vo_rng_u64(self.p)
    def f64(self) -> float: return lib().vo_rng_f64(self.p)
    def below(self, n: int) -> int: return lib().vo_rng_below(self.p, n)


class Lattice:
    """Lattice::{sc,bcc,fcc}(1.0).expand(x,y,z).drop_*() restated (src/input.rs:296-322)."""

    def __init__(self, unitcell: int, nx: int, ny: int, nz: int, pbc=(True, True, True)):
        self.s = LatticeS()
        rc = lib().vo_lattice_build(unitcell, nx, ny, nz, int(pbc[0]), int(pbc[1]), int(pbc[2]), C.byref(self.s))
        if rc:
            raise RuntimeError(f"vo_lattice_build failed: {rc}")
        self.n_sites, self.n_edges = self.s.n_sites, self.s.n_edges

    def edges(self):
        n = self.n_edges
        src = np.ctypeslib.as_array(self.s.src, (n,)).copy() if n else np.zeros(0, np.uint64)
        dst = np.ctypeslib.as_array(self.s.dst, (n,)).copy() if n else np.zeros(0, np.uint64)
        return src, dst

    def __del__(self):
        try:
            lib().vo_lattice_free(C.byref(self.s))
        except Exception:
            pass


class Csr:
    """Exchange::from_lattice / Exchange::new (src/energy.rs:171-187)."""

    def __init__(self):
        self.s = CsrS()
        self._keep = None

    @classmethod
    def from_lattice(cls, lat: Lattice, exchange: float = 1.0, literal: bool = False) -> "Csr":
        m = cls()
        rc = lib().vo_csr_from_lattice(C.byref(lat.s), exchange, int(literal), C.byref(m.s))
        if rc:
            raise RuntimeError("vo_csr_from_lattice failed")
        return m

    @classmethod
    def from_triplets(cls, n: int, rows, cols, vals) -> "Csr":
        m = cls()
        r = np.ascontiguousarray(rows, np.uint64); c = np.ascontiguousarray(cols, np.uint64)
        v = np.ascontiguousarray(vals, np.float64)
        rc = lib().vo_csr_from_triplets(n, len(r), _ptr(r), _ptr(c), _ptr(v), C.byref(m.s))
        if rc:
            raise RuntimeError("vo_csr_from_triplets failed")
        return m

    @property
    def n(self): return self.s.n

    def arrays(self):
        n = self.s.n
        rp = np.ctypeslib.as_array(self.s.row_ptr, (n + 1,)).copy()
        nnz = int(rp[-1])
        col = np.ctypeslib.as_array(self.s.col_idx, (max(nnz, 1),))[:nnz].copy()
        val = np.ctypeslib.as_array(self.s.values, (max(nnz, 1),))[:nnz].copy()
        return rp, col, val

    def __del__(self):
        try:
            lib().vo_csr_free(C.byref(self.s))
        except Exception:
            pass


class Hamiltonian:
    """hamiltonian!(...) compound (src/energy.rs:220-290); terms in the given order."""

    def __init__(self, model: int, terms, csr: Csr | None = None, gauge=0.0, aniso_k=0.0, aniso_axis=(0.0, 0.0, 1.0)):
        self.model = model
        self.csr = csr
        self.s = HamS()
        self.s.model = model
        self.s.n_terms = len(terms)
        for i, t in enumerate(terms):
            self.s.terms[i] = t
        self.s.gauge = gauge
        self.s.aniso_k = aniso_k
        for i in range(3):
            self.s.aniso_axis[i] = aniso_axis[i]
        self.s.exchange = C.pointer(csr.s) if csr is not None else None

    def thermostat(self, temperature: float, field_dir=(0.0, 0.0, 1.0), field_mag: float = 0.0) -> ThermoS:
        d = (f64 * 3)(*field_dir)
        return lib().vo_thermostat(temperature, C.cast(d, C.c_void_p), field_mag)

    def _state(self, state):
        if self.model == ISING:
            a = np.ascontiguousarray(state, np.int8); return a, a.size
        a = np.ascontiguousarray(state, np.float64).reshape(-1, 3); return a, a.shape[0]

    def energy(self, th, state, i):
        a, n = self._state(state)
        return lib().vo_energy(C.byref(self.s), C.byref(th), _ptr(a), n, i)

    def total_energy(self, th, state):
        a, n = self._state(state)
        return lib().vo_total_energy(C.byref(self.s), C.byref(th), _ptr(a), n)

    def site_energies(self, th, state):
        a, n = self._state(state)
        out = np.zeros(n)
        lib().vo_site_energies(C.byref(self.s), C.byref(th), _ptr(a), n, _ptr(out))
        return out

    def delta_energies(self, th, state, proposal=None):
        a, n = self._state(state)
        out = np.zeros(n)
        if proposal is None:
            lib().vo_delta_energies(C.byref(self.s), C.byref(th), _ptr(a), n, None, _ptr(out))
        else:
            p, _ = self._state(proposal)
            lib().vo_delta_energies(C.byref(self.s), C.byref(th), _ptr(a), n, _ptr(p), _ptr(out))
        return out

    def magnetization(self, state):
        a, n = self._state(state)
        xyz = np.zeros(3)
        mag = lib().vo_magnetization(self.model, _ptr(a), n, _ptr(xyz))
        return mag, xyz

    def rand_state(self, rng: OracleRng, n: int):
        a = np.zeros(n, np.int8) if self.model == ISING else np.zeros((n, 3))
        lib().vo_state_rand(self.model, rng.p, _ptr(a), n)
        return a

    def step(self, th, proposal: int, rng: OracleRng, state) -> int:
        """One Integrator::step in place (src/integrator.rs:66-92 / :109-138)."""
        assert state.flags["C_CONTIGUOUS"]
        n = state.size if self.model == ISING else state.shape[0]
        return lib().vo_metropolis_step(C.byref(self.s), C.byref(th), proposal, rng.p, _ptr(state), n)

    # ---- replay of the GPU's colour-ordered sweep with its Philox numbers
    def replay_ising_msc(self, th, proposal, seed, sweep, dims, state) -> int:
        return lib().vo_replay_ising_msc(C.byref(self.s), C.byref(th), proposal, seed, sweep, dims[0], dims[1], dims[2],
                                         _ptr(state))

    def replay_ising_sites(self, th, proposal, seed, sweep, colours, n_colours, state) -> int:
        col = np.ascontiguousarray(colours, np.uint8)
        return lib().vo_replay_ising_sites(C.byref(self.s), C.byref(th), proposal, seed, sweep, state.size, _ptr(col),
                                           n_colours, _ptr(state))

    def replay_heisenberg(self, th, proposal, f32, seed, sweep, colours, n_colours, state) -> int:
        col = np.ascontiguousarray(colours, np.uint8)
        return lib().vo_replay_heisenberg(C.byref(self.s), C.byref(th), proposal, int(f32), seed, sweep, state.shape[0],
                                          _ptr(col), n_colours, _ptr(state))


class Machine:
    """Machine + programs (src/machine.rs:44-125, src/program.rs:66-336) over the oracle integrator."""

    def __init__(self, ham: Hamiltonian, proposal: int, rng: OracleRng, state, n_sensors: int = 1):
        self.ham, self.rng, self.state = ham, rng, state
        self.m = MachineS()
        n = state.size if ham.model == ISING else state.shape[0]
        rc = lib().vo_machine_init(C.byref(self.m), C.byref(ham.s), proposal, rng.p, _ptr(state), n, n_sensors)
        if rc:
            raise RuntimeError("vo_machine_init failed")

    def set_thermostat(self, th: ThermoS): self.m.th = th
    def relax_for(self, steps): return lib().vo_relax_for(C.byref(self.m), steps)
    def measure_for(self, steps): return lib().vo_measure_for(C.byref(self.m), steps)
    def relax(self, steps, temperature): return lib().vo_program_relax(C.byref(self.m), steps, temperature)

    def cooldown(self, tmax, tmin, rate, relax, steps):
        return lib().vo_program_cooldown(C.byref(self.m), tmax, tmin, rate, relax, steps)

    def hysteresis(self, steps, relax, temperature, max_field, field_step):
        return lib().vo_program_hysteresis(C.byref(self.m), steps, relax, temperature, max_field, field_step)

    @property
    def attempts(self): return self.m.attempts

    def rows(self):
        return [(r.temperature, r.field, r.mean_e, r.cv, r.mean_m, r.chi, r.binder)
                for r in (self.m.rows[i] for i in range(self.m.rows_len))]

    def stat_lines(self):
        out = []
        buf = C.create_string_buffer(512)
        for i in range(self.m.rows_len):
            lib().vo_stat_line(C.byref(self.m.rows[i]), buf, 512)
            out.append(buf.value.decode())
        return out

    def observables(self):
        n = self.m.obs_len
        if n == 0:
            return np.zeros(0), np.zeros(0)
        return (np.ctypeslib.as_array(self.m.obs_energy, (n,)).copy(), np.ctypeslib.as_array(self.m.obs_mag, (n,)).copy())

    def __del__(self):
        try:
            lib().vo_machine_free(C.byref(self.m))
        except Exception:
            pass


def cooldown_points(tmax, tmin, rate):
    n = lib().vo_cooldown_points(tmax, tmin, rate, None, 0)
    out = np.zeros(n)
    lib().vo_cooldown_points(tmax, tmin, rate, _ptr(out), n)
    return out


def hysteresis_points(max_field, field_step):
    n = lib().vo_hysteresis_points(max_field, field_step, None, 0)
    out = np.zeros(n)
    lib().vo_hysteresis_points(max_field, field_step, _ptr(out), n)
    return out


PHILOX_ROUNDS = 7   # VO_PHILOX_ROUNDS (oracle/vegas_oracle.h) = PHILOX_ROUNDS of the GPU kernels (csrc/common.cuh)


def philox(ctr, key, rounds=10):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    if rounds == 10:
        lib().vo_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    else:
        lib().vo_philox4x32(_ptr(c), _ptr(k), int(rounds), _ptr(o))
    return o
