"""GpuMetropolis: host-side mirror of the reference's Integrator + Hamiltonian pair for the GPU sweep.

Like WolffIntegrator (src/integrator.rs:146-186) it is constructed WITH the model description and
does not introspect a generic `hamiltonian` argument.  One object plays three reference roles:
  * Integrator::step            (src/integrator.rs:40-49)   -> step() / step_host()
  * Hamiltonian::{energy,total_energy} (src/energy.rs:45-60) -> site_energies() / total_energy()
  * the per-step observers of src/instrument.rs:133-141      -> the (E, M) series step() returns
Everything goes through the C ABI (include/vegas_gpu.h); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

ISING, HEISENBERG = 0, 1
PROPOSE_FLIP, PROPOSE_RANDOM = 0, 1
F32, F64 = 0, 1
SC, BCC, FCC = 0, 1, 2
E_PHYSICAL, E_REFERENCE_COMPOUND, E_REFERENCE_EXCHANGE = 0, 1, 2


class VegasGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"vegas_gpu error {code}: {message}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lattice_adjacency(unitcell, size, pbc=(True, True, True), literal=False, exchange=1.0):
    """Host-only: the CSR Exchange::from_lattice builds for this lattice (src/energy.rs:176-187)."""
    lib = _lib.load()
    ld = _lib.LatticeDesc(unitcell, size[0], size[1], size[2], int(pbc[0]), int(pbc[1]), int(pbc[2]), int(literal), 0, 0)
    n, nnz = C.c_uint64(), C.c_uint64()
    rc = lib.vegas_gpu_lattice_adjacency(C.byref(ld), exchange, C.byref(n), C.byref(nnz), None, None, None)
    if rc:
        raise VegasGpuError(rc, (lib.vegas_gpu_last_error(None) or b"").decode())
    rp = np.zeros(n.value + 1, np.uint64); ci = np.zeros(nnz.value, np.uint32); va = np.zeros(nnz.value)
    lib.vegas_gpu_lattice_adjacency(C.byref(ld), exchange, None, None, _ptr(rp), _ptr(ci), _ptr(va))
    return rp, ci, va


def lattice_colours(unitcell, size, pbc=(True, True, True), literal=False):
    """Host-only: colour of every site as the general-adjacency sweep orders them."""
    lib = _lib.load()
    ld = _lib.LatticeDesc(unitcell, size[0], size[1], size[2], int(pbc[0]), int(pbc[1]), int(pbc[2]), int(literal), 0, 0)
    nb = {SC: 1, BCC: 2, FCC: 4}[unitcell]
    nc = C.c_int(); out = np.zeros(size[0] * size[1] * size[2] * nb, np.uint8)
    rc = lib.vegas_gpu_lattice_colours(C.byref(ld), C.byref(nc), _ptr(out))
    if rc:
        raise VegasGpuError(rc, (lib.vegas_gpu_last_error(None) or b"").decode())
    return nc.value, out


class GpuMetropolis:
    def __init__(self, model: int, *, unitcell: int = SC, size=None, pbc=(True, True, True), csr=None,
                 exchange: float | None = 1.0, zeeman: bool = True, anisotropy=None, gauge: float | None = None,
                 proposal: int | None = None, precision: int = F32, seed: int = 0, device: int = 0,
                 literal: bool = False, nz_global: int = 0, z_offset: int = 0, force_general: bool = False):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.model = model
        md = _lib.ModelDesc()
        md.model = model
        # src/input.rs:347-367: TOML Ising -> MetropolisFlipIntegrator, Heisenberg -> MetropolisIntegrator
        md.proposal = proposal if proposal is not None else (PROPOSE_FLIP if model == ISING else PROPOSE_RANDOM)
        md.precision = precision
        md.has_exchange = exchange is not None
        md.exchange = exchange if exchange is not None else 0.0
        md.has_zeeman = int(zeeman)
        if anisotropy is not None:
            axis, k = anisotropy
            md.has_anisotropy, md.anisotropy_k = 1, k
            for i in range(3):
                md.anisotropy_axis[i] = axis[i]
        else:
            md.anisotropy_axis[2] = 1.0
        if gauge is not None:
            md.has_gauge, md.gauge = 1, gauge
        md.seed, md.device, md.force_general = seed, device, int(force_general)
        self.seed = seed
        self.proposal = md.proposal
        self.precision = precision
        if csr is not None:
            row_ptr, col_idx, values = csr
            self._rp = np.ascontiguousarray(row_ptr, np.uint64)
            self._ci = np.ascontiguousarray(col_idx, np.uint32)
            self._va = None if values is None else np.ascontiguousarray(values, np.float64)
            cd = _lib.CsrDesc(len(self._rp) - 1, _ptr(self._rp), _ptr(self._ci), _ptr(self._va))
            rc = self._lib.vegas_gpu_create_csr(C.byref(md), C.byref(cd), C.byref(self._h))
        else:
            ld = _lib.LatticeDesc(unitcell, size[0], size[1], size[2], int(pbc[0]), int(pbc[1]), int(pbc[2]),
                                  int(literal), nz_global, z_offset)
            rc = self._lib.vegas_gpu_create_lattice(C.byref(md), C.byref(ld), C.byref(self._h))
        if rc != 0:
            msg = self._lib.vegas_gpu_last_error(None)
            self._h = C.c_void_p()
            raise VegasGpuError(rc, msg.decode() if msg else "")
        self.n_sites = self._lib.vegas_gpu_n_sites(self._h)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.vegas_gpu_last_error(self._h)
            raise VegasGpuError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vegas_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_family(self) -> str: return self._lib.vegas_gpu_kernel_family(self._h).decode()
    @property
    def n_colours(self) -> int: return self._lib.vegas_gpu_n_colours(self._h)
    @property
    def launches(self) -> int: return self._lib.vegas_gpu_launch_count(self._h)
    @property
    def stream(self) -> int: return self._lib.vegas_gpu_stream(self._h) or 0
    @property
    def state_transfer_bytes(self) -> int:
        """bytes one upload / download of the State moves over PCIe (n / 8 on the host-packed Ising path)"""
        return self._lib.vegas_gpu_state_transfer_bytes(self._h)

    def adjacency(self):
        n, nnz = C.c_uint64(), C.c_uint64()
        self._check(self._lib.vegas_gpu_adjacency(self._h, C.byref(n), C.byref(nnz), None, None, None))
        rp = np.zeros(n.value + 1, np.uint64); ci = np.zeros(nnz.value, np.uint32); va = np.zeros(nnz.value)
        self._check(self._lib.vegas_gpu_adjacency(self._h, None, None, _ptr(rp), _ptr(ci), _ptr(va)))
        return rp, ci, va

    def colours(self):
        out = np.zeros(self.n_sites, np.uint8)
        self._check(self._lib.vegas_gpu_colours(self._h, _ptr(out)))
        return out

    # ------------------------------------------------------------------ state (reference host layouts)
    def upload(self, state):
        if self.model == ISING:
            a = np.ascontiguousarray(state, np.int8)
            self._check(self._lib.vegas_gpu_upload_ising(self._h, _ptr(a), a.size))
        else:
            a = np.ascontiguousarray(state, np.float64).reshape(-1, 3)
            self._check(self._lib.vegas_gpu_upload_heisenberg(self._h, _ptr(a), a.shape[0]))

    def download(self):
        if self.model == ISING:
            a = np.zeros(self.n_sites, np.int8)
            self._check(self._lib.vegas_gpu_download_ising(self._h, _ptr(a), a.size))
        else:
            a = np.zeros((self.n_sites, 3))
            self._check(self._lib.vegas_gpu_download_heisenberg(self._h, _ptr(a), a.shape[0]))
        return a

    def download_into(self, out):
        """download() into a caller-owned buffer (e.g. pinned host memory): int8[n] or float64[n,3], C-contiguous."""
        assert out.flags["C_CONTIGUOUS"] and out.dtype == (np.int8 if self.model == ISING else np.float64)
        if self.model == ISING:
            assert out.size == self.n_sites
            self._check(self._lib.vegas_gpu_download_ising(self._h, _ptr(out), out.size))
        else:
            assert out.size == 3 * self.n_sites
            self._check(self._lib.vegas_gpu_download_heisenberg(self._h, _ptr(out), self.n_sites))
        return out

    def randomize(self): self._check(self._lib.vegas_gpu_randomize(self._h))
    def fill(self, up: bool = True): self._check(self._lib.vegas_gpu_fill(self._h, int(up)))

    # ------------------------------------------------------------------ thermostat
    def set_thermostat(self, temperature: float, field_dir=(0.0, 0.0, 1.0), field_mag: float = 0.0):
        d = np.asarray(field_dir, np.float64)
        self._check(self._lib.vegas_gpu_set_thermostat(self._h, temperature, _ptr(d), field_mag))

    def set_energy_convention(self, conv: int): self._check(self._lib.vegas_gpu_set_energy_convention(self._h, conv))

    # ------------------------------------------------------------------ hot path
    def step(self, n_steps: int = 1, observe: bool = True):
        """n_steps x Integrator::step on the device-resident state; returns (E[n], M[n,3]) when observe."""
        if not observe:
            self._check(self._lib.vegas_gpu_step(self._h, n_steps, None, None))
            return None
        e = np.zeros(n_steps); m = np.zeros((n_steps, 3))
        self._check(self._lib.vegas_gpu_step(self._h, n_steps, _ptr(e), _ptr(m)))
        return e, m

    def step_async(self, n_steps: int, record: bool = True): self._check(self._lib.vegas_gpu_step_async(self._h, n_steps, int(record)))

    def read_observables(self, n_steps: int):
        e = np.zeros(n_steps); m = np.zeros((n_steps, 3))
        self._check(self._lib.vegas_gpu_read_observables(self._h, n_steps, _ptr(e), _ptr(m)))
        return e, m

    def synchronize(self): self._check(self._lib.vegas_gpu_synchronize(self._h))

    def step_host(self, state):
        """Literal Integrator::step: host State in -> host State out (in place), plus (E, M)."""
        e = C.c_double(); m = np.zeros(3)
        if self.model == ISING:
            assert state.dtype == np.int8 and state.flags["C_CONTIGUOUS"]
            self._check(self._lib.vegas_gpu_step_host_ising(self._h, _ptr(state), state.size, C.byref(e), _ptr(m)))
        else:
            assert state.dtype == np.float64 and state.flags["C_CONTIGUOUS"]
            self._check(self._lib.vegas_gpu_step_host_heisenberg(self._h, _ptr(state), state.shape[0], C.byref(e), _ptr(m)))
        return state, e.value, m

    # ------------------------------------------------------------------ Hamiltonian role
    def total_energy(self) -> float:
        e = C.c_double()
        self._check(self._lib.vegas_gpu_total_energy(self._h, C.byref(e)))
        return e.value

    def magnetization(self):
        m = np.zeros(3)
        self._check(self._lib.vegas_gpu_magnetization(self._h, _ptr(m)))
        return m

    def site_energies(self):
        out = np.zeros(self.n_sites)
        self._check(self._lib.vegas_gpu_site_energies(self._h, _ptr(out)))
        return out

    def delta_energies(self, proposal=None):
        out = np.zeros(self.n_sites)
        if proposal is not None:
            proposal = np.ascontiguousarray(proposal, np.int8 if self.model == ISING else np.float64)
        self._check(self._lib.vegas_gpu_delta_energies(self._h, _ptr(proposal), _ptr(out)))
        return out

    def attempt_count(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.vegas_gpu_attempt_count(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def sweeps(self) -> int:
        s = C.c_uint64()
        self._check(self._lib.vegas_gpu_sweep_count(self._h, C.byref(s)))
        return s.value

    @sweeps.setter
    def sweeps(self, v: int): self._check(self._lib.vegas_gpu_set_sweep_count(self._h, v))

    def ising_thresholds(self):
        n = C.c_int(); thr = np.zeros(16, np.uint64); alw = np.zeros(16, np.uint8)
        self._check(self._lib.vegas_gpu_ising_thresholds(self._h, C.byref(n), _ptr(thr), _ptr(alw)))
        return n.value, thr.reshape(2, 8), alw.reshape(2, 8)

    # ------------------------------------------------------------------ slabs / timing
    def slab_export(self) -> bytes:
        buf = C.create_string_buffer(_lib.IPC_BYTES)
        self._check(self._lib.vegas_gpu_slab_export(self._h, buf))
        return buf.raw

    def slab_connect(self, lower: bytes, upper: bytes):
        self._check(self._lib.vegas_gpu_slab_connect(self._h, C.c_char_p(lower), C.c_char_p(upper)))

    def slab_connect_local(self, lower: "GpuMetropolis", upper: "GpuMetropolis"):
        self._check(self._lib.vegas_gpu_slab_connect_local(self._h, lower._h, upper._h))

    def set_tuning(self, key: str, value: int): self._check(self._lib.vegas_gpu_set_tuning(self._h, key.encode(), int(value)))

    @property
    def step_kernel(self) -> str: return self._lib.vegas_gpu_step_kernel(self._h).decode()

    def timer_start(self): self._check(self._lib.vegas_gpu_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.vegas_gpu_timer_stop(self._h, C.byref(ms)))
        return ms.value
