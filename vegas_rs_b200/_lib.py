"""ctypes signatures of include/vegas_gpu.h.  Loading fails loudly when the CUDA extension is
missing: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvegas_gpu.so")

IPC_BYTES = 256


class ModelDesc(C.Structure):
    _fields_ = [("model", C.c_int), ("proposal", C.c_int), ("precision", C.c_int), ("exchange", C.c_double),
                ("has_exchange", C.c_int), ("has_zeeman", C.c_int), ("has_anisotropy", C.c_int),
                ("anisotropy_k", C.c_double), ("anisotropy_axis", C.c_double * 3), ("has_gauge", C.c_int),
                ("gauge", C.c_double), ("seed", C.c_uint64), ("device", C.c_int), ("force_general", C.c_int)]


class LatticeDesc(C.Structure):
    _fields_ = [("unitcell", C.c_int), ("nx", C.c_uint64), ("ny", C.c_uint64), ("nz", C.c_uint64), ("pbc_x", C.c_int),
                ("pbc_y", C.c_int), ("pbc_z", C.c_int), ("literal_from_lattice_filter", C.c_int),
                ("nz_global", C.c_uint64), ("z_offset", C.c_uint64)]


class CsrDesc(C.Structure):
    _fields_ = [("n", C.c_uint64), ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("values", C.c_void_p)]


# every symbol include/vegas_gpu.h declares: (name, restype, argtypes)
_vp, _u64, _int, _dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
SYMBOLS = [
    ("vegas_gpu_create_lattice", _int, [C.POINTER(ModelDesc), C.POINTER(LatticeDesc), C.POINTER(_vp)]),
    ("vegas_gpu_create_csr", _int, [C.POINTER(ModelDesc), C.POINTER(CsrDesc), C.POINTER(_vp)]),
    ("vegas_gpu_destroy", None, [_vp]),
    ("vegas_gpu_last_error", C.c_char_p, [_vp]),
    ("vegas_gpu_version", C.c_char_p, []),
    ("vegas_gpu_n_sites", _u64, [_vp]),
    ("vegas_gpu_n_colours", _int, [_vp]),
    ("vegas_gpu_kernel_family", C.c_char_p, [_vp]),
    ("vegas_gpu_adjacency", _int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    ("vegas_gpu_colours", _int, [_vp, _vp]),
    ("vegas_gpu_lattice_adjacency", _int, [C.POINTER(LatticeDesc), _dbl, _vp, _vp, _vp, _vp, _vp]),
    ("vegas_gpu_lattice_colours", _int, [C.POINTER(LatticeDesc), _vp, _vp]),
    ("vegas_gpu_check_basis_tables", _int, []),
    ("vegas_gpu_wave_schedule", _int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("vegas_gpu_state_transfer_bytes", _u64, [_vp]),
    ("vegas_gpu_host_pack", _int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64]),
    ("vegas_gpu_host_unpack", _int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64]),
    ("vegas_gpu_basis_pair_structure", _int, [C.c_int]),
    ("vegas_gpu_basis_wave_schedule", _int, [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]),
    ("vegas_gpu_upload_ising", _int, [_vp, _vp, _u64]),
    ("vegas_gpu_upload_heisenberg", _int, [_vp, _vp, _u64]),
    ("vegas_gpu_download_ising", _int, [_vp, _vp, _u64]),
    ("vegas_gpu_download_heisenberg", _int, [_vp, _vp, _u64]),
    ("vegas_gpu_randomize", _int, [_vp]),
    ("vegas_gpu_fill", _int, [_vp, _int]),
    ("vegas_gpu_set_thermostat", _int, [_vp, _dbl, _vp, _dbl]),
    ("vegas_gpu_set_energy_convention", _int, [_vp, _int]),
    ("vegas_gpu_step", _int, [_vp, _u64, _vp, _vp]),
    ("vegas_gpu_step_async", _int, [_vp, _u64, _int]),
    ("vegas_gpu_read_observables", _int, [_vp, _u64, _vp, _vp]),
    ("vegas_gpu_synchronize", _int, [_vp]),
    ("vegas_gpu_step_host_ising", _int, [_vp, _vp, _u64, _vp, _vp]),
    ("vegas_gpu_step_host_heisenberg", _int, [_vp, _vp, _u64, _vp, _vp]),
    ("vegas_gpu_total_energy", _int, [_vp, _vp]),
    ("vegas_gpu_magnetization", _int, [_vp, _vp]),
    ("vegas_gpu_site_energies", _int, [_vp, _vp]),
    ("vegas_gpu_delta_energies", _int, [_vp, _vp, _vp]),
    ("vegas_gpu_attempt_count", _int, [_vp, _vp, _vp]),
    ("vegas_gpu_sweep_count", _int, [_vp, _vp]),
    ("vegas_gpu_set_sweep_count", _int, [_vp, _u64]),
    ("vegas_gpu_ising_thresholds", _int, [_vp, _vp, _vp, _vp]),
    ("vegas_gpu_slab_export", _int, [_vp, _vp]),
    ("vegas_gpu_slab_connect", _int, [_vp, _vp, _vp]),
    ("vegas_gpu_slab_connect_local", _int, [_vp, _vp, _vp]),
    ("vegas_gpu_set_tuning", _int, [_vp, C.c_char_p, C.c_long]),
    ("vegas_gpu_step_kernel", C.c_char_p, [_vp]),
    ("vegas_gpu_timer_start", _int, [_vp]),
    ("vegas_gpu_timer_stop", _int, [_vp, _vp]),
    ("vegas_gpu_launch_count", _u64, [_vp]),
    ("vegas_gpu_stream", _vp, [_vp]),
]

_lib = None


def load() -> C.CDLL:
    """Load libvegas_gpu.so (built in-tree by vegas_rs_b200.build).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m vegas_rs_b200.build` (needs nvcc). "
            "vegas_rs_b200 has no CPU fallback for the Metropolis sweep.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
