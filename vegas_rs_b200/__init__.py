"""vegas_rs_b200 -- B200-native Metropolis sweep for vegas-rs (Ising / Heisenberg, compound
Gauge + Exchange + Anisotropy + Zeeman Hamiltonian) behind the C ABI of include/vegas_gpu.h.

The product is the CUDA library `libvegas_gpu.so` (vegas_rs_b200/csrc); this package is the thin
host-side mirror of the reference's Integrator / Hamiltonian interface used by tests and bench.py.
Importing it never touches oracle/ and never falls back to a CPU path.
"""
from .gpu_metropolis import (ISING, HEISENBERG, PROPOSE_FLIP, PROPOSE_RANDOM, F32, F64, SC, BCC, FCC, E_PHYSICAL,
                             E_REFERENCE_COMPOUND, E_REFERENCE_EXCHANGE, GpuMetropolis, VegasGpuError)

__all__ = ["ISING", "HEISENBERG", "PROPOSE_FLIP", "PROPOSE_RANDOM", "F32", "F64", "SC", "BCC", "FCC", "E_PHYSICAL",
           "E_REFERENCE_COMPOUND", "E_REFERENCE_EXCHANGE", "GpuMetropolis", "VegasGpuError"]
