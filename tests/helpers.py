"""Shared helpers for the parity tests: build the oracle-side model that corresponds to a GpuMetropolis."""
import numpy as np

from oracle import binding as ob


def oracle_model(model, *, unitcell=0, size=None, pbc=(True, True, True), literal=False, csr=None, exchange=1.0,
                 zeeman=True, anisotropy=None, gauge=None):
    """Returns (Hamiltonian, Csr) restating hamiltonian!(Exchange, Zeeman[, Anisotropy][, Gauge])."""
    if csr is None:
        lat = ob.Lattice(unitcell, *size, pbc=pbc)
        m = ob.Csr.from_lattice(lat, exchange, literal)
    else:
        rp, ci, va = csr
        rows = np.repeat(np.arange(len(rp) - 1, dtype=np.uint64), np.diff(np.asarray(rp, np.int64)))
        vals = np.full(len(ci), exchange) if va is None else va
        m = ob.Csr.from_triplets(len(rp) - 1, rows, np.asarray(ci, np.uint64), vals)
    terms = []
    if exchange is not None:
        terms.append(ob.TERM_EXCHANGE)
    if zeeman:
        terms.append(ob.TERM_ZEEMAN)
    kw = {}
    if anisotropy is not None:
        terms.append(ob.TERM_ANISOTROPY)
        kw.update(aniso_axis=anisotropy[0], aniso_k=anisotropy[1])
    if gauge is not None:
        terms.append(ob.TERM_GAUGE)
        kw.update(gauge=gauge)
    return ob.Hamiltonian(model, terms, m, **kw), m


def random_state(model, n, seed):
    rng = np.random.default_rng(seed)
    if model == ob.ISING:
        return (2 * rng.integers(0, 2, n) - 1).astype(np.int8)
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def blocking_error(x, nblocks=20):
    """Standard error of the mean of a correlated series from block averages."""
    x = np.asarray(x, float)
    m = len(x) // nblocks
    b = x[: m * nblocks].reshape(nblocks, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nblocks)
