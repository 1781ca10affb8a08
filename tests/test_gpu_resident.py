"""GPU tests of the shared-memory-resident step kernels (csrc/resident.cuh): a batch of Monte Carlo steps of a small
general-family lattice in ONE launch must reproduce the launch-per-colour path bit for bit (same Philox counters), and
its fused per-step observers must equal the dedicated reductions and the oracle's Hamiltonian::total_energy."""
import numpy as np
import pytest

import vegas_rs_b200 as vg
from oracle import binding as ob
from helpers import oracle_model, random_state

pytestmark = pytest.mark.gpu

SMALL = [
    ("sc_10x10x10", dict(unitcell=vg.SC, size=(10, 10, 10))),                        # docs/metropolis.toml (config[0])
    ("sc_10x10x1", dict(unitcell=vg.SC, size=(10, 10, 1))),                          # README variant
    ("sc_open", dict(unitcell=vg.SC, size=(5, 4, 3), pbc=(False, True, False))),
    ("sc_odd_pbc", dict(unitcell=vg.SC, size=(5, 5, 3))),                            # three colours
    ("sc_literal", dict(unitcell=vg.SC, size=(4, 4, 4), literal=True)),
    ("bcc_literal", dict(unitcell=vg.BCC, size=(3, 3, 3), literal=True)),
    ("fcc_open", dict(unitcell=vg.FCC, size=(3, 3, 2), pbc=(True, False, False))),
    ("sc_20", dict(unitcell=vg.SC, size=(20, 20, 20))),                              # 8000 sites: several sites per thread
]


def n_sites(lat):
    nb = {vg.SC: 1, vg.BCC: 2, vg.FCC: 4}[lat["unitcell"]]
    return int(np.prod(lat["size"])) * nb


def pair(model, **kw):
    """(resident handle, launch-per-colour handle) of the same model and seed."""
    a = vg.GpuMetropolis(model, **kw)
    b = vg.GpuMetropolis(model, **kw)
    b.set_tuning("resident_max", 0)
    return a, b


@pytest.mark.parametrize("proposal", [vg.PROPOSE_FLIP, vg.PROPOSE_RANDOM], ids=["flip", "random"])
@pytest.mark.parametrize("name,lat", SMALL, ids=[n for n, _ in SMALL])
def test_resident_ising_identical_to_colour_passes(built, name, lat, proposal):
    a, b = pair(vg.ISING, proposal=proposal, seed=77, **lat)
    assert a.kernel_family == "ising_general" and a.step_kernel == "ising_resident" and b.step_kernel == "ising_general"
    s0 = random_state(ob.ISING, n_sites(lat), 5)
    for g in (a, b):
        g.upload(s0)
        g.set_thermostat(3.0, (0, 0, 1.0), 0.5)
    la = a.launches
    ea, ma = a.step(7)
    assert a.launches - la == 1                               # one launch for the whole batch
    eb, mb = b.step(7)
    assert np.array_equal(a.download(), b.download())
    assert np.array_equal(ea, eb) and np.array_equal(ma, mb)  # integer-valued sums: exact in any order
    a.step(5, observe=False); b.step(5, observe=False)
    ea, ma = a.step(3); eb, mb = b.step(3)
    assert np.array_equal(a.download(), b.download()) and np.array_equal(ea, eb) and np.array_equal(ma, mb)
    assert a.attempt_count() == b.attempt_count() and a.sweeps == b.sweeps == 15
    # the recorded energy is Hamiltonian::total_energy of the state after the step
    H, _ = oracle_model(ob.ISING, **lat)
    assert ea[-1] == H.total_energy(H.thermostat(3.0, (0, 0, 1.0), 0.5), a.download()) == a.total_energy()
    a.close(); b.close()


def test_resident_ising_replay_against_oracle(built):
    """config[0] lattice: every decision of a resident batch equals the oracle's replay of the reference rule."""
    lat = dict(unitcell=vg.SC, size=(10, 10, 10))
    g = vg.GpuMetropolis(vg.ISING, seed=3, **lat)
    assert g.step_kernel == "ising_resident"
    H, _ = oracle_model(ob.ISING, **lat)
    s = random_state(ob.ISING, 1000, 8)
    g.upload(s)
    g.set_thermostat(4.0, (0, 0, 1.0), 0.25)
    th = H.thermostat(4.0, (0, 0, 1.0), 0.25)
    col = g.colours()
    e, m = g.step(4)
    for sweep in range(4):
        H.replay_ising_sites(th, ob.PROPOSE_FLIP, 3, sweep, col, g.n_colours, s)
    assert np.array_equal(g.download(), s)
    assert e[-1] == H.total_energy(th, s) and m[-1, 2] == s.sum()
    g.close()


@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("name,lat", SMALL, ids=[n for n, _ in SMALL])
def test_resident_heisenberg_identical_to_colour_passes(built, name, lat, precision):
    kw = dict(anisotropy=((0.6, 0.0, 0.8), 0.25), precision=precision, seed=78, force_general=True, **lat)
    a, b = pair(vg.HEISENBERG, **kw)
    if name == "sc_20" and precision == vg.F64:   # 8000 x 24 B of spins + the neighbour table exceed one SM's shared memory
        assert a.step_kernel == "heis_general"
        a.close(); b.close()
        return
    assert a.kernel_family == "heis_general" and a.step_kernel == "heis_resident" and b.step_kernel == "heis_general"
    n = n_sites(lat)
    s0 = random_state(ob.HEISENBERG, n, 6)
    for g in (a, b):
        g.upload(s0)
        g.set_thermostat(1.5, (0, 0, 1.0), 0.7)
    ea, ma = a.step(6); eb, mb = b.step(6)
    a.step(4, observe=False); b.step(4, observe=False)
    assert np.array_equal(a.download(), b.download())         # same arithmetic on the same numbers: bitwise
    tol = 1e-12 if precision == vg.F64 else 1e-6              # the reductions differ in summation order only
    assert np.max(np.abs(ea - eb)) <= tol * 12 * n and np.max(np.abs(ma - mb)) <= tol * n
    assert a.attempt_count() == b.attempt_count()
    e, m = a.step(1)
    assert abs(e[0] - a.total_energy()) <= tol * 12 * n and np.max(np.abs(m[0] - a.magnetization())) <= tol * n
    a.close(); b.close()


def test_resident_csr_with_values(built):
    """Exchange::new(CsMat) with non-uniform couplings (src/energy.rs:171-173) on the resident path."""
    rng = np.random.default_rng(15)
    n = 500
    i = rng.integers(0, n, 600); j = rng.integers(0, n, 600)     # mean degree 2.4: a handful of colours
    keep = i != j
    i, j = i[keep], j[keep]
    w = rng.normal(size=len(i))
    m = ob.Csr.from_triplets(n, np.concatenate([i, j]), np.concatenate([j, i]), np.concatenate([w, w]))
    rp, ci, va = m.arrays()
    for model in (vg.ISING, vg.HEISENBERG):
        a, b = pair(model, csr=(rp, ci.astype(np.uint32), va), precision=vg.F64, seed=19)
        if a.n_colours > 8:
            assert a.step_kernel != "ising_resident"          # more colours than the resident plan holds
            a.close(); b.close()
            continue
        assert a.step_kernel.endswith("_resident")
        s = random_state(ob.ISING if model == vg.ISING else ob.HEISENBERG, n, 4)
        for g in (a, b):
            g.upload(s)
            g.set_thermostat(1.7, (0, 0, 1.0), 0.4)
        ea, _ = a.step(5); eb, _ = b.step(5)
        assert np.array_equal(a.download(), b.download())
        assert np.max(np.abs(ea - eb)) < 1e-9
        a.close(); b.close()


def test_resident_csr_uniform_uses_table(built):
    """A user CSR without values (uniform J, Exchange::new on a 0/1 pattern): the table variant, open chain + ring."""
    n = 257
    i = np.arange(n); j = (i + 1) % n
    m = ob.Csr.from_triplets(n, np.concatenate([i, j]), np.concatenate([j, i]), np.ones(2 * n))
    rp, ci, _ = m.arrays()
    for model in (vg.ISING, vg.HEISENBERG):
        a, b = pair(model, csr=(rp, ci.astype(np.uint32), None), exchange=0.75, precision=vg.F64, seed=23)
        assert a.step_kernel.endswith("_resident") and a.n_colours == 3     # odd ring
        s = random_state(ob.ISING if model == vg.ISING else ob.HEISENBERG, n, 9)
        for g in (a, b):
            g.upload(s)
            g.set_thermostat(0.9, (0, 0, 1.0), 0.25)
        ea, ma = a.step(9); eb, mb = b.step(9)
        assert np.array_equal(a.download(), b.download())
        assert np.max(np.abs(ea - eb)) < 1e-9 and np.max(np.abs(ma - mb)) < 1e-9
        a.close(); b.close()


def test_resident_threshold_follows_tuning_key(built):
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(10, 10, 10), seed=1)
    assert g.step_kernel == "ising_resident"
    g.set_tuning("resident_max", 999)
    assert g.step_kernel == "ising_general"
    g.close()
    big = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(30, 30, 30), seed=1)   # 27000 sites: launch per colour
    assert big.step_kernel == "ising_general"
    big.close()
