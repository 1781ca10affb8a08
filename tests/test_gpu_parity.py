"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): Ising bit-exact (per-site energy, dE, totals, and every sweep decision);
Heisenberg within 1e-5 relative for fp32 storage and 1e-12 for fp64.
"""
import numpy as np
import pytest

import vegas_rs_b200 as vg
from oracle import binding as ob
from helpers import oracle_model, random_state

pytestmark = pytest.mark.gpu

FIELD = dict(temperature=2.5, field_dir=(0.0, 0.0, 1.0), field_mag=0.75)

# (id, kwargs) -- lattices that exercise every kernel family and boundary rule
LATTICES = [
    ("sc_msc_3d", dict(unitcell=vg.SC, size=(256, 4, 6))),
    ("sc_msc_3d_64", dict(unitcell=vg.SC, size=(64, 6, 4))),
    ("sc_msc_2d_open_z", dict(unitcell=vg.SC, size=(256, 6, 1), pbc=(True, True, False))),
    ("sc_msc_2d_self_z", dict(unitcell=vg.SC, size=(256, 4, 1))),
    ("sc_stencil_small", dict(unitcell=vg.SC, size=(16, 4, 6))),
    ("sc_stencil_L2", dict(unitcell=vg.SC, size=(8, 2, 2))),
    ("sc_open", dict(unitcell=vg.SC, size=(5, 4, 3), pbc=(False, True, False))),
    ("sc_odd_pbc", dict(unitcell=vg.SC, size=(5, 5, 3))),
    ("sc_literal", dict(unitcell=vg.SC, size=(4, 4, 4), literal=True)),
    ("sc_10x10x10", dict(unitcell=vg.SC, size=(10, 10, 10))),
    ("bcc", dict(unitcell=vg.BCC, size=(3, 4, 3))),
    ("bcc_literal", dict(unitcell=vg.BCC, size=(3, 3, 3), literal=True)),
    ("fcc", dict(unitcell=vg.FCC, size=(3, 2, 4))),
    ("fcc_open", dict(unitcell=vg.FCC, size=(3, 3, 2), pbc=(True, False, False))),
    ("bcc_vec", dict(unitcell=vg.BCC, size=(8, 3, 4))),      # nx % 4 == 0: the 16-byte variant of the basis kernel
    ("fcc_vec", dict(unitcell=vg.FCC, size=(4, 3, 4))),
]


def n_sites(lat):
    nb = {vg.SC: 1, vg.BCC: 2, vg.FCC: 4}[lat["unitcell"]]
    return int(np.prod(lat["size"])) * nb


@pytest.mark.parametrize("name,lat", LATTICES, ids=[l[0] for l in LATTICES])
def test_ising_energies_bit_exact(built, name, lat):
    g = vg.GpuMetropolis(vg.ISING, exchange=1.0, zeeman=True, gauge=0.5, anisotropy=((0, 0, 1.0), 0.25), seed=7, **lat)
    H, _ = oracle_model(ob.ISING, exchange=1.0, zeeman=True, gauge=0.5, anisotropy=((0, 0, 1.0), 0.25), **lat)
    s = random_state(ob.ISING, n_sites(lat), 11)
    g.upload(s)
    assert np.array_equal(g.download(), s)
    assert np.array_equal(g.download_into(np.empty(n_sites(lat), np.int8)), s)
    for fmag, fdir in ((0.0, (0, 0, 1.0)), (0.75, (0, 0, 1.0)), (-1.5, (0, 0, -1.0))):
        g.set_thermostat(2.5, fdir, fmag)
        th = H.thermostat(2.5, fdir, fmag)
        assert np.array_equal(g.site_energies(), H.site_energies(th, s)), name
        assert np.array_equal(g.delta_energies(), H.delta_energies(th, s))
        prop = random_state(ob.ISING, n_sites(lat), 12)
        assert np.array_equal(g.delta_energies(prop), H.delta_energies(th, s, prop))
        g.set_energy_convention(vg.E_REFERENCE_COMPOUND)
        assert g.total_energy() == H.total_energy(th, s)
        mag, xyz = H.magnetization(s)
        assert np.array_equal(g.magnetization(), xyz)
    # Exchange alone, as `vegas bench` builds it (src/main.rs:26): total_energy halves the double count
    He = ob.Hamiltonian(ob.ISING, [ob.TERM_EXCHANGE], H.csr)
    g.set_energy_convention(vg.E_REFERENCE_EXCHANGE)
    assert g.total_energy() == He.total_energy(He.thermostat(2.5), s)
    g.close()


@pytest.mark.parametrize("precision,tol", [(vg.F64, 1e-12), (vg.F32, 1e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("name,lat", LATTICES, ids=[l[0] for l in LATTICES])
def test_heisenberg_energies(built, name, lat, precision, tol):
    kw = dict(exchange=0.8, zeeman=True, gauge=-0.3, anisotropy=((0.6, 0.0, 0.8), 0.4))
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=precision, seed=3, **kw, **lat)
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = n_sites(lat)
    s = random_state(ob.HEISENBERG, n, 21)
    g.upload(s)
    back = g.download()
    assert np.max(np.abs(back - s)) <= (0 if precision == vg.F64 else 1e-7)
    s_dev = back  # what the device actually holds (fp32-rounded for F32)
    fdir = (0.0, 0.6, 0.8)
    g.set_thermostat(1.3, fdir, 0.9)
    th = H.thermostat(1.3, fdir, 0.9)
    e_ref = H.site_energies(th, s_dev)
    scale = np.max(np.abs(e_ref)) + 1.0
    assert np.max(np.abs(g.site_energies() - e_ref)) <= tol * scale
    prop = random_state(ob.HEISENBERG, n, 22)
    assert np.max(np.abs(g.delta_energies(prop) - H.delta_energies(th, s_dev, prop))) <= tol * scale
    assert np.max(np.abs(g.delta_energies() - H.delta_energies(th, s_dev))) <= tol * scale
    g.set_energy_convention(vg.E_REFERENCE_COMPOUND)
    tot = H.total_energy(th, s_dev)
    assert abs(g.total_energy() - tot) <= tol * (abs(tot) + n)
    mag, xyz = H.magnetization(s_dev)
    assert np.max(np.abs(g.magnetization() - xyz)) <= tol * n
    g.close()


def test_csr_input_with_values(built):
    """Exchange::new(CsMat) with non-uniform couplings (src/energy.rs:171-173)."""
    rng = np.random.default_rng(5)
    n = 300
    a = rng.integers(0, n, 900); b = rng.integers(0, n, 900)
    keep = a != b
    a, b = a[keep], b[keep]
    w = rng.normal(size=len(a))
    m = ob.Csr.from_triplets(n, np.concatenate([a, b]), np.concatenate([b, a]), np.concatenate([w, w]))
    rp, ci, va = m.arrays()
    for model in (ob.ISING, ob.HEISENBERG):
        g = vg.GpuMetropolis(model, csr=(rp, ci.astype(np.uint32), va), precision=vg.F64, seed=9)
        H, _ = oracle_model(model, csr=(rp, ci, va))
        s = random_state(model, n, 31)
        g.upload(s)
        g.set_thermostat(1.7, (0, 0, 1.0), 0.4)
        th = H.thermostat(1.7, (0, 0, 1.0), 0.4)
        assert np.max(np.abs(g.site_energies() - H.site_energies(th, s))) < 1e-12 * 10
        assert abs(g.total_energy() - H.total_energy(th, s)) < 1e-9
        # colours are a proper colouring
        col = g.colours()
        rows = np.repeat(np.arange(n), np.diff(rp.astype(np.int64)))
        assert not np.any((col[rows] == col[ci.astype(np.int64)]) & (rows != ci.astype(np.int64)))
        # one sweep equals the replay of the reference rule with the same random numbers
        before = s.copy()
        g.step(1, observe=False)
        after = g.download()
        if model == ob.ISING:
            H.replay_ising_sites(th, ob.PROPOSE_FLIP, 9, 0, col, g.n_colours, before)
            assert np.array_equal(after, before)
        else:
            H.replay_heisenberg(th, ob.PROPOSE_RANDOM, False, 9, 0, col, g.n_colours, before)
            assert np.max(np.abs(after - before)) < 1e-12
        g.close()


# ------------------------------------------------------------------------------------------ sweeps
@pytest.mark.parametrize("proposal", [vg.PROPOSE_FLIP, vg.PROPOSE_RANDOM], ids=["flip", "random"])
@pytest.mark.parametrize("lat,fmag", [
    (dict(unitcell=vg.SC, size=(256, 4, 6)), 0.0),
    (dict(unitcell=vg.SC, size=(64, 6, 2)), 0.0),
    (dict(unitcell=vg.SC, size=(256, 4, 6)), 0.625),  # dyadic field: every partial sum of the reference fold is exact
    (dict(unitcell=vg.SC, size=(512, 6, 1), pbc=(True, True, False)), 0.0),
    (dict(unitcell=vg.SC, size=(256, 4, 1)), -0.375),
], ids=["3d", "3d_64_L2", "3d_field", "2d", "2d_field_selfz"])
def test_ising_msc_sweep_replay_bit_exact(built, lat, fmag, proposal):
    """Every decision of the multi-spin-coded sweep equals the reference rule evaluated by the oracle."""
    seed = 1234
    g = vg.GpuMetropolis(vg.ISING, proposal=proposal, seed=seed, **lat)
    assert g.kernel_family == "ising_msc"
    H, _ = oracle_model(ob.ISING, **lat)
    s = random_state(ob.ISING, n_sites(lat), 41)
    g.upload(s)
    cpu = s.copy()
    for T in (4.5, 2.2, 0.7):
        g.set_thermostat(T, (0, 0, 1.0), fmag)
        th = H.thermostat(T, (0, 0, 1.0), fmag)
        for _ in range(2):
            sweep = g.sweeps
            e, m = g.step(1)
            H.replay_ising_msc(th, proposal, seed, sweep, lat["size"], cpu)
            assert np.array_equal(g.download(), cpu)
            # fused observables of the last colour pass == the reference's per-step observers
            assert e[0] == H.total_energy(th, cpu)
            assert m[0, 2] == H.magnetization(cpu)[1][2]
            assert g.total_energy() == e[0]
    g.close()


@pytest.mark.parametrize("name,lat", [l for l in LATTICES if "msc" not in l[0]],
                         ids=[l[0] for l in LATTICES if "msc" not in l[0]])
def test_ising_general_sweep_replay_bit_exact(built, name, lat):
    seed = 99
    for proposal in (vg.PROPOSE_FLIP, vg.PROPOSE_RANDOM):
        g = vg.GpuMetropolis(vg.ISING, proposal=proposal, seed=seed, force_general=True, **lat)
        assert g.kernel_family == "ising_general"
        H, _ = oracle_model(ob.ISING, **lat)
        s = random_state(ob.ISING, n_sites(lat), 43)
        g.upload(s)
        cpu = s.copy()
        col = g.colours()
        g.set_thermostat(3.0, (0, 0, 1.0), 0.25)
        th = H.thermostat(3.0, (0, 0, 1.0), 0.25)
        for _ in range(3):
            sweep = g.sweeps
            e, m = g.step(1)
            H.replay_ising_sites(th, proposal, seed, sweep, col, g.n_colours, cpu)
            assert np.array_equal(g.download(), cpu), name
            assert e[0] == H.total_energy(th, cpu)
        g.close()


@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("general", [False, True], ids=["stencil", "general"])
def test_heisenberg_sweep_replay(built, precision, general):
    lat = dict(unitcell=vg.SC, size=(16, 4, 6))
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.0, 0.0, 1.0), -0.3))
    seed = 77
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=precision, seed=seed, force_general=general, **kw, **lat)
    assert g.kernel_family == ("heis_general" if general else "heis_stencil")
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    s = random_state(ob.HEISENBERG, n_sites(lat), 51)
    g.upload(s)
    cpu = g.download()
    col = g.colours()
    g.set_thermostat(1.2, (0, 0, 1.0), 0.5)
    th = H.thermostat(1.2, (0, 0, 1.0), 0.5)
    n = n_sites(lat)
    for _ in range(3):
        sweep = g.sweeps
        e, m = g.step(1)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, precision == vg.F32, seed, sweep, col, 2, cpu)
        dev = g.download()
        diff = np.max(np.abs(dev - cpu), axis=1)
        if precision == vg.F64:
            assert np.max(diff) < 1e-12
            assert abs(e[0] - H.total_energy(th, cpu)) < 1e-12 * n * 10
        else:
            # fp32 proposals/decisions can differ from the f64 replay at knife edges: allow a few sites
            bad = np.sum(diff > 1e-5)
            assert bad <= max(2, n // 200), bad
            assert abs(e[0] - H.total_energy(th, dev)) < 1e-5 * n * 6
            cpu = dev.copy()  # resynchronise so that rare flips do not cascade
        assert np.max(np.abs(np.linalg.norm(dev, axis=1) - 1.0)) < (1e-12 if precision == vg.F64 else 3e-6)
    g.close()


@pytest.mark.parametrize("name", ["bcc", "fcc", "fcc_open", "bcc_literal", "bcc_vec", "fcc_vec"])
@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
def test_heisenberg_basis_lattices_recorded_step(built, name, precision):
    """bcc / fcc: the recorded step reduces E and M inside the colour passes (each bond once, towards the lower colours);
    they must equal the dedicated reduction over the final state and the oracle's Hamiltonian::total_energy."""
    lat = dict(LATTICES)[name]
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.6, 0.0, 0.8), 0.25))
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=precision, seed=12, **kw, **lat)
    # periodic bcc / fcc take the basis-split kernel with compile-time neighbour tables, the rest the general one
    assert g.kernel_family == ("heis_basis" if name in ("bcc", "fcc", "bcc_vec", "fcc_vec") else "heis_general")
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = n_sites(lat)
    s0 = random_state(ob.HEISENBERG, n, 3)
    g.upload(s0)
    assert np.max(np.abs(g.download() - s0)) <= (0 if precision == vg.F64 else 1e-7)
    g.set_thermostat(1.5, (0, 0, 1.0), 0.7)
    th = H.thermostat(1.5, (0, 0, 1.0), 0.7)
    tol = 1e-12 if precision == vg.F64 else 1e-5
    # every decision of a sweep against the oracle's replay with Hamiltonian::energy
    cpu = g.download(); col = g.colours()
    for _ in range(2):
        sweep = g.sweeps
        g.step(1)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, precision == vg.F32, 12, sweep, col, g.n_colours, cpu)
        dev = g.download()
        diff = np.max(np.abs(dev - cpu), axis=1)
        if precision == vg.F64:
            assert np.max(diff) < 1e-12
        else:
            assert np.sum(diff > 1e-5) <= max(2, n // 200)
            cpu = dev.copy()
    for conv in (vg.E_REFERENCE_COMPOUND, vg.E_PHYSICAL):
        g.set_energy_convention(conv)
        e, m = g.step(2)
        dev = g.download()
        assert abs(e[-1] - g.total_energy()) <= tol * n * 10
        assert np.max(np.abs(m[-1] - g.magnetization())) <= tol * n
        if conv == vg.E_REFERENCE_COMPOUND:
            assert abs(e[-1] - H.total_energy(th, dev)) <= tol * n * 10
    g.close()


def test_stencil_equals_general_ising_state_evolution(built):
    """ising_msc and ising_general are both exact restatements of the rule; on a field-free run their
    equilibrium statistics must agree (different random-number mappings, so not bitwise)."""
    lat = dict(unitcell=vg.SC, size=(256, 8, 8))
    res = []
    for general in (False, True):
        g = vg.GpuMetropolis(vg.ISING, seed=5, force_general=general, **lat)
        g.randomize()
        g.set_thermostat(5.0)
        g.step(200, observe=False)
        e, m = g.step(1500)
        res.append((e.mean(), np.abs(m[:, 2]).mean(), e.std() / np.sqrt(1500 / 8)))
        g.close()
    assert abs(res[0][0] - res[1][0]) < 6 * max(res[0][2], res[1][2])


# ------------------------------------------------------------------------------------------ physics
def exact_4x4(T, pbc=True):
    """Exact enumeration of the 4x4 Ising model, physical energy convention."""
    L = 4
    idx = np.arange(16).reshape(4, 4)
    bonds = []
    for y in range(4):
        for x in range(4):
            if x + 1 < L or pbc: bonds.append((idx[y, x], idx[y, (x + 1) % L]))
            if y + 1 < L or pbc: bonds.append((idx[y, x], idx[(y + 1) % L, x]))
    states = ((np.arange(1 << 16)[:, None] >> np.arange(16)) & 1) * 2 - 1
    E = np.zeros(1 << 16)
    for a, b in bonds:
        E -= states[:, a] * states[:, b]
    M = np.abs(states.sum(axis=1))
    w = np.exp(-(E - E.min()) / T); w /= w.sum()
    return (w * E).sum(), (w * M).sum()


@pytest.mark.parametrize("T", [1.5, 2.269185314, 4.0])
def test_ising_4x4_exact_enumeration(built, T):
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(4, 4, 1), pbc=(True, True, False), seed=2024)
    g.set_energy_convention(vg.E_PHYSICAL)
    g.randomize()
    g.set_thermostat(T)
    g.step(2000, observe=False)
    es, ms = [], []
    for _ in range(50):
        e, m = g.step(4000)
        es.append(e.mean()); ms.append(np.abs(m[:, 2]).mean())
    e_exact, m_exact = exact_4x4(T)
    se_e = np.std(es, ddof=1) / np.sqrt(len(es)); se_m = np.std(ms, ddof=1) / np.sqrt(len(ms))
    assert abs(np.mean(es) - e_exact) < 4 * se_e + 1e-3, (np.mean(es), e_exact, se_e)
    assert abs(np.mean(ms) - m_exact) < 4 * se_m + 1e-3, (np.mean(ms), m_exact, se_m)
    g.close()


def test_ising_2d_onsager_magnetisation(built):
    """Multi-spin-coded 2D sweep, 512x512 pbc at T=2.0 < Tc: |m| -> (1 - sinh(2/T)^-4)^(1/8), u -> Onsager."""
    from scipy.special import ellipk
    L, T = 512, 2.0
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, 1), pbc=(True, True, False), seed=17)
    assert g.kernel_family == "ising_msc"
    g.set_energy_convention(vg.E_PHYSICAL)
    g.fill(True)
    g.set_thermostat(T)
    g.step(1500, observe=False)
    e, m = g.step(3000)
    m_exact = (1 - np.sinh(2 / T) ** -4) ** 0.125
    K = 1 / T
    k = 2 * np.sinh(2 * K) / np.cosh(2 * K) ** 2
    u_exact = -(1 / np.tanh(2 * K)) * (1 + (2 / np.pi) * (2 * np.tanh(2 * K) ** 2 - 1) * ellipk(k * k))
    assert abs(np.abs(m[:, 2]).mean() / L**2 - m_exact) < 2e-3
    assert abs(e.mean() / L**2 - u_exact) < 2e-3
    g.close()


def test_ising_2d_cooldown_peaks_at_onsager_tc(built):
    """BASELINE config[1] in miniature (north_star check 3): a CoolDown of the 2D Ising model across Tc; the specific heat
    Var(E)/(N T^2) and the susceptibility Var(|M|)/(N T) (StatSensor formulas, src/instrument.rs:98-131) must peak at
    Onsager's Tc = 2/ln(1+sqrt 2) up to the finite-size shift of a 64 x 64 lattice (Tc(L) - Tc ~ 0.4 Tc/L for Cv,
    ~ 1.1 Tc / L for chi') and the 0.05 temperature grid."""
    L, Tc = 64, 2.269185314213022
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, 1), pbc=(True, True, False), seed=41)
    g.set_energy_convention(vg.E_PHYSICAL)
    g.randomize()
    temps = [2.7 - 0.05 * i for i in range(17)]              # 2.70 ... 1.90, annealed like CoolDown (program.rs:203-211)
    cv, chi = [], []
    for T in temps:
        g.set_thermostat(T)
        g.step(5000, observe=False)
        e, m = g.step(60000)
        mag = np.abs(m[:, 2])
        cv.append(e.var() / (L * L * T * T)); chi.append(mag.var() / (L * L * T))
    t_cv, t_chi = temps[int(np.argmax(cv))], temps[int(np.argmax(chi))]
    assert abs(t_cv - Tc) <= 0.085, (t_cv, cv)
    assert -0.03 <= t_chi - Tc <= 0.135, (t_chi, chi)
    assert max(cv) > 1.5 and max(cv) > 2.0 * cv[0] and max(cv) > 2.0 * cv[-1]   # a peak, not a slope
    g.close()


@pytest.mark.parametrize("precision", [vg.F32, vg.F64], ids=["f32", "f64"])
def test_heisenberg_free_spins_langevin(built, precision):
    """No exchange: independent spins in a field, reference sign +|H| s.o (src/energy.rs:147-151)
    => <s.o> = -(coth(h/T) - T/h)."""
    h, T = 1.5, 0.8
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.SC, size=(16, 8, 8), exchange=None, zeeman=True, precision=precision, seed=8)
    g.randomize()
    g.set_thermostat(T, (0, 0, 1.0), h)
    g.step(200, observe=False)
    e, m = g.step(2000)
    mz = m[:, 2].mean() / g.n_sites
    exact = -(1 / np.tanh(h / T) - T / h)
    assert abs(mz - exact) < 3e-3, (mz, exact)
    g.close()


def test_heisenberg_stencil_vs_general_statistics(built):
    lat = dict(unitcell=vg.SC, size=(16, 16, 16))
    res = []
    for general in (False, True):
        g = vg.GpuMetropolis(vg.HEISENBERG, seed=5, force_general=general, **lat)
        g.set_energy_convention(vg.E_PHYSICAL)
        g.randomize()
        g.set_thermostat(1.0)
        g.step(300, observe=False)
        e, m = g.step(1500)
        res.append(e.mean() / g.n_sites)
        g.close()
    assert abs(res[0] - res[1]) < 5e-3, res


# ------------------------------------------------------------------------------------------ fused two-colour step
FUSED_CASES = [
    # (size, ty, cz): several y-tiles, z-chunks that do and do not divide Lz, minimal Ly = ty + 4
    ((16, 8, 6), 4, 0),
    ((16, 8, 6), 4, 2),
    ((32, 12, 8), 4, 3),
    ((24, 10, 4), 2, 1),
    ((64, 16, 6), 6, 0),
]


@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("size,ty,cz", FUSED_CASES, ids=[f"{c[0][0]}x{c[0][1]}x{c[0][2]}_ty{c[1]}_cz{c[2]}" for c in FUSED_CASES])
def test_heisenberg_fused_step_replay(built, precision, size, ty, cz):
    """heis_fused_kernel (one launch per step, both colours, ping-pong buffers) against the oracle's replay of
    the reference rule with Hamiltonian::energy, and against the two-pass kernels bit for bit."""
    lat = dict(unitcell=vg.SC, size=size)
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.0, 0.6, 0.8), -0.3))
    seed = 4242
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=precision, seed=seed, **kw, **lat)
    g.set_tuning("heis_fused", 1); g.set_tuning("heis_fused_ty", ty); g.set_tuning("heis_fused_cz", cz)
    assert g.step_kernel == "heis_fused"
    two = vg.GpuMetropolis(vg.HEISENBERG, precision=precision, seed=seed, **kw, **lat)
    two.set_tuning("heis_fused", 0)
    assert two.step_kernel == "heis_stencil"
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = n_sites(lat)
    s = random_state(ob.HEISENBERG, n, 5)
    g.upload(s); two.upload(s)
    cpu = g.download()
    col = g.colours()
    for T, hmag in ((1.2, 0.5), (0.4, 0.0)):
        g.set_thermostat(T, (0, 0.6, 0.8), hmag); two.set_thermostat(T, (0, 0.6, 0.8), hmag)
        th = H.thermostat(T, (0, 0.6, 0.8), hmag)
        for _ in range(3):
            sweep = g.sweeps
            e, m = g.step(1)
            e2, m2 = two.step(1)
            dev = g.download()
            assert np.array_equal(dev, two.download())          # same keys, same arithmetic order
            assert abs(e[0] - e2[0]) <= 1e-6 * n and np.max(np.abs(m[0] - m2[0])) <= 1e-6 * n
            H.replay_heisenberg(th, ob.PROPOSE_RANDOM, precision == vg.F32, seed, sweep, col, 2, cpu)
            diff = np.max(np.abs(dev - cpu), axis=1)
            if precision == vg.F64:
                assert np.max(diff) < 1e-12
                assert abs(e[0] - H.total_energy(th, cpu)) < 1e-12 * n * 10
                assert np.max(np.abs(m[0] - cpu.sum(axis=0))) < 1e-12 * n
            else:
                assert np.sum(diff > 1e-5) <= max(2, n // 200)
                assert abs(e[0] - H.total_energy(th, dev)) < 1e-5 * n * 6
                assert np.max(np.abs(m[0] - dev.sum(axis=0))) < 1e-5 * n
                cpu = dev.copy()
        a1, acc1 = g.attempt_count(); a2, acc2 = two.attempt_count()
        assert (a1, acc1) == (a2, acc2)
    # unrecorded steps take the same trajectory, and the measure-only path reads the current buffers
    g.step(3, observe=False); two.step(3, observe=False)
    assert np.array_equal(g.download(), two.download())
    assert abs(g.total_energy() - two.total_energy()) <= 1e-9 * n
    g.close(); two.close()


def test_heisenberg_wave_chunked_passes_identical(built):
    """tuning key heis_wave_c: colour passes interleaved in z-chunks (L2 reuse experiment) give the same trajectory."""
    lat = dict(unitcell=vg.SC, size=(16, 8, 12))
    res = []
    for c in (0, 2, 3, 5):
        g = vg.GpuMetropolis(vg.HEISENBERG, precision=vg.F32, seed=5, anisotropy=((0, 0, 1.0), 0.1), **lat)
        g.set_tuning("heis_wave_c", c)
        g.randomize(); g.set_thermostat(0.9, (0, 0, 1.0), 0.3)
        e, m = g.step(3)
        res.append((g.download(), e))
        g.close()
    for d, e in res[1:]:
        assert np.array_equal(d, res[0][0]) and np.allclose(e, res[0][1], rtol=1e-6)


@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
def test_heisenberg_wave_kernel_identical(built, precision):
    """heis_wave_kernel (both colour passes in one persistent launch, dependency-counted wave order) takes the same
    trajectory as the two separate passes, for several chunk sizes / lags, recorded and unrecorded steps."""
    lat = dict(unitcell=vg.SC, size=(32, 12, 12))
    kw = dict(precision=precision, seed=21, anisotropy=((0, 0.6, 0.8), 0.2))
    ref = vg.GpuMetropolis(vg.HEISENBERG, **kw, **lat)
    ref.randomize(); ref.set_thermostat(0.8, (0, 0, 1.0), 0.4)
    e0, m0 = ref.step(3)
    ref.step(2, observe=False)
    e1, m1 = ref.step(1)
    want = ref.download(); acc = ref.attempt_count()
    ref.close()
    # (planes per chunk, lag in chunks, steps fused into one launch)
    for planes, lag, k in ((1, 1, 1), (2, 1, 1), (3, 2, 1), (4, 2, 1), (5, 7, 1), (1, 3, 2), (2, 4, 2), (1, 3, 3), (3, 5, 4), (1, 4, 4),
                           (6, 3, 2)):
        g = vg.GpuMetropolis(vg.HEISENBERG, **kw, **lat)
        g.set_tuning("heis_wave", 1); g.set_tuning("heis_wave_planes", planes); g.set_tuning("heis_wave_lag", lag)
        g.set_tuning("heis_wave_steps", k)
        assert g.step_kernel == "heis_wave"
        g.randomize(); g.set_thermostat(0.8, (0, 0, 1.0), 0.4)
        e, m = g.step(3)
        g.step(2, observe=False)
        e2, m2 = g.step(1)
        g.synchronize()
        assert np.array_equal(g.download(), want), (planes, lag, k)
        assert np.allclose(e, e0, rtol=1e-6) and np.allclose(e2, e1, rtol=1e-6) and np.allclose(m, m0, rtol=1e-5, atol=1e-3)
        assert g.attempt_count() == acc
        g.close()


@pytest.mark.parametrize("k", [2, 3, 4])
def test_heisenberg_wave_multi_step_launch_identical(built, k):
    """Several steps fused into one persistent launch (2k colour-pass phases in rotated wave order, many tiles per
    chunk, all CTAs waiting on each other's chunk counters): same trajectory and the same per-step E, M rows as single
    steps; a batch that is not a multiple of k ends with a shorter launch."""
    lat = dict(unitcell=vg.SC, size=(128, 64, 24))
    kw = dict(precision=vg.F32, seed=33, anisotropy=((0, 0, 1.0), 0.1))
    res = []
    for steps_per_launch in (1, k):
        g = vg.GpuMetropolis(vg.HEISENBERG, **kw, **lat)
        g.set_tuning("heis_wave", 1); g.set_tuning("heis_wave_planes", 2); g.set_tuning("heis_wave_lag", 3)
        g.set_tuning("heis_wave_steps", steps_per_launch)
        assert g.step_kernel == "heis_wave"
        g.randomize(); g.set_thermostat(1.2, (0, 0, 1.0), 0.5)
        e, m = g.step(2 * k + 1)
        g.step(k, observe=False)
        g.synchronize()
        res.append((g.download(), e, m, g.attempt_count(), g.launches))
        g.close()
    (s1, e1, m1, a1, l1), (sk, ek, mk, ak, lk) = res
    assert np.array_equal(s1, sk)
    assert np.allclose(e1, ek, rtol=1e-6) and np.allclose(m1, mk, rtol=1e-5, atol=1e-2)
    assert a1 == ak
    assert lk < l1                                            # fewer launches for the same steps


def test_heisenberg_fused_flip_proposal_and_larger(built):
    """Flip proposal (MetropolisFlipIntegrator, src/integrator.rs:109-138) and an auto-planned tile on 64x64x32."""
    lat = dict(unitcell=vg.SC, size=(64, 64, 32))
    for proposal in (vg.PROPOSE_FLIP, vg.PROPOSE_RANDOM):
        pair = []
        for fused in (1, 0):
            g = vg.GpuMetropolis(vg.HEISENBERG, precision=vg.F32, seed=9, proposal=proposal, anisotropy=((0, 0, 1.0), 0.1), **lat)
            g.set_tuning("heis_fused", fused)
            # 32 planes: without the fused kernel the auto choice is the pipelined TMA kernel (same trajectory)
            assert g.step_kernel == ("heis_fused" if fused else "heis_pipe")
            g.randomize()
            g.set_thermostat(1.0, (0, 0, 1.0), 1.0)
            e, m = g.step(4)
            pair.append((g.download(), e, m, g.attempt_count()))
            g.close()
        assert np.array_equal(pair[0][0], pair[1][0])
        assert np.allclose(pair[0][1], pair[1][1], rtol=1e-6, atol=1e-3) and np.allclose(pair[0][2], pair[1][2], rtol=1e-6, atol=1e-2)
        assert pair[0][3] == pair[1][3]


@pytest.mark.parametrize("uc,size", [(vg.FCC, (16, 6, 5)), (vg.BCC, (8, 7, 6)), (vg.FCC, (4, 2, 2))], ids=["fcc", "bcc", "fcc_tiny"])
@pytest.mark.parametrize("precision", [vg.F32, vg.F64], ids=["f32", "f64"])
def test_heisenberg_basis_vector_kernel_identical(built, uc, size, precision):
    """heis_basis_vec_kernel (16-byte loads, shifted rows with a scalar carry) walks the neighbour table in the order of
    the scalar kernel: same sums, same random numbers, the same trajectory bit for bit; recorded E, M agree to rounding."""
    res = []
    for vec in (1, 0):
        g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=uc, size=size, precision=precision, seed=55, anisotropy=((0.6, 0, 0.8), 0.3))
        assert g.kernel_family == "heis_basis"
        g.set_tuning("basis_vec", vec)
        g.randomize(); g.set_thermostat(2.0, (0, 0, 1.0), 0.4)
        e, m = g.step(4)
        g.step(3, observe=False)
        res.append((g.download(), e, m, g.attempt_count(), g.total_energy()))
        g.close()
    (s1, e1, m1, a1, t1), (s0, e0, m0, a0, t0) = res
    assert np.array_equal(s1, s0) and a1 == a0
    tol = 1e-5 if precision == vg.F32 else 1e-12
    assert np.allclose(e1, e0, rtol=tol, atol=tol * len(s0)) and np.allclose(m1, m0, rtol=tol, atol=tol * len(s0))
    assert abs(t1 - t0) <= tol * len(s0) * 12


# ------------------------------------------------------------------------------------------ slabs
SLAB_KINDS = {
    # kind: (model, unitcell, (nx, ny, nz) in cells, sites per cell)
    "ising": (vg.ISING, vg.SC, (256, 4, 8), 1),
    "heisenberg": (vg.HEISENBERG, vg.SC, (16, 4, 8), 1),
    "heisenberg_bcc": (vg.HEISENBERG, vg.BCC, (5, 3, 8), 2),
    "heisenberg_fcc": (vg.HEISENBERG, vg.FCC, (4, 3, 8), 4),       # nx % 4 == 0: 16-byte variant (peer stores as vectors)
    "heisenberg_fcc_scalar": (vg.HEISENBERG, vg.FCC, (6, 3, 8), 4),
    "heisenberg_bcc_vec": (vg.HEISENBERG, vg.BCC, (12, 5, 8), 2),
}


@pytest.mark.parametrize("kind", list(SLAB_KINDS))
@pytest.mark.parametrize("nslab", [2, 4])
def test_slab_decomposition_bit_identical(built, kind, nslab):
    """z-slabs with peer-written halos reproduce the single-handle run bit for bit (same Philox keys): sc stencil
    kernels (Ising, Heisenberg) and the bcc / fcc basis kernel."""
    model, uc, (Lx, Ly, Lz), nb = SLAB_KINDS[kind]
    kw = dict(seed=314, precision=vg.F32)
    whole = vg.GpuMetropolis(model, unitcell=uc, size=(Lx, Ly, Lz), **kw)
    s = random_state(model, Lx * Ly * Lz * nb, 61)
    whole.upload(s)
    nz = Lz // nslab
    slabs = [vg.GpuMetropolis(model, unitcell=uc, size=(Lx, Ly, nz), nz_global=Lz, z_offset=r * nz, **kw) for r in range(nslab)]
    assert all(sl.kernel_family == whole.kernel_family for sl in slabs)
    plane = Lx * Ly * nb
    for r, sl in enumerate(slabs):
        sl.upload(s[r * nz * plane:(r + 1) * nz * plane])
    for r, sl in enumerate(slabs):
        sl.slab_connect_local(slabs[(r - 1) % nslab], slabs[(r + 1) % nslab])
    for T in (4.0, 1.5):
        whole.set_thermostat(T, (0, 0, 1.0), 0.2)
        for sl in slabs:
            sl.set_thermostat(T, (0, 0, 1.0), 0.2)
        e, m = whole.step(3)
        for _ in range(3):
            for sl in slabs:
                sl.step_async(1, True)
            parts = [sl.read_observables(1) for sl in slabs]
        ref = whole.download()
        got = np.concatenate([sl.download() for sl in slabs])
        assert np.array_equal(ref, got)
        e_sum = sum(p[0][0] for p in parts)
        assert abs(e_sum - e[-1]) <= (1e-14 * abs(e[-1]) if model == vg.ISING else 1e-6 * abs(e[-1]) + 1e-6)   # Ising: integer sums, one rounding of the field term per slab
        m_sum = sum(p[1][0] for p in parts)
        assert np.max(np.abs(m_sum - m[-1])) <= (0 if model == vg.ISING else 1e-4)
        # the measure-only entry points of a slab reduce its own sites (halo planes are read, not counted)
        e_now = sum(sl.total_energy() for sl in slabs)
        assert abs(e_now - whole.total_energy()) <= (1e-14 * abs(e[-1]) if model == vg.ISING else 1e-6 * abs(e[-1]) + 1e-6)
    # a fresh random state keyed by the GLOBAL site index is the same with and without slabs
    whole.randomize()
    for sl in slabs:
        sl.randomize()
    assert np.array_equal(whole.download(), np.concatenate([sl.download() for sl in slabs]))
    for sl in slabs:
        sl.close()
    whole.close()


# ------------------------------------------------------------------------------------------ full sizes
def test_full_size_ising_2d_8192(built):
    """BASELINE config 2 size: known-answer and conservation properties (no O(N) host oracle)."""
    L = 8192
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, 1), pbc=(True, True, False), seed=1)
    N = L * L
    g.fill(True)
    g.set_thermostat(2.0, (0, 0, 1.0), 0.5)
    assert g.total_energy() == N * (-4 + 0.5)            # all-up compound total = N(-z + |H|)
    assert g.magnetization()[2] == N
    g.set_thermostat(2.269185314)
    g.randomize()
    assert abs(g.magnetization()[2]) < 6 * np.sqrt(N)
    e, m = g.step(3)
    assert g.total_energy() == e[-1] and g.magnetization()[2] == m[-1, 2]
    assert e[-1] < e[0] < 0                              # quench from T=inf lowers the energy
    a, acc = g.attempt_count()
    assert a == 3 * N and 0 < acc < a
    g.close()


def test_full_size_ising_3d_1024(built):
    L = 1024
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, L), seed=2)
    N = L**3
    g.fill(False)
    g.set_thermostat(4.5, (0, 0, 1.0), 0.25)
    assert g.total_energy() == N * (-6 - 0.25)           # downs: Zeeman energy() is +|H| s.o
    g.randomize()
    g.set_thermostat(4.5)
    e, m = g.step(2)
    assert g.total_energy() == e[-1] and g.magnetization()[2] == m[-1, 2]
    g.close()


def test_full_size_heisenberg_512(built):
    L = 512
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.SC, size=(L, L, L), anisotropy=((0, 0, 1.0), 0.1), seed=3)
    N = L**3
    g.fill(True)
    g.set_thermostat(1.0, (0, 0, 1.0), 1.0)
    assert abs(g.total_energy() - N * (-6 + 1.0 + 0.1)) < 1e-5 * N
    g.randomize()
    e, m = g.step(2)
    assert abs(g.total_energy() - e[-1]) < 1e-5 * abs(e[-1]) + 1e-5 * N
    assert np.max(np.abs(g.magnetization() - m[-1])) < 1e-5 * N
    g.close()


def test_full_size_heisenberg_fcc_384(built):
    """BASELINE config[4] size: fcc 384^3 cells = 226,492,416 sites, z = 12, four colours (heis_basis kernel)."""
    L = 384
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(L, L, L), seed=4)
    assert g.kernel_family == "heis_basis" and g.n_colours == 4
    N = 4 * L**3
    assert g.n_sites == N
    g.fill(True)
    g.set_thermostat(3.2, (0, 0, 1.0), 0.5)
    # all-up: energy(i) = -12 J + |H| (Exchange fold over z = 12 neighbours, Zeeman +|H| s.o); compound total = sum_i
    assert abs(g.total_energy() - N * (-12 + 0.5)) < 1e-5 * N
    g.set_energy_convention(vg.E_PHYSICAL)
    assert abs(g.total_energy() - N * (-6 - 0.5)) < 1e-5 * N
    assert abs(g.magnetization()[2] - N) < 1e-5 * N
    g.set_energy_convention(vg.E_REFERENCE_COMPOUND)
    g.randomize()
    assert np.max(np.abs(g.magnetization())) < 6 * np.sqrt(N)
    e, m = g.step(2)
    assert abs(g.total_energy() - e[-1]) < 1e-5 * abs(e[-1]) + 1e-5 * N
    assert np.max(np.abs(g.magnetization() - m[-1])) < 1e-5 * N
    assert e[-1] < e[0] < 0                                   # quench from T = inf towards T = 3.2
    a, acc = g.attempt_count()
    assert a == 2 * N and 0 < acc < a
    g.close()


def test_runtime_error_paths_return_status_and_keep_the_handle_usable(built):
    """Misuse returns a status with a message (never aborts, never touches the state) and the handle keeps working."""
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(64, 4, 4), seed=9)
    s = random_state(ob.ISING, 1024, 2)
    g.upload(s)
    for bad, text in ((lambda: g.upload(s[:100]), "wrong model or size"),
                      (lambda: g._check(g._lib.vegas_gpu_upload_heisenberg(g._h, vg.gpu_metropolis._ptr(np.zeros((1024, 3))), 1024)),
                       "upload_heisenberg: wrong model or size"),
                      (lambda: g.set_thermostat(float("nan")), "temperature is NaN"),
                      (lambda: g.step_async(5000, True), "at most 4096 steps"),
                      (lambda: g.set_tuning("no_such_key", 1), "unknown tuning key"),
                      (lambda: g.slab_export(), "not a slab")):
        with pytest.raises(vg.VegasGpuError) as ei:
            bad()
        assert text in str(ei.value) and ei.value.code in (-1, -4)
    assert np.array_equal(g.download(), s)                    # untouched by the failed calls
    g.set_thermostat(0.0)                                     # Thermostat clamps T to f64::EPSILON (thermostat.rs:29-40)
    e, m = g.step(2)
    assert e[-1] <= e[0] and g.attempt_count()[0] == 2 * 1024
    g.close()
    h = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.SC, size=(16, 4, 4), seed=9)
    with pytest.raises(vg.VegasGpuError):                     # Ising state into a Heisenberg handle
        h._check(h._lib.vegas_gpu_upload_ising(h._h, vg.gpu_metropolis._ptr(s[:256].copy()), 256))
    with pytest.raises(vg.VegasGpuError):
        h.ising_thresholds()
    h.close()
    with pytest.raises(vg.VegasGpuError) as ei:               # z-slab of a lattice the stencil path cannot take
        vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(10, 10, 5), nz_global=10, z_offset=0)
    assert "z-slab decomposition needs" in str(ei.value)
