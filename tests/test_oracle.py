"""CPU tests of the oracle (oracle/): the restatement of the reference must reproduce every known answer the
reference's own unit tests pin (tests/golden/reference_known_answers.json) and the exact results derived
from the cited formulas.  No GPU."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from helpers import oracle_model

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))
EXACT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ising_4x4_exact.json")))
TERM = {"gauge": ob.TERM_GAUGE, "anisotropy": ob.TERM_ANISOTROPY, "zeeman": ob.TERM_ZEEMAN, "exchange": ob.TERM_EXCHANGE}


@pytest.mark.parametrize("case", GOLD["energy_rs_tests"], ids=[c["cite"] + "/" + c["state"] for c in GOLD["energy_rs_tests"]])
def test_reference_energy_unit_tests(case):
    state = np.tile([0.0, 0.0, 1.0 if case["state"].startswith("up") else -1.0], (10, 1))
    H = ob.Hamiltonian(ob.HEISENBERG, [TERM[t] for t in case["terms"]], gauge=case.get("gauge", 0.0), aniso_k=case.get("aniso_k", 0.0))
    th = H.thermostat(0.0, (0.0, 0.0, 1.0), case["field"])
    assert H.total_energy(th, state) == case["total_energy"]


def test_reference_state_unit_tests():
    g = GOLD["state_rs_tests"]
    H = ob.Hamiltonian(ob.ISING, [ob.TERM_ZEEMAN])
    th = H.thermostat(1.0, (0, 0, 1.0), 1.0)  # Zeeman energy(i) = dot(s, up) * 1
    assert H.energy(th, np.array([1], np.int8), 0) == g["ising_dot"]["up_up"]
    assert H.energy(th, np.array([-1], np.int8), 0) == g["ising_dot"]["up_down"]
    thd = H.thermostat(1.0, (0, 0, -1.0), 1.0)
    assert H.energy(thd, np.array([1], np.int8), 0) == g["ising_dot"]["down_up"]
    assert H.energy(thd, np.array([-1], np.int8), 0) == g["ising_dot"]["down_down"]
    mag, xyz = H.magnetization(np.ones(10, np.int8))
    assert mag == g["ising_magnetization_up10"]["magnitude"] and xyz[2] > 0
    rng = ob.OracleRng(3)
    Hh = ob.Hamiltonian(ob.HEISENBERG, [ob.TERM_GAUGE])
    s = Hh.rand_state(rng, 100)
    assert np.max(np.abs((s * s).sum(axis=1) - 1.0)) < 1e-15 * 4  # marsaglia, util.rs:21-34
    assert len(np.unique(s[:, 0])) == 100


def test_philox_known_answers():
    """Random123's kat_vectors lines for philox4x32 with 10 rounds (the library default) and with 7 (the Crush-resistant
    minimum the kernels and the replay draw with: PHILOX_ROUNDS / VO_PHILOX_ROUNDS)."""
    for k in GOLD["derived"]["philox4x32_10_kat"]:
        out = ob.philox(k["ctr"], k["key"])
        assert [f"{int(x):08x}" for x in out] == k["out"]
    for k in GOLD["derived"]["philox4x32_7_kat"]:
        out = ob.philox(k["ctr"], k["key"], rounds=7)
        assert [f"{int(x):08x}" for x in out] == k["out"]
    # the constant is the same on both sides of the parity tests
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "vegas_rs_b200", "csrc", "common.cuh")).read()
    hdr = open(os.path.join(root, "oracle", "vegas_oracle.h")).read()
    import re
    assert int(re.search(r"#define VEGAS_PHILOX_ROUNDS (\d+)", src).group(1)) == int(re.search(r"#define VO_PHILOX_ROUNDS (\d+)", hdr).group(1)) == ob.PHILOX_ROUNDS


def test_program_schedules():
    for c in GOLD["derived"]["cooldown_points"]:
        pts = ob.cooldown_points(c["tmax"], c["tmin"], c["rate"])
        assert len(pts) == c["n"]
        if "last" in c:
            assert pts[-1] == c["last"]
    h = GOLD["derived"]["hysteresis_points"]
    pts = ob.hysteresis_points(h["max_field"], h["field_step"])
    assert len(pts) == h["n"] and pts.max() == h["max"] and pts.min() == h["min"]


@pytest.mark.parametrize("L", [(4, 4, 4), (6, 4, 2)])
@pytest.mark.parametrize("field", [0.0, 0.5, 2.0])
def test_all_up_sc_known_answers(L, field):
    H, m = oracle_model(ob.ISING, unitcell=ob.SC, size=L)
    n = int(np.prod(L))
    s = np.ones(n, np.int8)
    th = H.thermostat(2.0, (0, 0, 1.0), field)
    assert np.all(H.site_energies(th, s) == -6 + field)
    assert H.total_energy(th, s) == n * (-6 + field)
    assert np.all(H.delta_energies(th, s) == 12 - 2 * field)
    He = ob.Hamiltonian(ob.ISING, [ob.TERM_EXCHANGE], m)
    assert He.total_energy(th, s) == -3 * n


@pytest.mark.parametrize("uc,z,nb", [(ob.BCC, 8, 2), (ob.FCC, 12, 4)], ids=["bcc", "fcc"])
@pytest.mark.parametrize("model", [ob.ISING, ob.HEISENBERG], ids=["ising", "heisenberg"])
def test_all_up_bcc_fcc_known_answers(uc, z, nb, model):
    """Derived known answers on the other two unit cells of src/input.rs:296-322 (z = 8, 12): all spins up, J = 1,
    hamiltonian!(Exchange, Zeeman): energy(i) = -z + |H|, compound total = N (-z + |H|) (exchange counted twice,
    App. A Q3), Exchange::total_energy alone = -z N / 2, and flipping any spin costs 2 z - 2 |H|."""
    size, field = (4, 3, 5), 0.75
    H, m = oracle_model(model, unitcell=uc, size=size)
    n = int(np.prod(size)) * nb
    s = np.ones(n, np.int8) if model == ob.ISING else np.tile([0.0, 0.0, 1.0], (n, 1))
    th = H.thermostat(2.0, (0, 0, 1.0), field)
    assert np.all(H.site_energies(th, s) == -z + field)
    assert H.total_energy(th, s) == n * (-z + field)
    assert np.all(H.delta_energies(th, s) == 2 * z - 2 * field)
    He = ob.Hamiltonian(model, [ob.TERM_EXCHANGE], m)
    assert He.total_energy(th, s) == -z * n / 2
    # one spin down: its own energy changes sign, each of its z neighbours loses two units of bond energy
    if model == ob.ISING:
        s[7] = -1
    else:
        s[7] = [0.0, 0.0, -1.0]
    e = H.site_energies(th, s)
    assert e[7] == z - field and sorted(set(e.tolist())) == sorted({z - field, -z + 2 + field, -z + field})
    assert int(np.sum(e == -z + 2 + field)) == z


def test_lattice_coordination_and_csr_shape():
    for uc, z, nb in ((ob.SC, 6, 1), (ob.BCC, 8, 2), (ob.FCC, 12, 4)):
        lat = ob.Lattice(uc, 4, 3, 5)
        assert lat.n_sites == 60 * nb and lat.n_edges == 60 * nb * z // 2
        rp, col, val = ob.Csr.from_lattice(lat, 1.5, False).arrays()
        assert np.all(np.diff(rp.astype(np.int64)) == z) and np.all(val == 1.5)
        rows = np.repeat(np.arange(lat.n_sites), z)
        assert np.all(np.diff(col.astype(np.int64))[np.diff(rows) == 0] > 0)            # sorted, no duplicates
        pairs = set(zip(rows.tolist(), col.tolist()))
        assert all((c, r) in pairs for r, c in pairs)                                    # symmetric
    # drop_* removes the bonds that cross the open boundary (input.rs:312-320)
    lat = ob.Lattice(ob.SC, 4, 4, 4, pbc=(False, True, True))
    assert lat.n_edges == 3 * 64 - 16
    # expand(.,.,1) with pbc z: self edge -> 2J diagonal (energy.rs:181-182, sprs sums duplicates)
    rp, col, val = ob.Csr.from_lattice(ob.Lattice(ob.SC, 4, 4, 1), 1.0, False).arrays()
    assert all(col[rp[i]:rp[i + 1]].tolist().count(i) == 1 for i in range(16))
    assert all(val[rp[i]:rp[i + 1]][col[rp[i]:rp[i + 1]] == i][0] == 2.0 for i in range(16))
    # extent 2 with pbc: both bonds join the same pair -> 2J entry
    rp, col, val = ob.Csr.from_lattice(ob.Lattice(ob.SC, 2, 1, 1, pbc=(True, False, False)), 1.0, False).arrays()
    assert col.tolist() == [1, 0] and val.tolist() == [2.0, 2.0]


def test_literal_from_lattice_filter():
    """energy.rs:180 keeps an edge only when source <= target: on the once-per-bond edge list this drops the
    periodic wrap bonds of an sc lattice (documented in DESIGN.md; adjacency parity is unpinned)."""
    lat = ob.Lattice(ob.SC, 4, 4, 4)
    rp_l, _, _ = ob.Csr.from_lattice(lat, 1.0, True).arrays()
    rp_o, _, _ = ob.Csr.from_lattice(ob.Lattice(ob.SC, 4, 4, 4, pbc=(False, False, False)), 1.0, False).arrays()
    assert np.array_equal(rp_l, rp_o)


def test_delta_matches_two_energy_calls_heisenberg():
    H, _ = oracle_model(ob.HEISENBERG, unitcell=ob.FCC, size=(2, 2, 2), anisotropy=((0.0, 0.6, 0.8), 0.3), gauge=1.0)
    rng = np.random.default_rng(1)
    s = rng.normal(size=(32, 3)); s /= np.linalg.norm(s, axis=1, keepdims=True)
    p = rng.normal(size=(32, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    th = H.thermostat(1.0, (0.0, 0.0, 1.0), 0.7)
    d = H.delta_energies(th, s, p)
    for i in range(32):
        t = s.copy(); t[i] = p[i]
        assert d[i] == H.energy(th, t, i) - H.energy(th, s, i)
    # compound total = sum_i energy(i): exchange counted twice (energy.rs:55-59)
    assert abs(H.total_energy(th, s) - H.site_energies(th, s).sum()) < 1e-12


def test_accumulator_and_stat_line():
    H, _ = oracle_model(ob.ISING, unitcell=ob.SC, size=(4, 4, 1), pbc=(True, True, False))
    rng = ob.OracleRng(5)
    s = H.rand_state(rng, 16)
    m = ob.Machine(H, ob.PROPOSE_FLIP, rng, s, n_sensors=2)
    assert m.cooldown(3.0, 2.0, 0.5, 10, 50) == 0
    rows, lines = m.rows(), m.stat_lines()
    assert len(rows) == 3 and [r[0] for r in rows] == [3.0, 2.5, 2.0]
    e, mag = m.observables()
    assert len(e) == 3 * 60                                      # ObservableSensor records relax AND measure steps
    es = e[10:60]
    assert abs(rows[0][2] - es.mean()) < 1e-12
    assert abs(rows[0][3] - es.var() / (16 * 9.0)) < 1e-12       # Cv = Var(E) / (N T^2), population variance
    assert len(lines[0].split(" ")) == 7 and all(len(x.split(".")[1]) == 16 for x in lines[0].split(" "))
    assert m.attempts == 3 * 60 * 16
    # program validation errors (program.rs:105-110,190-201,289-300)
    assert m.relax(0, 1.0) == 1 and m.relax(10, 0.0) == 2
    assert m.cooldown(1.0, 2.0, 0.1, 1, 1) == 3 and m.cooldown(2.0, 1.0, 0.0, 1, 1) == 4
    assert m.hysteresis(1, 1, 1.0, 0.0, 0.1) == 5 and m.hysteresis(1, 1, 1.0, 1.0, 0.0) == 6


@pytest.mark.parametrize("boundary", ["pbc", "open"])
@pytest.mark.parametrize("T", [2.0, 2.269185314213022, 4.0])
def test_oracle_metropolis_vs_exact_enumeration(T, boundary):
    """The restated MetropolisFlipIntegrator samples the Boltzmann distribution of the 4x4 model (periodic: 32 bonds;
    open: 24 bonds, the adjacency the `drop_*` calls of src/input.rs:296-322 leave)."""
    ex = [r for r in EXACT[boundary]["rows"] if abs(r["T"] - T) < 1e-9][0]
    H, m = oracle_model(ob.ISING, unitcell=ob.SC, size=(4, 4, 1), pbc=(boundary == "pbc", boundary == "pbc", False))
    He = ob.Hamiltonian(ob.ISING, [ob.TERM_EXCHANGE], m)  # physical energy: Exchange::total_energy alone
    rng = ob.OracleRng(11)
    s = He.rand_state(rng, 16)
    mach = ob.Machine(He, ob.PROPOSE_FLIP, rng, s, n_sensors=2)
    mach.set_thermostat(He.thermostat(T))
    mach.relax_for(2000)
    means_e, means_m, cvs, chis = [], [], [], []
    for _ in range(20):
        mach.m.obs_len = 0
        mach.measure_for(5000)
        e, mg = mach.observables()
        means_e.append(e.mean()); means_m.append(mg.mean())
        cvs.append(e.var() / (16 * T * T)); chis.append(mg.var() / (16 * T))   # StatSensor, src/instrument.rs:98-131
    se = lambda x: np.std(x, ddof=1) / np.sqrt(len(x))
    assert abs(np.mean(means_e) - ex["E"]) < 4 * se(means_e) + 1e-3
    assert abs(np.mean(means_m) - ex["M"]) < 4 * se(means_m) + 1e-3
    assert abs(np.mean(cvs) - ex["Cv"]) < 4 * se(cvs) + 5e-3       # block variances are biased low by O(tau / block)
    assert abs(np.mean(chis) - ex["chi"]) < 4 * se(chis) + 5e-3


def test_oracle_heisenberg_single_spin_langevin():
    """Zeeman only, reference sign +|H| s.o (energy.rs:147-151): <s.o> = -(coth(h/T) - T/h)."""
    H = ob.Hamiltonian(ob.HEISENBERG, [ob.TERM_ZEEMAN])
    rng = ob.OracleRng(2)
    s = H.rand_state(rng, 64)
    mach = ob.Machine(H, ob.PROPOSE_RANDOM, rng, s, n_sensors=0)
    th = H.thermostat(0.8, (0, 0, 1.0), 1.5)
    mach.set_thermostat(th)
    mach.relax_for(200)
    acc = []
    for _ in range(3000):
        mach.relax_for(1)
        acc.append(s[:, 2].mean())
    exact = -(1 / np.tanh(1.5 / 0.8) - 0.8 / 1.5)
    assert abs(np.mean(acc) - exact) < 5e-3


def test_oracle_heisenberg_open_chain_exact_energy():
    """Exact layer (SURVEY section 4 ii): open classical Heisenberg chain, J = 1, no field.  The partition function
    factorises over the bonds, so <s_i.s_{i+1}> = coth(J/T) - T/J and the physical energy per bond is its negative; the
    restated MetropolisIntegrator (random site, Marsaglia proposal, src/integrator.rs:66-92) over the restated
    Exchange (src/energy.rs:164-214) must reproduce it -- the reference itself has no Exchange or integrator test."""
    n, T = 48, 0.7
    lat = ob.Lattice(ob.SC, n, 1, 1, pbc=(False, False, False))
    csr = ob.Csr.from_lattice(lat, 1.0, False)
    H = ob.Hamiltonian(ob.HEISENBERG, [ob.TERM_EXCHANGE], csr)     # Exchange::total_energy alone: every bond once
    rng = ob.OracleRng(7)
    s = H.rand_state(rng, n)
    mach = ob.Machine(H, ob.PROPOSE_RANDOM, rng, s, n_sensors=2)
    mach.set_thermostat(H.thermostat(T))
    mach.relax_for(3000)
    means = []
    for _ in range(20):
        mach.m.obs_len = 0
        mach.measure_for(4000)
        e, _ = mach.observables()
        means.append(e.mean() / (n - 1))
    exact = -(1 / np.tanh(1 / T) - T)
    se = np.std(means, ddof=1) / np.sqrt(len(means))
    assert abs(np.mean(means) - exact) < 4 * se + 1e-3, (np.mean(means), exact, se)
    # the nearest-neighbour correlation measured directly on the final state series agrees as well
    assert abs((s[:-1] * s[1:]).sum(axis=1).mean() + exact) < 0.15


GPU_FIXTURES = ["ising_msc_3d_field", "ising_resident_sc10", "heis_stencil_f64", "heis_fcc_vec_f64"]


@pytest.mark.parametrize("name", GPU_FIXTURES)
def test_oracle_replays_committed_gpu_trajectories(built, name):
    """tests/golden/gpu_replay_*.npz were recorded on a B200 (tests/golden/make_gpu_replay_fixtures.py): initial state,
    per-sweep energies and final state of the CUDA kernels.  The oracle replays the same colour-ordered sweeps with the
    same Philox numbers but its own restatement of Hamiltonian::energy and of the accept rule (src/integrator.rs:77-88,
    :123-134): the trajectories must coincide -- bit for bit for Ising, to 1e-12 for fp64 Heisenberg -- and every recorded
    energy must equal Hamiltonian::total_energy (the compound's trait default, src/energy.rs:55-59) of the replayed state.
    This pins the GPU path against the CPU restatement in the CPU-only suite."""
    import json
    from helpers import oracle_model
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"gpu_replay_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    size, seed = tuple(meta["size"]), meta["seed"]
    model = ob.ISING if meta["model"] == "ising" else ob.HEISENBERG
    kw = {}
    if meta.get("anisotropy"):
        kw["anisotropy"] = (tuple(meta["anisotropy"][0]), meta["anisotropy"][1])
    H, _ = oracle_model(model, unitcell=meta["unitcell"], size=size, **kw)
    th = H.thermostat(meta["T"], (0.0, 0.0, 1.0), meta["H"])
    state = z["initial"].copy()
    colours, nc = z["colours"], meta["n_colours"]
    for sweep in range(meta["sweeps"]):
        if meta["kernel_family"] == "ising_msc":
            H.replay_ising_msc(th, ob.PROPOSE_FLIP, seed, sweep, size, state)
        elif model == ob.ISING:
            H.replay_ising_sites(th, ob.PROPOSE_FLIP, seed, sweep, colours, nc, state)
        else:
            H.replay_heisenberg(th, ob.PROPOSE_RANDOM, False, seed, sweep, colours, nc, state)
        e_ref = H.total_energy(th, state)
        if model == ob.ISING:
            assert z["energy"][sweep] == e_ref
            assert z["magnetization"][sweep, 2] == state.sum()
        else:
            assert abs(z["energy"][sweep] - e_ref) < 1e-9
            assert np.max(np.abs(z["magnetization"][sweep] - state.sum(axis=0))) < 1e-9
    if model == ob.ISING:
        assert np.array_equal(z["final"], state)
    else:
        assert np.max(np.abs(z["final"] - state)) < 1e-12
        assert np.max(np.abs(np.linalg.norm(z["final"], axis=1) - 1.0)) < 1e-12
    assert not np.array_equal(z["final"], z["initial"])       # the sweeps did move spins


@pytest.mark.parametrize("model", [ob.ISING, ob.HEISENBERG], ids=["ising", "heisenberg"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_consistency_on_random_weighted_graphs(model, seed):
    """Properties every Hamiltonian of src/energy.rs must satisfy, on random symmetric CSR graphs with per-bond values
    (Exchange::new), anisotropy and gauge: the compound total is the sum of the per-site energies (trait default,
    :55-59); delta_energies equals energy-after minus energy-before of an explicit move (integrator.rs:77-81); a sweep at
    T -> infinity accepts every proposal; at T -> 0 (clamped to EPSILON, thermostat.rs:29-40) the energy never rises."""
    rng = np.random.default_rng(seed)
    n = 60
    i = rng.integers(0, n, 150); j = rng.integers(0, n, 150)
    keep = i != j
    i, j = i[keep], j[keep]
    w = rng.integers(-8, 9, len(i)) / 4.0                                   # dyadic couplings: sums are exact
    m = ob.Csr.from_triplets(n, np.concatenate([i, j]), np.concatenate([j, i]), np.concatenate([w, w]))
    terms = [ob.TERM_EXCHANGE, ob.TERM_ZEEMAN, ob.TERM_ANISOTROPY, ob.TERM_GAUGE]
    H = ob.Hamiltonian(model, terms, m, gauge=0.5, aniso_k=0.25, aniso_axis=(0.0, 0.0, 1.0))
    orng = ob.OracleRng(seed)
    s = H.rand_state(orng, n)
    th = H.thermostat(1.5, (0, 0, 1.0), 0.75)
    e_sites = H.site_energies(th, s)
    assert abs(H.total_energy(th, s) - e_sites.sum()) < 1e-9
    assert all(H.energy(th, s, k) == e_sites[k] for k in (0, 17, n - 1))
    d = H.delta_energies(th, s)                                              # flip of every site, one at a time
    for k in (3, 29, 58):
        t = s.copy()
        t[k] = -t[k]
        assert abs((H.energy(th, t, k) - H.energy(th, s, k)) - d[k]) < 1e-12
    # T -> infinity: every attempt of a step is accepted
    hot = H.thermostat(1e300)
    proposal = ob.PROPOSE_FLIP if model == ob.ISING else ob.PROPOSE_RANDOM
    assert H.step(hot, proposal, orng, s) == n
    # T -> 0: only moves with e_new - e_old <= 0 pass, so the energy whose single-site differences the integrator uses
    # (every bond ONCE: the compound total counts the exchange twice, App. A Q3, so subtract it once) never rises
    He = ob.Hamiltonian(model, [ob.TERM_EXCHANGE], m)
    cold = H.thermostat(0.0, (0, 0, 1.0), 0.75)
    physical = lambda: H.total_energy(cold, s) - He.total_energy(cold, s)
    e_prev = physical()
    for _ in range(8):
        H.step(cold, proposal, orng, s)
        e_now = physical()
        assert e_now <= e_prev + 1e-9
        e_prev = e_now
