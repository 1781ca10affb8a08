"""K3p, the phase-pipelined TMA Heisenberg kernel (vegas_rs_b200/csrc/heis_pipe.cu): same trajectories as the two colour
passes it replaces (src/integrator.rs:66-92, :109-138), decision-by-decision equal to the oracle replay, all through the
C ABI.  GPU only."""
import numpy as np
import pytest

import vegas_rs_b200 as vg
from oracle import binding as ob
from helpers import oracle_model, random_state

pytestmark = pytest.mark.gpu


def _two_pass(**kw):
    g = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    g.set_tuning("heis_pipe", 0); g.set_tuning("heis_wave", 0)
    assert g.step_kernel == "heis_stencil"
    return g


def _pipe(tiles=0, stages=0, own=0, vec=0, **kw):
    g = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    g.set_tuning("heis_pipe", 1)
    if vec:
        g.set_tuning("heis_pipe_vec", vec)
    if tiles:
        g.set_tuning("heis_pipe_tiles", tiles)
    if stages:
        g.set_tuning("heis_pipe_stages", stages); g.set_tuning("heis_pipe_own", own)
    assert g.step_kernel == "heis_pipe"
    return g


# (size, precision, band count, other-ring stages, own-ring stages, sites per thread): even and uneven bands (14 rows over
# 4 bands = 4+4+3+3), one-row bands, a single band (its own y-neighbour), minimal and deep rings, more planes than ring
# slots, whole and half 16-byte vectors per thread
CASES = [
    ((64, 14, 12), vg.F32, 4, 0, 0, 0),
    ((64, 14, 12), vg.F32, 4, 4, 1, 4),
    ((64, 12, 9 + 1), vg.F32, 12, 5, 2, 2),
    ((64, 8, 8), vg.F32, 1, 4, 2, 4),
    ((128, 30, 16), vg.F32, 0, 0, 0, 4),
    ((128, 230, 8), vg.F32, 0, 6, 3, 2),
    ((256, 36, 40), vg.F32, 3, 5, 4, 2),
    ((32, 14, 12), vg.F64, 4, 0, 0, 0),
    ((64, 10, 20), vg.F64, 3, 4, 1, 2),
    ((64, 10, 20), vg.F64, 5, 6, 3, 1),
]


@pytest.mark.parametrize("size,precision,tiles,stages,own,vec", CASES)
@pytest.mark.parametrize("proposal", [vg.PROPOSE_RANDOM, vg.PROPOSE_FLIP], ids=["random", "flip"])
def test_pipe_kernel_identical_to_two_passes(built, size, precision, tiles, stages, own, vec, proposal):
    kw = dict(unitcell=vg.SC, size=size, precision=precision, seed=21, anisotropy=((0, 0.6, 0.8), 0.2), proposal=proposal)
    ref = _two_pass(**kw)
    ref.randomize(); ref.set_thermostat(0.8, (0, 0, 1.0), 0.4)
    e0, m0 = ref.step(3)
    ref.step(2, observe=False)
    e1, m1 = ref.step(1)
    want = ref.download(); acc = ref.attempt_count()
    ref.close()
    g = _pipe(tiles, stages, own, vec, **kw)
    g.randomize(); g.set_thermostat(0.8, (0, 0, 1.0), 0.4)
    e, m = g.step(3)
    g.step(2, observe=False)
    e2, m2 = g.step(1)
    g.synchronize()
    assert np.array_equal(g.download(), want)
    n = size[0] * size[1] * size[2]
    tol = 1e-12 if precision == vg.F64 else 1e-6
    assert np.allclose(e, e0, rtol=tol, atol=tol * n) and np.allclose(e2, e1, rtol=tol, atol=tol * n)
    assert np.allclose(m, m0, rtol=10 * tol, atol=10 * tol * n) and np.allclose(m2, m1, rtol=10 * tol, atol=10 * tol * n)
    assert g.attempt_count() == acc
    g.close()


@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
def test_pipe_kernel_replays_the_reference_rule(built, precision):
    """Every decision of the pipelined step against the oracle: dE from its restatement of Hamiltonian::energy
    (src/energy.rs:63-214), accept rule of src/integrator.rs:82-88, same Philox numbers; fused E and M equal
    total_energy / magnetization of the replayed state."""
    size = (32, 6, 10) if precision == vg.F64 else (64, 6, 10)
    lat = dict(unitcell=vg.SC, size=size)
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.0, 0.0, 1.0), -0.3))
    seed = 77
    g = _pipe(3, precision=precision, seed=seed, **kw, **lat)
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = size[0] * size[1] * size[2]
    g.upload(random_state(ob.HEISENBERG, n, 51))
    cpu = g.download(); col = g.colours()
    g.set_thermostat(1.2, (0, 0, 1.0), 0.5)
    th = H.thermostat(1.2, (0, 0, 1.0), 0.5)
    for _ in range(3):
        sweep = g.sweeps
        e, m = g.step(1)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, precision == vg.F32, seed, sweep, col, 2, cpu)
        dev = g.download()
        diff = np.max(np.abs(dev - cpu), axis=1)
        if precision == vg.F64:
            assert np.max(diff) < 1e-12
            assert abs(e[0] - H.total_energy(th, cpu)) < 1e-12 * n * 10
            assert np.max(np.abs(m[0] - cpu.sum(axis=0))) < 1e-12 * n
        else:
            assert np.sum(diff > 1e-5) <= max(2, n // 200)
            assert abs(e[0] - H.total_energy(th, dev)) < 1e-5 * n * 6
            cpu = dev.copy()
    g.close()


def test_pipe_kernel_is_the_default_for_big_lattices(built):
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.SC, size=(128, 64, 32), seed=5)
    assert g.step_kernel == "heis_pipe"
    g.randomize(); g.set_thermostat(1.0, (0, 0, 1.0), 0.3)
    e, m = g.step(4)
    assert abs(g.total_energy() - e[-1]) < 1e-5 * abs(e[-1]) + 1e-5 * g.n_sites
    assert np.max(np.abs(g.magnetization() - m[-1])) < 1e-5 * g.n_sites
    g.close()
    # lattices the kernel cannot take (row of 48 bytes) keep the older kernels
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.SC, size=(24, 64, 32), seed=5)
    assert g.step_kernel in ("heis_wave", "heis_stencil")
    g.close()


# ---------------------------------------------------------------------------------------------------------------------
# K4p: all colour passes of a periodic bcc / fcc step in one phase-pipelined launch (vegas_rs_b200/csrc/basis_pipe.cu)
# ---------------------------------------------------------------------------------------------------------------------
BASIS_CASES = [
    (vg.FCC, (16, 6, 9), vg.F32, 0, 0, 0),
    (vg.FCC, (16, 14, 12), vg.F32, 4, 6, 1),        # uneven bands (4 + 4 + 3 + 3 rows), minimal lead
    (vg.FCC, (8, 9, 8), vg.F64, 9, 0, 2),           # one-row bands, progress published every second plane
    (vg.FCC, (4, 2, 8), vg.F32, 1, 0, 0),           # a single band: its own y-neighbour
    (vg.BCC, (8, 7, 10), vg.F32, 0, 0, 0),
    (vg.BCC, (6, 10, 16), vg.F64, 3, 0, 0),
    (vg.FCC, (384, 40, 8), vg.F32, 0, 0, 0),        # config[4] rows: two rounds of items per thread
]


@pytest.mark.parametrize("uc,size,precision,tiles,lead,pub", BASIS_CASES)
@pytest.mark.parametrize("proposal", [vg.PROPOSE_RANDOM, vg.PROPOSE_FLIP], ids=["random", "flip"])
def test_basis_pipe_identical_to_colour_passes(built, uc, size, precision, tiles, lead, pub, proposal):
    kw = dict(unitcell=uc, size=size, precision=precision, seed=31, anisotropy=((0.6, 0, 0.8), 0.15), proposal=proposal)
    ref = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    ref.set_tuning("basis_pipe", 0); ref.set_tuning("basis_wave", 0); ref.set_tuning("basis_pair", 0)
    assert ref.step_kernel == "heis_basis"
    ref.randomize(); ref.set_thermostat(1.4, (0, 0, 1.0), 0.4)
    e0, m0 = ref.step(3)
    ref.step(2, observe=False)
    e1, m1 = ref.step(1)
    want = ref.download(); acc = ref.attempt_count()
    ref.close()
    g = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    g.set_tuning("basis_pipe", 1)
    for k, v in (("basis_pipe_tiles", tiles), ("basis_pipe_lead", lead), ("basis_pipe_pub", pub)):
        if v:
            g.set_tuning(k, v)
    assert g.step_kernel == "basis_pipe"
    g.randomize(); g.set_thermostat(1.4, (0, 0, 1.0), 0.4)
    e, m = g.step(3)
    g.step(2, observe=False)
    e2, m2 = g.step(1)
    g.synchronize()
    assert np.array_equal(g.download(), want)
    n = g.n_sites
    tol = 1e-12 if precision == vg.F64 else 1e-6
    assert np.allclose(e, e0, rtol=tol, atol=tol * n) and np.allclose(e2, e1, rtol=tol, atol=tol * n)
    assert np.allclose(m, m0, rtol=10 * tol, atol=10 * tol * n) and np.allclose(m2, m1, rtol=10 * tol, atol=10 * tol * n)
    assert g.attempt_count() == acc
    g.close()


def test_basis_pipe_replays_the_reference_rule(built):
    """fcc, fp64: every decision of the pipelined step against the oracle replay (Hamiltonian::energy of src/energy.rs,
    accept rule of src/integrator.rs:82-88); the fused E and M equal total_energy / magnetization of the replayed state."""
    lat = dict(unitcell=vg.FCC, size=(8, 6, 8))
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.6, 0.0, 0.8), 0.25))
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=vg.F64, seed=12, **kw, **lat)
    g.set_tuning("basis_pipe", 1)
    assert g.step_kernel == "basis_pipe"
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = g.n_sites
    g.upload(random_state(ob.HEISENBERG, n, 3))
    cpu = g.download(); col = g.colours()
    g.set_thermostat(1.5, (0, 0, 1.0), 0.7)
    th = H.thermostat(1.5, (0, 0, 1.0), 0.7)
    for _ in range(3):
        sweep = g.sweeps
        e, m = g.step(1)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, False, 12, sweep, col, g.n_colours, cpu)
        assert np.max(np.abs(g.download() - cpu)) < 1e-12
        assert abs(e[0] - H.total_energy(th, cpu)) < 1e-12 * n * 10
        assert np.max(np.abs(m[0] - cpu.sum(axis=0))) < 1e-12 * n
    g.close()


# ---------------------------------------------------------------------------------------------------------------------
# K4w: the colour passes of a periodic bcc / fcc step as one persistent launch in wave order (vegas_rs_b200/csrc/basis_wave.cu)
# ---------------------------------------------------------------------------------------------------------------------
# (unit cell, size, precision, slots between colours, items per thread, CTA cap): many tiles per unit, one tile per unit, fewer
# CTAs than items of a slot (every CTA walks several units, the waits are real), minimal and wide lags, ragged last tile
WAVE_CASES = [
    (vg.FCC, (16, 6, 14), vg.F32, 0, 0, 0),
    (vg.FCC, (64, 40, 12), vg.F32, 1, 1, 0),
    (vg.FCC, (64, 40, 16), vg.F32, 2, 1, 3),
    (vg.FCC, (64, 40, 16), vg.F32, 1, 2, 1),         # a single CTA: pure item order
    (vg.FCC, (8, 9, 20), vg.F64, 3, 1, 5),
    (vg.FCC, (4, 2, 8), vg.F32, 1, 0, 0),
    (vg.BCC, (8, 7, 10), vg.F32, 0, 0, 0),
    (vg.BCC, (36, 10, 16), vg.F64, 2, 1, 7),
    (vg.FCC, (384, 40, 12), vg.F32, 0, 0, 0),        # config[4] rows
]


@pytest.mark.parametrize("uc,size,precision,lag,ipt,grid", WAVE_CASES)
@pytest.mark.parametrize("proposal", [vg.PROPOSE_RANDOM, vg.PROPOSE_FLIP], ids=["random", "flip"])
def test_basis_wave_identical_to_colour_passes(built, uc, size, precision, lag, ipt, grid, proposal):
    kw = dict(unitcell=uc, size=size, precision=precision, seed=33, anisotropy=((0.6, 0, 0.8), 0.15), proposal=proposal)
    ref = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    ref.set_tuning("basis_wave", 0); ref.set_tuning("basis_pair", 0)
    assert ref.step_kernel == "heis_basis"
    ref.randomize(); ref.set_thermostat(1.4, (0, 0, 1.0), 0.4)
    e0, m0 = ref.step(3)
    ref.step(2, observe=False)
    e1, m1 = ref.step(1)
    want = ref.download(); acc = ref.attempt_count()
    ref.close()
    g = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    g.set_tuning("basis_wave", 1)
    for k, v in (("basis_wave_lag", lag), ("basis_wave_ipt", ipt), ("basis_wave_grid", grid)):
        if v:
            g.set_tuning(k, v)
    assert g.step_kernel == "basis_wave"
    g.randomize(); g.set_thermostat(1.4, (0, 0, 1.0), 0.4)
    e, m = g.step(3)
    g.step(2, observe=False)
    e2, m2 = g.step(1)
    g.synchronize()
    assert np.array_equal(g.download(), want)
    n = g.n_sites
    tol = 1e-12 if precision == vg.F64 else 1e-6
    assert np.allclose(e, e0, rtol=tol, atol=tol * n) and np.allclose(e2, e1, rtol=tol, atol=tol * n)
    assert np.allclose(m, m0, rtol=10 * tol, atol=10 * tol * n) and np.allclose(m2, m1, rtol=10 * tol, atol=10 * tol * n)
    assert g.attempt_count() == acc
    g.close()


def test_basis_wave_replays_the_reference_rule(built):
    """fcc, fp64: every decision of the wave-ordered step against the oracle replay (Hamiltonian::energy of src/energy.rs,
    accept rule of src/integrator.rs:82-88); the fused E and M equal total_energy / magnetization of the replayed state."""
    lat = dict(unitcell=vg.FCC, size=(8, 6, 12))
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.6, 0.0, 0.8), 0.25))
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=vg.F64, seed=12, **kw, **lat)
    g.set_tuning("basis_wave", 1); g.set_tuning("basis_wave_grid", 9)
    assert g.step_kernel == "basis_wave"
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = g.n_sites
    g.upload(random_state(ob.HEISENBERG, n, 3))
    cpu = g.download(); col = g.colours()
    g.set_thermostat(1.5, (0, 0, 1.0), 0.7)
    th = H.thermostat(1.5, (0, 0, 1.0), 0.7)
    for _ in range(3):
        sweep = g.sweeps
        e, m = g.step(1)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, False, 12, sweep, col, g.n_colours, cpu)
        assert np.max(np.abs(g.download() - cpu)) < 1e-12
        assert abs(e[0] - H.total_energy(th, cpu)) < 1e-12 * n * 10
        assert np.max(np.abs(m[0] - cpu.sum(axis=0))) < 1e-12 * n
    g.close()


def test_basis_wave_is_opt_in(built):
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(64, 32, 32), seed=5)
    assert g.step_kernel == "heis_basis"          # a State that fits in L2: colour launches; the wave kernel is opt-in
    g.set_tuning("basis_wave", 1)
    assert g.step_kernel == "basis_wave"
    g.randomize(); g.set_thermostat(1.0, (0, 0, 1.0), 0.3)
    e, m = g.step(4)
    assert abs(g.total_energy() - e[-1]) < 1e-5 * abs(e[-1]) + 1e-5 * g.n_sites
    assert np.max(np.abs(g.magnetization() - m[-1])) < 1e-5 * g.n_sites
    g.close()
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(64, 32, 4), seed=5)     # too few planes for the wave order
    g.set_tuning("basis_wave", 1)
    assert g.step_kernel == "heis_basis"
    g.close()


# ---------------------------------------------------------------------------------------------------------------------
# K4f: the fcc step as two PAIR launches (heis_basis_pair_kernel in vegas_rs_b200/csrc/heis_basis.cuh), arrays S -> D
# ---------------------------------------------------------------------------------------------------------------------
# (size, precision, rows per CTA): one tile per plane (the recomputed row is the tile's own first row), ragged last tile,
# one-row tiles, tiles larger than the plane, config[4] rows
PAIR_CASES = [
    ((16, 6, 5), vg.F32, 0),
    ((16, 6, 5), vg.F32, 4),
    ((8, 9, 3), vg.F64, 2),
    ((4, 2, 2), vg.F32, 1),
    ((24, 7, 1), vg.F32, 3),
    ((12, 5, 4), vg.F64, 64),
    ((384, 40, 3), vg.F32, 0),
]


@pytest.mark.parametrize("size,precision,rows", PAIR_CASES)
@pytest.mark.parametrize("proposal", [vg.PROPOSE_RANDOM, vg.PROPOSE_FLIP], ids=["random", "flip"])
def test_basis_pair_identical_to_colour_passes(built, size, precision, rows, proposal):
    kw = dict(unitcell=vg.FCC, size=size, precision=precision, seed=35, anisotropy=((0.6, 0, 0.8), 0.15), proposal=proposal)
    ref = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    ref.set_tuning("basis_pair", 0)
    assert ref.step_kernel == "heis_basis"
    g = vg.GpuMetropolis(vg.HEISENBERG, **kw)
    g.set_tuning("basis_pair", 1)
    if rows:
        g.set_tuning("basis_pair_rows", rows); g.set_tuning("basis_pair_chunk", 1 + rows % 3)
    assert g.step_kernel == "basis_pair"
    n = g.n_sites
    s0 = random_state(ob.HEISENBERG, n, 9)
    tol = 1e-12 if precision == vg.F64 else 1e-6
    for h in (ref, g):
        h.upload(s0); h.set_thermostat(1.4, (0, 0, 1.0), 0.4)
    # odd and even numbers of steps (the two array sets swap after every step), recorded and not, measure-only entry points
    for k, observe in ((1, True), (2, False), (3, True), (1, False)):
        ra = ref.step(k, observe=observe); ga = g.step(k, observe=observe)
        assert np.array_equal(g.download(), ref.download())
        if observe:
            assert np.allclose(ga[0], ra[0], rtol=tol, atol=tol * n) and np.allclose(ga[1], ra[1], rtol=10 * tol, atol=10 * tol * n)
        assert abs(g.total_energy() - ref.total_energy()) <= tol * n * 10
        assert np.allclose(g.magnetization(), ref.magnetization(), rtol=10 * tol, atol=10 * tol * n)
    assert g.attempt_count() == ref.attempt_count()
    # a new State after an odd number of steps lands in the current set
    s1 = random_state(ob.HEISENBERG, n, 10)
    for h in (ref, g):
        h.upload(s1)
    ref.step(2, observe=False); g.step(2, observe=False)
    assert np.array_equal(g.download(), ref.download())
    ref.close(); g.close()


def test_basis_pair_replays_the_reference_rule(built):
    """fcc, fp64: every decision of the pair launches against the oracle replay (Hamiltonian::energy of src/energy.rs, accept
    rule of src/integrator.rs:82-88); the fused E and M equal total_energy / magnetization of the replayed state."""
    lat = dict(unitcell=vg.FCC, size=(8, 6, 4))
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.6, 0.0, 0.8), 0.25))
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=vg.F64, seed=12, **kw, **lat)
    g.set_tuning("basis_pair", 1); g.set_tuning("basis_pair_rows", 4)
    assert g.step_kernel == "basis_pair"
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = g.n_sites
    g.upload(random_state(ob.HEISENBERG, n, 3))
    cpu = g.download(); col = g.colours()
    g.set_thermostat(1.5, (0, 0, 1.0), 0.7)
    th = H.thermostat(1.5, (0, 0, 1.0), 0.7)
    for _ in range(3):
        sweep = g.sweeps
        e, m = g.step(1)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, False, 12, sweep, col, g.n_colours, cpu)
        assert np.max(np.abs(g.download() - cpu)) < 1e-12
        assert abs(e[0] - H.total_energy(th, cpu)) < 1e-12 * n * 10
        assert np.max(np.abs(m[0] - cpu.sum(axis=0))) < 1e-12 * n
    g.close()


def test_basis_pair_is_the_default_beyond_l2(built):
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(128, 128, 32), seed=5)     # 2 Mi sites x 12 B = 24 MiB < L2
    assert g.step_kernel == "heis_basis"
    g.close()
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(256, 128, 96), seed=5)     # 12.6 M sites: 151 MB
    assert g.step_kernel == "basis_pair"
    g.randomize(); g.set_thermostat(1.0, (0, 0, 1.0), 0.3)
    e, m = g.step(3)
    assert abs(g.total_energy() - e[-1]) < 1e-5 * abs(e[-1]) + 1e-5 * g.n_sites
    assert np.max(np.abs(g.magnetization() - m[-1])) < 1e-5 * g.n_sites
    g.set_tuning("basis_pair", 0)
    assert g.step_kernel == "heis_basis"
    g.close()


@pytest.mark.parametrize("nslab,precision", [(2, vg.F32), (3, vg.F64), (4, vg.F32)])
def test_basis_pair_on_slabs_bit_identical(built, nslab, precision):
    """z-slabs stepping with the pair launches (each slab keeps two array sets in its one allocation, all slabs swap in lock step,
    boundary planes go into the neighbours' halos of the set being written) reproduce the single handle's colour launches."""
    Lx, Ly, nz = 16, 10, 4
    Lz = nz * nslab
    kw = dict(unitcell=vg.FCC, seed=41, precision=precision, anisotropy=((0.6, 0, 0.8), 0.15))
    whole = vg.GpuMetropolis(vg.HEISENBERG, size=(Lx, Ly, Lz), **kw)
    whole.set_tuning("basis_pair", 0)
    s = random_state(ob.HEISENBERG, Lx * Ly * Lz * 4, 17)
    whole.upload(s)
    slabs = [vg.GpuMetropolis(vg.HEISENBERG, size=(Lx, Ly, nz), nz_global=Lz, z_offset=r * nz, **kw) for r in range(nslab)]
    plane = Lx * Ly * 4
    for r, sl in enumerate(slabs):
        sl.set_tuning("basis_pair", 1); sl.set_tuning("basis_pair_rows", 4)
        sl.upload(s[r * nz * plane:(r + 1) * nz * plane])
    for r, sl in enumerate(slabs):
        sl.slab_connect_local(slabs[(r - 1) % nslab], slabs[(r + 1) % nslab])
    assert all(sl.step_kernel == "basis_pair" for sl in slabs)
    whole.set_thermostat(1.3, (0, 0, 1.0), 0.3)
    for sl in slabs:
        sl.set_thermostat(1.3, (0, 0, 1.0), 0.3)
    tol = 1e-12 if precision == vg.F64 else 1e-6
    for k in (1, 2, 2):   # odd and even numbers of steps: the sets swap every step (5 in all: the upload below meets set 1)
        e, m = whole.step(k)
        for _ in range(k):
            for sl in slabs:
                sl.step_async(1, True)
            parts = [sl.read_observables(1) for sl in slabs]
        assert np.array_equal(np.concatenate([sl.download() for sl in slabs]), whole.download())
        e_sum = sum(p[0][0] for p in parts)
        assert abs(e_sum - e[-1]) <= tol * abs(e[-1]) * 10 + tol
        assert abs(sum(sl.total_energy() for sl in slabs) - whole.total_energy()) <= tol * abs(e[-1]) * 10 + tol
    # a new State after an odd number of steps: upload pushes the boundary planes into the neighbours' CURRENT set
    s2 = random_state(ob.HEISENBERG, Lx * Ly * Lz * 4, 18)
    whole.upload(s2)
    for r, sl in enumerate(slabs):
        sl.upload(s2[r * nz * plane:(r + 1) * nz * plane])
    whole.step(2, observe=False)
    for _ in range(2):
        for sl in slabs:
            sl.step_async(1, False)
    assert np.array_equal(np.concatenate([sl.download() for sl in slabs]), whole.download())
    for sl in slabs:
        sl.close()
    whole.close()
