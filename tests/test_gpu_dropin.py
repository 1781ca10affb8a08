"""Round-2 parity additions, all through the C ABI on the GPU:
  * the literal Integrator::step drop-in vegas_gpu_step_host_{ising,heisenberg} (src/integrator.rs:40-49: State in,
    State out) against the oracle replay of the reference rule;
  * BASELINE config[1] at its full 8192^2 size against Onsager's exact u(T) and |m|(T) on both sides of Tc."""
import numpy as np
import pytest

import vegas_rs_b200 as vg
from oracle import binding as ob
from helpers import oracle_model, random_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,family", [((64, 6, 4), "ising_msc"), ((10, 10, 10), "ising_general")])
def test_step_host_ising_is_integrator_step(built, size, family):
    """host State in -> ONE step -> host State out, three times in a row, each equal to the oracle's replay of
    MetropolisFlipIntegrator::step (src/integrator.rs:109-138) on the same host State; E and M returned with it equal
    Hamiltonian::total_energy / State::magnetization of the returned State (bit-exact)."""
    seed = 4242
    lat = dict(unitcell=vg.SC, size=size)
    g = vg.GpuMetropolis(vg.ISING, seed=seed, **lat)
    assert g.kernel_family == family
    H, _ = oracle_model(ob.ISING, **lat)
    n = size[0] * size[1] * size[2]
    g.set_thermostat(4.2, (0, 0, 1.0), 0.5)
    th = H.thermostat(4.2, (0, 0, 1.0), 0.5)
    host = random_state(ob.ISING, n, 9)          # the caller's State (reference layout: +1 / -1 per site)
    cpu = host.copy()
    col = g.colours()
    for k in range(3):
        sweep = g.sweeps
        out, e, m = g.step_host(host)
        assert out is host
        if family == "ising_msc":
            H.replay_ising_msc(th, ob.PROPOSE_FLIP, seed, sweep, size, cpu)
        else:
            H.replay_ising_sites(th, ob.PROPOSE_FLIP, seed, sweep, col, g.n_colours, cpu)
        assert np.array_equal(host, cpu), k
        assert e == H.total_energy(th, cpu)
        assert m[2] == cpu.astype(np.int64).sum() and m[0] == 0 and m[1] == 0
    # the State really is the caller's: a State edited on the host between two calls is the one that gets stepped
    host[:] = 1
    cpu[:] = 1
    sweep = g.sweeps
    g.step_host(host)
    if family == "ising_msc":
        H.replay_ising_msc(th, ob.PROPOSE_FLIP, seed, sweep, size, cpu)
    else:
        H.replay_ising_sites(th, ob.PROPOSE_FLIP, seed, sweep, col, g.n_colours, cpu)
    assert np.array_equal(host, cpu)
    g.close()


@pytest.mark.parametrize("precision", [vg.F64, vg.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("unitcell,size", [(vg.SC, (16, 4, 6)), (vg.FCC, (4, 4, 4))], ids=["sc", "fcc"])
def test_step_host_heisenberg_is_integrator_step(built, precision, unitcell, size):
    """The same for MetropolisIntegrator::step (src/integrator.rs:66-92) on [f64; 3] host spins: 1e-12 (fp64 device
    storage) / 1e-5 with a few knife-edge sites (fp32)."""
    seed = 99
    lat = dict(unitcell=unitcell, size=size)
    kw = dict(exchange=1.0, zeeman=True, anisotropy=((0.0, 0.6, 0.8), 0.2))
    g = vg.GpuMetropolis(vg.HEISENBERG, precision=precision, seed=seed, **kw, **lat)
    H, _ = oracle_model(ob.HEISENBERG, **kw, **lat)
    n = g.n_sites
    g.set_thermostat(0.9, (0, 0, 1.0), 0.3)
    th = H.thermostat(0.9, (0, 0, 1.0), 0.3)
    host = random_state(ob.HEISENBERG, n, 5)
    g.upload(host); host = g.download()          # fp32: start from device-representable spins
    cpu = host.copy()
    col = g.colours()
    for k in range(3):
        sweep = g.sweeps
        out, e, m = g.step_host(host)
        H.replay_heisenberg(th, ob.PROPOSE_RANDOM, precision == vg.F32, seed, sweep, col, g.n_colours, cpu)
        diff = np.max(np.abs(host - cpu), axis=1)
        if precision == vg.F64:
            assert np.max(diff) < 1e-12, k
            assert abs(e - H.total_energy(th, cpu)) < 1e-12 * n * 10
            assert np.max(np.abs(m - cpu.sum(axis=0))) < 1e-12 * n
        else:
            assert np.sum(diff > 1e-5) <= max(2, n // 200), k
            assert abs(e - H.total_energy(th, host)) < 1e-5 * n * 6
            cpu = host.copy()
    g.close()


def _onsager(T):
    from scipy.special import ellipk
    K = 1.0 / T
    k = 2 * np.sinh(2 * K) / np.cosh(2 * K) ** 2
    u = -(1 / np.tanh(2 * K)) * (1 + (2 / np.pi) * (2 * np.tanh(2 * K) ** 2 - 1) * ellipk(k * k))
    m = (1 - np.sinh(2 * K) ** -4) ** 0.125 if T < 2.269185314213022 else 0.0
    return u, m


@pytest.mark.parametrize("size,chunk", [((128, 16, 16), 4096), ((64, 6, 4), 1 << 26), ((512, 32, 16), 16384)])
def test_host_packed_state_transfer(built, size, chunk):
    """Big ising_msc lattices move their State over PCIe as a sign bitmap packed by host threads (host_pack.cpp) and split into
    the colour arrays by a kernel; forced here on small lattices (host_pack_min = 0), with many pipelined chunks.  Upload,
    download and the literal Integrator::step entry point must be indistinguishable from the byte-per-spin path."""
    n = size[0] * size[1] * size[2]
    s = random_state(ob.ISING, n, 5)
    ref = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=size, seed=77)
    ref.set_tuning("host_pack_min", -1)
    assert ref.state_transfer_bytes == n
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=size, seed=77)
    g.set_tuning("host_pack_min", 0); g.set_tuning("host_pack_chunk", chunk)
    assert g.kernel_family == "ising_msc" and g.state_transfer_bytes == n // 8
    ref.upload(s); g.upload(s)
    assert np.array_equal(g.download(), s) and np.array_equal(ref.download(), s)
    for h in (ref, g):
        h.set_thermostat(4.0, (0, 0, 1.0), 0.25)
    a, b = s.copy(), s.copy()
    for _ in range(3):
        _, ea, ma = ref.step_host(a)
        _, eb, mb = g.step_host(b)
        assert ea == eb and np.array_equal(ma, mb) and np.array_equal(a, b)
    assert not np.array_equal(a, s)
    # the bitmap path accepts any byte > 0 as Up and anything else as Down, like the byte path's pack kernel
    ref.close(); g.close()


def test_full_size_ising_2d_8192_matches_onsager(built):
    """BASELINE config[1] at its full size (67 M spins, multi-spin-coded kernel): the energy per site and |m| per site
    of the equilibrated lattice equal Onsager's exact infinite-lattice values within 1e-3 at T = 2.0 (ordered) and T = 2.6
    (disordered: |m| -> 0 as N^-1/2).  Finite-size corrections at L = 8192 are far below the bar away from Tc."""
    L = 8192
    N = L * L
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(L, L, 1), pbc=(True, True, False), seed=8192)
    assert g.kernel_family == "ising_msc"
    g.set_energy_convention(vg.E_PHYSICAL)
    for T, start_up in ((2.0, True), (2.6, False)):
        if start_up:
            g.fill(True)
        else:
            g.randomize()
        g.set_thermostat(T)
        g.step(1500, observe=False)              # correlation length 2.3 (T = 2.0) / 5.5 (T = 2.6) lattice units
        e, m = g.step(500)
        u_exact, m_exact = _onsager(T)
        assert abs(e.mean() / N - u_exact) < 1e-3, (T, e.mean() / N, u_exact)
        assert abs(np.abs(m[:, 2]).mean() / N - m_exact) < 1e-3, (T, np.abs(m[:, 2]).mean() / N, m_exact)
    g.close()
