"""north_star check 2: <E>, <|M|>, specific heat and susceptibility versus T of the GPU Metropolis agree with the CPU
Metropolis (the oracle's port of src/integrator.rs:66-138 driven by its port of Machine / CoolDown, src/machine.rs:91-125,
src/program.rs:182-214) within 3 sigma of their statistical errors, for Ising and Heisenberg (fp32 AND fp64), across Tc.

Both sides run the SAME CoolDown program through their Machine; the per-step (E, |M|) series come from the
ObservableSensor (src/instrument.rs:254-262) and are reduced exactly as StatSensor does (src/instrument.rs:98-131):
    Cv = Var(E) / (N T^2),   chi = Var(|M|) / (N T).
Errors: block averages for the means, block jackknife for the two variances.  The trajectories differ by construction
(random-site selection with rand_pcg-like streams on the CPU, colour-ordered Philox on the GPU): only distributions can agree."""
import multiprocessing as mp

import numpy as np
import pytest

import vegas_rs_b200 as vg
from oracle import binding as ob
from helpers import oracle_model

pytestmark = pytest.mark.gpu

L = 12
N = L ** 3
NBLOCKS = 20
# model -> CoolDown program (both sides of Tc: 4.51 for 3-D Ising, 1.44 for sc Heisenberg) and Hamiltonian terms
PROGRAMS = {
    "ising": dict(tmax=5.7, tmin=3.3, rate=0.4, relax=1500, steps=12000, T0=6.5, kw=dict(exchange=1.0, zeeman=True)),
    "heisenberg": dict(tmax=2.0, tmin=0.8, rate=0.2, relax=1500, steps=10000, T0=2.5,
                       kw=dict(exchange=1.0, zeeman=True, anisotropy=((0.0, 0.0, 1.0), 0.1))),
}


def _oracle_cooldown(name):
    """Worker process: the oracle Machine runs the program; returns per temperature point the (E, |M|) series."""
    p = PROGRAMS[name]
    model = ob.ISING if name == "ising" else ob.HEISENBERG
    H, _ = oracle_model(model, unitcell=ob.SC, size=(L, L, L), **p["kw"])
    rng = ob.OracleRng(20261017)
    s = H.rand_state(rng, N)
    m = ob.Machine(H, ob.PROPOSE_FLIP if name == "ising" else ob.PROPOSE_RANDOM, rng, s, n_sensors=2)
    m.relax(p["relax"], p["T0"])
    m.cooldown(p["tmax"], p["tmin"], p["rate"], p["relax"], p["steps"])
    e, mag = m.observables()
    rows = m.rows()
    out, k = [], p["relax"]
    for row in rows:
        k += p["relax"]
        out.append((row[0], e[k:k + p["steps"]].copy(), mag[k:k + p["steps"]].copy()))
        k += p["steps"]
    assert k == len(e)
    return out


def _gpu_cooldown(name, precision):
    from vegas_rs_b200.machine import Machine
    p = PROGRAMS[name]
    g = vg.GpuMetropolis(vg.ISING if name == "ising" else vg.HEISENBERG, unitcell=vg.SC, size=(L, L, L), seed=977,
                         precision=precision, **p["kw"])
    g.randomize()
    m = Machine(g)
    batches = []
    m.add_observable_sensor(lambda relax, stage, n, T, f, e, mag: batches.append((relax, stage, T, e, mag)))
    m.relax(p["relax"], p["T0"])
    m.cooldown(p["tmax"], p["tmin"], p["rate"], p["relax"], p["steps"])
    m.close(); g.close()
    out = {}
    for relax, stage, T, e, mag in batches:   # a stage may arrive in several batches (observable ring of 4096 rows)
        if relax:
            continue
        out.setdefault((stage, T), [[], []])
        out[(stage, T)][0].append(e); out[(stage, T)][1].append(mag)
    res = [(T, np.concatenate(v[0]), np.concatenate(v[1])) for (stage, T), v in sorted(out.items())]
    assert all(len(r[1]) == p["steps"] for r in res)
    return res


def _mean_err(x):
    b = x[: len(x) // NBLOCKS * NBLOCKS].reshape(NBLOCKS, -1).mean(axis=1)
    return x.mean(), b.std(ddof=1) / np.sqrt(NBLOCKS)


def _var_err(x):
    """variance and its block-jackknife error"""
    m = len(x) // NBLOCKS
    xb = x[: m * NBLOCKS].reshape(NBLOCKS, m)
    full = x.var()
    jk = np.array([np.delete(xb, i, axis=0).var() for i in range(NBLOCKS)])
    return full, np.sqrt((NBLOCKS - 1) / NBLOCKS * np.sum((jk - jk.mean()) ** 2))


@pytest.fixture(scope="module")
def oracle_series(built):
    """Both oracle programs run in worker processes while the GPU side is being measured."""
    pool = mp.get_context("fork").Pool(2)
    pending = {name: pool.apply_async(_oracle_cooldown, (name,)) for name in PROGRAMS}
    yield pending
    pool.terminate()


@pytest.mark.parametrize("name,precision", [("ising", vg.F32), ("heisenberg", vg.F32), ("heisenberg", vg.F64)],
                         ids=["ising", "heisenberg_f32", "heisenberg_f64"])
def test_observables_vs_temperature_match_the_cpu_metropolis(built, oracle_series, name, precision):
    gpu = _gpu_cooldown(name, precision)
    cpu = oracle_series[name].get(timeout=600)
    assert len(gpu) == len(cpu) >= 6
    tc = 4.51 if name == "ising" else 1.44
    assert min(r[0] for r in gpu) < tc < max(r[0] for r in gpu)
    worst = 0.0
    for (Tg, eg, mg), (Tc_, ec, mc) in zip(gpu, cpu):
        assert abs(Tg - Tc_) < 1e-12
        checks = {
            "<E>": (_mean_err(eg), _mean_err(ec), 1.0),
            "<|M|>": (_mean_err(mg), _mean_err(mc), 1.0),
            "Cv": (_var_err(eg), _var_err(ec), 1.0 / (N * Tg * Tg)),       # src/instrument.rs:118
            "chi": (_var_err(mg), _var_err(mc), 1.0 / (N * Tg)),           # src/instrument.rs:120
        }
        for what, ((a, da), (b, db), scale) in checks.items():
            sigma = np.hypot(da, db)
            z = abs(a - b) / sigma
            worst = max(worst, z)
            assert z < 3.0, f"{name} T={Tg:.2f} {what}: gpu {a * scale:.6g} +- {da * scale:.2g}, cpu {b * scale:.6g} +- {db * scale:.2g} ({z:.2f} sigma)"
    print(f"{name}: worst deviation {worst:.2f} sigma over {4 * len(gpu)} comparisons")
