"""Host layer (include/vegas_host.h): Machine / instruments / programs over the GPU sweep, and the TOML +
parquet front end, against the oracle's restatement of the same reference code."""
import io
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_symbols_exported(built):
    from vegas_rs_b200 import machine
    lib = machine._load()
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vegas_host.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(vegas_(?:machine|program)_[a-z_]+)\s*\(", text)))
    assert len(names) == 15 and "vegas_machine_set_group" in names
    bound = {s[0] for s in machine.HOST_SYMBOLS}
    for n in names:
        assert hasattr(lib, n) and n in bound, n


def test_toml_schema_parsing():
    from vegas_rs_b200 import run
    cfg = run.parse_input(open(os.path.join(ROOT, "tests", "golden", "cfg0_ising_sc10.toml")).read())
    assert cfg["model"] == "Ising" and cfg["algorithm"] == "Metropolis" and cfg["unitcell"] == "sc"
    assert cfg["size"] == (10, 10, 10) and cfg["pbc"] == (True, True, True) and cfg["exchange"] is None
    assert [s["program"] for s in cfg["stages"]] == ["Relax", "CoolDown"]
    assert cfg["output"]["state"]["frequency"] == 1000
    with pytest.raises(run.InputError):
        run.parse_input('model="Ising"\nalgorithm="Metropolis"\n[sample.unitcell]\nname="sc"\n[sample.size]\nx=1\ny=1\nz=1\n'
                        '[sample.pbc]\nx=true\ny=true\nz=true\n[[stages]]\nprogram="Relax"\nsteps=10\n')  # temperature missing
    with pytest.raises(run.InputError):
        run.parse_input('model="Potts"\nalgorithm="Metropolis"\n')


@pytest.mark.gpu
def test_machine_hooks_and_counters(built):
    import vegas_rs_b200 as vg
    from vegas_rs_b200.machine import Machine, ProgramError
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(10, 10, 10), seed=4)
    g.randomize()
    m = Machine(g)
    assert m.thermostat() == (2.8, 0.0)                       # src/input.rs:274
    lines, batches, dumps = [], [], []
    m.add_stat_sensor(lambda line, row: lines.append((line, row)))
    m.add_observable_sensor(lambda *a: batches.append(a))
    m.add_state_sensor(7, lambda *a: dumps.append(a))
    m.relax(20, 6.0)
    m.cooldown(3.0, 2.0, 0.5, 10, 25)
    assert m.steps_done == 20 + 3 * 35
    assert len(lines) == 3 and [r[1][0] for r in lines] == [3.0, 2.5, 2.0]
    assert all(len(l[0].split(" ")) == 7 for l in lines)
    # ObservableSensor: one batch per relax and per measure stage, stage counter increments on both
    assert [(b[0], b[1], len(b[5])) for b in batches] == [(True, 0, 20), (True, 1, 10), (False, 2, 25), (True, 3, 10),
                                                         (False, 4, 25), (True, 5, 10), (False, 6, 25)]
    # StatSensor statistics are those of the measure batch (population variance, totals)
    e, mag = batches[2][5], batches[2][6]
    row = lines[0][1]
    assert abs(row[2] - e.mean()) < 1e-9 and abs(row[3] - e.var() / (1000 * 9.0)) < 1e-9
    assert abs(row[4] - mag.mean()) < 1e-9 and abs(row[5] - mag.var() / (1000 * 3.0)) < 1e-9
    # StateSensor: step.is_multiple_of(7), counter restarts per stage
    assert [(d[0], d[1], d[2]) for d in dumps[:5]] == [(True, 0, 0), (True, 0, 7), (True, 0, 14), (True, 1, 0), (True, 1, 7)]
    assert dumps[0][5].shape == (1000,) and set(np.unique(dumps[0][5])) <= {-1, 1}
    # dumped state is the state after that step: its energy equals the recorded per-step energy
    from oracle import binding as ob
    from helpers import oracle_model
    H, _ = oracle_model(ob.ISING, unitcell=ob.SC, size=(10, 10, 10))
    assert H.total_energy(H.thermostat(6.0), dumps[1][5]) == batches[0][5][7]
    with pytest.raises(ProgramError):
        m.cooldown(1.0, 2.0, 0.1, 1, 1)
    with pytest.raises(ProgramError):
        m.relax(0, 1.0)
    with pytest.raises(ProgramError):
        m.hysteresis(1, 1, 1.0, 0.0, 0.1)
    m.close(); g.close()


@pytest.mark.gpu
def test_cooldown_statistics_match_oracle_machine(built):
    """config[0] shape (docs/metropolis.toml, shortened): GPU checkerboard Machine vs the oracle's
    random-site Machine, same program; <E>, <|M|> agree within 4 sigma (blocking errors) at every T."""
    import vegas_rs_b200 as vg
    from vegas_rs_b200.machine import Machine
    from oracle import binding as ob
    from helpers import oracle_model, blocking_error
    prog = dict(tmax=5.5, tmin=3.5, rate=1.0, relax=400, steps=3000)
    g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(10, 10, 10), seed=21)
    g.randomize()
    m = Machine(g)
    gb = []
    m.add_observable_sensor(lambda relax, stage, n, T, f, e, mag: gb.append((relax, T, e, mag)))
    m.relax(400, 6.0)
    m.cooldown(prog["tmax"], prog["tmin"], prog["rate"], prog["relax"], prog["steps"])
    m.close(); g.close()
    H, _ = oracle_model(ob.ISING, unitcell=ob.SC, size=(10, 10, 10))
    rng = ob.OracleRng(5)
    s = H.rand_state(rng, 1000)
    om = ob.Machine(H, ob.PROPOSE_FLIP, rng, s, n_sensors=2)
    om.relax(400, 6.0)
    om.cooldown(prog["tmax"], prog["tmin"], prog["rate"], prog["relax"], prog["steps"])
    oe, omag = om.observables()
    k = 400
    for relax, T, e, mag in [b for b in gb if not b[0]]:
        k += prog["relax"]
        ce, cm = oe[k:k + prog["steps"]], omag[k:k + prog["steps"]]
        k += prog["steps"]
        for a, b in ((e, ce), (mag, cm)):
            err = np.hypot(blocking_error(a), blocking_error(b))
            assert abs(a.mean() - b.mean()) < 4 * err + 1e-9, (T, a.mean(), b.mean(), err)


@pytest.mark.gpu
def test_run_toml_end_to_end(built, tmp_path):
    """`vegas run docs/metropolis.toml` shape: stdout lines + parquet files with the reference schemas."""
    import pyarrow.parquet as pq
    from vegas_rs_b200 import run
    text = open(os.path.join(ROOT, "tests", "golden", "cfg0_ising_sc10.toml")).read()
    text = text.replace("steps = 20000", "steps = 60").replace("relax = 1000", "relax = 20").replace("steps = 1000", "steps = 30")
    text = text.replace("cool_rate = 0.05", "cool_rate = 1.0").replace("frequency = 1000", "frequency = 25")
    text = text.replace("./output.parquet", str(tmp_path / "output.parquet")).replace("./state.parquet", str(tmp_path / "state.parquet"))
    cfg = run.parse_input(text)
    out = io.StringIO()
    run.run_input(cfg, seed=7, out=out)
    lines = out.getvalue().strip().split("\n")
    assert len(lines) == 6 and lines[0].startswith("6.0000000000000000 0.0000000000000000 ")
    assert not os.path.exists(tmp_path / "output.parquet.tmp")
    obs = pq.read_table(tmp_path / "output.parquet")
    assert obs.schema.names == ["relax", "stage", "step", "n", "temperature", "field", "energy", "magnetization"]
    assert [str(t) for t in obs.schema.types] == ["bool", "uint64", "uint64", "uint64", "double", "double", "double", "double"]
    assert obs.num_rows == 30 + 6 * 80 and set(obs.column("n").to_pylist()) == {1000}
    assert pq.ParquetFile(tmp_path / "output.parquet").metadata.row_group(0).column(0).compression == "SNAPPY"
    st = pq.read_table(tmp_path / "state.parquet")
    assert st.schema.names == ["relax", "stage", "step", "temperature", "field", "id", "sx", "sy", "sz"]
    assert st.num_rows % 1000 == 0 and set(st.column("sz").to_pylist()) <= {1.0, -1.0}


@pytest.mark.gpu
def test_sharded_cooldown_points_through_machine(built):
    """distributed.sharded_cooldown: a rank's temperature points as one-point CoolDowns on the real Machine."""
    import vegas_rs_b200 as vg
    from vegas_rs_b200 import distributed as vd
    from vegas_rs_b200.machine import Machine
    pts = vd.cooldown_temperatures(3.0, 2.0, 0.25)
    lines_by_rank = []
    for rank in range(2):  # what two ranks would do, one after the other on this GPU
        g = vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(64, 8, 8), seed=40 + rank)
        g.randomize()
        m = Machine(g)
        lines = []
        m.add_stat_sensor(lambda line, row: lines.append(line))
        mine = vd.shard_points(pts, rank, 2)
        vd.sharded_cooldown(m, mine, 0.25, 30, 40)
        assert [float(l.split()[0]) for l in lines] == mine and m.steps_done == len(mine) * 70
        lines_by_rank.append(lines)
        m.close(); g.close()
    assert len(lines_by_rank[0]) == 3 and len(lines_by_rank[1]) == 2


@pytest.mark.gpu
def test_machine_on_fcc_heisenberg_reports_vector_magnetisation(built):
    """config[4] shape through the Machine: periodic fcc takes the heis_basis family, whose |M| is the norm of three
    projections (HeisenbergSpin::from_projections, src/state.rs:150-160) and whose StateSensor dump is [f64;3] per site."""
    import vegas_rs_b200 as vg
    from vegas_rs_b200.machine import Machine
    g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=vg.FCC, size=(4, 4, 4), precision=vg.F64, seed=8)
    assert g.kernel_family == "heis_basis"
    v = np.tile(np.array([[0.6, 0.8, 0.0]]), (g.n_sites, 1))      # magnetised in the xy plane: Mz = 0, |M| = N
    g.upload(v)
    m = Machine(g)
    batches, dumps = [], []
    m.add_observable_sensor(lambda *a: batches.append(a))
    m.add_state_sensor(5, lambda *a: dumps.append(a))
    m.relax(6, 0.05)                                              # cold: the state stays close to the initial one
    mag = batches[0][6]
    assert mag.shape == (6,) and np.all(mag > 0.9 * g.n_sites)    # |M| from all three projections, not |Mz|
    assert dumps[0][5].shape == (g.n_sites, 3)
    assert np.max(np.abs(np.linalg.norm(dumps[1][5], axis=1) - 1.0)) < 1e-12
    # the dumped state is the state after that step: its energy is the recorded one
    from oracle import binding as ob
    from helpers import oracle_model
    H, _ = oracle_model(ob.HEISENBERG, unitcell=ob.FCC, size=(4, 4, 4))
    assert abs(H.total_energy(H.thermostat(0.05), dumps[1][5]) - batches[0][5][5]) < 1e-9
    m.close(); g.close()


@pytest.mark.gpu
def test_run_toml_hysteresis_heisenberg(built, tmp_path):
    """config[3] shape in miniature: Heisenberg sc HysteresisLoop through the TOML front end; the field column is
    |magnitude| (Field::magnitude, src/state.rs:219-221) and the point list overshoots max_field (App. A Q11)."""
    import pyarrow.parquet as pq
    from vegas_rs_b200 import run
    text = f"""
model = "Heisenberg"
algorithm = "Metropolis"
exchange = 1.0

[sample]
unitcell = {{ name = "sc" }}
size = {{ x = 16, y = 8, z = 8 }}
pbc = {{ x = true, y = true, z = true }}

[[stages]]
program = "Hysteresis"
steps = 20
relax = 10
temperature = 1.0
max_field = 1.0
field_step = 0.5

[output]
observables = "{tmp_path / 'hyst.parquet'}"
"""
    cfg = run.parse_input(text)
    out = io.StringIO()
    run.run_input(cfg, seed=3, out=out)
    lines = out.getvalue().strip().split("\n")
    fields = [float(l.split()[1]) for l in lines]
    # 0, .5, 1 | 1.5, 1, .5, 0, -.5, -1 | -1.5, -1, ..., 1  with |.| applied (program.rs:303-332)
    assert fields == [0.0, 0.5, 1.0, 1.5, 1.0, 0.5, 0.0, 0.5, 1.0, 1.5, 1.0, 0.5, 0.0, 0.5, 1.0]
    obs = pq.read_table(tmp_path / "hyst.parquet")
    assert obs.num_rows == len(fields) * 30 and set(obs.column("n").to_pylist()) == {1024}
    # reference sign (+|H| s.o, src/energy.rs:147-151): the spins turn AGAINST the field orientation
    mags = np.array(obs.column("magnetization").to_pylist())
    assert np.all(mags >= 0)


def test_parquet_sinks_match_the_reference_schemas(tmp_path):
    """ObservableParquetOutput / StateParquetOutput (src/output.rs:33-51, :122-141, :101-113): column names, types,
    non-nullable fields, SNAPPY, written to `<stem>.parquet.tmp` and renamed on close.  No GPU involved."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from vegas_rs_b200 import run
    for schema, names, path in ((run.observable_schema(), ["relax", "stage", "step", "n", "temperature", "field", "energy", "magnetization"],
                                 tmp_path / "obs.parquet"),
                                (run.state_schema(), ["relax", "stage", "step", "temperature", "field", "id", "sx", "sy", "sz"],
                                 tmp_path / "deep" / ".." / "state.out")):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        assert schema.names == names and not any(f.nullable for f in schema)
        assert [str(f.type) for f in schema][:3] == ["bool", "uint64", "uint64"] and all(str(f.type) == "double" or f.name in ("relax", "stage", "step", "n", "id") for f in schema)
        sink = run.ParquetSink(str(path), schema)
        stem = os.path.splitext(str(path))[0]
        assert os.path.exists(stem + ".parquet.tmp") and not os.path.exists(str(path))   # Path::with_extension("parquet.tmp")
        k = 5
        cols = [pa.array(np.arange(k) % 2 == 0)] + [pa.array(np.arange(k, dtype=np.uint64) + i) if str(f.type) == "uint64"
                                                    else pa.array(np.linspace(0, 1, k) + i) for i, f in enumerate(schema) if f.name != "relax"]
        sink.write(cols); sink.write(cols)
        sink.close(); sink.close()                                # idempotent, as Drop + explicit close in the reference
        assert os.path.exists(str(path)) and not os.path.exists(stem + ".parquet.tmp")
        t = pq.read_table(str(path))
        assert t.num_rows == 2 * k and t.schema.names == names and not any(f.nullable for f in t.schema)
        md = pq.ParquetFile(str(path)).metadata
        assert all(md.row_group(0).column(c).compression == "SNAPPY" for c in range(md.num_columns))
