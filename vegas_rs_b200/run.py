"""`python -m vegas_rs_b200.run input.toml [--seed S]` -- the reference's `vegas run` (src/main.rs:61-75,
src/input.rs:264-367) on the GPU sweep: same TOML schema, same StatSensor stdout lines, same parquet
schemas (src/output.rs:37-46, :126-136; SNAPPY; written to *.parquet.tmp and renamed when finished).

Front-end glue only: tomllib + pyarrow.  Stages, hooks and sensors run in the C++ host layer
(include/vegas_host.h); the sweep in CUDA (include/vegas_gpu.h).
"""
from __future__ import annotations

import argparse
import os
import sys
import tomllib

import numpy as np

from . import BCC, FCC, HEISENBERG, ISING, SC, E_REFERENCE_COMPOUND, GpuMetropolis
from .machine import Machine


class InputError(ValueError):
    pass


def parse_input(text: str) -> dict:
    """serde schema of src/input.rs:19-170 (all stage fields are required, App. A Q19)."""
    d = tomllib.loads(text)
    model = d.get("model")
    if model not in ("Ising", "Heisenberg"):
        raise InputError(f"unknown model {model!r}")
    algo = d.get("algorithm")
    if algo not in ("Metropolis", "Wolff"):
        raise InputError(f"unknown algorithm {algo!r}")
    sample = d.get("sample")
    if sample is None:
        raise InputError("missing field `sample`")
    uc = sample.get("unitcell", {})
    if "path" in uc:
        raise NotImplementedError("unitcell.path is todo!() in the reference (src/input.rs:303)")
    name = uc.get("name")
    if name not in ("sc", "bcc", "fcc"):
        raise InputError(f"unknown unit cell {name!r}")
    size, pbc = sample["size"], sample["pbc"]
    stages = []
    for st in d.get("stages", []):
        prog = st.get("program")
        req = {"Relax": ("steps", "temperature"),
               "CoolDown": ("max_temperature", "min_temperature", "cool_rate", "relax", "steps"),
               "Hysteresis": ("steps", "relax", "temperature", "max_field", "field_step")}.get(prog)
        if req is None:
            raise InputError(f"unknown program {prog!r}")
        for k in req:
            if k not in st:
                raise InputError(f"missing field `{k}` in {prog} stage")
        stages.append(st)
    return dict(model=model, algorithm=algo, exchange=d.get("exchange"), unitcell=name,
                size=(size["x"], size["y"], size["z"]), pbc=(pbc["x"], pbc["y"], pbc["z"]), stages=stages,
                output=d.get("output"))


class ParquetSink:
    """Arrow RecordBatch -> Parquet (SNAPPY) into path.with_extension("parquet.tmp"), renamed on close
    (src/output.rs:35, :101-113)."""

    def __init__(self, path: str, schema):
        import pyarrow.parquet as pq
        root, _ = os.path.splitext(path)
        self.path, self.tmp = path, root + ".parquet.tmp"
        self.writer = pq.ParquetWriter(self.tmp, schema, compression="snappy")
        self.schema = schema

    def write(self, columns):
        import pyarrow as pa
        self.writer.write_batch(pa.record_batch(columns, schema=self.schema))

    def close(self):
        if self.writer is not None:
            self.writer.close()
            self.writer = None
            os.replace(self.tmp, self.path)


def observable_schema():
    import pyarrow as pa
    return pa.schema([pa.field("relax", pa.bool_(), False), pa.field("stage", pa.uint64(), False),
                      pa.field("step", pa.uint64(), False), pa.field("n", pa.uint64(), False),
                      pa.field("temperature", pa.float64(), False), pa.field("field", pa.float64(), False),
                      pa.field("energy", pa.float64(), False), pa.field("magnetization", pa.float64(), False)])


def state_schema():
    import pyarrow as pa
    return pa.schema([pa.field("relax", pa.bool_(), False), pa.field("stage", pa.uint64(), False),
                      pa.field("step", pa.uint64(), False), pa.field("temperature", pa.float64(), False),
                      pa.field("field", pa.float64(), False), pa.field("id", pa.uint64(), False),
                      pa.field("sx", pa.float64(), False), pa.field("sy", pa.float64(), False),
                      pa.field("sz", pa.float64(), False)])


def run_stages(cfg: dict, m, out=sys.stdout):
    """run_with_spin's instrument wiring and stage loop (src/input.rs:273-292, :324-345) on a Machine: StatSensor ->
    `out`, ObservableSensor -> observables parquet, StateSensor -> state parquet; then every stage in order.  Closes the
    parquet sinks (tmp file renamed) whether or not a stage fails."""
    sinks = []
    # Input::instruments, src/input.rs:324-345: StatSensor(stdout) [+ ObservableSensor] [+ StateSensor]
    m.add_stat_sensor(lambda line, row: print(line, file=out, flush=True))
    output = cfg.get("output") or {}
    if output.get("observables"):
        import pyarrow as pa
        sink = ParquetSink(output["observables"], observable_schema()); sinks.append(sink)

        def on_batch(relax, stage, n, T, field, e, mag):
            k = len(e)
            sink.write([pa.array(np.full(k, relax)), pa.array(np.full(k, stage, np.uint64)), pa.array(np.arange(k, dtype=np.uint64)),
                        pa.array(np.full(k, n, np.uint64)), pa.array(np.full(k, T)), pa.array(np.full(k, field)), pa.array(e), pa.array(mag)])
        m.add_observable_sensor(on_batch)
    if output.get("state"):
        import pyarrow as pa
        gather = cfg.get("_gather")
        ssink = None
        if output["state"]["path"] is not None:
            ssink = ParquetSink(output["state"]["path"], state_schema()); sinks.append(ssink)

        def on_state(relax, stage, step, T, field, s):
            if gather is not None:   # slab group: the State is the slabs in rank (= z) order
                dist, group, rank, world = gather
                parts = [None] * world if rank == 0 else None
                dist.gather_object(s, parts, dst=0, group=group)
                if rank != 0:
                    return
                s = np.concatenate(parts)
            k = len(s)
            if s.ndim == 1:   # IsingSpin::{sx,sy,sz}, src/state.rs:103-121
                sx = sy = np.zeros(k); sz = s.astype(np.float64)
            else:
                sx, sy, sz = s[:, 0].copy(), s[:, 1].copy(), s[:, 2].copy()
            ssink.write([pa.array(np.full(k, relax)), pa.array(np.full(k, stage, np.uint64)), pa.array(np.full(k, step, np.uint64)),
                         pa.array(np.full(k, T)), pa.array(np.full(k, field)), pa.array(np.arange(k, dtype=np.uint64)),
                         pa.array(sx), pa.array(sy), pa.array(sz)])
        m.add_state_sensor(int(output["state"]["frequency"]), on_state)
    try:
        for st in cfg["stages"]:
            if st["program"] == "Relax":
                m.relax(st["steps"], st["temperature"])
            elif st["program"] == "CoolDown":
                m.cooldown(st["max_temperature"], st["min_temperature"], st["cool_rate"], st["relax"], st["steps"])
            else:
                m.hysteresis(st["steps"], st["relax"], st["temperature"], st["max_field"], st["field_step"])
    finally:
        for s in sinks:
            s.close()


def group_reduce(dist, device: int, group=None):
    """In-place sum over the ranks of a float64 numpy array (the Machine's slab-group reduction): 4 doubles per step."""
    import torch
    on_gpu = dist.get_backend(group) == "nccl"

    def reduce_sum(values):
        t = torch.from_numpy(values)
        if on_gpu:
            d = t.to(f"cuda:{device}")
            dist.all_reduce(d, group=group)
            t.copy_(d)
        else:
            dist.all_reduce(t, group=group)
    return reduce_sum


def run_input(cfg: dict, seed: int | None = None, out=sys.stdout, device: int = 0, literal: bool = False, dist=None, group=None):
    """Input::run (src/input.rs:347-367) -> run_with_spin (:264-294).

    With an initialised `dist` (torch.distributed, one process per GPU) the lattice is cut into z-slabs, one per rank;
    every rank runs the same stages on its slab through its own Machine, the per-step (E, M) partial sums are all-reduced,
    and rank 0 owns stdout and the parquet files (a StateSensor dump gathers the slabs on rank 0)."""
    if cfg["algorithm"] == "Wolff":
        raise NotImplementedError("the Wolff cluster integrator is outside the GPU sweep's scope (SURVEY section 2)")
    model = ISING if cfg["model"] == "Ising" else HEISENBERG
    world = dist.get_world_size(group) if dist is not None else 1
    rank = dist.get_rank(group) if dist is not None else 0
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")  # Pcg64::from_rng(&mut rand::rng()), src/main.rs:70-73
        if world > 1:  # one Philox key for the whole lattice
            box = [seed]
            dist.broadcast_object_list(box, src=0, group=group)
            seed = box[0]
    uc = {"sc": SC, "bcc": BCC, "fcc": FCC}[cfg["unitcell"]]
    exchange = cfg["exchange"] if cfg["exchange"] is not None else 1.0  # src/input.rs:352
    size, kw = tuple(cfg["size"]), {}
    if world > 1:
        from . import distributed as vd
        nz, zoff = vd.slab_extent(size[2], rank, world)
        kw = dict(nz_global=size[2], z_offset=zoff)
        size = (size[0], size[1], nz)
    # hamiltonian!(Exchange::from_lattice(exchange, &lattice), Zeeman::new()), src/input.rs:271
    g = GpuMetropolis(model, unitcell=uc, size=size, pbc=cfg["pbc"], exchange=exchange, zeeman=True, seed=seed,
                      device=device, literal=literal, **kw)
    try:
        g.set_energy_convention(E_REFERENCE_COMPOUND)
        if world > 1:
            vd.connect_slabs(g, dist, group)
        g.randomize()  # State::rand_with_size, src/input.rs:278 (keyed by the GLOBAL site index: slab independent)
        if world > 1:
            dist.barrier(group)   # every slab has pushed its boundary planes into the neighbours' halos
        m = Machine(g)
        try:
            if world > 1:
                m.set_group(group_reduce(dist, device, group), g.n_sites * world)
                if rank != 0:   # same stages, same hooks, no output: rank 0 owns stdout and the files
                    cfg = dict(cfg, output=_gather_only(cfg.get("output")))
                    out = open(os.devnull, "w")
                cfg = dict(cfg, _gather=(dist, group, rank, world))
            run_stages(cfg, m, out)
        finally:
            m.close()
    finally:
        g.close()


def _gather_only(output):
    """Ranks > 0 of a slab group write nothing, but a StateSensor must still fire on them (its dump is gathered on rank 0)."""
    if not output or not output.get("state"):
        return None
    return {"state": dict(output["state"], path=None)}


def main(argv=None):
    ap = argparse.ArgumentParser(prog="vegas_rs_b200.run", description="vegas run on the B200 sweep (one GPU, or one z-slab per "
                                 "GPU under `python -m torch.distributed.run --nproc-per-node N -m vegas_rs_b200.run input.toml`)")
    ap.add_argument("input", help="input TOML file, or - for stdin")
    ap.add_argument("-s", "--seed", type=int, default=None)
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--literal-lattice", action="store_true", help="apply the source<=target filter of Exchange::from_lattice")
    a = ap.parse_args(argv)
    text = sys.stdin.read() if a.input == "-" else open(a.input).read()
    dist, device = None, a.device or 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:   # launched by torchrun: one process per GPU, z-slabs
        import torch
        import torch.distributed as dist
        device = int(os.environ.get("LOCAL_RANK", "0")) if a.device is None else a.device
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{device}"))
    try:
        run_input(parse_input(text), a.seed, device=device, literal=a.literal_lattice, dist=dist)
    except Exception as e:  # check_error, src/main.rs:84-89
        print(f"Error: {e}", file=sys.stderr)
        sys.exit(1)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
