"""One process per GPU: z-slab decomposition helpers on top of torch.distributed (plumbing only).

The data path never goes through torch or NCCL: after `connect_slabs` each rank's sweep kernels store
their boundary planes straight into the z-neighbours' halo buffers (CUDA IPC mapped peer memory over
NVLink) and signal with release/acquire flags.  torch.distributed only carries the 256-byte IPC blobs
at start-up and, optionally, the all-reduce of the per-step (E, M) partial sums.
"""
from __future__ import annotations

import numpy as np


def slab_extent(nz_global: int, rank: int, world: int):
    """Contiguous z-planes [z_offset, z_offset+nz) of rank `rank`; every slab must be even and equal."""
    if nz_global % world:
        raise ValueError(f"nz_global={nz_global} is not divisible by world={world}")
    nz = nz_global // world
    if nz % 2:
        raise ValueError("slabs need an even number of planes (checkerboard parity)")
    return nz, rank * nz


def neighbours(rank: int, world: int):
    """(lower, upper) z-neighbour ranks of a periodic slab ring."""
    return (rank - 1) % world, (rank + 1) % world


def exchange_blobs(blob: bytes, dist, group=None):
    """all_gather of the per-rank IPC blobs (any backend: the payload is 256 opaque bytes)."""
    world = dist.get_world_size(group)
    blobs = [None] * world
    dist.all_gather_object(blobs, blob, group=group)
    return blobs


def connect_slabs(g, dist, group=None):
    """Exchange IPC handles, map the neighbours' halos, and synchronise before the first sweep."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    blobs = exchange_blobs(g.slab_export(), dist, group)
    lo, hi = neighbours(rank, world)
    dist.barrier(group)
    g.slab_connect(blobs[lo], blobs[hi])
    dist.barrier(group)


def reduce_observables(energy: np.ndarray, mag: np.ndarray, dist, device, group=None):
    """Sum of the slab partials of the per-step scalars (4 doubles per step)."""
    import torch
    t = torch.tensor(np.concatenate([energy, mag.ravel()]), device=device)
    dist.all_reduce(t, group=group)
    out = t.cpu().numpy()
    n = len(energy)
    return out[:n], out[n:].reshape(n, 3)


# ---------------------------------------------------------------------------------------------------
# Point sharding: independent temperature points of a CoolDown, one subset per GPU.
# The reference anneals (the state is carried from point to point, src/program.rs:203-211); sharded points
# are equilibrium estimates that start from the rank's own state and need their own `relax`.
# HysteresisLoop points are history dependent and are NOT sharded (shard whole loops / seeds instead).
# ---------------------------------------------------------------------------------------------------
def cooldown_temperatures(max_temperature: float, min_temperature: float, cool_rate: float):
    """The point list CoolDown visits: `T -= cool_rate` in f64, stop once T < min (src/program.rs:202-211).
    6.0 -> 1.0 @ 0.05 gives 101 points (last 1.0000000000000133); 4.0 -> 0.1 @ 0.1 gives 39 (the 0.1 point is lost)."""
    if cool_rate <= 0.0:
        raise ValueError("cool_rate must be positive")  # ProgramError::ZeroCoolRate
    if max_temperature < min_temperature:
        raise ValueError("max_temperature < min_temperature")  # ProgramError::MaxTemperatureLessThanMin
    out, t = [], float(max_temperature)
    while True:
        out.append(t)
        t -= cool_rate
        if t < min_temperature:
            return out


def shard_points(points, rank: int, world: int):
    """Interleaved shard (rank, rank+world, ...) so that every rank spans the whole temperature range."""
    return list(points[rank::world])


def sharded_cooldown(machine, temperatures, cool_rate: float, relax: int, steps: int):
    """Run this rank's points through the host Machine: each is a one-point CoolDown (relax, then measure),
    so the StatSensor / ObservableSensor hooks fire exactly as in the reference program."""
    for t in temperatures:
        machine.cooldown(t, t, cool_rate, relax, steps)


def gather_lines(lines, dist, group=None):
    """All ranks' StatSensor lines on rank 0, sorted by temperature descending (the order CoolDown prints them)."""
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, list(lines), group=group)
    merged = [ln for p in parts for ln in p]
    merged.sort(key=lambda ln: -float(ln.split()[0]))
    return merged
