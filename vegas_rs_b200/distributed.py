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
