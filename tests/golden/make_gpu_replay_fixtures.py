"""Generates tests/golden/gpu_replay_*.npz ON A B200: small trajectories of the CUDA sweep kernels (inputs, per-sweep
energies, final state) so that the CPU suite can check the oracle's replay of the reference Metropolis rule against
RECORDED GPU output without a GPU (tests/test_oracle.py::test_oracle_replays_committed_gpu_trajectories).

    gpurun -- 'python tests/golden/make_gpu_replay_fixtures.py gpurun_out/golden'   # then copy the .npz files here
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vegas_rs_b200 as vg  # noqa: E402

CASES = [
    dict(name="ising_msc_3d_field", model="ising", unitcell=0, size=(64, 6, 4), seed=101, T=4.0, H=0.625, sweeps=3),
    dict(name="ising_resident_sc10", model="ising", unitcell=0, size=(10, 10, 10), seed=102, T=4.5, H=0.25, sweeps=4),
    dict(name="heis_stencil_f64", model="heisenberg", unitcell=0, size=(16, 4, 4), seed=103, T=1.0, H=0.5, sweeps=2),
    dict(name="heis_fcc_vec_f64", model="heisenberg", unitcell=2, size=(4, 3, 4), seed=104, T=1.5, H=0.7, sweeps=2,
         anisotropy=((0.6, 0.0, 0.8), 0.25)),
]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for c in CASES:
        rng = np.random.default_rng(c["seed"])
        nb = {0: 1, 1: 2, 2: 4}[c["unitcell"]]
        n = int(np.prod(c["size"])) * nb
        if c["model"] == "ising":
            g = vg.GpuMetropolis(vg.ISING, unitcell=c["unitcell"], size=c["size"], seed=c["seed"])
            s0 = (2 * rng.integers(0, 2, n) - 1).astype(np.int8)
        else:
            g = vg.GpuMetropolis(vg.HEISENBERG, unitcell=c["unitcell"], size=c["size"], seed=c["seed"], precision=vg.F64,
                                 anisotropy=c.get("anisotropy"))
            v = rng.normal(size=(n, 3))
            s0 = v / np.linalg.norm(v, axis=1, keepdims=True)
        g.upload(s0)
        g.set_thermostat(c["T"], (0.0, 0.0, 1.0), c["H"])
        colours = g.colours()
        e, m = g.step(c["sweeps"])
        final = g.download()
        meta = dict(c, kernel_family=g.kernel_family, step_kernel=g.step_kernel, n_colours=g.n_colours,
                    library=vg._lib.load().vegas_gpu_version().decode())
        np.savez_compressed(os.path.join(out_dir, f"gpu_replay_{c['name']}.npz"), meta=json.dumps(meta), initial=s0,
                            colours=colours, energy=e, magnetization=m, final=final)
        print(c["name"], g.kernel_family, g.step_kernel, "E", e, flush=True)
        g.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
