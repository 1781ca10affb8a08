"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/vegas_gpu.h
declares, refuses to run without a CUDA device, and its host-side lattice logic agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vegas_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vegas_gpu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from vegas_rs_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 35
    bound = {s[0] for s in _lib.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vegas_gpu.h but not exported"
        assert n in bound, f"{n} has no ctypes signature in vegas_rs_b200/_lib.py"
    assert b"sm_100a" in lib.vegas_gpu_version()


def test_extension_is_sm100a_native(built):
    """The shipped library carries sm_100a SASS (no PTX-JIT fallback for another arch)."""
    import subprocess
    from vegas_rs_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(built):
    import torch
    import vegas_rs_b200 as vg
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(vg.VegasGpuError) as ei:
        vg.GpuMetropolis(vg.ISING, unitcell=vg.SC, size=(64, 4, 4))
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vegas_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle/", "ORACLE_DIR_MENTION") or \
                    not re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle", text, flags=re.M), f


@pytest.mark.parametrize("uc", [0, 1, 2])
def test_host_lattice_matches_oracle(built, uc):
    """vegas_rs_b200/csrc/lattice.hpp (product) and oracle/vegas_oracle.c are independent restatements of
    Exchange::from_lattice; they must produce the same CSR for every boundary rule."""
    from oracle import binding as ob
    from vegas_rs_b200.gpu_metropolis import lattice_adjacency, lattice_colours
    for size in ((4, 4, 4), (3, 5, 2), (1, 1, 1), (2, 2, 2), (5, 1, 3), (4, 6, 1)):
        for pbc in ((1, 1, 1), (0, 1, 1), (1, 0, 0), (0, 0, 0)):
            for lit in (False, True):
                rp, ci, va = lattice_adjacency(uc, size, pbc, lit, 1.5)
                orp, oci, ova = ob.Csr.from_lattice(ob.Lattice(uc, *size, pbc=pbc), 1.5, lit).arrays()
                assert np.array_equal(rp, orp) and np.array_equal(ci.astype(np.uint64), oci) and np.array_equal(va, ova)
                nc, col = lattice_colours(uc, size, pbc, lit)
                rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp.astype(np.int64)))
                assert not np.any((col[rows] == col[ci]) & (rows != ci)), (uc, size, pbc, lit)
                assert nc <= 4 and col.max() < nc


def test_basis_kernel_neighbour_tables_match_lattice(built):
    """heis_basis.cuh bakes the bcc / fcc unit-cell bonds into the kernel at compile time; the same library checks
    them (on the host, no GPU) against the edge list its adjacency export and the oracle comparison are built from."""
    from vegas_rs_b200 import _lib
    assert _lib.load().vegas_gpu_check_basis_tables() == 0


def test_bench_reference_arm_line(built):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    on the reference's own config[0] lattice so that it finishes in seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", "ising_sc10_cfg0"], capture_output=True, text=True, timeout=300, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "spin-flip attempts/sec" and line["unit"] == "attempts/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["value"] > 1e5 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "ising_sc10_cfg0"


@pytest.mark.parametrize("n_chunks,lag,steps", [(2, 3, 1), (3, 3, 2), (4, 3, 4), (12, 3, 3), (12, 5, 1), (128, 5, 1), (128, 3, 2),
                                                (64, 4, 4), (7, 9, 2), (5, 1, 3)])
def test_wave_schedule_dependencies_precede_their_users(built, n_chunks, lag, steps):
    """The persistent wave kernel deals its work units to co-resident CTAs in list order and a unit spins until the
    previous colour-pass phase is complete on its chunk and the two neighbouring chunks (csrc/heis.cuh, K3w).  That can
    only terminate if every one of those three units comes EARLIER in the list; each (phase, chunk) must appear once."""
    from vegas_rs_b200 import _lib
    lib = _lib.load()
    count = C.c_uint64()
    assert lib.vegas_gpu_wave_schedule(n_chunks, lag, steps, None, 0, C.byref(count)) == 0
    assert count.value == 2 * steps * n_chunks
    units = np.zeros(count.value, np.uint32)
    assert lib.vegas_gpu_wave_schedule(n_chunks, lag, steps, units.ctypes.data_as(C.c_void_p), units.size, C.byref(count)) == 0
    phase, chunk = units >> 24, units & 0xFFFFFF
    position = {(int(p), int(c)): i for i, (p, c) in enumerate(zip(phase, chunk))}
    assert len(position) == units.size and set(position) == {(p, c) for p in range(2 * steps) for c in range(n_chunks)}
    for (p, c), i in position.items():
        if p == 0:
            continue
        for dep in ((c - 1) % n_chunks, c, (c + 1) % n_chunks):
            assert position[(p - 1, dep)] < i, (p, c, dep)
    # phase p starts at chunk p (rotated order): the wrap-around chunk is the last one a phase visits
    for p in range(2 * steps):
        mine = [int(c) for q, c in zip(phase, chunk) if q == p]
        assert mine == [(p + k) % n_chunks for k in range(n_chunks)]
    assert lib.vegas_gpu_wave_schedule(0, 3, 1, None, 0, C.byref(count)) != 0
    assert lib.vegas_gpu_wave_schedule(8, 3, 5, None, 0, C.byref(count)) != 0


@pytest.mark.parametrize("uc,nb", [(1, 2), (2, 4)], ids=["bcc", "fcc"])
@pytest.mark.parametrize("nz,lag", [(384, 3), (32, 1), (12, 3), (7, 1), (40, 5)])
def test_basis_wave_schedule_dependencies_precede_their_users(built, uc, nb, nz, lag):
    """basis_wave.cu deals (colour, plane) units to co-resident CTAs in list order; a unit spins until every bonded LOWER
    colour is complete on the planes its bonds reach.  Checked against the adjacency itself (vegas_gpu_lattice_adjacency):
    every neighbour of a site of colour b on plane z that belongs to a lower colour lies on plane z or z + 1, the schedule
    waits for exactly those (need mask), and the unit it waits for comes EARLIER in the list."""
    from vegas_rs_b200 import _lib
    lib = _lib.load()
    count = C.c_uint64()
    need = np.zeros(4, np.uint32)
    assert lib.vegas_gpu_basis_wave_schedule(uc, nz, lag, None, 0, C.byref(count), need.ctypes.data_as(C.c_void_p)) == 0
    assert count.value == nb * nz
    units = np.zeros(count.value, np.uint32)
    assert lib.vegas_gpu_basis_wave_schedule(uc, nz, lag, units.ctypes.data_as(C.c_void_p), units.size, C.byref(count), need.ctypes.data_as(C.c_void_p)) == 0
    colour, plane = units >> 24, units & 0xFFFFFF
    position = {(int(b), int(z)): i for i, (b, z) in enumerate(zip(colour, plane))}
    assert len(position) == units.size and set(position) == {(b, z) for b in range(nb) for z in range(nz)}
    # what the bonds really reach: a 4 x 4 x 6 lattice of the same unit cell (site = cell * nb + basis)
    d = _lib.LatticeDesc(uc, 4, 4, 6, 1, 1, 1, 0, 0, 0)
    n, nnz = C.c_uint64(), C.c_uint64()
    assert lib.vegas_gpu_lattice_adjacency(C.byref(d), 1.0, C.byref(n), C.byref(nnz), None, None, None) == 0
    rp = np.zeros(n.value + 1, np.uint64); col = np.zeros(nnz.value, np.uint32)
    assert lib.vegas_gpu_lattice_adjacency(C.byref(d), 1.0, C.byref(n), C.byref(nnz), rp.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), None) == 0
    reach = {}
    for i in range(n.value):
        b, zi = i % nb, (i // nb) // 16
        for j in col[int(rp[i]):int(rp[i + 1])]:
            a, zj = int(j) % nb, (int(j) // nb) // 16
            reach.setdefault((b, a), set()).add((zj - zi + 3) % 6 - 3)
    for b in range(nb):
        want = 0
        for a in range(b):
            dzs = reach.get((b, a), set())
            assert dzs <= {0, 1}, (b, a, dzs)          # lower colours only at dz >= 0: the structure the scheme relies on
            for r in dzs:
                want |= 1 << (2 * a + r)
        assert int(need[b]) == want, (b, hex(int(need[b])), hex(want))
    for (b, z), i in position.items():
        for a in range(b):
            for r in range(2):
                if (int(need[b]) >> (2 * a + r)) & 1:
                    assert position[(a, (z + r) % nz)] < i, (b, z, a, r)
    # each colour walks the planes in order
    for b in range(nb):
        assert [int(z) for q, z in zip(colour, plane) if q == b] == list(range(nz))
    assert lib.vegas_gpu_basis_wave_schedule(0, nz, lag, None, 0, C.byref(count), need.ctypes.data_as(C.c_void_p)) != 0
    assert lib.vegas_gpu_basis_wave_schedule(uc, 3, 3, None, 0, C.byref(count), need.ctypes.data_as(C.c_void_p)) == 0 and count.value == 0


@pytest.mark.parametrize("n_words,threads,chunk", [(1, 1, 1), (37, 1, 8), (1000, 3, 64), (4096, 4, 1 << 20), (8192, 3, 1000), (20000, 5, 700)])
def test_host_pack_round_trip(built, n_words, threads, chunk):
    """The host side of the bitmap State transfer (csrc/host_pack.cpp): words[i] bit b = (s[32 i + b] > 0) for the reference's
    +1 / -1 bytes (src/state.rs:60-63), chunked over worker threads; unpack restores the State byte for byte."""
    from vegas_rs_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n_words)
    s = (2 * rng.integers(0, 2, 32 * n_words) - 1).astype(np.int8)
    words = np.zeros(n_words, np.uint32)
    assert lib.vegas_gpu_host_pack(s.ctypes.data, words.ctypes.data, n_words, threads, chunk) == 0
    want = np.packbits((s > 0).reshape(n_words, 32), axis=1, bitorder="little").view(np.uint32).ravel()
    assert np.array_equal(words, want)
    back = np.zeros_like(s)
    assert lib.vegas_gpu_host_unpack(words.ctypes.data, back.ctypes.data, n_words, threads, chunk) == 0
    assert np.array_equal(back, s)
    # an unaligned State buffer (numpy slices) takes the unaligned store path
    buf = np.zeros(32 * n_words + 1, np.int8)
    assert lib.vegas_gpu_host_unpack(words.ctypes.data, buf[1:].ctypes.data, n_words, threads, chunk) == 0
    assert np.array_equal(buf[1:], s)
    assert lib.vegas_gpu_host_pack(None, words.ctypes.data, n_words, threads, chunk) != 0


def test_basis_pair_structure_matches_the_adjacency(built):
    """heis_basis_pair_kernel updates colours 2k and 2k + 1 in one launch without inter-CTA synchronisation; that is only right if
    every neighbour of a colour-(2k + 1) site inside colour 2k lies in the SAME cell plane and in row y or y + 1.  Checked here
    against the adjacency the lattice generator exports (vegas_gpu_lattice_adjacency), independently of the kernel's tables."""
    from vegas_rs_b200 import _lib
    lib = _lib.load()
    for uc, nb, want in ((2, 4, 1), (1, 2, 0)):
        d = _lib.LatticeDesc(uc, 5, 6, 7, 1, 1, 1, 0, 0, 0)
        n, nnz = C.c_uint64(), C.c_uint64()
        assert lib.vegas_gpu_lattice_adjacency(C.byref(d), 1.0, C.byref(n), C.byref(nnz), None, None, None) == 0
        rp = np.zeros(n.value + 1, np.uint64); col = np.zeros(nnz.value, np.uint32)
        assert lib.vegas_gpu_lattice_adjacency(C.byref(d), 1.0, C.byref(n), C.byref(nnz), rp.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), None) == 0
        ok = True
        for i in range(n.value):
            b, cell = i % nb, i // nb
            if b % 2 == 0:
                continue
            yi, zi = (cell // 5) % 6, cell // 30
            for j in col[int(rp[i]):int(rp[i + 1])]:
                a, cj = int(j) % nb, int(j) // nb
                if a != b - 1:
                    continue
                yj, zj = (cj // 5) % 6, cj // 30
                dy, dz = (yj - yi + 3) % 6 - 3, (zj - zi + 3) % 7 - 3
                ok = ok and dz == 0 and dy in (0, 1)
        assert int(ok) == want == lib.vegas_gpu_basis_pair_structure(uc), (uc, ok)
    assert lib.vegas_gpu_basis_pair_structure(0) == 0


def test_rust_shim_struct_layout_matches_header():
    """bindings/rust/gpu.rs cannot be compiled here; at least its #[repr(C)] structs must list the fields of
    include/vegas_gpu.h in the same order, and every extern function it declares must exist in the headers."""
    rs = open(os.path.join(ROOT, "bindings", "rust", "gpu.rs")).read()
    hdr = open(os.path.join(ROOT, "include", "vegas_gpu.h")).read() + open(os.path.join(ROOT, "include", "vegas_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name in ("vegas_model_desc", "vegas_lattice_desc"):
        end = hdr.index("} " + name + ";")
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if " " in decl else decl
            c_fields += [re.sub(r"\[.*\]", "", n).strip() for n in names.split(",")]
        rs_body = re.search(r"pub struct " + name + r" \{(.*?)\n\}", rs, flags=re.S).group(1)
        rs_fields = re.findall(r"pub (\w+):", rs_body)
        assert rs_fields == c_fields, (name, rs_fields, c_fields)
    for fn in re.findall(r"\bfn (vegas_\w+)\(", rs):
        assert re.search(r"\b" + fn + r"\s*\(", hdr), fn


def test_create_rejects_bad_arguments_before_touching_the_device(built):
    """Argument validation comes before the CUDA device check and never aborts: status VEGAS_ERR_INVALID (-1) with a
    message (the reference returns Result everywhere, src/error.rs) -- testable without a GPU."""
    import vegas_rs_b200 as vg

    def code_of(**kw):
        with pytest.raises(vg.VegasGpuError) as ei:
            vg.GpuMetropolis(kw.pop("model", vg.ISING), **kw)
        return ei.value.code, str(ei.value)

    assert code_of(unitcell=vg.SC, size=(0, 4, 4)) == (-1, "vegas_gpu error -1: empty lattice")
    assert code_of(unitcell=vg.SC, size=(4, 4, 0))[0] == -1
    assert code_of(unitcell=7, size=(4, 4, 4)) == (-1, "vegas_gpu error -1: unknown unit cell")
    assert code_of(model=5, unitcell=vg.SC, size=(4, 4, 4)) == (-1, "vegas_gpu error -1: unknown model")
    assert code_of(unitcell=vg.SC, size=(4, 4, 4), proposal=9)[1].endswith("unknown proposal")
    assert code_of(model=vg.HEISENBERG, unitcell=vg.SC, size=(4, 4, 4), precision=3)[1].endswith("unknown precision")
    # Exchange::new(CsMat): malformed CSR
    rp = np.array([0, 2, 3], np.uint64); ci = np.array([1, 5, 0], np.uint32)
    assert code_of(csr=(rp, ci, None))[1].endswith("column index out of range")
    assert code_of(csr=(np.array([0, 2, 1], np.uint64), np.array([1, 0], np.uint32), None))[1].endswith("row_ptr not monotone")
    wide = np.array([0, 40, 40], np.uint64)
    assert code_of(csr=(wide, np.ones(40, np.uint32), None))[1].endswith("rows longer than 32 are not supported")


def test_documented_tuning_keys_exist_in_the_library_source():
    """include/vegas_gpu.h documents the kernel-selection knobs; each must be handled by vegas_gpu_set_tuning and vice versa."""
    hdr = open(os.path.join(ROOT, "include", "vegas_gpu.h")).read()
    block = hdr[hdr.index("kernel selection knobs"):hdr.index("int vegas_gpu_set_tuning")]
    documented = set(re.findall(r'"([a-z_]+)"', block))
    src = open(os.path.join(ROOT, "vegas_rs_b200", "csrc", "vegas_gpu.cu")).read()
    body = src[src.index("int vegas_gpu_set_tuning("):src.index("const char* vegas_gpu_step_kernel(")]
    handled = set(re.findall(r'k == "([a-z_]+)"', body))
    assert documented == handled, (sorted(documented - handled), sorted(handled - documented))
