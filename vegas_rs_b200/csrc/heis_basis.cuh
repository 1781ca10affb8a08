// heis_basis.cuh -- K4b: Heisenberg colour pass for periodic bcc / fcc lattices (2 / 4 sites per cell, z = 8 / 12).
//
// Replaces MetropolisIntegrator::step / MetropolisFlipIntegrator::step (src/integrator.rs:66-92, :109-138) over the
// compound Hamiltonian (src/energy.rs:63-257) for the lattices `Lattice::bcc / fcc (..).expand(x, y, z)` builds
// (src/input.rs:296-322).  Colour = basis index (the sublattices are independent sets), one launch per colour.
//
// Layout: basis-split SoA, s[basis][component][cell] with cell = (iz*ny + iy)*nx + ix, so that a warp (32 consecutive
// ix of one row) reads every neighbour sublattice as one coalesced segment; the unit-cell neighbour table is a
// compile-time constant (template on the unit cell and the basis), the nine (dy, dz) row bases are uniform per CTA, and
// no index array is ever read (24 B/attempt algorithmic).  Same random numbers as the general-adjacency
// kernel (heis_rand keyed by the NATURAL site index cell*nb + basis), same heis_attempt.
#pragma once
#include <utility>

#include "heis.cuh"

namespace vg {

struct BasisNb { int tb, dx, dy, dz; };

// Unit-cell bonds in the order of lattice.hpp (vgl::unitcell_edges): q-th bond of the cell, UC 1 = bcc, 2 = fcc.
template <int UC> struct BasisCell;
template <> struct BasisCell<1> {
    static constexpr int NB = 2, NE = 8, Z = 8;
    // A(0,0,0) - B(1/2,1/2,1/2): dx, dy, dz in {0,-1}, dx fastest
    static constexpr __host__ __device__ void edge(int q, int& s, int& t, int& dx, int& dy, int& dz) {
        s = 0; t = 1; dx = -(q & 1); dy = -((q >> 1) & 1); dz = -((q >> 2) & 1);
    }
};
template <> struct BasisCell<2> {
    static constexpr int NB = 4, NE = 24, Z = 12;
    // pairs A-B, A-C, A-D, B-C, B-D, C-D; each 2x2 offsets on the two half-integer axes of the pair (first axis fastest)
    static constexpr __host__ __device__ void edge(int q, int& s, int& t, int& dx, int& dy, int& dz) {
        constexpr int S[6] = {0, 0, 0, 1, 1, 2}, T[6] = {1, 2, 3, 2, 3, 3};
        constexpr int AX0[6] = {0, 0, 1, 1, 0, 0}, LO0[6] = {-1, -1, -1, 0, 0, 0};
        constexpr int AX1[6] = {1, 2, 2, 2, 2, 1}, LO1[6] = {-1, -1, -1, -1, -1, -1};
        const int pr = q >> 2, a = LO0[pr] + (q & 1), b = LO1[pr] + ((q >> 1) & 1);
        int d[3] = {0, 0, 0};
        d[AX0[pr]] = a; d[AX1[pr]] = b;
        s = S[pr]; t = T[pr]; dx = d[0]; dy = d[1]; dz = d[2];
    }
};

// q-th neighbour of basis B: the bonds with source B in table order, then the bonds with target B reversed
// (the order StructuredNb / vgl::for_each_neighbour enumerate them in).
template <int UC, int B>
constexpr __host__ __device__ BasisNb basis_neighbour(int q) {
    int k = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int e = 0; e < BasisCell<UC>::NE; ++e) {
            int s = 0, t = 0, dx = 0, dy = 0, dz = 0;
            BasisCell<UC>::edge(e, s, t, dx, dy, dz);
            if (pass == 0 ? s != B : t != B) continue;
            if (k++ == q) return pass == 0 ? BasisNb{t, dx, dy, dz} : BasisNb{s, -dx, -dy, -dz};
        }
    return BasisNb{0, 0, 0, 0};
}

template <typename F, size_t... Q>
__device__ __forceinline__ void basis_for_each(F&& f, std::index_sequence<Q...>) {
    (f(std::integral_constant<int, (int)Q>{}), ...);
}

template <typename real>
struct BasisPtrs { real* s[4][3]; };  // [basis][component][cell]

// z-slab (multi-GPU): the handle owns cell planes [z_offset, z_offset + nz) of nz_global; every array then carries one
// halo plane below (index -1) and above (index nz), filled by the z-neighbours, and `ext` = (nz + 2) * ny * nx is the
// element distance between consecutive arrays of the single allocation (used for the stores into the peers' halos).
struct BasisGeom { uint32_t nx, ny, nz, ncells, z_offset, nz_global; size_t ext; };

template <typename real>
struct BasisPeers { real* lo; real* hi; };  // bases of the lower / upper neighbour's allocation (null: not a slab)

// MODE 0: update.  1: update + observables of this colour: obs[1..3] += s, obs[4] += (s.a)^2 and the exchange bonds
// towards the LOWER colours (final by now) so that a step counts every bond once; obs[0] receives twice that sum
// (layout of general_reduce_kernel: sum_i sum_j J s_i.s_j, every bond twice).  2: the same reductions, no update.
// obs[5] += accepted.
#ifndef BASIS_MINB
#define BASIS_MINB 1   // no register cap: capping the recorded-step variants at 40 registers (12 CTAs per SM) spills and
#endif                 // was measured slower (4.02 vs 3.42 ms per fcc 384^3 step), 32 registers slower still (4.27 ms)
template <typename real, int UC, int B, bool FLIP, int MODE, bool SLAB = false>
__global__ void __launch_bounds__(128, BASIS_MINB)
heis_basis_kernel(BasisPtrs<real> P, BasisPeers<real> peers, BasisGeom g, uint32_t rows_per_cta, uint32_t z_begin, uint32_t z_step,
                  HeisParams<real> p, uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    constexpr int NB = BasisCell<UC>::NB, Z = BasisCell<UC>::Z;
    __shared__ double s_red[6 * 32];
    // a thread owns one ix and marches over rows_per_cta rows of plane iz: one block reduction per CTA
    // plane of this CTA: z_begin + blockIdx.z * z_step (z_step > 1: the two boundary planes of a slab in one launch)
    const uint32_t ix = blockIdx.x * blockDim.x + threadIdx.x, iz = z_begin + blockIdx.z * z_step;
    const uint32_t y0 = blockIdx.y * rows_per_cta, y1 = min(y0 + rows_per_cta, g.ny);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (ix < g.nx) {
        // slab: planes -1 and nz are the halo planes of the same array (signed plane index); else periodic wrap
        const int zs[3] = {SLAB ? (int)iz - 1 : (int)(iz == 0 ? g.nz - 1 : iz - 1), (int)iz,
                           SLAB ? (int)iz + 1 : (int)(iz + 1 == g.nz ? 0u : iz + 1)};
        const uint32_t xs[3] = {ix == 0 ? g.nx - 1 : ix - 1, ix, ix + 1 == g.nx ? 0u : ix + 1};
        real fs[5] = {0, 0, 0, 0, 0};
        int accepted = 0;
        for (uint32_t iy = y0; iy < y1; ++iy) {
            const uint32_t ys[3] = {iy == 0 ? g.ny - 1 : iy - 1, iy, iy + 1 == g.ny ? 0u : iy + 1};  // uniform per CTA
            const uint32_t cell = (iz * g.ny + iy) * g.nx + ix;
            real nx = 0, ny = 0, nz = 0, lx = 0, ly = 0, lz = 0;
            auto gather = [&](auto qtag) {
                constexpr int Q = decltype(qtag)::value;
                constexpr BasisNb nb = basis_neighbour<UC, B>(Q);
                const int j = (zs[nb.dz + 1] * (int)g.ny + (int)ys[nb.dy + 1]) * (int)g.nx + (int)xs[nb.dx + 1];
                const real u = P.s[nb.tb][0][j], v = P.s[nb.tb][1][j], w = P.s[nb.tb][2][j];
                nx += u; ny += v; nz += w;
                if (MODE != 0 && nb.tb < B) { lx += u; ly += v; lz += w; }
            };
            basis_for_each(gather, std::make_index_sequence<Z>{});
            real x = P.s[B][0][cell], y = P.s[B][1][cell], z = P.s[B][2][cell];
            if (MODE != 2) {
                HeisRand<real> rnd;
                const uint64_t gcell = ((uint64_t)(iz + g.z_offset) * g.ny + iy) * g.nx + ix;  // global cell: slab-independent keys
                heis_rand(gcell * NB + B, sweep, pk, rnd);
                const bool ok = heis_attempt<real, FLIP>(x, y, z, p.J * nx - p.h[0], p.J * ny - p.h[1], p.J * nz - p.h[2], p, rnd);
                if (ok) { P.s[B][0][cell] = x; P.s[B][1][cell] = y; P.s[B][2][cell] = z; }
                if (SLAB) {  // boundary planes also go straight into the neighbours' halo planes (peer memory over NVLink)
                    const size_t in_plane = (size_t)iy * g.nx + ix, pl = (size_t)g.ny * g.nx;
                    if (iz == 0 && peers.lo != nullptr) {
                        real* q = peers.lo + (size_t)(B * 3) * g.ext + (size_t)(g.nz + 1) * pl + in_plane;  // their plane "nz"
                        q[0] = x; q[g.ext] = y; q[2 * g.ext] = z;
                    }
                    if (iz + 1 == g.nz && peers.hi != nullptr) {
                        real* q = peers.hi + (size_t)(B * 3) * g.ext + in_plane;                            // their plane "-1"
                        q[0] = x; q[g.ext] = y; q[2 * g.ext] = z;
                    }
                }
                accepted += ok ? 1 : 0;
            }
            if (MODE != 0) {
                fs[0] += x * lx + y * ly + z * lz;
                fs[1] += x; fs[2] += y; fs[3] += z;
                const real d = x * p.a[0] + y * p.a[1] + z * p.a[2];
                fs[4] += d * d;
            }
        }
        if (MODE != 0) {
            acc[0] = 2.0 * (double)p.J * (double)fs[0];
#pragma unroll
            for (int i = 1; i < 5; ++i) acc[i] = (double)fs[i];
        }
        acc[5] = (double)accepted;
    }
    if (MODE == 0) {
        double a1[1] = {acc[5]};
        block_atomic_add<double, 1>(a1, s_red, obs + 5);
    } else {
        block_atomic_add<double, 6>(acc, s_red, obs);
    }
}

// ---- host layout <-> basis-split SoA, fills, random state ------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) basis_pack_kernel(const double* __restrict__ aos, BasisPtrs<real> P, uint32_t nb, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // natural site index cell*nb + b
    if (i >= n) return;
    const size_t cell = i / nb; const uint32_t b = (uint32_t)(i - cell * nb);
#pragma unroll
    for (int c = 0; c < 3; ++c) P.s[b][c][cell] = (real)aos[3 * i + c];
}

// stride 3 with o = aos, aos+1, aos+2 writes AoS; stride 1 writes natural-order SoA (per-site energy path)
template <typename real, typename outT>
__global__ void __launch_bounds__(256) basis_unpack_kernel(outT* __restrict__ ox, outT* __restrict__ oy, outT* __restrict__ oz,
                                                           size_t stride, BasisPtrs<real> P, uint32_t nb, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t cell = i / nb; const uint32_t b = (uint32_t)(i - cell * nb);
    ox[i * stride] = (outT)P.s[b][0][cell]; oy[i * stride] = (outT)P.s[b][1][cell]; oz[i * stride] = (outT)P.s[b][2][cell];
}

template <typename real>
__global__ void __launch_bounds__(256) basis_randomize_kernel(BasisPtrs<real> P, uint32_t nb, size_t n, uint64_t site_offset, PhiloxKey pk) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t cell = i / nb; const uint32_t b = (uint32_t)(i - cell * nb);
    real x, y, z;
    heis_random_spin<real>((uint64_t)i + site_offset, pk, x, y, z);  // keyed by the natural GLOBAL site index, as the general kernels
    P.s[b][0][cell] = x; P.s[b][1][cell] = y; P.s[b][2][cell] = z;
}

}  // namespace vg
