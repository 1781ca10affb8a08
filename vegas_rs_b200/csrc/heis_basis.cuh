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
                const bool ok = heis_attempt<real, FLIP>(x, y, z, heis_field(p.J, nx, p.h[0]), heis_field(p.J, ny, p.h[1]), heis_field(p.J, nz, p.h[2]), p, rnd);
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

// ---------------------------------------------------------------------------------------
// K4v: the same colour pass with 16-byte accesses (nx % N == 0, N = 4 floats / 2 doubles).  One work item = N
// consecutive cells of a row; the items of a plane are flattened (row, vector) and dealt to the CTA's threads in order, so
// a warp reads consecutive vectors whatever nx is.  A neighbour row (partner basis, dy, dz) is loaded once as a vector:
// the bonds with dx = 0 take its elements as they are, the bonds with dx = -1 / +1 the row shifted by one cell with one
// scalar carry from the adjacent vector (periodic in x).  In the unit-cell tables the two x offsets of a row are adjacent
// entries, so the walk below (table order, exactly the summation order of the scalar kernel and of the general-adjacency
// kernel) re-uses the vector the previous entry loaded.  Same random numbers, same heis_attempt, same observables.
// ---------------------------------------------------------------------------------------
#ifndef BASIS_VEC_MINB
#define BASIS_VEC_MINB 8   // 64 registers.  Uncapped the compiler hoists all 39 loads of an item (200 registers, 2 CTAs per
#endif                     // SM); measured ms per fcc 384^3 step by cap 1/3/4/5/6/8: 2.79 / 2.85 / 2.69 / 2.84 / 2.56 / 2.51
                           // (profiles/r01y_fcc_vec_probe.txt)
// One work item of K4v: the N cells x0 .. x0+N-1 of row iy of plane iz (w = iy * VX + x0 / N).  Shared by the per-colour
// launches (heis_basis_vec_kernel) and the wave-ordered persistent step (basis_wave.cu): same loads, same summation
// order, same random numbers.
// W: where the updated spins go (the pair kernel below writes a second set of arrays; everywhere else W is P).  counted = false:
// a redundantly computed row -- same update, but it enters neither the observables nor the accepted count.
template <typename real, int UC, int B, bool FLIP, int MODE, bool SLAB>
__device__ __forceinline__ void basis_vec_item(const BasisPtrs<real>& P, const BasisPtrs<real>& W, const BasisPeers<real>& peers, const BasisGeom& g,
                                               uint32_t iz, const int (&zs)[3], uint32_t w, uint32_t VX, const HeisParams<real>& p, uint64_t sweep,
                                               const PhiloxKey& pk, real (&fs)[5], int& accepted, bool counted = true) {
    constexpr int NB = BasisCell<UC>::NB, Z = BasisCell<UC>::Z, N = VecOf<real>::N;
    const uint32_t iy = w / VX, x0 = (w - iy * VX) * N;
    const uint32_t ys[3] = {iy == 0 ? g.ny - 1 : iy - 1, iy, iy + 1 == g.ny ? 0u : iy + 1};
    const uint32_t xl = x0 == 0 ? g.nx - 1 : x0 - 1, xr = x0 + N == g.nx ? 0u : x0 + N;   // carries (periodic in x)
    const int cell = ((int)iz * (int)g.ny + (int)iy) * (int)g.nx + (int)x0;
    real n[3][N], l[3][N];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < N; ++e) { n[c][e] = 0; l[c][e] = 0; }
    auto gather = [&](auto qtag) {
        constexpr int Q = decltype(qtag)::value;
        constexpr BasisNb nb = basis_neighbour<UC, B>(Q);
        const int row = (zs[nb.dz + 1] * (int)g.ny + (int)ys[nb.dy + 1]) * (int)g.nx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            real v[N];
            vec_load(P.s[nb.tb][c] + (row + (int)x0), v);   // identical address for the two x offsets of a row: loaded once
            real u[N];
            if (nb.dx == 0) {
#pragma unroll
                for (int e = 0; e < N; ++e) u[e] = v[e];
            } else if (nb.dx < 0) {
                u[0] = P.s[nb.tb][c][row + (int)xl];
#pragma unroll
                for (int e = 1; e < N; ++e) u[e] = v[e - 1];
            } else {
#pragma unroll
                for (int e = 0; e + 1 < N; ++e) u[e] = v[e + 1];
                u[N - 1] = P.s[nb.tb][c][row + (int)xr];
            }
#pragma unroll
            for (int e = 0; e < N; ++e) {
                n[c][e] += u[e];
                if (MODE != 0 && nb.tb < B) l[c][e] += u[e];
            }
        }
    };
    basis_for_each(gather, std::make_index_sequence<Z>{});
    real sx[N], sy[N], sz[N];
    vec_load(P.s[B][0] + cell, sx); vec_load(P.s[B][1] + cell, sy); vec_load(P.s[B][2] + cell, sz);
    if (MODE != 2) {
        const uint64_t gcell = ((uint64_t)(iz + g.z_offset) * g.ny + iy) * g.nx + x0;  // global cell: slab-independent keys
#pragma unroll
        for (int e = 0; e < N; ++e) {
            HeisRand<real> rnd;
            heis_rand((gcell + e) * NB + B, sweep, pk, rnd);
            const bool ok = heis_attempt<real, FLIP>(sx[e], sy[e], sz[e], heis_field(p.J, n[0][e], p.h[0]), heis_field(p.J, n[1][e], p.h[1]),
                                                     heis_field(p.J, n[2][e], p.h[2]), p, rnd);
            accepted += (ok && counted) ? 1 : 0;
        }
        vec_store(W.s[B][0] + cell, sx); vec_store(W.s[B][1] + cell, sy); vec_store(W.s[B][2] + cell, sz);
        if (SLAB) {  // boundary planes also go straight into the neighbours' halo planes (peer memory over NVLink)
            const size_t in_plane = (size_t)iy * g.nx + x0, pl = (size_t)g.ny * g.nx;
            if (iz == 0 && peers.lo != nullptr) {
                real* q = peers.lo + (size_t)(B * 3) * g.ext + (size_t)(g.nz + 1) * pl + in_plane;  // their plane "nz"
                vec_store(q, sx); vec_store(q + g.ext, sy); vec_store(q + 2 * g.ext, sz);
            }
            if (iz + 1 == g.nz && peers.hi != nullptr) {
                real* q = peers.hi + (size_t)(B * 3) * g.ext + in_plane;                            // their plane "-1"
                vec_store(q, sx); vec_store(q + g.ext, sy); vec_store(q + 2 * g.ext, sz);
            }
        }
    }
    if (MODE != 0 && counted) {
#pragma unroll
        for (int e = 0; e < N; ++e) {
            fs[0] += sx[e] * l[0][e] + sy[e] * l[1][e] + sz[e] * l[2][e];
            fs[1] += sx[e]; fs[2] += sy[e]; fs[3] += sz[e];
            const real d = sx[e] * p.a[0] + sy[e] * p.a[1] + sz[e] * p.a[2];
            fs[4] += d * d;
        }
    }
}

template <typename real, int UC, int B, bool FLIP, int MODE, bool SLAB = false>
__global__ void __launch_bounds__(128, BASIS_VEC_MINB)
heis_basis_vec_kernel(BasisPtrs<real> P, BasisPeers<real> peers, BasisGeom g, uint32_t items_per_thread, uint32_t z_begin,
                      uint32_t z_step, HeisParams<real> p, uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    constexpr int N = VecOf<real>::N;
    __shared__ double s_red[6 * 32];
    const uint32_t iz = z_begin + blockIdx.z * z_step;
    const uint32_t VX = g.nx / N, items = VX * g.ny;
    const int zs[3] = {SLAB ? (int)iz - 1 : (int)(iz == 0 ? g.nz - 1 : iz - 1), (int)iz,
                       SLAB ? (int)iz + 1 : (int)(iz + 1 == g.nz ? 0u : iz + 1)};
    real fs[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    const uint32_t w0 = blockIdx.x * (items_per_thread * blockDim.x) + threadIdx.x;
    for (uint32_t it = 0; it < items_per_thread; ++it) {
        const uint32_t w = w0 + it * blockDim.x;
        if (w >= items) break;
        basis_vec_item<real, UC, B, FLIP, MODE, SLAB>(P, P, peers, g, iz, zs, w, VX, p, sweep, pk, fs, accepted);
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (MODE != 0) {
        acc[0] = 2.0 * (double)p.J * (double)fs[0];
#pragma unroll
        for (int i = 1; i < 5; ++i) acc[i] = (double)fs[i];
    }
    acc[5] = (double)accepted;
    if (MODE == 0) {
        double a1[1] = {acc[5]};
        block_atomic_add<double, 1>(a1, s_red, obs + 5);
    } else {
        block_atomic_add<double, 6>(acc, s_red, obs);
    }
}

// ---------------------------------------------------------------------------------------
// K4f: TWO consecutive colours (B0, B0 + 1) of a periodic fcc step in one launch, without any inter-CTA synchronisation.
// As four launches every colour pass reads all four sublattices and writes one (5 array sweeps per pass, 62 B/attempt measured
// against 24 algorithmic); a pair launch reads four and writes two (36 B/attempt for the step).  What makes it possible:
//   * in the fcc unit-cell table the bonds between colours 2k and 2k + 1 have dz = 0 and, seen from the second colour,
//     dy in {0, +1} (checked on the host before this kernel is chosen): the second colour on rows [r0, r1) of a plane needs
//     the first colour's NEW spins on rows [r0, r1] of the same plane only;
//   * a CTA owns rows [r0, r1) of one plane.  It updates the first colour on rows [r0, r1] -- row r1 belongs to the next
//     CTA, which computes the identical update (same inputs, site-keyed random numbers), so it is computed twice and
//     counted once -- then, after a CTA barrier, the second colour on its own rows;
//   * the step goes from one set of arrays to another (S = before, D = after; swapped by the host after every step), so
//     that a redundantly updated row never sees a neighbour that another CTA has already moved on: every read of an "old"
//     spin comes from S, which no launch of this step writes, every read of a "new" spin of a LOWER colour from D.
// Same work item as the colour launches (basis_vec_item): bit-identical trajectories.
// ---------------------------------------------------------------------------------------
// resident CTAs per SM the pair kernel is compiled for: fewer, fatter CTAs keep the window between a CTA's two colours inside
// L2 and more loads in flight per thread.  fcc 384^3, 32 rows per tile, colours alternating every 4 rows: 8 CTAs (64 registers)
// 2.69 ms, 4 (128) 2.48, 3 (170) 2.22, 2 (221) 2.31 (profiles/r02/p4.sh)
#ifndef BASIS_PAIR_MINB
#define BASIS_PAIR_MINB 3
#endif
// SLAB: planes -1 and nz of every array are halo planes the z-neighbours store into; `peers` are the bases of THEIR copy of the
// array set this launch writes (D), so that my boundary planes land in their halos; planes z_begin, z_begin + z_step, ...
template <typename real, int UC, int B0, bool FLIP, int MODE, bool SLAB = false>
__global__ void __launch_bounds__(128, BASIS_PAIR_MINB)
heis_basis_pair_kernel(BasisPtrs<real> S, BasisPtrs<real> D, BasisPeers<real> peers, BasisGeom g, uint32_t rows_per_cta, uint32_t chunk_rows,
                       uint32_t z_begin, uint32_t z_step, HeisParams<real> p, uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    constexpr int N = VecOf<real>::N, NB = BasisCell<UC>::NB;
    static_assert(B0 + 1 < NB, "a pair is (B0, B0 + 1)");
    __shared__ double s_red[6 * 32];
    const uint32_t iz = z_begin + blockIdx.z * z_step, VX = g.nx / N;
    const uint32_t r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, g.ny);
    const int zs[3] = {SLAB ? (int)iz - 1 : (int)(iz == 0 ? g.nz - 1 : iz - 1), (int)iz, SLAB ? (int)iz + 1 : (int)(iz + 1 == g.nz ? 0u : iz + 1)};
    // what colour B reads: lower colours are final (D), itself and higher colours are still the old state (S)
    BasisPtrs<real> R0, R1;
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) { R0.s[b][c] = b < B0 ? D.s[b][c] : S.s[b][c]; R1.s[b][c] = b <= B0 ? D.s[b][c] : S.s[b][c]; }
    real fs[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    // The two colours alternate in chunks of `chunk_rows` rows so that the second colour finds the partner rows the first one
    // has just read in L1 / L2 (a whole tile per colour puts 520 MB between the two uses: measured 14.6 GB per step, as much
    // as four launches).  First colour: row r0, then per chunk the rows (ra, rb] -- rb may be r1, the row after my tile
    // (periodic in y), recomputed here and counted by its owner.
    auto first = [&](uint32_t ra, uint32_t rb) {   // rows [ra, rb) in tile coordinates that may run one past r1
        const uint32_t n = (rb - ra) * VX;
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
            const uint32_t dr = t / VX, y = ra + dr, row = y == g.ny ? 0u : y;
            basis_vec_item<real, UC, B0, FLIP, MODE, SLAB>(R0, D, peers, g, iz, zs, row * VX + (t - dr * VX), VX, p, sweep, pk, fs, accepted, y < r1);
        }
    };
    first(r0, r0 + 1u);
    for (uint32_t ra = r0; ra < r1; ra += chunk_rows) {
        const uint32_t rb = min(ra + chunk_rows, r1);
        first(ra + 1u, rb + 1u);
        __syncthreads();   // the first colour's new spins on rows [ra, rb] were stored by threads of this CTA
        const uint32_t n2 = (rb - ra) * VX;
        for (uint32_t t = threadIdx.x; t < n2; t += blockDim.x)
            basis_vec_item<real, UC, B0 + 1, FLIP, MODE, SLAB>(R1, D, peers, g, iz, zs, ra * VX + t, VX, p, sweep, pk, fs, accepted);
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (MODE != 0) {
        acc[0] = 2.0 * (double)p.J * (double)fs[0];
#pragma unroll
        for (int i = 1; i < 5; ++i) acc[i] = (double)fs[i];
    }
    acc[5] = (double)accepted;
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
