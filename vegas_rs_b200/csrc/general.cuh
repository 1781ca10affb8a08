// general.cuh -- K2/K4/K5: colour-by-colour Metropolis sweep, reductions and per-site energies for
// arbitrary adjacency.  Spins are stored in natural site order (Ising: int8 +1/-1, Heisenberg: SoA).
//
// Two neighbour policies feed the same kernels:
//   StructuredNb  implicit unit-cell stencil for sc/bcc/fcc x expansion with per-axis pbc and the
//                 optional `source <= target` filter of Exchange::from_lattice (src/energy.rs:176-187);
//                 no index arrays are stored (24 B instead of 76 B per fcc attempt).
//   CsrNb         explicit CSR rows as Exchange::new(CsMat) holds them (src/energy.rs:171-173).
// Both enumerate exactly the entries of the reference's CSR row (duplicates appear as repeats,
// the 2J diagonal of a self edge as two (i,i) entries).
#pragma once
#include "common.cuh"
#include "heis.cuh"

namespace vg {

struct NbEntry { int8_t tb, dx, dy, dz, fwd, pad[3]; };

struct StructuredNb {
    uint32_t nx, ny, nz;
    int nb;                 // basis sites per cell
    int pbc[3];
    int literal;            // apply source <= target filter
    int count[4];           // entries per basis
    NbEntry e[4][12];
    double J;

    template <typename F>
    __device__ __forceinline__ void for_each(uint32_t i, F&& f) const {
        const uint32_t cell = i / (uint32_t)nb, b = i - cell * (uint32_t)nb;
        const uint32_t ix = cell % nx, iy = (cell / nx) % ny, iz = cell / (nx * ny);
        for (int q = 0; q < count[b]; ++q) {
            const NbEntry en = e[b][q];
            int tx = (int)ix + en.dx, ty = (int)iy + en.dy, tz = (int)iz + en.dz;
            if (tx < 0 || tx >= (int)nx) { if (!pbc[0]) continue; tx = tx < 0 ? tx + (int)nx : tx - (int)nx; }
            if (ty < 0 || ty >= (int)ny) { if (!pbc[1]) continue; ty = ty < 0 ? ty + (int)ny : ty - (int)ny; }
            if (tz < 0 || tz >= (int)nz) { if (!pbc[2]) continue; tz = tz < 0 ? tz + (int)nz : tz - (int)nz; }
            const uint32_t j = (((uint32_t)tz * ny + (uint32_t)ty) * nx + (uint32_t)tx) * (uint32_t)nb + (uint32_t)en.tb;
            if (literal && (en.fwd ? !(i <= j) : !(j <= i))) continue;
            f(j, J);
        }
    }
    // same enumeration, also handing out the neighbour's basis index (= its colour for bcc / fcc)
    template <typename F>
    __device__ __forceinline__ void for_each_b(uint32_t i, F&& f) const {
        const uint32_t cell = i / (uint32_t)nb, b = i - cell * (uint32_t)nb;
        const uint32_t ix = cell % nx, iy = (cell / nx) % ny, iz = cell / (nx * ny);
        for (int q = 0; q < count[b]; ++q) {
            const NbEntry en = e[b][q];
            int tx = (int)ix + en.dx, ty = (int)iy + en.dy, tz = (int)iz + en.dz;
            if (tx < 0 || tx >= (int)nx) { if (!pbc[0]) continue; tx = tx < 0 ? tx + (int)nx : tx - (int)nx; }
            if (ty < 0 || ty >= (int)ny) { if (!pbc[1]) continue; ty = ty < 0 ? ty + (int)ny : ty - (int)ny; }
            if (tz < 0 || tz >= (int)nz) { if (!pbc[2]) continue; tz = tz < 0 ? tz + (int)nz : tz - (int)nz; }
            const uint32_t j = (((uint32_t)tz * ny + (uint32_t)ty) * nx + (uint32_t)tx) * (uint32_t)nb + (uint32_t)en.tb;
            if (literal && (en.fwd ? !(i <= j) : !(j <= i))) continue;
            f(j, J, (int)en.tb);
        }
    }
};

struct CsrNb {
    const unsigned long long* row_ptr;
    const uint32_t* col;
    const double* val;  // null: uniform J
    double J;
    template <typename F>
    __device__ __forceinline__ void for_each(uint32_t i, F&& f) const {
        const unsigned long long a = row_ptr[i], b = row_ptr[i + 1];
        for (unsigned long long p = a; p < b; ++p) f(col[p], val ? val[p] : J);
    }
};

// ---------------------------------------------------------------------------------------
// Ising
// ---------------------------------------------------------------------------------------
constexpr int ISING_ZMAX = 32;  // table covers |sum_nb s_j| <= 32 (uniform J)

struct IsingGeneralParams {
    const unsigned long long* thr;  // [2][2*ZMAX+1] thresholds by (spin up?, m + ZMAX), m = sum_nb s_j
    const uint8_t* code;            // same shape: 0 never, 1 compare, 2 always
    int uniform;                    // 1: table; 0: evaluate exp(-dE/T) per site from CSR values
    double h_o;                     // |H| * (orientation . up) = +-|H|
    double invT;
};

// Uniform-J decision for a spin si whose neighbour sum m = sum_j s_j is known: threshold table lookup with the 64-bit
// uniform U (thr / code may live in global or shared memory).
__device__ __forceinline__ bool ising_table_decision(int si, int m, unsigned long long U, const unsigned long long* thr,
                                                     const uint8_t* code) {
    const int idx = (si > 0 ? 1 : 0) * (2 * ISING_ZMAX + 1) + (m + ISING_ZMAX);
    const uint8_t c = code[idx];
    return c == 2 || (c == 1 && U < thr[idx]);
}

// One attempt on site i (the body of src/integrator.rs:121-135 / :75-89 for IsingSpin); `s` may live in global or
// shared memory.  Returns true when the move counts as accepted.
template <typename NB, bool RANDPROP>
__device__ __forceinline__ bool ising_general_attempt(int8_t* s, const NB& nb, uint32_t i, const IsingGeneralParams& p,
                                                      uint64_t site_offset, uint64_t sweep, const PhiloxKey& pk) {
    const int si = s[i];
    uint32_t r[4];
    philox_at((uint64_t)i + site_offset, sweep, 0u, pk, r);
    const unsigned long long U = ((unsigned long long)r[0] << 32) | r[1];
    bool proposed = true;
    if (RANDPROP) proposed = ((r[2] & 1u) ? 1 : -1) != si;  // IsingSpin::rand src/state.rs:76-84
    bool ok;
    if (p.uniform) {
        int m = 0;
        nb.for_each(i, [&](uint32_t j, double) { if (j != i) m += s[j]; });
        ok = ising_table_decision(si, m, U, p.thr, p.code);
    } else {
        double ex = 0.0;  // Exchange::energy fold src/energy.rs:197-201 without the constant diagonal
        nb.for_each(i, [&](uint32_t j, double Jij) { if (j != i) ex = ex + (-Jij * (double)(si * s[j])); });
        const double dE = -2.0 * (ex + p.h_o * (double)si);
        const double pr = exp(-dE * p.invT);
        ok = !(pr < 1.0) || U < __double2ull_rd(pr * 18446744073709551616.0);
    }
    if (!proposed) ok = true;
    if (ok && proposed) s[i] = (int8_t)-si;
    return ok;
}

template <typename NB, bool RANDPROP>
__global__ void __launch_bounds__(256)
ising_general_sweep_kernel(int8_t* __restrict__ s, NB nb, const uint32_t* __restrict__ sites, uint32_t count,
                           IsingGeneralParams p, uint64_t site_offset, uint64_t sweep, PhiloxKey pk,
                           unsigned long long* __restrict__ obs) {
    __shared__ unsigned long long s_red[32];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc[1] = {0ull};
    if (t < count) acc[0] = ising_general_attempt<NB, RANDPROP>(s, nb, sites[t], p, site_offset, sweep, pk) ? 1ull : 0ull;
    block_atomic_add<unsigned long long, 1>(acc, s_red, obs + 2);
}

// ---------------------------------------------------------------------------------------
// Heisenberg
// ---------------------------------------------------------------------------------------
// The Metropolis decision for site i once its neighbour sum n = sum_j J_ij s_j is known (shared by every gather
// policy); the SoA arrays may live in global or shared memory.
template <typename real, bool FLIP>
__device__ __forceinline__ bool heis_site_update(real* sx, real* sy, real* sz, uint32_t i, real nx, real ny, real nz,
                                                 const HeisParams<real>& p, uint64_t site_offset, uint64_t sweep,
                                                 const PhiloxKey& pk) {
    real x = sx[i], y = sy[i], z = sz[i];
    HeisRand<real> rnd;
    heis_rand((uint64_t)i + site_offset, sweep, pk, rnd);
    const bool ok = heis_attempt<real, FLIP>(x, y, z, nx - p.h[0], ny - p.h[1], nz - p.h[2], p, rnd);
    if (ok) { sx[i] = x; sy[i] = y; sz[i] = z; }
    return ok;
}

// One attempt on site i for HeisenbergSpin, neighbours enumerated by the policy NB.
template <typename NB, typename real, bool FLIP>
__device__ __forceinline__ bool heis_general_attempt(real* sx, real* sy, real* sz, const NB& nb, uint32_t i,
                                                     const HeisParams<real>& p, uint64_t site_offset, uint64_t sweep,
                                                     const PhiloxKey& pk) {
    real nx = 0, ny = 0, nz = 0;
    nb.for_each(i, [&](uint32_t j, double Jij) {
        if (j != i) { const real w = (real)Jij; nx += w * sx[j]; ny += w * sy[j]; nz += w * sz[j]; }
    });
    return heis_site_update<real, FLIP>(sx, sy, sz, i, nx, ny, nz, p, site_offset, sweep, pk);
}

template <typename NB, typename real, bool FLIP>
__global__ void __launch_bounds__(128)
heis_general_sweep_kernel(real* __restrict__ sx, real* __restrict__ sy, real* __restrict__ sz, NB nb,
                          const uint32_t* __restrict__ sites, uint32_t count, HeisParams<real> p, uint64_t site_offset,
                          uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    __shared__ double s_red[32];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0.0};
    if (t < count) acc[0] = heis_general_attempt<NB, real, FLIP>(sx, sy, sz, nb, sites[t], p, site_offset, sweep, pk) ? 1.0 : 0.0;
    block_atomic_add<double, 1>(acc, s_red, obs + 5);
}

// Recorded step on a bcc / fcc lattice (colour = basis index): the colour pass also reduces its own colour --
// obs[1..3] += s, obs[4] += (s.a)^2 -- and the exchange bonds towards the LOWER colours, which are final by now, so
// that every bond is counted once over the step; obs[0] receives twice that (the layout general_reduce_kernel fills:
// sum_i sum_j J_ij s_i.s_j with every bond twice).  No separate reduction launch.
template <typename real, bool FLIP>
__global__ void __launch_bounds__(128)
heis_basis_sweep_obs_kernel(real* __restrict__ sx, real* __restrict__ sy, real* __restrict__ sz, StructuredNb nb,
                            const uint32_t* __restrict__ sites, uint32_t count, int colour, HeisParams<real> p,
                            uint64_t site_offset, uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    __shared__ double s_red[6 * 32];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (t < count) {
        const uint32_t i = sites[t];
        real nx = 0, ny = 0, nz = 0, lx = 0, ly = 0, lz = 0;
        nb.for_each_b(i, [&](uint32_t j, double Jij, int tb) {
            if (j != i) {
                const real w = (real)Jij, u = w * sx[j], v = w * sy[j], q = w * sz[j];
                nx += u; ny += v; nz += q;
                if (tb < colour) { lx += u; ly += v; lz += q; }
            }
        });
        real x = sx[i], y = sy[i], z = sz[i];
        HeisRand<real> rnd;
        heis_rand((uint64_t)i + site_offset, sweep, pk, rnd);
        const bool ok = heis_attempt<real, FLIP>(x, y, z, nx - p.h[0], ny - p.h[1], nz - p.h[2], p, rnd);
        if (ok) { sx[i] = x; sy[i] = y; sz[i] = z; }
        acc[0] = 2.0 * ((double)x * lx + (double)y * ly + (double)z * lz);
        acc[1] = x; acc[2] = y; acc[3] = z;
        const double d = (double)x * p.a[0] + (double)y * p.a[1] + (double)z * p.a[2];
        acc[4] = d * d;
        acc[5] = ok ? 1.0 : 0.0;
    }
    block_atomic_add<double, 6>(acc, s_red, obs);
}

// ---------------------------------------------------------------------------------------
// K5 reductions over all sites: obs[0] += sum_i sum_j J_ij s_i.s_j (every bond twice, diagonal as stored),
// obs[1..3] += sum s, obs[4] += sum (s.a)^2.  Spin access through a functor so that Ising and
// Heisenberg share the code.
// ---------------------------------------------------------------------------------------
struct IsingSpins {
    const int8_t* s;
    __device__ __forceinline__ void get(uint32_t i, double& x, double& y, double& z) const { x = 0; y = 0; z = (double)s[i]; }
};
template <typename real>
struct HeisSpins {
    const real* sx; const real* sy; const real* sz;
    __device__ __forceinline__ void get(uint32_t i, double& x, double& y, double& z) const {
        x = (double)sx[i]; y = (double)sy[i]; z = (double)sz[i];
    }
};

struct EnergyParams {
    double h[3];       // |H| * orientation (Ising: (0,0,+-|H|))
    double k, a[3];    // anisotropy (Ising: a = (0,0,+-1))
    double gauge;
    int has_exchange, has_zeeman, has_aniso, has_gauge;
};

// Site i's share of the K5 sums (acc[0..4] as general_reduce_kernel documents them).
template <typename NB, typename SP>
__device__ __forceinline__ void general_site_terms(const NB& nb, const SP& sp, uint32_t i, double ax, double ay, double az,
                                                   double (&acc)[5]) {
    double x, y, z;
    sp.get(i, x, y, z);
    double e = 0.0;
    nb.for_each(i, [&](uint32_t j, double Jij) {
        double u, v, w;
        sp.get(j, u, v, w);
        e += Jij * (x * u + y * v + z * w);
    });
    acc[0] += e; acc[1] += x; acc[2] += y; acc[3] += z;
    const double d = x * ax + y * ay + z * az;
    acc[4] += d * d;
}

template <typename NB, typename SP>
__global__ void __launch_bounds__(256)
general_reduce_kernel(NB nb, SP sp, uint32_t n, double ax, double ay, double az, double* __restrict__ obs) {
    __shared__ double s_red[5 * 32];
    double acc[5] = {0, 0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        general_site_terms(nb, sp, i, ax, ay, az, acc);
    block_atomic_add<double, 5>(acc, s_red, obs);
}

// Hamiltonian::energy(i) of the compound (src/energy.rs:254-256) for every site, reference signs:
// Exchange fold of -J_ij s_i.s_j (:197-201), Zeeman +|H| s.o (:147-151), anisotropy k (s.a)^2 (:108-112),
// gauge (:75-79).  If prop != null also e_new - e_old for the proposed spin (src/integrator.rs:77-81).
template <typename NB, typename SP>
__global__ void __launch_bounds__(256)
site_energy_kernel(NB nb, SP sp, uint32_t n, EnergyParams ep, const double* __restrict__ prop /* AoS [3n] or null */,
                   int flip, double* __restrict__ out_e, double* __restrict__ out_de) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    sp.get(i, x, y, z);
    auto energy = [&](double px, double py, double pz) {
        double e = 0.0;
        bool first = true;
        auto add = [&](double v) { e = first ? v : e + v; first = false; };
        if (ep.has_exchange) {
            double ex = 0.0;
            nb.for_each(i, [&](uint32_t j, double Jij) {
                double u, v, w;
                if (j == i) { u = px; v = py; w = pz; } else sp.get(j, u, v, w);
                ex = ex + (-Jij * (((0.0 + px * u) + py * v) + pz * w));
            });
            add(ex);
        }
        if (ep.has_zeeman) add(((0.0 + px * ep.h[0]) + py * ep.h[1]) + pz * ep.h[2]);
        if (ep.has_aniso) { const double d = ((0.0 + px * ep.a[0]) + py * ep.a[1]) + pz * ep.a[2]; add(d * d * ep.k); }
        if (ep.has_gauge) add(ep.gauge);
        return e;
    };
    const double e_old = energy(x, y, z);
    if (out_e) out_e[i] = e_old;
    if (out_de) {
        double px, py, pz;
        if (flip) { px = -x; py = -y; pz = -z; }
        else { px = prop[3 * (size_t)i]; py = prop[3 * (size_t)i + 1]; pz = prop[3 * (size_t)i + 2]; }
        out_de[i] = energy(px, py, pz) - e_old;
    }
}

// natural-order helpers ------------------------------------------------------------------
__global__ void __launch_bounds__(256) ising_general_randomize_kernel(int8_t* s, uint32_t n, uint64_t site_offset,
                                                                      PhiloxKey pk) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[4];
    philox_at(((uint64_t)i + site_offset) | (1ull << 61), ~0ull, 0xFEu, pk, r);
    s[i] = (r[0] & 1u) ? 1 : -1;
}

template <typename real>
__global__ void __launch_bounds__(256) heis_general_randomize_kernel(real* sx, real* sy, real* sz, uint32_t n,
                                                                     uint64_t site_offset, PhiloxKey pk) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    real x, y, z;
    heis_random_spin<real>((uint64_t)i + site_offset, pk, x, y, z);
    sx[i] = x; sy[i] = y; sz[i] = z;
}

template <typename real>
__global__ void __launch_bounds__(256) aos_to_soa_kernel(const double* __restrict__ aos, real* sx, real* sy, real* sz,
                                                         uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sx[i] = (real)aos[3 * (size_t)i]; sy[i] = (real)aos[3 * (size_t)i + 1]; sz[i] = (real)aos[3 * (size_t)i + 2];
}

template <typename real>
__global__ void __launch_bounds__(256) soa_to_aos_kernel(double* __restrict__ aos, const real* sx, const real* sy,
                                                         const real* sz, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    aos[3 * (size_t)i] = (double)sx[i]; aos[3 * (size_t)i + 1] = (double)sy[i]; aos[3 * (size_t)i + 2] = (double)sz[i];
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* p, size_t n, T v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace vg
