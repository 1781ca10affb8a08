// resident.cuh -- K2r: shared-memory-resident Metropolis steps for SMALL lattices of the general family.
//
// The reference's own workloads are small (docs/metropolis.toml: 10x10x10 = 1000 sites, 2.1e6 steps): on those a step
// of colour-pass launches is pure launch latency (three launches and a memset, ~16 us for 1000 attempts).  Here ONE CTA
// keeps the whole State in shared memory and runs a batch of up to 4096 Monte Carlo steps (Integrator::step,
// src/integrator.rs:66-138) in one launch: colours in ascending order with a CTA barrier between them, the per-step
// observers of src/instrument.rs:133-141,254-262 reduced in the CTA and stored to the step's observable row.
//
// TABLE variant (uniform J): at launch every site's neighbour list -- the entries the reference's CSR row holds, in row
// order, self entries counted apart -- is written ONCE into a shared-memory table (u16 indices, q-major so that the
// threads of a warp read consecutive elements; missing neighbours of open boundaries point at a zero spin stored at
// index n), together with the colour-ordered site list and the acceptance thresholds.  A step then touches shared
// memory only: no index arithmetic, no integer divisions, no global loads.
// Direct variant (CSR with per-bond values): the neighbour policy is enumerated every attempt, as the colour-pass
// kernels do.
//
// Either way the attempts use the device functions of the colour-pass kernels (general.cuh) with the same Philox
// counters (site, sweep) and the same order of floating-point operations, so the trajectory is bit-identical to the
// launch-per-colour path (tests/test_gpu_resident.py).
#pragma once
#include "general.cuh"

namespace vg {

constexpr int RES_MAX_COLOURS = 8;
constexpr int RES_MAX_Z = 16;    // widest neighbour row the table variant holds
constexpr int RES_RED = 6 * 32 + 2;  // doubles of reduction scratch at the start of the dynamic shared memory (the first two: integer sums)
constexpr int RES_THR = 2 * (2 * ISING_ZMAX + 1);  // threshold table entries

struct ResidentPlan {
    const uint32_t* sites[RES_MAX_COLOURS];  // per-colour site lists (device), as the colour-pass kernels read them
    uint32_t counts[RES_MAX_COLOURS];
    int n_colours;
    uint32_t n;
    int zmax;  // table columns (0: direct variant)
};

__host__ __device__ inline size_t res_align16(size_t v) { return (v + 15) & ~(size_t)15; }
// dynamic shared memory of a launch: reduction scratch | thresholds | spins (n + 1 per component) | table | self counts | order
inline size_t resident_smem_bytes(uint32_t n, int zmax, size_t spin_bytes_per_site) {
    size_t b = RES_RED * sizeof(double) + res_align16(RES_THR * 8) + res_align16(RES_THR);
    b += res_align16(spin_bytes_per_site * (n + 1));
    if (zmax > 0) b += res_align16((size_t)2 * zmax * n) + res_align16(n) + res_align16((size_t)2 * n);
    return b;
}

// Block sum of NV doubles per thread; the totals are valid in thread 0 on return.  Every thread must call it.
template <int NV>
__device__ __forceinline__ void resident_block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) red[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(lane < nwarp ? red[i * 32 + lane] : 0.0);
    }
    __syncthreads();  // red may be rewritten by the next step
}

// Shared-memory carve-up common to both models (pointers into the dynamic shared memory).
struct ResidentSmem {
    double* red;
    unsigned long long* thr;
    uint8_t* code;
    unsigned char* spins;
    uint16_t* tab;     // [zmax][n] neighbour indices (n = the zero spin)
    uint8_t* nself;    // [n] self entries of the row (periodic axis of extent 1)
    uint16_t* order;   // [n] sites in colour order
};

__device__ __forceinline__ ResidentSmem resident_carve(unsigned char* base, uint32_t n, int zmax, size_t spin_bytes_per_site) {
    ResidentSmem m;
    m.red = reinterpret_cast<double*>(base); base += RES_RED * sizeof(double);
    m.thr = reinterpret_cast<unsigned long long*>(base); base += res_align16(RES_THR * 8);
    m.code = base; base += res_align16(RES_THR);
    m.spins = base; base += res_align16(spin_bytes_per_site * (n + 1));
    m.tab = reinterpret_cast<uint16_t*>(base); base += res_align16((size_t)2 * zmax * n);
    m.nself = base; base += res_align16(n);
    m.order = reinterpret_cast<uint16_t*>(base);
    return m;
}

// Neighbour table and colour-ordered site list, built once per launch (every thread of the CTA calls this).
template <typename NB>
__device__ __forceinline__ void resident_build(const NB& nb, const ResidentPlan& rp, const ResidentSmem& m) {
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) {
        int q = 0, ns = 0;
        nb.for_each(i, [&](uint32_t j, double) {
            if (j == i) ++ns;
            else if (q < rp.zmax) m.tab[(uint32_t)(q++) * rp.n + i] = (uint16_t)j;
        });
        for (; q < rp.zmax; ++q) m.tab[(uint32_t)q * rp.n + i] = (uint16_t)rp.n;
        m.nself[i] = (uint8_t)ns;
    }
    uint32_t base = 0;
    for (int c = 0; c < rp.n_colours; ++c) {
        const uint32_t* __restrict__ sites = rp.sites[c];
        for (uint32_t t = threadIdx.x; t < rp.counts[c]; t += blockDim.x) m.order[base + t] = (uint16_t)sites[t];
        base += rp.counts[c];
    }
}

// Neighbour sums over a table row.  The common row widths are compile-time trip counts so that all the index loads
// issue back to back (a run-time loop serialises the two dependent shared-memory loads of every neighbour).
template <int Z>
__device__ __forceinline__ int res_msum_fixed(const int8_t* s, const uint16_t* tab, uint32_t n, uint32_t i, int zmax) {
    int m = 0;
#pragma unroll
    for (int q = 0; q < (Z > 0 ? Z : zmax); ++q) m += s[tab[(uint32_t)q * n + i]];
    return m;
}
__device__ __forceinline__ int res_msum(const int8_t* s, const uint16_t* tab, uint32_t n, uint32_t i, int zmax) {
    switch (zmax) {
        case 4: return res_msum_fixed<4>(s, tab, n, i, zmax);
        case 6: return res_msum_fixed<6>(s, tab, n, i, zmax);
        case 8: return res_msum_fixed<8>(s, tab, n, i, zmax);
        case 12: return res_msum_fixed<12>(s, tab, n, i, zmax);
        default: return res_msum_fixed<0>(s, tab, n, i, zmax);
    }
}
// the fold of heis_general_attempt (same order, a missing neighbour adds w * 0)
template <int Z, typename real>
__device__ __forceinline__ void res_nsum_fixed(const real* sx, const real* sy, const real* sz, const uint16_t* tab, uint32_t n,
                                               uint32_t i, int zmax, real w, real& nx, real& ny, real& nz) {
    nx = 0; ny = 0; nz = 0;
#pragma unroll
    for (int q = 0; q < (Z > 0 ? Z : zmax); ++q) {
        const uint32_t j = tab[(uint32_t)q * n + i];
        nx += w * sx[j]; ny += w * sy[j]; nz += w * sz[j];
    }
}
template <typename real>
__device__ __forceinline__ void res_nsum(const real* sx, const real* sy, const real* sz, const uint16_t* tab, uint32_t n,
                                         uint32_t i, int zmax, real w, real& nx, real& ny, real& nz) {
    switch (zmax) {
        case 4: res_nsum_fixed<4>(sx, sy, sz, tab, n, i, zmax, w, nx, ny, nz); break;
        case 6: res_nsum_fixed<6>(sx, sy, sz, tab, n, i, zmax, w, nx, ny, nz); break;
        case 8: res_nsum_fixed<8>(sx, sy, sz, tab, n, i, zmax, w, nx, ny, nz); break;
        case 12: res_nsum_fixed<12>(sx, sy, sz, tab, n, i, zmax, w, nx, ny, nz); break;
        default: res_nsum_fixed<0>(sx, sy, sz, tab, n, i, zmax, w, nx, ny, nz); break;
    }
}

// Block sum of three ints with REDUX + shared atomics; totals valid in thread 0 on return (every thread calls it).
// `acc` = three zeroed ints in shared memory; thread 0 zeroes them again before it returns.
__device__ __forceinline__ void resident_block_sum_int3(int (&v)[3], int* acc) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        v[i] = __reduce_add_sync(0xffffffffu, v[i]);
        if ((threadIdx.x & 31u) == 0 && v[i] != 0) atomicAdd(acc + i, v[i]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { v[i] = acc[i]; acc[i] = 0; }
    }
}

// One colour phase of the Ising table variant; Z = compile-time row width (0: run-time zmax).  The width is chosen once
// per phase (uniform switch in the kernel), not once per attempt.
template <int Z, bool RANDPROP, bool TRACK>
__device__ __forceinline__ void ising_table_phase(int8_t* s, const ResidentSmem& m, uint32_t n, int zmax, uint32_t base,
                                                  uint32_t count, uint64_t sweep, const PhiloxKey& pk, int& accepted, int& dS,
                                                  int& dM) {
    for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
        const uint32_t i = m.order[base + t];
        const int si = s[i];
        uint32_t r[4];
        philox_at((uint64_t)i, sweep, 0u, pk, r);
        const unsigned long long U = ((unsigned long long)r[0] << 32) | r[1];
        bool proposed = true;
        if (RANDPROP) proposed = ((r[2] & 1u) ? 1 : -1) != si;  // IsingSpin::rand src/state.rs:76-84
        const int msum = res_msum_fixed<Z>(s, m.tab, n, i, zmax);
        bool ok = ising_table_decision(si, msum, U, m.thr, m.code);
        if (!proposed) ok = true;
        if (ok && proposed) {
            s[i] = (int8_t)-si;
            if (TRACK) { dS -= 4 * si * msum; dM -= 2 * si; }
        }
        accepted += ok ? 1 : 0;
    }
}

// SYMM: the adjacency is symmetric (every structured lattice: Exchange::from_lattice adds (s,t) and (t,s),
// src/energy.rs:181-184), so a recorded batch reduces sum_i sum_j s_i s_j and sum s ONCE at launch and then follows
// them exactly, in integers, through the accepted flips: flipping s_i changes the double sum by -4 s_i m_i.
template <typename NB, bool RANDPROP, bool TABLE, bool SYMM>
__global__ void __launch_bounds__(1024, 1)
ising_resident_kernel(int8_t* __restrict__ s_glob, NB nb, ResidentPlan rp, IsingGeneralParams p, double J, uint64_t sweep0,
                      uint32_t n_steps, PhiloxKey pk, unsigned long long* __restrict__ obs /* rows or null */,
                      int obs_w, unsigned long long* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    const ResidentSmem m = resident_carve(res_smem, rp.n, TABLE ? rp.zmax : 0, 1);
    int8_t* s = reinterpret_cast<int8_t*>(m.spins);
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) s[i] = s_glob[i];
    if (threadIdx.x == 0) s[rp.n] = 0;  // the zero spin missing neighbours point at
    if (TABLE) {
        resident_build(nb, rp, m);
        for (uint32_t i = threadIdx.x; i < RES_THR; i += blockDim.x) { m.thr[i] = p.thr[i]; m.code[i] = p.code[i]; }
    }
    constexpr bool TRACK = TABLE && SYMM;
    int* iacc = reinterpret_cast<int*>(m.red);  // three ints of the integer block sums (the doubles use m.red + 2 up)
    if (threadIdx.x < 3) iacc[threadIdx.x] = 0;
    __syncthreads();
    const IsingSpins sp{s};
    long long S = 0, M = 0;  // thread 0: sum_i (s_i m_i + self entries), sum_i s_i of the current state
    if (TRACK && obs != nullptr) {
        int v[3] = {0, 0, 0};
        for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) {
            const int si = s[i];
            v[0] += si * res_msum(s, m.tab, rp.n, i, rp.zmax) + (int)m.nself[i];
            v[1] += si;
        }
        resident_block_sum_int3(v, iacc);
        S = v[0]; M = v[1];
        __syncthreads();
    }
    unsigned long long unrecorded = 0;
    for (uint32_t step = 0; step < n_steps; ++step) {
        int accepted = 0, dS = 0, dM = 0;
        uint32_t base = 0;
        for (int c = 0; c < rp.n_colours; ++c) {
            const uint32_t count = rp.counts[c];
            if (TABLE) {
#define VG_PHASE(Z) ising_table_phase<Z, RANDPROP, TRACK>(s, m, rp.n, rp.zmax, base, count, sweep0 + step, pk, accepted, dS, dM)
                switch (rp.zmax) {
                    case 4: VG_PHASE(4); break;
                    case 6: VG_PHASE(6); break;
                    case 8: VG_PHASE(8); break;
                    case 12: VG_PHASE(12); break;
                    default: VG_PHASE(0); break;
                }
#undef VG_PHASE
            } else {
                const uint32_t* __restrict__ sites = rp.sites[c];
                for (uint32_t t = threadIdx.x; t < count; t += blockDim.x)
                    accepted += ising_general_attempt<NB, RANDPROP>(s, nb, sites[t], p, 0, sweep0 + step, pk) ? 1 : 0;
            }
            base += count;
            __syncthreads();
        }
        if (obs != nullptr && TRACK) {
            int v[3] = {dS, dM, accepted};
            resident_block_sum_int3(v, iacc);
            if (threadIdx.x == 0) {  // row layout of general_reduce_kernel + the accepted counter of the sweep kernel
                S += v[0]; M += v[1];
                unsigned long long* row = obs + (size_t)step * obs_w;
                double* d = reinterpret_cast<double*>(row);
                d[0] = J * (double)S; d[1] = 0.0; d[2] = 0.0; d[3] = (double)M; d[4] = (double)rp.n;
                row[6] = (unsigned long long)v[2];
            }
        } else if (obs != nullptr) {
            double acc[5] = {0, 0, 0, 0, 0};
            if (TABLE) {  // the sums of general_site_terms: sum_j J s_i s_j over the row (self entries: J each), sum s
                for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) {
                    const int si = s[i];
                    acc[0] += J * (double)(si * res_msum(s, m.tab, rp.n, i, rp.zmax) + (int)m.nself[i]);
                    acc[3] += (double)si;
                }
            } else {
                for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) general_site_terms(nb, sp, i, 0.0, 0.0, 1.0, acc);
            }
            double v[3] = {acc[0], acc[3], (double)accepted};
            resident_block_sum<3>(v, m.red + 2);
            if (threadIdx.x == 0) {
                unsigned long long* row = obs + (size_t)step * obs_w;
                double* d = reinterpret_cast<double*>(row);
                d[0] = v[0]; d[1] = 0.0; d[2] = 0.0; d[3] = v[1]; d[4] = (double)rp.n;
                row[6] = (unsigned long long)v[2];
            }
        } else {
            unrecorded += (unsigned long long)accepted;
        }
    }
    if (obs == nullptr) {
        double v[1] = {(double)unrecorded};
        resident_block_sum<1>(v, m.red + 2);
        if (threadIdx.x == 0) scratch[6] += (unsigned long long)v[0];
    }
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) s_glob[i] = s[i];
}

template <typename NB, typename real, bool FLIP, bool TABLE>
__global__ void __launch_bounds__(1024, 1)
heis_resident_kernel(real* __restrict__ gx, real* __restrict__ gy, real* __restrict__ gz, NB nb, ResidentPlan rp,
                     HeisParams<real> p, double J, double ax, double ay, double az, uint64_t sweep0, uint32_t n_steps,
                     PhiloxKey pk, double* __restrict__ obs /* rows or null */, int obs_w, double* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    const ResidentSmem m = resident_carve(res_smem, rp.n, TABLE ? rp.zmax : 0, 3 * sizeof(real));
    real* sx = reinterpret_cast<real*>(m.spins);
    real* sy = sx + (rp.n + 1);
    real* sz = sy + (rp.n + 1);
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) { sx[i] = gx[i]; sy[i] = gy[i]; sz[i] = gz[i]; }
    if (threadIdx.x == 0) { sx[rp.n] = 0; sy[rp.n] = 0; sz[rp.n] = 0; }
    if (TABLE) resident_build(nb, rp, m);
    __syncthreads();
    const HeisSpins<real> sp{sx, sy, sz};
    const real w = (real)J;
    double unrecorded = 0.0;
    for (uint32_t step = 0; step < n_steps; ++step) {
        int accepted = 0;
        uint32_t base = 0;
        for (int c = 0; c < rp.n_colours; ++c) {
            const uint32_t count = rp.counts[c];
            if (TABLE) {
                for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
                    const uint32_t i = m.order[base + t];
                    real nx, ny, nz;
                    res_nsum<real>(sx, sy, sz, m.tab, rp.n, i, rp.zmax, w, nx, ny, nz);
                    accepted += heis_site_update<real, FLIP>(sx, sy, sz, i, nx, ny, nz, p, 0, sweep0 + step, pk) ? 1 : 0;
                }
            } else {
                const uint32_t* __restrict__ sites = rp.sites[c];
                for (uint32_t t = threadIdx.x; t < count; t += blockDim.x)
                    accepted += heis_general_attempt<NB, real, FLIP>(sx, sy, sz, nb, sites[t], p, 0, sweep0 + step, pk) ? 1 : 0;
            }
            base += count;
            __syncthreads();
        }
        if (obs != nullptr) {
            double acc[5] = {0, 0, 0, 0, 0};
            if (TABLE) {
                for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) {
                    const double x = (double)sx[i], y = (double)sy[i], z = (double)sz[i];
                    double e = 0.0;
                    for (int q = 0; q < rp.zmax; ++q) {
                        const uint32_t j = m.tab[(uint32_t)q * rp.n + i];
                        e += J * (x * (double)sx[j] + y * (double)sy[j] + z * (double)sz[j]);
                    }
                    e += (double)m.nself[i] * (J * (x * x + y * y + z * z));
                    acc[0] += e; acc[1] += x; acc[2] += y; acc[3] += z;
                    const double d = x * ax + y * ay + z * az;
                    acc[4] += d * d;
                }
            } else {
                for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) general_site_terms(nb, sp, i, ax, ay, az, acc);
            }
            double v[6] = {acc[0], acc[1], acc[2], acc[3], acc[4], (double)accepted};
            resident_block_sum<6>(v, m.red);
            if (threadIdx.x == 0) {
                double* d = obs + (size_t)step * obs_w;
#pragma unroll
                for (int i = 0; i < 6; ++i) d[i] = v[i];
            }
        } else {
            unrecorded += (double)accepted;
        }
    }
    if (obs == nullptr) {
        double v[1] = {unrecorded};
        resident_block_sum<1>(v, m.red);
        if (threadIdx.x == 0) scratch[5] += v[0];
    }
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) { gx[i] = sx[i]; gy[i] = sy[i]; gz[i] = sz[i]; }
}

}  // namespace vg
