// resident.cuh -- K2r: shared-memory-resident Metropolis steps for SMALL lattices of the general family.
//
// The reference's own workloads are small (docs/metropolis.toml: 10x10x10 = 1000 sites, 2.1e6 steps): on those a step
// of colour-pass launches is pure launch latency (3 launches ~ 8 us for 1000 attempts).  Here ONE CTA keeps the whole
// State in shared memory and runs a batch of up to 4096 Monte Carlo steps (Integrator::step, src/integrator.rs:66-138)
// in one launch: colours in ascending order with a CTA barrier between them, the per-step observers of
// src/instrument.rs:133-141,254-262 reduced in the CTA and stored to the step's observable row.  The attempts are the
// very device functions of the colour-pass kernels (general.cuh) with the same Philox counters (site, sweep), so the
// trajectory is bit-identical to the launch-per-colour path (tests/test_gpu_parity.py::test_resident_*).
#pragma once
#include "general.cuh"

namespace vg {

constexpr int RES_MAX_COLOURS = 8;
constexpr int RES_RED = 6 * 32;  // doubles of reduction scratch at the start of the dynamic shared memory

struct ResidentPlan {
    const uint32_t* sites[RES_MAX_COLOURS];  // per-colour site lists (device), as the colour-pass kernels read them
    uint32_t counts[RES_MAX_COLOURS];
    int n_colours;
    uint32_t n;
};

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

template <typename NB, bool RANDPROP>
__global__ void __launch_bounds__(1024, 1)
ising_resident_kernel(int8_t* __restrict__ s_glob, NB nb, ResidentPlan rp, IsingGeneralParams p, uint64_t sweep0,
                      uint32_t n_steps, PhiloxKey pk, unsigned long long* __restrict__ obs /* rows or null */,
                      int obs_w, unsigned long long* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    double* red = reinterpret_cast<double*>(res_smem);
    int8_t* s = reinterpret_cast<int8_t*>(res_smem + RES_RED * sizeof(double));
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) s[i] = s_glob[i];
    __syncthreads();
    const IsingSpins sp{s};
    unsigned long long unrecorded = 0;
    for (uint32_t step = 0; step < n_steps; ++step) {
        int accepted = 0;
        for (int c = 0; c < rp.n_colours; ++c) {
            const uint32_t* __restrict__ sites = rp.sites[c];
            const uint32_t count = rp.counts[c];
            for (uint32_t t = threadIdx.x; t < count; t += blockDim.x)
                accepted += ising_general_attempt<NB, RANDPROP>(s, nb, sites[t], p, 0, sweep0 + step, pk) ? 1 : 0;
            __syncthreads();
        }
        if (obs != nullptr) {
            double acc[5] = {0, 0, 0, 0, 0};
            for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) general_site_terms(nb, sp, i, 0.0, 0.0, 1.0, acc);
            double v[4] = {acc[0], acc[3], acc[4], (double)accepted};
            resident_block_sum<4>(v, red);
            if (threadIdx.x == 0) {  // row layout of general_reduce_kernel + the accepted counter of the sweep kernel
                unsigned long long* row = obs + (size_t)step * obs_w;
                double* d = reinterpret_cast<double*>(row);
                d[0] = v[0]; d[1] = 0.0; d[2] = 0.0; d[3] = v[1]; d[4] = v[2];
                row[6] = (unsigned long long)v[3];
            }
        } else {
            unrecorded += (unsigned long long)accepted;
        }
    }
    if (obs == nullptr) {
        double v[1] = {(double)unrecorded};
        resident_block_sum<1>(v, red);
        if (threadIdx.x == 0) scratch[6] += (unsigned long long)v[0];
    }
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) s_glob[i] = s[i];
}

template <typename NB, typename real, bool FLIP>
__global__ void __launch_bounds__(1024, 1)
heis_resident_kernel(real* __restrict__ gx, real* __restrict__ gy, real* __restrict__ gz, NB nb, ResidentPlan rp,
                     HeisParams<real> p, double ax, double ay, double az, uint64_t sweep0, uint32_t n_steps, PhiloxKey pk,
                     double* __restrict__ obs /* rows or null */, int obs_w, double* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    double* red = reinterpret_cast<double*>(res_smem);
    real* sx = reinterpret_cast<real*>(res_smem + RES_RED * sizeof(double));
    real* sy = sx + rp.n;
    real* sz = sy + rp.n;
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) { sx[i] = gx[i]; sy[i] = gy[i]; sz[i] = gz[i]; }
    __syncthreads();
    const HeisSpins<real> sp{sx, sy, sz};
    double unrecorded = 0.0;
    for (uint32_t step = 0; step < n_steps; ++step) {
        int accepted = 0;
        for (int c = 0; c < rp.n_colours; ++c) {
            const uint32_t* __restrict__ sites = rp.sites[c];
            const uint32_t count = rp.counts[c];
            for (uint32_t t = threadIdx.x; t < count; t += blockDim.x)
                accepted += heis_general_attempt<NB, real, FLIP>(sx, sy, sz, nb, sites[t], p, 0, sweep0 + step, pk) ? 1 : 0;
            __syncthreads();
        }
        if (obs != nullptr) {
            double acc[5] = {0, 0, 0, 0, 0};
            for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) general_site_terms(nb, sp, i, ax, ay, az, acc);
            double v[6] = {acc[0], acc[1], acc[2], acc[3], acc[4], (double)accepted};
            resident_block_sum<6>(v, red);
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
        resident_block_sum<1>(v, red);
        if (threadIdx.x == 0) scratch[5] += v[0];
    }
    for (uint32_t i = threadIdx.x; i < rp.n; i += blockDim.x) { gx[i] = sx[i]; gy[i] = sy[i]; gz[i] = sz[i]; }
}

}  // namespace vg
