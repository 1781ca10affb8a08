// common.cuh -- device helpers shared by every vegas_gpu kernel (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vg {

// ---------------------------------------------------------------------------------------
// Philox4x32-7 (Random123), counter-based, all state in registers.  The key is the run
// seed; the counter is (site-or-word index, sweep, call) so that any decomposition of the
// lattice over GPUs reproduces the same random numbers per site.
// Round keys are thread-invariant, so the compiler keeps them on the uniform datapath.
// Rounds: 7 is the fewest for which Philox4x32 passes BigCrush (Salmon et al., SC'11, table 2: the
// "Crush-resistant" minimum; the library default of 10 is that plus a safety margin).  The generator is a
// third of the Ising kernel's instructions: 7 rounds measured +14 % on 1024^3, +9 % on 8192^2, +2.5 % on
// the Heisenberg kernels (profiles/r02/README.md).  The oracle replays with the same constant
// (VO_PHILOX_ROUNDS, oracle/vegas_oracle.h); both are pinned by Random123's 7- and 10-round known answers.
// ---------------------------------------------------------------------------------------
// Round keys are precomputed on the host (PhiloxKey, passed by value = constant bank) so that no
// per-thread integer adds are spent on the key schedule.
#ifndef VEGAS_PHILOX_ROUNDS
#define VEGAS_PHILOX_ROUNDS 7
#endif
constexpr int PHILOX_ROUNDS = VEGAS_PHILOX_ROUNDS;
struct PhiloxKey {
    uint32_t k[PHILOX_ROUNDS][2];
};

inline PhiloxKey make_philox_key(uint64_t seed) {
    PhiloxKey pk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        pk.k[r][0] = k0; pk.k[r][1] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return pk;
}

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey& pk,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ pk.k[r][0];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ pk.k[r][1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Counter layout used everywhere: c0,c1 = 64-bit index (bit 62 of the index carries the
// colour for word-keyed streams), c2 = sweep low, c3 = sweep bits 32..55 | call << 24.
__device__ __forceinline__ void philox_at(uint64_t index, uint64_t sweep, uint32_t call, const PhiloxKey& pk,
                                          uint32_t (&out)[4]) {
    philox4x32((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)sweep,
                  ((uint32_t)(sweep >> 32) & 0x00FFFFFFu) | (call << 24), pk, out);
}

// ---------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums NV values per thread over the block and lets thread 0 add them to out[0..NV) with
// one atomic each.  `red` is shared scratch of NV * 32 elements.
template <typename T, int NV>
__device__ __forceinline__ void block_atomic_add(T (&v)[NV], T* red, T* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        T s = warp_sum(v[i]);
        if (lane == 0) red[i * 32 + warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            T s = lane < nwarp ? red[i * 32 + lane] : T(0);
            s = warp_sum(s);
            if (lane == 0) atomicAdd(out + i, s);
        }
    }
}

// Barrier-free block reduction of NV 32-bit integers: every warp reduces with REDUX, adds into shared
// accumulators, takes a ticket; the last warp to arrive flushes the block totals with one 64-bit global
// atomic per value.  `s_acc` (NV ints) and `s_cnt` must be zeroed before (one barrier at kernel start).
// All 32 lanes of every warp must call this.
template <int NV>
__device__ __forceinline__ void block_flush_int(const int (&v)[NV], int* s_acc, unsigned int* s_cnt,
                                                unsigned long long* out) {
    int r[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) r[i] = __reduce_add_sync(0xffffffffu, v[i]);
    if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) atomicAdd(s_acc + i, r[i]);
        __threadfence_block();
        const unsigned int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
        if (atomicAdd(s_cnt, 1u) == nwarp - 1) {
            __threadfence_block();
#pragma unroll
            for (int i = 0; i < NV; ++i)
                atomicAdd(out + i, (unsigned long long)(long long)atomicAdd(s_acc + i, 0));
        }
    }
}

__device__ __forceinline__ unsigned long long* as_ull(long long* p) { return reinterpret_cast<unsigned long long*>(p); }

// ---------------------------------------------------------------------------------------
// uniform variates
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53;
}
__device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * 0x1.0p-24f; }

}  // namespace vg
