// heis_pipe.cu -- K3p: the Heisenberg checkerboard step on sc lattices as ONE persistent, phase-pipelined launch fed by TMA.
//
// Replaces MetropolisIntegrator::step (src/integrator.rs:66-92; MetropolisFlipIntegrator :109-138 when FLIP) for
// HeisenbergSpin with the compound energy of src/energy.rs:63-214 (closed form in heis.cuh).
//
// Structure.  The lattice is cut into `tiles` bands of rows (full x extent).  CTA (phase, tile) -- phase = colour for a
// one-step launch -- owns its band for the whole march over the z planes and is resident for the whole launch (one CTA
// per SM, cooperative launch).  Inside a CTA one producer thread feeds two shared-memory rings with
// cp.async.bulk.tensor (TMA) guarded by mbarriers:
//   * `other` ring: the other colour's plane tiles with one halo row below and above (rows + 2 rows x 3 components);
//     a plane is fetched ONCE and serves as z+1, z and z-1 neighbour of three consecutive march steps;
//   * `own` ring: the plane tile being updated.
// The consumer warps (one thread per 16-byte vector of the band) compute from shared memory only and store the new
// spins straight to global memory.  Phase p trails phase p-1 by a few planes: before the producer of (p, t) fetches
// the other colour's plane z it waits until the CTAs (p-1, t-1..t+1) have published that plane (per-CTA progress
// counters, release/acquire at gpu scope), so the second colour pass finds both the first pass's output and its own
// old spins in L2: DRAM traffic is the compulsory 24 B/attempt (fp32) instead of 36 for two separate passes.  The wait
// also covers the anti-dependency (phase p-1 has consumed the old values phase p overwrites).
// Phase p starts its march at plane p (colour 1: planes 1, 2, ..., Lz-1, 0) because its last plane needs phase p-1's
// first one across the periodic boundary.
//
// Connected z-slabs (HALO): the planes below / above the local range come from the halo buffers the neighbours store
// into over NVLink; boundary planes are also stored into the neighbours' halos and signalled with system-scope
// counters, and the producers wait on the counters written by the neighbours before they fetch a halo plane.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "heis_pipe.hpp"

namespace vg {

namespace {

// ---------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, fences, scoped loads
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, uint32_t c0, uint32_t c1, uint32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long PIPE_TIMEOUT_NS = 4000000000ull;   // a wait that long means a broken dependency: give up, never hang

enum PipeError : unsigned int { PIPE_ERR_FULL = 1, PIPE_ERR_EMPTY = 2, PIPE_ERR_GATE = 3, PIPE_ERR_PEER = 4 };

// Waits for phase `parity` of an mbarrier.  False when the launch is being abandoned (abort flag) or on time-out.
__device__ __forceinline__ bool wait_bar(uint64_t* b, uint32_t parity, volatile uint32_t* abort_flag, unsigned int* gerr, unsigned int code) {
    if (mbar_try_wait(b, parity)) return true;
    const unsigned long long t0 = global_timer();
    uint32_t n = 0;
    while (!mbar_try_wait(b, parity)) {
        if ((++n & 63u) == 0) {
            if (*abort_flag) return false;
            if (global_timer() - t0 > PIPE_TIMEOUT_NS) { *abort_flag = 1u; atomicExch(gerr, code); return false; }
        }
    }
    return true;
}

// Spins until *p >= target (SYS: the word is written by another GPU).
template <bool SYS>
__device__ __forceinline__ bool wait_counter(const unsigned long long* p, unsigned long long target, volatile uint32_t* abort_flag,
                                             unsigned int* gerr, unsigned int code) {
    if ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) >= target) return true;
    const unsigned long long t0 = global_timer();
    uint32_t n = 0;
    while ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) < target) {
        __nanosleep(32);
        if ((++n & 63u) == 0) {
            if (*abort_flag) return false;
            if (global_timer() - t0 > PIPE_TIMEOUT_NS) { *abort_flag = 1u; atomicExch(gerr, code); return false; }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------
// tensor maps: index ((array * 3 + component) * 3 + kind); array 0, 1 = the colour arrays, 2 + (colour * 2 + hi) = halo
// planes of a slab; kind 0: box of `rows` rows, 1: rows - 1 rows, 2: one row.
constexpr int PIPE_MAPS = 6 * 3 * 3;

template <typename real>
struct PipeArgs {
    real* arr[2][3];
    real* peer[2][2][3];                 // [colour][to lower / to upper][component] (HALO)
    const CUtensorMap* maps;
    HeisGeom g;
    uint32_t tiles, rows, tiles_long, S, SO, n_cw;
    unsigned long long* prog;            // [2][tiles] planes x consumer warps finished, monotone over the launches
    unsigned long long base;             // value of every progress counter when this launch starts
    const unsigned long long* flags;     // HALO: [lower, upper][colour] boundary-plane CTAs that have stored into my halos
    unsigned long long* peer_flags[2];   // HALO: the word block of the lower / upper neighbour I add to
    unsigned long long flag_base;        // HALO: launches so far * tiles
    unsigned int* error;
    HeisParams<real> p;
    uint64_t sweep;
    PhiloxKey pk;
    double* obs;
};

struct PlaneRef { uint32_t array, z; };   // tensor-map array index and plane coordinate inside it

// March bookkeeping shared by the producer and the consumers of a CTA.
//   planes of the `other` sequence q = 0, 1, ...: step i consumes q = qbase(i), +1, +2 (z-1, z, z+1)
template <bool HALO>
struct March {
    uint32_t Lz, phase;
    __device__ __forceinline__ uint32_t steps() const { return Lz; }
    __device__ __forceinline__ uint32_t z_of(uint32_t i) const { const uint32_t z = phase + i; return z >= Lz ? z - Lz : z; }
    // HALO colour 1: planes 1 .. Lz-1 use q = i .. i+2 over [0 .. Lz-1, HI]; the last step (plane 0) uses [LO, 0, 1] = q Lz+1 ..
    __device__ __forceinline__ uint32_t qbase(uint32_t i) const { return (HALO && phase == 1 && i + 1 == Lz) ? Lz + 1 : i; }
    __device__ __forceinline__ uint32_t n_q() const { return (HALO && phase == 1) ? Lz + 4 : Lz + 2; }
    // plane of sequence entry q: local z, or -1 (lower halo) / Lz (upper halo) for a slab
    __device__ __forceinline__ int plane_of(uint32_t q) const {
        if (!HALO) { const uint32_t z = phase + Lz - 1 + q; return (int)(z % Lz); }
        if (phase == 0) return (int)q - 1;                     // -1, 0, ..., Lz
        return q <= Lz ? (int)q : (int)(q - Lz) - 2;           // 0 .. Lz-1, Lz (HI) | -1 (LO), 0, 1
    }
};

template <typename real, bool FLIP, bool RECORD, bool HALO, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) heis_pipe_kernel(const __grid_constant__ PipeArgs<real> A) {
    constexpr int N = VecOf<real>::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const HeisGeom& g = A.g;
    const uint32_t phase = blockIdx.x / A.tiles, tile = blockIdx.x - phase * A.tiles;
    const int colour = (int)(phase & 1u);
    const uint32_t rows = A.rows;
    const uint32_t nr = tile < A.tiles_long ? rows : rows - 1;
    const uint32_t y0 = tile < A.tiles_long ? tile * rows : A.tiles_long * rows + (tile - A.tiles_long) * (rows - 1);
    const uint32_t Hx = g.Hx, Ly = g.Ly, Lz = g.Lz;
    const uint32_t orow = rows + 2;
    const uint32_t stage_o = 3 * orow * Hx, stage_w = 3 * rows * Hx;   // elements
    const uint32_t S = A.S, SO = A.SO, n_cw = A.n_cw;
    real* ring_o = reinterpret_cast<real*>(smem_raw);
    real* ring_w = ring_o + (size_t)S * stage_o;
    uint64_t* full_o = reinterpret_cast<uint64_t*>(ring_w + (size_t)SO * stage_w);
    uint64_t* empty_o = full_o + S;
    uint64_t* full_w = empty_o + S;
    uint64_t* empty_w = full_w + SO;
    double* s_acc = reinterpret_cast<double*>(empty_w + SO);              // 6 doubles
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(s_acc + 6);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(full_o + s, 1u); mbar_init(empty_o + s, n_cw); }
        for (uint32_t s = 0; s < SO; ++s) { mbar_init(full_w + s, 1u); mbar_init(empty_w + s, n_cw); }
        *abort_flag = 0u;
        fence_barrier_init();
        fence_proxy_async();
    }
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;
    __syncthreads();

    March<HALO> mz{Lz, phase};
    const uint32_t n_phases = gridDim.x / A.tiles;

    if (warp == n_cw) {
        // ===================== producer: one thread issues every TMA load of this CTA =====================
        if (lane == 0) {
            const uint32_t oc = (uint32_t)(1 - colour);
            const uint32_t kind = nr == rows ? 0u : 1u;
            const uint32_t ym = y0 == 0 ? Ly - 1 : y0 - 1, yp = y0 + nr == Ly ? 0u : y0 + nr;
            const uint32_t bytes_o = 3u * (nr + 2u) * Hx * (uint32_t)sizeof(real), bytes_w = 3u * nr * Hx * (uint32_t)sizeof(real);
            const uint32_t tm = tile == 0 ? A.tiles - 1 : tile - 1, tp = tile + 1 == A.tiles ? 0u : tile + 1;
            const unsigned long long* dep = phase > 0 ? A.prog + (size_t)(phase - 1) * A.tiles : nullptr;
            bool ok = true;
            auto load_other = [&](uint32_t q) {
                const uint32_t slot = q % S, use = q / S;
                if (use > 0 && !wait_bar(empty_o + slot, (use - 1u) & 1u, abort_flag, A.error, PIPE_ERR_EMPTY)) { ok = false; return; }
                const int pl = mz.plane_of(q);
                uint32_t array = oc, zc = (uint32_t)pl;
                if (HALO && (pl < 0 || pl >= (int)Lz)) {
                    // a neighbour's boundary plane: wait until all its CTAs have stored it.  Other colour = phase - 1's
                    // output of THIS step for colour 1, the previous step's colour 1 for colour 0.
                    const uint32_t hi = pl < 0 ? 0u : 1u;
                    const unsigned long long target = A.flag_base + (colour == 1 ? (unsigned long long)A.tiles : 0ull);
                    if (!wait_counter<true>(A.flags + hi * 2u + oc, target, abort_flag, A.error, PIPE_ERR_PEER)) { ok = false; return; }
                    array = 2u + oc * 2u + hi; zc = 0u;
                    fence_proxy_async();
                } else if (phase > 0) {
                    // position of this plane in the previous phase's march (it starts at plane phase - 1; slabs at 0)
                    const uint32_t first = HALO ? 0u : (phase - 1u) % Lz;
                    const uint32_t j = zc >= first ? zc - first : zc + Lz - first;
                    const unsigned long long target = A.base + (unsigned long long)(j + 1u) * n_cw;
                    if (!wait_counter<false>(dep + tm, target, abort_flag, A.error, PIPE_ERR_GATE) ||
                        !wait_counter<false>(dep + tile, target, abort_flag, A.error, PIPE_ERR_GATE) ||
                        !wait_counter<false>(dep + tp, target, abort_flag, A.error, PIPE_ERR_GATE)) { ok = false; return; }
                    fence_proxy_async();
                }
                mbar_expect_tx(full_o + slot, bytes_o);
                real* dst = ring_o + (size_t)slot * stage_o;
                const CUtensorMap* m = A.maps + (size_t)array * 9;
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) {
                    tma_load_3d(dst + (c * orow) * Hx, m + c * 3 + 2, full_o + slot, 0u, ym, zc);
                    tma_load_3d(dst + (c * orow + 1u) * Hx, m + c * 3 + kind, full_o + slot, 0u, y0, zc);
                    tma_load_3d(dst + (c * orow + 1u + nr) * Hx, m + c * 3 + 2, full_o + slot, 0u, yp, zc);
                }
            };
            auto load_own = [&](uint32_t i) {
                const uint32_t slot = i % SO, use = i / SO;
                if (use > 0 && !wait_bar(empty_w + slot, (use - 1u) & 1u, abort_flag, A.error, PIPE_ERR_EMPTY)) { ok = false; return; }
                mbar_expect_tx(full_w + slot, bytes_w);
                real* dst = ring_w + (size_t)slot * stage_w;
                const CUtensorMap* m = A.maps + (size_t)colour * 9;
                const uint32_t z = mz.z_of(i);
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) tma_load_3d(dst + (c * rows) * Hx, m + c * 3 + kind, full_w + slot, 0u, y0, z);
            };
            // sequence entries in the order the consumers need them: q <= qbase(i) + 2 and own(i) before step i
            uint32_t q_next = 0;
            for (uint32_t i = 0; ok && i < Lz; ++i) {
                const uint32_t q_need = mz.qbase(i) + 2u;
                while (ok && q_next <= q_need) load_other(q_next++);
                if (ok) load_own(i);
            }
        }
    } else {
        // ===================== consumers: one thread per 16-byte vector of the band =====================
        const uint32_t r = threadIdx.x / g.Gx, gx = threadIdx.x - r * g.Gx;
        const bool active = r < nr;
        const uint32_t y = y0 + (active ? r : 0u);
        const uint32_t plane = Ly * Hx;
        const uint32_t el = y * Hx + gx * N;                    // offset inside a plane
        const uint32_t so_row = ((active ? r : 0u) + 1u) * Hx + gx * N;   // my row inside a component block of an `other` stage
        const uint32_t sw_row = (active ? r : 0u) * Hx + gx * N;
        const uint32_t cx_right = (gx + 1 == g.Gx) ? 0u : (gx + 1) * N, cx_left = (gx == 0 ? g.Gx : gx) * N - 1;
        real* const own0 = colour ? A.arr[1][0] : A.arr[0][0];
        real* const own1 = colour ? A.arr[1][1] : A.arr[0][1];
        real* const own2 = colour ? A.arr[1][2] : A.arr[0][2];
        unsigned long long* const my_prog = A.prog + (size_t)phase * A.tiles + tile;
        const bool publish = phase + 1 < n_phases;
        const bool energy = colour == 1;
        real facc[5] = {0, 0, 0, 0, 0};
        int accepted = 0;
        bool ok = true;
        uint32_t q_waited = 0;                                  // `other` entries < q_waited have landed
        for (uint32_t i = 0; i < Lz; ++i) {
            const uint32_t qb = mz.qbase(i);
            for (; q_waited <= qb + 2u; ++q_waited)
                ok = ok && wait_bar(full_o + q_waited % S, (q_waited / S) & 1u, abort_flag, A.error, PIPE_ERR_FULL);
            const uint32_t slot_w = i % SO;
            ok = ok && wait_bar(full_w + slot_w, (i / SO) & 1u, abort_flag, A.error, PIPE_ERR_FULL);
            if (!__all_sync(0xffffffffu, ok)) break;
            const uint32_t z = mz.z_of(i);
            if (active) {
                const uint32_t zg = z + g.z_offset;
                const uint32_t rp = (y + zg + (uint32_t)colour) & 1u;
                const real* pl = ring_o + (size_t)(qb % S) * stage_o + so_row;
                const real* pn = ring_o + (size_t)((qb + 1u) % S) * stage_o + so_row;
                const real* ph = ring_o + (size_t)((qb + 2u) % S) * stage_o + so_row;
                const real* pc = ring_o + (size_t)((qb + 1u) % S) * stage_o + (so_row - gx * N) + (rp ? cx_right : cx_left);
                const real* pw = ring_w + (size_t)slot_w * stage_w + sw_row;
                real s[3][N], nsum[3][N];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    real n0[N], a[N], b[N], lo[N], hi[N];
                    const uint32_t co = (uint32_t)c * orow * Hx;
                    vec_load(pw + (uint32_t)c * rows * Hx, s[c]);
                    vec_load(pn + co, n0);
                    vec_load(pn + co - Hx, a);
                    vec_load(pn + co + Hx, b);
                    const real carry = pc[co];
                    vec_load(pl + co, lo);
                    vec_load(ph + co, hi);
                    // same association order as heis_march (heis.cuh): bit-identical neighbour sums
#pragma unroll
                    for (int e = 0; e < N; ++e) nsum[c][e] = n0[e] + (a[e] + b[e]);
                    if (rp) {
#pragma unroll
                        for (int e = 0; e < N; ++e) nsum[c][e] += e + 1 < N ? n0[(e + 1) % N] : carry;
                    } else {
#pragma unroll
                        for (int e = 0; e < N; ++e) nsum[c][e] += e > 0 ? n0[(e + N - 1) % N] : carry;
                    }
#pragma unroll
                    for (int e = 0; e < N; ++e) nsum[c][e] += lo[e] + hi[e];
                }
                HeisRand<real> rnd[N];
                const uint64_t site0 = (uint64_t)(zg * Ly + y) * g.Lx + 2u * (gx * N) + rp;  // element e: site0 + 2e
                if (sizeof(real) == 4) {
#pragma unroll
                    for (int e = 0; e < N; e += 2) {
                        uint32_t rr[4];
                        philox_at(site0 + 2u * e, A.sweep, 0u, A.pk, rr);
                        reinterpret_cast<HeisRand<float>&>(rnd[e]) = heis_rand_words(rr[0], rr[1]);
                        reinterpret_cast<HeisRand<float>&>(rnd[e + 1]) = heis_rand_words(rr[2], rr[3]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < N; ++e) heis_rand(site0 + 2u * e, A.sweep, A.pk, rnd[e]);
                }
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    const bool acc = heis_attempt<real, FLIP>(s[0][e], s[1][e], s[2][e], A.p.J * nsum[0][e] - A.p.h[0],
                                                              A.p.J * nsum[1][e] - A.p.h[1], A.p.J * nsum[2][e] - A.p.h[2], A.p, rnd[e]);
                    accepted += acc ? 1 : 0;
                    if (RECORD) {
                        if (energy) facc[0] -= A.p.J * (s[0][e] * nsum[0][e] + s[1][e] * nsum[1][e] + s[2][e] * nsum[2][e]);
                        facc[1] += s[0][e]; facc[2] += s[1][e]; facc[3] += s[2][e];
                        const real d1 = s[0][e] * A.p.a[0] + s[1][e] * A.p.a[1] + s[2][e] * A.p.a[2];
                        facc[4] += d1 * d1;
                    }
                }
                const uint32_t e0 = z * plane + el;
                vec_store(own0 + e0, s[0]); vec_store(own1 + e0, s[1]); vec_store(own2 + e0, s[2]);
                if (HALO) {
                    if (z == 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) vec_store(A.peer[colour][0][c] + el, s[c]);
                    }
                    if (z + 1 == Lz) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) vec_store(A.peer[colour][1][c] + el, s[c]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                // the `other` entries the next step no longer needs, and my own-ring slot
                const uint32_t q_rel_end = i + 1 < Lz ? mz.qbase(i + 1) : qb + 1u;
                for (uint32_t q = qb; q < q_rel_end; ++q) mbar_arrive(empty_o + q % S);
                mbar_arrive(empty_w + slot_w);
                if (publish) { __threadfence(); atomicAdd(my_prog, 1ull); }
            }
            if (HALO && (z == 0 || z + 1 == Lz)) {
                // every consumer warp has stored its part of a boundary plane into the neighbour's halo: one signal per CTA
                asm volatile("bar.sync 1, %0;" ::"r"(n_cw * 32u) : "memory");
                if (threadIdx.x == 0) {
                    __threadfence_system();
                    // plane 0 feeds the lower neighbour's UPPER halo: its "from upper" words [2 + colour]; plane Lz-1 the upper
                    // neighbour's "from lower" words [colour]
                    if (z == 0) atomicAdd_system(A.peer_flags[0] + 2 + colour, 1ull);
                    if (z + 1 == Lz) atomicAdd_system(A.peer_flags[1] + colour, 1ull);
                }
            }
            if (RECORD && (i & 15u) == 15u) heis_flush(facc, s_acc);
        }
        if (RECORD) heis_flush(facc, s_acc);
        const int a = __reduce_add_sync(0xffffffffu, accepted);
        if (lane == 0 && a != 0) atomicAdd(&s_acc[5], (double)a);
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_acc[threadIdx.x] != 0.0) atomicAdd(A.obs + threadIdx.x, s_acc[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

}  // namespace

struct HeisPipeState {
    HeisPipeDesc d;
    uint32_t Hx = 0, Gx = 0, tiles = 0, rows = 0, tiles_long = 0, S = 0, SO = 0, n_cw = 0, threads = 0;
    size_t smem = 0;
    CUtensorMap* d_maps = nullptr;
    unsigned long long* d_prog = nullptr;
    unsigned int* d_error = nullptr;
    unsigned long long launches = 0;
    std::string text;
};

namespace {

template <typename real, bool HALO, int MAXT>
const void* pipe_kernel_ptr(bool flip, bool record) {
    if (flip) return record ? (const void*)heis_pipe_kernel<real, true, true, HALO, MAXT> : (const void*)heis_pipe_kernel<real, true, false, HALO, MAXT>;
    return record ? (const void*)heis_pipe_kernel<real, false, true, HALO, MAXT> : (const void*)heis_pipe_kernel<real, false, false, HALO, MAXT>;
}
template <typename real>
const void* pipe_kernel(bool flip, bool record, bool halo, uint32_t threads) {
    if (threads <= 512) return halo ? pipe_kernel_ptr<real, true, 512>(flip, record) : pipe_kernel_ptr<real, false, 512>(flip, record);
    return halo ? pipe_kernel_ptr<real, true, 1024>(flip, record) : pipe_kernel_ptr<real, false, 1024>(flip, record);
}
const void* pipe_kernel_any(bool f64, bool flip, bool record, bool halo, uint32_t threads) {
    return f64 ? pipe_kernel<double>(flip, record, halo, threads) : pipe_kernel<float>(flip, record, halo, threads);
}

}  // namespace

HeisPipeState* heis_pipe_create(const HeisPipeDesc& d, std::string& why) {
    const size_t sz = d.f64 ? 8 : 4;
    const uint32_t N = (uint32_t)(16 / sz);
    const uint32_t Hx = d.Lx / 2;
    if (d.Lx % 2 || Hx == 0 || Hx > 256 || (Hx * sz) % 128 != 0) { why = "row length: Lx/2 elements must be <= 256 and a multiple of 128 bytes"; return nullptr; }
    if (d.Lz < 8 || d.Ly < 2) { why = "needs at least 8 planes and 2 rows"; return nullptr; }
    int sms = 0, smem_max = 0, coop = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d.device) != cudaSuccess ||
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.device) != cudaSuccess ||
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d.device) != cudaSuccess || !coop) {
        cudaGetLastError();
        why = "device attributes / cooperative launch unavailable";
        return nullptr;
    }
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) { why = "cuTensorMapEncodeTiled is not available from this driver"; return nullptr; }
    HeisPipeState* st = new HeisPipeState();
    st->d = d;
    st->Hx = Hx; st->Gx = Hx / N;
    // bands: one CTA per SM, half of them per colour
    uint32_t tiles = d.tiles ? d.tiles : (uint32_t)(sms / 2);
    tiles = std::max(1u, std::min(tiles, d.Ly));
    tiles = std::min<uint32_t>(tiles, (uint32_t)(sms / 2));
    const uint32_t rows = (d.Ly + tiles - 1) / tiles;
    tiles = (d.Ly + rows - 1) / rows;                       // no empty bands
    st->tiles = tiles; st->rows = rows;
    st->tiles_long = d.Ly - tiles * (rows - 1);             // bands of `rows` rows; the others have rows - 1
    const uint32_t cthreads = (rows * st->Gx + 31u) / 32u * 32u;
    st->n_cw = cthreads / 32; st->threads = cthreads + 32;
    if (st->threads > 1024) { why = "band needs more than 1024 threads"; delete st; return nullptr; }
    const size_t stage_o = (size_t)3 * (rows + 2) * Hx * sz, stage_w = (size_t)3 * rows * Hx * sz;
    const uint32_t choices[][2] = {{6, 3}, {5, 3}, {5, 2}, {4, 2}, {4, 1}};
    auto smem_for = [&](uint32_t S, uint32_t SO) { return S * stage_o + SO * stage_w + (size_t)(2 * S + 2 * SO) * 8 + 6 * 8 + 16; };
    if (d.stages_other >= 4 && d.stages_own >= 1) { st->S = d.stages_other; st->SO = d.stages_own; }
    else
        for (auto& c : choices)
            if (smem_for(c[0], c[1]) <= (size_t)smem_max) { st->S = c[0]; st->SO = c[1]; break; }
    if (st->S == 0 || smem_for(st->S, st->SO) > (size_t)smem_max) { why = "band does not fit in shared memory"; delete st; return nullptr; }
    st->smem = smem_for(st->S, st->SO);
    // every variant must be able to hold one CTA per SM with this configuration, and the grid must be co-resident
    for (int v = 0; v < 8; ++v) {
        const void* k = pipe_kernel_any(d.f64, v & 1, v & 2, v & 4, st->threads);
        int per_sm = 0;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, (int)st->threads, st->smem) != cudaSuccess || per_sm < 1 ||
            (uint32_t)(per_sm * sms) < 2 * tiles) {
            cudaGetLastError();
            why = "kernel cannot be made co-resident (registers / shared memory)";
            delete st;
            return nullptr;
        }
    }
    // tensor maps
    std::vector<CUtensorMap> maps(PIPE_MAPS);
    memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
    auto encode_one = [&](CUtensorMap* m, void* base, uint32_t nz, uint32_t box_rows) -> bool {
        const cuuint64_t dims[3] = {Hx, d.Ly, nz};
        const cuuint64_t strides[2] = {(cuuint64_t)Hx * sz, (cuuint64_t)Hx * d.Ly * sz};
        const cuuint32_t box[3] = {Hx, box_rows, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        return encode(m, d.f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool okm = true;
    for (uint32_t a = 0; a < 6 && okm; ++a)
        for (uint32_t c = 0; c < 3 && okm; ++c) {
            void* base = a < 2 ? d.arr[a][c] : d.halo[(a - 2) / 2][(a - 2) % 2][c];
            if (!base) continue;                             // halo arrays of a single handle
            const uint32_t nz = a < 2 ? d.Lz : 1u;
            CUtensorMap* m = &maps[(a * 3 + c) * 3];
            okm = encode_one(m + 0, base, nz, rows) && encode_one(m + 1, base, nz, std::max(1u, rows - 1)) && encode_one(m + 2, base, nz, 1);
        }
    if (!okm) { why = "cuTensorMapEncodeTiled failed"; delete st; return nullptr; }
    if (cudaMalloc(&st->d_maps, maps.size() * sizeof(CUtensorMap)) != cudaSuccess ||
        cudaMalloc(&st->d_prog, (size_t)2 * tiles * 8) != cudaSuccess || cudaMalloc(&st->d_error, 4) != cudaSuccess) {
        cudaGetLastError();
        why = "cudaMalloc failed";
        heis_pipe_destroy(st);
        return nullptr;
    }
    cudaMemcpy(st->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    cudaMemset(st->d_prog, 0, (size_t)2 * tiles * 8);
    cudaMemset(st->d_error, 0, 4);
    char buf[256];
    snprintf(buf, sizeof buf, "heis_pipe: %u bands x 2 colours, %u rows/band, %u threads, other ring %u, own ring %u, %zu B smem%s",
             tiles, rows, st->threads, st->S, st->SO, st->smem, d.slab ? ", slab" : "");
    st->text = buf;
    return st;
}

void heis_pipe_destroy(HeisPipeState* st) {
    if (!st) return;
    cudaFree(st->d_maps); cudaFree(st->d_prog); cudaFree(st->d_error);
    delete st;
}

const char* heis_pipe_describe(const HeisPipeState* st) { return st ? st->text.c_str() : ""; }

template <typename real>
int heis_pipe_step(HeisPipeState* st, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                   double* obs_row, cudaStream_t stream, std::string& err) {
    const HeisPipeDesc& d = st->d;
    PipeArgs<real> A;
    memset(&A, 0, sizeof A);
    for (int col = 0; col < 2; ++col)
        for (int c = 0; c < 3; ++c) {
            A.arr[col][c] = (real*)d.arr[col][c];
            A.peer[col][0][c] = (real*)d.peer[col][0][c];
            A.peer[col][1][c] = (real*)d.peer[col][1][c];
        }
    A.maps = st->d_maps;
    A.g.Hx = st->Hx; A.g.Gx = st->Gx; A.g.Ly = d.Ly; A.g.Lz = d.Lz; A.g.z_offset = d.z_offset; A.g.Lx = d.Lx;
    A.tiles = st->tiles; A.rows = st->rows; A.tiles_long = st->tiles_long; A.S = st->S; A.SO = st->SO; A.n_cw = st->n_cw;
    A.prog = st->d_prog;
    A.base = st->launches * (unsigned long long)d.Lz * st->n_cw;
    A.flags = d.flags;
    A.peer_flags[0] = d.peer_flags[0]; A.peer_flags[1] = d.peer_flags[1];
    A.flag_base = st->launches * (unsigned long long)st->tiles;
    A.error = st->d_error;
    A.p = p; A.sweep = sweep; A.pk = pk; A.obs = obs_row;
    const void* k = pipe_kernel<real>(flip, record, d.slab, st->threads);
    void* args[] = {&A};
    const cudaError_t e = cudaLaunchCooperativeKernel(k, dim3(2 * st->tiles), dim3(st->threads), args, st->smem, stream);
    if (e != cudaSuccess) {
        err = std::string("heis_pipe_kernel launch failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return -1;
    }
    st->launches++;
    return 0;
}
template int heis_pipe_step<float>(HeisPipeState*, const HeisParams<float>&, bool, bool, uint64_t, const PhiloxKey&, double*, cudaStream_t, std::string&);
template int heis_pipe_step<double>(HeisPipeState*, const HeisParams<double>&, bool, bool, uint64_t, const PhiloxKey&, double*, cudaStream_t, std::string&);

int heis_pipe_check(HeisPipeState* st, std::string& err) {
    if (!st) return 0;
    unsigned int e = 0;
    if (cudaMemcpy(&e, st->d_error, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { err = "heis_pipe: cannot read the error flag"; return -1; }
    if (e == 0) return 0;
    static const char* const what[] = {"", "a TMA load never completed", "a ring slot was never released", "a colour-phase dependency wait timed out",
                                       "a neighbour slab never signalled its boundary plane"};
    err = std::string("heis_pipe_kernel: ") + what[e < 5 ? e : 0] + " (results invalid)";
    // re-arm: counters back to a consistent state
    cudaMemset(st->d_prog, 0, (size_t)2 * st->tiles * 8);
    cudaMemset(st->d_error, 0, 4);
    st->launches = 0;
    return -1;
}

}  // namespace vg
