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
#include "pipe_ptx.cuh"

namespace vg {

namespace {

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
    uint32_t tiles, rows, tiles_long, S, SO, n_cw, lead, pub_every;
    uint32_t backoff_consumer, backoff_helper;   // ns slept between failed barrier polls (0: re-poll at once)
    uint32_t l2_hints;                           // != 0: TMA loads / stores carry L2 eviction-priority hints
    unsigned long long* prog;            // [2][tiles] planes finished (published), monotone over the launches
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

// March bookkeeping shared by the producer and the consumers of a CTA.
//   planes of the `other` sequence q = 0, 1, ...: step i consumes q = qbase(i), +1, +2 (z-1, z, z+1)
template <bool HALO>
struct March {
    uint32_t Lz, phase;
    __device__ __forceinline__ uint32_t z_of(uint32_t i) const { const uint32_t z = phase + i; return z >= Lz ? z - Lz : z; }
    // HALO colour 1: planes 1 .. Lz-1 use q = i .. i+2 over [0 .. Lz-1, HI]; the last step (plane 0) uses [LO, 0, 1] = q Lz+1 ..
    __device__ __forceinline__ uint32_t qbase(uint32_t i) const { return (HALO && phase == 1 && i + 1 == Lz) ? Lz + 1 : i; }
    // plane of sequence entry q: local z, or -1 (lower halo) / Lz (upper halo) for a slab
    __device__ __forceinline__ int plane_of(uint32_t q) const {
        if (!HALO) { const uint32_t z = phase + Lz - 1 + q; return (int)(z % Lz); }
        if (phase == 0) return (int)q - 1;                     // -1, 0, ..., Lz
        return q <= Lz ? (int)q : (int)(q - Lz) - 2;           // 0 .. Lz-1, Lz (HI) | -1 (LO), 0, 1
    }
};

template <typename real, int V> struct PackOf;
template <> struct PackOf<float, 4> { typedef float4 type; };
template <> struct PackOf<float, 2> { typedef float2 type; };
template <> struct PackOf<double, 2> { typedef double2 type; };
template <> struct PackOf<double, 1> { typedef double type; };
template <typename real, int V>
__device__ __forceinline__ void pack_load(const real* p, real (&v)[V]) {
    typedef typename PackOf<real, V>::type T;
    const T q = *reinterpret_cast<const T*>(p);
    const real* e = reinterpret_cast<const real*>(&q);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = e[i];
}
template <typename real, int V>
__device__ __forceinline__ void pack_store(real* p, const real (&v)[V]) {
    typedef typename PackOf<real, V>::type T;
    T q;
    real* e = reinterpret_cast<real*>(&q);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = v[i];
    *reinterpret_cast<T*>(p) = q;
}

constexpr uint32_t PIPE_DONE_SLOTS = 8;   // > own-ring depth (<= 4): the consumers are never that far ahead of the publisher, which
                                          // co-owns the own-ring slots (it arrives on empty_w after publishing the plane)

// V = sites per consumer thread (one 16-byte vector, or half of one for twice the warps per band)
// The consumers write the new spins back into the own-ring slot and the publisher stores the tile with TMA (no generic
// global stores in flight, so the release that publishes a plane has nothing to drain).
// AXZ: the anisotropy axis is (0, 0, a_z): s.a = s_z a_z (bit-identical to the general dot product, 7 instructions fewer per site)
template <typename real, int V, bool FLIP, bool RECORD, bool HALO, bool AXZ, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) heis_pipe_kernel(const __grid_constant__ PipeArgs<real> A) {
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
    uint64_t* done = empty_w + SO;                                        // PIPE_DONE_SLOTS: plane stored by every consumer warp
    double* s_acc = reinterpret_cast<double*>(done + PIPE_DONE_SLOTS);    // 6 doubles
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(s_acc + 6);
    uint32_t* ready = const_cast<uint32_t*>(abort_flag) + 1;   // planes whose TMA stores are complete (publisher -> releaser)
    uint32_t* bnd = ready + 1;   // HALO: [0] boundary planes the consumers have stored into the neighbours' halos so far, [1], [2] their z

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n_phases = gridDim.x / A.tiles;
    const bool publish = true;     // every phase: the next one waits for it, and the first one may not run too far ahead of the last
    const bool has_publisher = publish || HALO;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(full_o + s, 1u); mbar_init(empty_o + s, n_cw); }
        for (uint32_t s = 0; s < SO; ++s) { mbar_init(full_w + s, 1u); mbar_init(empty_w + s, 1u); }
        for (uint32_t s = 0; s < PIPE_DONE_SLOTS; ++s) mbar_init(done + s, n_cw);
        *abort_flag = 0u;
        *ready = 0u; bnd[0] = 0u;
        fence_barrier_init();
        fence_proxy_async();
    }
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;
    __syncthreads();

    March<HALO> mz{Lz, phase};

    if (warp == n_cw) {
        // ===================== producer warp: lane 0 waits (ring slots, phase dependencies), the lanes issue the TMA loads =====
        // (one thread issuing the 12 loads of a plane one after the other took longer than the consumers need per plane)
        const uint32_t oc = (uint32_t)(1 - colour);
        const uint32_t kind = nr == rows ? 0u : 1u;
        const uint32_t ym = y0 == 0 ? Ly - 1 : y0 - 1, yp = y0 + nr == Ly ? 0u : y0 + nr;
        const uint32_t bytes_o = 3u * (nr + 2u) * Hx * (uint32_t)sizeof(real);
        const uint32_t tm = tile == 0 ? A.tiles - 1 : tile - 1, tp = tile + 1 == A.tiles ? 0u : tile + 1;
        const unsigned long long* dep = phase > 0 ? A.prog + (size_t)(phase - 1) * A.tiles : nullptr;
        // lane l < 9 loads part l % 3 (halo row below, band rows, halo row above) of component l / 3 of an `other` plane;
        // lane l < 3 loads component l of an own plane
        const uint32_t my_c = lane / 3u, my_part = lane - my_c * 3u;
        const uint32_t my_y = my_part == 0 ? ym : (my_part == 1 ? y0 : yp);
        const uint32_t my_row = my_part == 0 ? 0u : (my_part == 1 ? 1u : 1u + nr);
        const uint32_t my_kind = my_part == 1 ? kind : 2u;
        bool ok = true;
        // L2 hints: the first phase's `other` planes are the last phase's own planes a few planes later (keep them); every other
        // phase reads its `other` planes (the previous phase's output) for the last time in this step
        const uint64_t pol_other = phase + 1 < n_phases ? l2_policy_evict_last() : l2_policy_evict_first();
        RingPos po;              // next `other` slot to fill; parity = of the fill being made
        unsigned long long seen[3] = {0, 0, 0};   // lane 0: last progress values read (monotone counters)
        uint32_t q_next = 0, z_next = (phase + Lz - 1) % Lz;     // non-slab: plane of entry q_next (advances with wrap)
        auto load_other = [&]() {
            const uint32_t q = q_next++;
            const uint32_t slot = po.slot, parity = po.parity;
            po.advance(S);
            int pl;
            if (HALO) pl = mz.plane_of(q);
            else { pl = (int)z_next; z_next = z_next + 1 == Lz ? 0u : z_next + 1; }
            uint32_t array = oc, zc = (uint32_t)pl;
            const bool is_halo = HALO && (pl < 0 || pl >= (int)Lz);
            if (is_halo) { array = 2u + oc * 2u + (pl < 0 ? 0u : 1u); zc = 0u; }
            uint32_t good = 1u, fenced = 0u;
            if (lane == 0) {
                if (q >= S && !wait_bar(empty_o + slot, parity ^ 1u, abort_flag, A.error, PIPE_ERR_EMPTY, A.backoff_helper)) good = 0u;
                if (good && is_halo) {
                    // a neighbour's boundary plane: wait until all its CTAs have stored it.  Other colour = phase - 1's
                    // output of THIS step for colour 1, the previous step's colour 1 for colour 0.
                    const unsigned long long target = A.flag_base + (colour == 1 ? (unsigned long long)A.tiles : 0ull);
                    unsigned long long seen_peer = 0;
                    if (!wait_counter<true>(A.flags + (pl < 0 ? 0u : 2u) + oc, target, seen_peer, abort_flag, A.error, PIPE_ERR_PEER)) good = 0u;
                    fenced = 1u;
                } else if (good && phase > 0) {
                    // position of this plane in the previous phase's march (it starts at plane phase - 1; slabs at 0)
                    const uint32_t first = HALO ? 0u : (phase - 1u) % Lz;
                    const uint32_t j = zc >= first ? zc - first : zc + Lz - first;
                    const unsigned long long target = A.base + (unsigned long long)(j + 1u);
                    // progress is published in groups of planes: most of the time the last values seen already cover the target
                    if (seen[0] < target || seen[1] < target || seen[2] < target) {
                        if (!wait_counter<false>(dep + tm, target, seen[0], abort_flag, A.error, PIPE_ERR_GATE) ||
                            !wait_counter<false>(dep + tile, target, seen[1], abort_flag, A.error, PIPE_ERR_GATE) ||
                            !wait_counter<false>(dep + tp, target, seen[2], abort_flag, A.error, PIPE_ERR_GATE)) good = 0u;
                        fenced = 1u;
                    }
                }
                if (good) mbar_expect_tx(full_o + slot, bytes_o);
            }
            const uint32_t both = __shfl_sync(0xffffffffu, good | (fenced << 1), 0);
            if (!(both & 1u)) { ok = false; return; }
            if (lane < 9) {
                if (both & 2u) fence_proxy_async();   // data written through the generic proxy by other SMs, read by TMA
                real* dst = ring_o + (size_t)slot * stage_o + (my_c * orow + my_row) * Hx;
                if (A.l2_hints) tma_load_3d_hint(dst, A.maps + (size_t)array * 9 + my_c * 3 + my_kind, full_o + slot, 0u, my_y, zc, pol_other);
                else tma_load_3d(dst, A.maps + (size_t)array * 9 + my_c * 3 + my_kind, full_o + slot, 0u, my_y, zc);
            }
        };
        // sequence entries in the order the consumers need them: q <= qbase(i) + 2 before step i
        for (uint32_t i = 0; ok && i < Lz; ++i) {
            const uint32_t q_need = mz.qbase(i) + 2u;
            while (ok && q_next <= q_need) load_other();
        }
    } else if (warp == n_cw + 3) {
        // ===================== releaser: publishes the band's progress (planes whose stores are complete) at gpu scope ==========
        // Runs at its own pace: when a release takes longer than a plane, the next one simply covers several planes.
        if (lane == 0) {
            unsigned long long* const my_prog = A.prog + (size_t)phase * A.tiles + tile;
            uint32_t last = 0, bnd_seen = 0;
            const unsigned long long t0 = global_timer();
            while (last < Lz || (HALO && bnd_seen < 2u)) {
                uint32_t now;
                if (HALO) {
                    uint32_t nb;
                    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(nb) : "r"(smem_u32(bnd)) : "memory");
                    if (nb != bnd_seen) {
                        __threadfence_system();
                        for (; bnd_seen < nb; ++bnd_seen) {
                            // plane 0 feeds the lower neighbour's UPPER halo: its "from upper" words [2 + colour]; plane Lz-1 the
                            // upper neighbour's "from lower" words [colour]
                            const uint32_t z = ((volatile uint32_t*)bnd)[1 + bnd_seen];
                            if (z == 0) atomicAdd_system(A.peer_flags[0] + 2 + colour, 1ull);
                            else atomicAdd_system(A.peer_flags[1] + colour, 1ull);
                        }
                    }
                }
                asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(now) : "r"(smem_u32(ready)) : "memory");
                if (now == last) {
                    __nanosleep(100);
                    if (*abort_flag || global_timer() - t0 > 8 * PIPE_TIMEOUT_NS) break;
                    continue;
                }
                last = now;
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(my_prog), "l"(A.base + (unsigned long long)now) : "memory");
            }
        }
    } else if (warp == n_cw + 2) {
        // ===================== own-ring producer warp (independent of the other ring: neither holds the other up) =========
        const uint32_t kind = nr == rows ? 0u : 1u;
        const uint32_t bytes_w = 3u * nr * Hx * (uint32_t)sizeof(real);
        // the first phase may lead the last one by at most `lead` planes: everything in between stays in L2
        const unsigned long long* const tail = A.prog + (size_t)(n_phases - 1) * A.tiles + tile;
        unsigned long long seen_tail = 0;
        const uint64_t pol_own = l2_policy_evict_first();
        RingPos pw;
        for (uint32_t i = 0; i < Lz; ++i) {
            const uint32_t slot = pw.slot, parity = pw.parity;
            pw.advance(SO);
            uint32_t good = 1u;
            if (lane == 0) {
                if (phase == 0 && i >= A.lead && seen_tail < A.base + (unsigned long long)(i - A.lead) + 1ull &&
                    !wait_counter<false>(tail, A.base + (unsigned long long)(i - A.lead) + 1ull, seen_tail, abort_flag, A.error, PIPE_ERR_GATE)) good = 0u;
                if (good && i >= SO && !wait_bar(empty_w + slot, parity ^ 1u, abort_flag, A.error, PIPE_ERR_EMPTY, A.backoff_helper)) good = 0u;
                if (good) mbar_expect_tx(full_w + slot, bytes_w);
            }
            good = __shfl_sync(0xffffffffu, good, 0);
            if (!good) break;
            if (lane < 3) {
                real* dst = ring_w + (size_t)slot * stage_w + (lane * rows) * Hx;
                // own planes are read once per step: never worth keeping
                if (A.l2_hints) tma_load_3d_hint(dst, A.maps + (size_t)colour * 9 + lane * 3 + kind, full_w + slot, 0u, y0, mz.z_of(i), pol_own);
                else tma_load_3d(dst, A.maps + (size_t)colour * 9 + lane * 3 + kind, full_w + slot, 0u, y0, mz.z_of(i));
            }
        }
    } else if (warp == n_cw + 1) {
        // ===================== publisher: makes finished planes visible to the next phase / the neighbour slabs ==========
        // (keeps the gpu-scope fence off the consumers' critical path; the consumers' stores are ordered before it by
        // their arrive on the `done` barrier)
        if (lane == 0 && has_publisher) {
            RingPos pd, pown;
            const uint64_t pol_store = phase + 1 < n_phases ? l2_policy_evict_last() : l2_policy_evict_first();
            uint32_t since_pub = 0;
            const uint32_t kind = nr == rows ? 0u : 1u;
            uint32_t n_bnd = 0;
            // a boundary plane sits in the neighbour's halo once every consumer warp has arrived: the releaser signals it (the
            // system-scope fence waits for the NVLink stores and must not hold up the own-ring slots)
            auto signal_peers = [&](uint32_t z) {
                if (z == 0 || z + 1 == Lz) {
                    bnd[1 + n_bnd] = z;
                    ++n_bnd;
                    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(bnd)), "r"(n_bnd) : "memory");
                }
            };
            for (uint32_t i = 0; i < Lz; ++i) {
                if (!wait_bar(done + pd.slot, pd.parity, abort_flag, A.error, PIPE_ERR_FULL, A.backoff_helper)) break;
                pd.advance(PIPE_DONE_SLOTS);
                // plane i sits updated in its own-ring slot: store it; then retire plane i - 1 (slot readable again once the
                // store has read it, progress published once its writes are complete)
                const real* src = ring_w + (size_t)pown.slot * stage_w;
                const CUtensorMap* m = A.maps + (size_t)colour * 9;
                const uint32_t z = mz.z_of(i);
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) {
                    // the updated tile is the next phase's `other` plane within a few planes (keep it); the last phase's output
                    // is not read again before the next step
                    if (A.l2_hints) tma_store_3d_hint(m + c * 3 + kind, src + (c * rows) * Hx, 0u, y0, z, pol_store);
                    else tma_store_3d(m + c * 3 + kind, src + (c * rows) * Hx, 0u, y0, z);
                }
                tma_store_commit();
                tma_store_wait_read<0>();               // the store has read the slot: the producer may refill it
                mbar_arrive(empty_w + pown.slot);
                pown.advance(SO);
                // planes <= i - PUB_LAG are complete once at most PUB_LAG stores are pending: waiting for older stores only
                // keeps the publisher (which also frees the own-ring slots) from ever blocking on a store's completion latency
                constexpr uint32_t PUB_LAG = 2;
                if (i >= PUB_LAG && ++since_pub == A.pub_every) {
                    since_pub = 0;
                    tma_store_wait<PUB_LAG>();
                    // hand the count to the releaser warp: the gpu-scope release costs about a plane's time and must not sit here
                    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(ready)), "r"(i + 1u - PUB_LAG) : "memory");
                }
                if (HALO) signal_peers(z);
            }
            tma_store_wait<0>();
            asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(ready)), "r"(Lz) : "memory");
        }
    } else {
        // ===================== consumers: one thread per V sites of the band =====================
        const uint32_t Tx = Hx / V;                                   // threads per row
        const uint32_t r = threadIdx.x / Tx, gx = threadIdx.x - r * Tx;
        const bool active = r < nr;
        const uint32_t y = y0 + (active ? r : 0u);
        const uint32_t el = y * Hx + gx * V;                          // offset inside a plane
        const uint32_t so_row = ((active ? r : 0u) + 1u) * Hx + gx * V;   // my row inside a component block of an `other` stage
        const uint32_t sw_row = (active ? r : 0u) * Hx + gx * V;
        const uint32_t cx_right = (gx + 1 == Tx) ? 0u : (gx + 1) * V, cx_left = (gx == 0 ? Tx : gx) * V - 1;
        const bool energy = colour == 1;
        real facc[5] = {0, 0, 0, 0, 0};
        float2 facc2[5];   // fp32: per-lane partial sums of the packed path, folded into facc at every flush
#pragma unroll
        for (int k = 0; k < 5; ++k) facc2[k] = make_float2(0.0f, 0.0f);
        auto fold = [&]() {
            if constexpr (sizeof(real) == 4) {
#pragma unroll
                for (int k = 0; k < 5; ++k) { facc[k] += facc2[k].x + facc2[k].y; facc2[k] = make_float2(0.0f, 0.0f); }
            }
        };
        int accepted = 0;
        bool ok = true;
        RingPos p_lo, p_wait, p_own, p_done;    // q = qbase(i); next `other` entry to wait for; own ring; done ring
        uint32_t q_waited = 0;                  // `other` entries < q_waited have landed
        for (uint32_t i = 0; i < Lz; ++i) {
            const uint32_t qb = mz.qbase(i);
            for (; q_waited <= qb + 2u; ++q_waited) {
                ok = ok && wait_bar(full_o + p_wait.slot, p_wait.parity, abort_flag, A.error, PIPE_ERR_FULL, A.backoff_consumer);
                p_wait.advance(S);
            }
            ok = ok && wait_bar(full_w + p_own.slot, p_own.parity, abort_flag, A.error, PIPE_ERR_FULL, A.backoff_consumer);
            if (!__all_sync(0xffffffffu, ok)) break;
            const uint32_t z = mz.z_of(i);
            const uint32_t slot_lo = p_lo.slot, slot_n0 = slot_lo + 1 == S ? 0u : slot_lo + 1, slot_hi = slot_n0 + 1 == S ? 0u : slot_n0 + 1;
            if (active) {
                const uint32_t zg = z + g.z_offset;
                const uint32_t rp = (y + zg + (uint32_t)colour) & 1u;
                const real* pl = ring_o + slot_lo * stage_o + so_row;
                const real* pn = ring_o + slot_n0 * stage_o + so_row;
                const real* ph = ring_o + slot_hi * stage_o + so_row;
                const real* pc = ring_o + slot_n0 * stage_o + (so_row - gx * V) + (rp ? cx_right : cx_left);
                const real* pw = ring_w + p_own.slot * stage_w + sw_row;
                real s[3][V];
                if constexpr (sizeof(real) == 4) {
                    // fp32: two sites per instruction (FADD2 / FMUL2 / FFMA2); lane arithmetic identical to the scalar kernels
                    constexpr int NP = V / 2;
                    float2 s2[3][NP], n2[3][NP];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float n0[V], a[V], b[V], lo[V], hi[V], sv[V];
                        const uint32_t co = (uint32_t)c * orow * Hx;
                        pack_load<float, V>(pw + (uint32_t)c * rows * Hx, sv);
                        pack_load<float, V>(pn + co, n0);
                        pack_load<float, V>(pn + co - Hx, a);
                        pack_load<float, V>(pn + co + Hx, b);
                        const float carry = pc[co];
                        pack_load<float, V>(pl + co, lo);
                        pack_load<float, V>(ph + co, hi);
                        // same association order as heis_march (heis.cuh): ((n0 + (a + b)) + x-shifted) + (lo + hi)
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const int e = 2 * q;
                            float2 t = l_add(make_float2(n0[e], n0[e + 1]), l_add(make_float2(a[e], a[e + 1]), make_float2(b[e], b[e + 1])));
                            float2 sh;
                            if (rp) {   // warp-uniform: a warp lies inside one row
                                asm volatile("");
                                sh = make_float2(n0[e + 1], e + 2 < V ? n0[(e + 2) % V] : carry);
                            } else {
                                asm volatile("");
                                sh = make_float2(e > 0 ? n0[(e + V - 1) % V] : carry, n0[e]);
                            }
                            t = l_add(t, sh);
                            n2[c][q] = l_add(t, l_add(make_float2(lo[e], lo[e + 1]), make_float2(hi[e], hi[e + 1])));
                            s2[c][q] = make_float2(sv[e], sv[e + 1]);
                        }
                    }
                    const uint64_t site0 = (uint64_t)(zg * Ly + y) * g.Lx + 2u * (gx * V) + rp;  // element e: site0 + 2e
#pragma unroll
                    for (int q = 0; q < NP; ++q) {   // bit 1 of site0 is clear: elements 2q, 2q + 1 share a Philox call
                        uint32_t rr[4];
                        philox_at(site0 + 4u * q, A.sweep, 0u, A.pk, rr);
                        const HeisRand<float> r0 = heis_rand_words(rr[0], rr[1]), r1 = heis_rand_words(rr[2], rr[3]);
                        bool a0, a1;
                        heis_attempt2<FLIP, AXZ>(s2[0][q], s2[1][q], s2[2][q], heis_field(A.p.J, n2[0][q], A.p.h[0]),
                                                 heis_field(A.p.J, n2[1][q], A.p.h[1]), heis_field(A.p.J, n2[2][q], A.p.h[2]), A.p, r0, r1, a0, a1);
                        accepted += (a0 ? 1 : 0) + (a1 ? 1 : 0);
                        if (RECORD) {
                            if (energy) facc2[0] = l_fma(l_bc<float2>(-A.p.J), l_fma(s2[2][q], n2[2][q], l_fma(s2[1][q], n2[1][q], l_mul(s2[0][q], n2[0][q]))), facc2[0]);
                            facc2[1] = l_add(facc2[1], s2[0][q]); facc2[2] = l_add(facc2[2], s2[1][q]); facc2[3] = l_add(facc2[3], s2[2][q]);
                            const float2 d1 = AXZ ? l_mul(s2[2][q], l_bc<float2>(A.p.a[2]))
                                                  : l_fma(s2[2][q], l_bc<float2>(A.p.a[2]), l_fma(s2[1][q], l_bc<float2>(A.p.a[1]), l_mul(s2[0][q], l_bc<float2>(A.p.a[0]))));
                            facc2[4] = l_fma(d1, d1, facc2[4]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int q = 0; q < NP; ++q) { s[c][2 * q] = s2[c][q].x; s[c][2 * q + 1] = s2[c][q].y; }
                } else {
                    real nsum[3][V];
    #pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        real n0[V], a[V], b[V], lo[V], hi[V];
                        const uint32_t co = (uint32_t)c * orow * Hx;
                        pack_load<real, V>(pw + (uint32_t)c * rows * Hx, s[c]);
                        pack_load<real, V>(pn + co, n0);
                        pack_load<real, V>(pn + co - Hx, a);
                        pack_load<real, V>(pn + co + Hx, b);
                        const real carry = pc[co];
                        pack_load<real, V>(pl + co, lo);
                        pack_load<real, V>(ph + co, hi);
                        // same association order as heis_march (heis.cuh): bit-identical neighbour sums
    #pragma unroll
                        for (int e = 0; e < V; ++e) nsum[c][e] = n0[e] + (a[e] + b[e]);
                        if (rp) {   // warp-uniform (a warp lies inside one row); the empty asm keeps the arms from being if-converted into selects
                            asm volatile("");
    #pragma unroll
                            for (int e = 0; e < V; ++e) nsum[c][e] += e + 1 < V ? n0[(e + 1) % V] : carry;
                        } else {
                            asm volatile("");
    #pragma unroll
                            for (int e = 0; e < V; ++e) nsum[c][e] += e > 0 ? n0[(e + V - 1) % V] : carry;
                        }
    #pragma unroll
                        for (int e = 0; e < V; ++e) nsum[c][e] += lo[e] + hi[e];
                    }
                    HeisRand<real> rnd[V];
                    const uint64_t site0 = (uint64_t)(zg * Ly + y) * g.Lx + 2u * (gx * V) + rp;  // element e: site0 + 2e
                    if (sizeof(real) == 4) {
    #pragma unroll
                        for (int e = 0; e < V; e += 2) {   // bit 1 of site0 is clear: elements e, e + 1 share a Philox call
                            uint32_t rr[4];
                            philox_at(site0 + 2u * e, A.sweep, 0u, A.pk, rr);
                            reinterpret_cast<HeisRand<float>&>(rnd[e]) = heis_rand_words(rr[0], rr[1]);
                            reinterpret_cast<HeisRand<float>&>(rnd[e + 1]) = heis_rand_words(rr[2], rr[3]);
                        }
                    } else {
    #pragma unroll
                        for (int e = 0; e < V; ++e) heis_rand(site0 + 2u * e, A.sweep, A.pk, rnd[e]);
                    }
    #pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const bool acc = heis_attempt<real, FLIP, AXZ>(s[0][e], s[1][e], s[2][e], heis_field(A.p.J, nsum[0][e], A.p.h[0]),
                                                                  heis_field(A.p.J, nsum[1][e], A.p.h[1]), heis_field(A.p.J, nsum[2][e], A.p.h[2]), A.p, rnd[e]);
                        accepted += acc ? 1 : 0;
                        if (RECORD) {
                            if (energy) facc[0] -= A.p.J * (s[0][e] * nsum[0][e] + s[1][e] * nsum[1][e] + s[2][e] * nsum[2][e]);
                            facc[1] += s[0][e]; facc[2] += s[1][e]; facc[3] += s[2][e];
                            const real d1 = AXZ ? s[2][e] * A.p.a[2] : s[0][e] * A.p.a[0] + s[1][e] * A.p.a[1] + s[2][e] * A.p.a[2];
                            facc[4] += d1 * d1;
                        }
                    }
                }
                real* pws = ring_w + p_own.slot * stage_w + sw_row;
#pragma unroll
                for (int c = 0; c < 3; ++c) pack_store<real, V>(pws + (uint32_t)c * rows * Hx, s[c]);
                fence_proxy_async_smem();     // my shared-memory writes before the publisher's TMA store reads them
                if (HALO) {
                    if (z == 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) pack_store<real, V>(A.peer[colour][0][c] + el, s[c]);
                    }
                    if (z + 1 == Lz) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) pack_store<real, V>(A.peer[colour][1][c] + el, s[c]);
                    }
                }
            }
            __syncwarp();
            // the `other` entries the next step no longer needs, my own-ring slot, and "this warp has stored plane i"
            const uint32_t n_rel = i + 1 < Lz ? mz.qbase(i + 1) - qb : 1u;
            for (uint32_t k = 0; k < n_rel; ++k) {
                if (lane == 0) mbar_arrive(empty_o + p_lo.slot);
                p_lo.advance(S);
            }
            if (lane == 0) {
                if (has_publisher) mbar_arrive(done + p_done.slot);
            }
            p_own.advance(SO);
            p_done.advance(PIPE_DONE_SLOTS);
            if (RECORD && (i & 15u) == 15u) { fold(); heis_flush(facc, s_acc); }
        }
        if (RECORD) { fold(); heis_flush(facc, s_acc); }
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
    uint32_t Hx = 0, V = 0, tiles = 0, rows = 0, tiles_long = 0, S = 0, SO = 0, n_cw = 0, threads = 0;
    size_t smem = 0;
    CUtensorMap* d_maps = nullptr;
    unsigned long long* d_prog = nullptr;
    unsigned int* d_error = nullptr;
    unsigned long long launches = 0;
    std::string text;
};

namespace {

template <typename real, int V, bool HALO, bool AXZ, int MAXT>
const void* pipe_kernel_ptr(bool flip, bool record) {
    if (flip) return record ? (const void*)heis_pipe_kernel<real, V, true, true, HALO, AXZ, MAXT> : (const void*)heis_pipe_kernel<real, V, true, false, HALO, AXZ, MAXT>;
    return record ? (const void*)heis_pipe_kernel<real, V, false, true, HALO, AXZ, MAXT> : (const void*)heis_pipe_kernel<real, V, false, false, HALO, AXZ, MAXT>;
}
template <typename real, int V, int MAXT>
const void* pipe_kernel_v(bool axz, bool flip, bool record, bool halo) {
    if (axz) return halo ? pipe_kernel_ptr<real, V, true, true, MAXT>(flip, record) : pipe_kernel_ptr<real, V, false, true, MAXT>(flip, record);
    return halo ? pipe_kernel_ptr<real, V, true, false, MAXT>(flip, record) : pipe_kernel_ptr<real, V, false, false, MAXT>(flip, record);
}
// variants: a whole 16-byte vector per thread with up to 576 or 1024 threads, or half a vector (always 1024)
template <typename real>
const void* pipe_kernel(uint32_t V, bool axz, bool flip, bool record, bool halo, uint32_t threads) {
    constexpr int N = VecOf<real>::N;
    if (V != (uint32_t)N) return pipe_kernel_v<real, N / 2, 1024>(axz, flip, record, halo);
    return threads <= 576 ? pipe_kernel_v<real, N, 576>(axz, flip, record, halo) : pipe_kernel_v<real, N, 1024>(axz, flip, record, halo);
}
const void* pipe_kernel_any(bool f64, uint32_t V, bool axz, bool flip, bool record, bool halo, uint32_t threads) {
    return f64 ? pipe_kernel<double>(V, axz, flip, record, halo, threads) : pipe_kernel<float>(V, axz, flip, record, halo, threads);
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
    st->Hx = Hx;
    // bands: one CTA per SM, half of them per colour
    uint32_t tiles = d.tiles ? d.tiles : (uint32_t)(sms / 2);
    tiles = std::max(1u, std::min(tiles, d.Ly));
    tiles = std::min<uint32_t>(tiles, (uint32_t)(sms / 2));
    const uint32_t rows = (d.Ly + tiles - 1) / tiles;
    tiles = (d.Ly + rows - 1) / rows;                       // no empty bands
    st->tiles = tiles; st->rows = rows;
    st->tiles_long = d.Ly - tiles * (rows - 1);             // bands of `rows` rows; the others have rows - 1
    // sites per consumer thread: a whole 16-byte vector, or half of one (tuning key: twice the warps, more instructions per site)
    auto threads_for = [&](uint32_t V) { return (rows * (Hx / V) + 31u) / 32u * 32u + 128u; };   // + two producer warps, the publisher and the releaser
    uint32_t V = d.vec ? d.vec : N;   // measured on 512^3 fp32: whole vectors 0.88 ms per step, half vectors (twice the warps) 1.04
    if (V != N && V != N / 2) { why = "sites per thread must be a whole or half 16-byte vector"; delete st; return nullptr; }
    if (!d.vec && threads_for(V) > 1024) { why = "band needs more than 1024 threads"; delete st; return nullptr; }
    st->V = V;
    st->threads = threads_for(V);
    st->n_cw = st->threads / 32 - 4;
    if (st->threads > 1024) { why = "band needs more than 1024 threads"; delete st; return nullptr; }
    const size_t stage_o = (size_t)3 * (rows + 2) * Hx * sz, stage_w = (size_t)3 * rows * Hx * sz;
    // ring depths: an own plane holds its slot from the load until the TMA store has read the updated tile
    const uint32_t choices[][2] = {{6, 3}, {5, 3}, {4, 3}, {4, 2}, {4, 1}};   // measured on 512^3 fp32: 0.862 / 0.878 ms per step for the first two
    auto smem_for = [&](uint32_t S, uint32_t SO) { return S * stage_o + SO * stage_w + (size_t)(2 * S + 2 * SO + PIPE_DONE_SLOTS) * 8 + 6 * 8 + 48; };
    if (d.stages_other >= 4 && d.stages_own >= 1) { st->S = d.stages_other; st->SO = std::min(4u, d.stages_own); }
    else
        for (auto& c : choices)
            if (smem_for(c[0], c[1]) <= (size_t)smem_max) { st->S = c[0]; st->SO = c[1]; break; }
    if (st->S == 0 || smem_for(st->S, st->SO) > (size_t)smem_max) { why = "band does not fit in shared memory"; delete st; return nullptr; }
    st->smem = smem_for(st->S, st->SO);
    // every variant must be able to hold one CTA per SM with this configuration, and the grid must be co-resident
    for (int v = 0; v < 16; ++v) {
        const void* k = pipe_kernel_any(d.f64, st->V, v & 8, v & 1, v & 2, v & 4, st->threads);
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
    snprintf(buf, sizeof buf, "heis_pipe: %u bands x 2 colours, %u rows/band, %u sites/thread, %u threads, other ring %u, own ring %u, %zu B smem%s",
             tiles, rows, st->V, st->threads, st->S, st->SO, st->smem, d.slab ? ", slab" : "");
    st->text = buf;
    return st;
}

void heis_pipe_destroy(HeisPipeState* st) {
    if (!st) return;
    cudaFree(st->d_maps); cudaFree(st->d_prog); cudaFree(st->d_error);
    delete st;
}

const char* heis_pipe_describe(const HeisPipeState* st) { return st ? st->text.c_str() : ""; }
uint32_t heis_pipe_tiles(const HeisPipeState* st) { return st ? st->tiles : 0u; }

template <typename real>
int heis_pipe_step(HeisPipeState* st, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                   double* obs_row, uint64_t slab_steps, cudaStream_t stream, std::string& err) {
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
    A.g.Hx = st->Hx; A.g.Gx = st->Hx / st->V; A.g.Ly = d.Ly; A.g.Lz = d.Lz; A.g.z_offset = d.z_offset; A.g.Lx = d.Lx;
    A.tiles = st->tiles; A.rows = st->rows; A.tiles_long = st->tiles_long; A.S = st->S; A.SO = st->SO; A.n_cw = st->n_cw;
    // Defaults measured on 512^3 fp32 (profiles/r02/README.md): publishing every plane makes the fronts wait on each other's
    // release latency (pub 1 / lead 32: 0.95 ms per step), every 4th with a lead of 32 planes 0.85 ms.  With the 7-round
    // generator the consumers are faster and want a longer lead: pub 4 / lead 48 0.826 ms and 4.03 GB of DRAM traffic per
    // 512^3 step, pub 8 / lead 64 0.829 ms and 4.33 GB, pub 4 / lead 32 0.861 ms (profiles/r02/h7.sh; the lead is a cap, the
    // second colour normally follows pub + 2..4 planes behind the first).  Leads below 2 pub + 8 can deadlock (publication
    // lags the update by the store's completion and the releaser's turn-around)
    A.pub_every = std::max(1u, d.pub_every ? d.pub_every : 4u);
    A.lead = std::max(2u * A.pub_every + 8u, d.lead ? d.lead : 12u * A.pub_every);
    A.backoff_consumer = d.backoff_consumer; A.backoff_helper = d.backoff_helper; A.l2_hints = d.l2_hints;
    A.prog = st->d_prog;
    A.base = st->launches * (unsigned long long)d.Lz;
    A.flags = d.flags;
    A.peer_flags[0] = d.peer_flags[0]; A.peer_flags[1] = d.peer_flags[1];
    A.flag_base = slab_steps * (unsigned long long)st->tiles;
    A.error = st->d_error;
    A.p = p; A.sweep = sweep; A.pk = pk; A.obs = obs_row;
    const bool axz = p.a[0] == (real)0 && p.a[1] == (real)0;
    const void* k = pipe_kernel<real>(st->V, axz, flip, record, d.slab, st->threads);
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
template int heis_pipe_step<float>(HeisPipeState*, const HeisParams<float>&, bool, bool, uint64_t, const PhiloxKey&, double*, uint64_t, cudaStream_t, std::string&);
template int heis_pipe_step<double>(HeisPipeState*, const HeisParams<double>&, bool, bool, uint64_t, const PhiloxKey&, double*, uint64_t, cudaStream_t, std::string&);

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
