// basis_pipe.cu -- K4p: the 2 (bcc) / 4 (fcc) colour passes of a periodic Heisenberg step as ONE persistent, phase-pipelined
// launch.
//
// Replaces MetropolisIntegrator::step (src/integrator.rs:66-92; MetropolisFlipIntegrator :109-138 when FLIP) for
// HeisenbergSpin on `Lattice::bcc / fcc (..).expand(x, y, z)` (src/input.rs:296-322), compound energy of
// src/energy.rs:63-257.  Colour = basis index.
//
// Why.  As four separate launches every pass re-reads the three partner sublattices from DRAM: 62 B/attempt measured for
// fcc 384^3 against 24 B/attempt algorithmic (profiles/r01z_heis_basis_vec.*).  Here CTA (basis b, band t) owns a band of
// rows for the whole march over the cell planes and all CTAs are co-resident (cooperative launch, one per SM).  Basis b
// trails basis b-1 by two or three planes: before it updates plane z it waits until every lower basis a < b has
// published the planes of a that b's bonds reach (z and, for the pairs with a z offset, z + 1) on the bands t-1, t, t+1.
// That one rule covers the true dependencies (b sees the NEW spins of the lower colours, exactly as in colour-ordered
// launches) and, bonds being symmetric, the anti-dependencies (a has finished reading the OLD spins b overwrites).  The
// first basis may not lead the last one by more than `lead` planes, so everything between the fronts stays in L2 and
// DRAM sees each sublattice once in and once out per step.
//
// A consumer thread gathers its 8 / 12 neighbours with 16-byte loads (same walk over the compile-time unit-cell table
// as heis_basis_vec_kernel: same summation order, bit-identical), makes the attempt and writes the new spins into a
// shared-memory tile; a publisher warp stores the tile with bulk async copies (cp.async.bulk, TMA engine: no generic
// global stores in flight, so publishing a plane has nothing to drain) and releases the band's progress counter.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "basis_pipe.hpp"
#include "pipe_ptx.cuh"

namespace vg {

namespace {

template <typename real>
struct BasisPipeArgs {
    BasisPtrs<real> P;
    uint32_t nx, ny, nz;
    uint32_t tiles, items_per_band, lead, pub_every;   // bands per colour; work items (16-byte vectors) of a plane per band
    unsigned long long* prog;      // [NB][tiles] planes finished, monotone over the launches
    unsigned long long base;
    unsigned int* error;
    HeisParams<real> p;
    uint64_t sweep;
    PhiloxKey pk;
    double* obs;
};

// largest dz among the bonds from basis B to basis A: B at plane z needs A's planes z .. z + reach
template <int UC, int B, int A>
constexpr __host__ __device__ int basis_reach() {
    int m = -1;
    for (int q = 0; q < BasisCell<UC>::Z; ++q) {
        const BasisNb nb = basis_neighbour<UC, B>(q);
        if (nb.tb == A && nb.dz > m) m = nb.dz;
    }
    return m;   // -1: no bond
}

constexpr uint32_t BP_STAGES = 2;   // output tiles per warp: one being filled, one being stored
constexpr uint32_t BP_CNT = 8;      // per-plane completion counters in flight; a warp may be at most BP_AHEAD planes ahead of the slowest
constexpr uint32_t BP_AHEAD = 4;

// One work item: NV consecutive cells of row iy of plane iz.  Neighbour walk copied from heis_basis_vec_kernel (heis_basis.cuh).
template <typename real, int UC, int B, bool FLIP, bool RECORD>
__device__ __forceinline__ void basis_item(const BasisPipeArgs<real>& A, uint32_t iz, uint32_t iy, uint32_t x0, real* out /* [3][out_comp] */,
                                           uint32_t out_off, uint32_t out_comp, real (&fs)[5], int& accepted) {
    constexpr int NB = BasisCell<UC>::NB, Z = BasisCell<UC>::Z, N = VecOf<real>::N;
    const uint32_t nx = A.nx, ny = A.ny, nz = A.nz;
    const int zs[3] = {(int)(iz == 0 ? nz - 1 : iz - 1), (int)iz, (int)(iz + 1 == nz ? 0u : iz + 1)};
    const uint32_t ys[3] = {iy == 0 ? ny - 1 : iy - 1, iy, iy + 1 == ny ? 0u : iy + 1};
    const uint32_t xl = x0 == 0 ? nx - 1 : x0 - 1, xr = x0 + N == nx ? 0u : x0 + N;   // carries (periodic in x)
    const int cell = ((int)iz * (int)ny + (int)iy) * (int)nx + (int)x0;
    real n[3][N], l[3][N];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < N; ++e) { n[c][e] = 0; l[c][e] = 0; }
    auto gather = [&](auto qtag) {
        constexpr int Q = decltype(qtag)::value;
        constexpr BasisNb nb = basis_neighbour<UC, B>(Q);
        const int row = (zs[nb.dz + 1] * (int)ny + (int)ys[nb.dy + 1]) * (int)nx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            real v[N];
            vec_load(A.P.s[nb.tb][c] + (row + (int)x0), v);   // identical address for the two x offsets of a row: loaded once
            real u[N];
            if (nb.dx == 0) {
#pragma unroll
                for (int e = 0; e < N; ++e) u[e] = v[e];
            } else if (nb.dx < 0) {
                u[0] = A.P.s[nb.tb][c][row + (int)xl];
#pragma unroll
                for (int e = 1; e < N; ++e) u[e] = v[e - 1];
            } else {
#pragma unroll
                for (int e = 0; e + 1 < N; ++e) u[e] = v[e + 1];
                u[N - 1] = A.P.s[nb.tb][c][row + (int)xr];
            }
#pragma unroll
            for (int e = 0; e < N; ++e) {
                n[c][e] += u[e];
                if (RECORD && nb.tb < B) l[c][e] += u[e];
            }
        }
    };
    basis_for_each(gather, std::make_index_sequence<Z>{});
    real sx[N], sy[N], sz[N];
    vec_load(A.P.s[B][0] + cell, sx); vec_load(A.P.s[B][1] + cell, sy); vec_load(A.P.s[B][2] + cell, sz);
    const uint64_t gcell = ((uint64_t)iz * ny + iy) * nx + x0;
#pragma unroll
    for (int e = 0; e < N; ++e) {
        HeisRand<real> rnd;
        heis_rand((gcell + e) * NB + B, A.sweep, A.pk, rnd);
        const bool ok = heis_attempt<real, FLIP>(sx[e], sy[e], sz[e], heis_field(A.p.J, n[0][e], A.p.h[0]), heis_field(A.p.J, n[1][e], A.p.h[1]),
                                                 heis_field(A.p.J, n[2][e], A.p.h[2]), A.p, rnd);
        accepted += ok ? 1 : 0;
    }
    vec_store(out + out_off, sx); vec_store(out + out_comp + out_off, sy); vec_store(out + 2 * out_comp + out_off, sz);
    if (RECORD) {
#pragma unroll
        for (int e = 0; e < N; ++e) {
            fs[0] += sx[e] * l[0][e] + sy[e] * l[1][e] + sz[e] * l[2][e];
            fs[1] += sx[e]; fs[2] += sy[e]; fs[3] += sz[e];
            const real d = sx[e] * A.p.a[0] + sy[e] * A.p.a[1] + sz[e] * A.p.a[2];
            fs[4] += d * d;
        }
    }
}

// shared control words of a CTA
struct BandShared {
    uint32_t gate_open;          // planes z < gate_open may be updated (dependencies on the other colours' bands known to hold)
    uint32_t gate_lock;          // one warp at a time polls the progress counters
    uint32_t cta_done;           // planes every warp of this CTA has stored completely
    uint32_t abort_flag;
    uint32_t plane_cnt[BP_CNT];  // warps whose stores of plane (z % BP_CNT) are complete
};

__device__ __forceinline__ bool st_abort(const BandShared* sh) {   // warp-uniform view of the abort flag
    return __any_sync(0xffffffffu, *(volatile const uint32_t*)&sh->abort_flag != 0u);
}

// The march of one warp of CTA (B, band): every thread owns ONE item (N cells) of the band in every plane.  No CTA-wide
// barrier, no helper warp: a warp stores its own 32 items with bulk async copies, counts its finished planes in shared
// memory, and whoever completes a plane's count publishes the band's progress.
template <typename real, int UC, int B, bool FLIP, bool RECORD>
__device__ __forceinline__ void basis_march(const BasisPipeArgs<real>& A, uint32_t band, uint32_t item0, uint32_t n_items, real* out_ring,
                                            uint32_t out_comp, BandShared* sh, double* s_acc) {
    constexpr int NB = BasisCell<UC>::NB, N = VecOf<real>::N;
    const uint32_t nz = A.nz, VX = A.nx / N;
    const uint32_t tm = band == 0 ? A.tiles - 1 : band - 1, tp = band + 1 == A.tiles ? 0u : band + 1;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const bool active = threadIdx.x < n_items;
    const uint32_t item = item0 + (active ? threadIdx.x : 0u);
    const uint32_t iy = item / VX, x0 = (item - iy * VX) * N;
    // my warp's 32 items are contiguous in the band, hence in every component array
    const uint32_t w_first = warp * 32u, w_count = w_first < n_items ? min(32u, n_items - w_first) : 0u;
    const uint32_t w_bytes = w_count * N * (uint32_t)sizeof(real);
    const size_t w_goff = (size_t)(item0 + w_first) * N;          // offset of my warp's first cell inside a plane
    const size_t plane_elems = (size_t)A.ny * A.nx;
    unsigned long long* const my_prog = A.prog + (size_t)B * A.tiles + band;
    volatile uint32_t* const v_gate = &sh->gate_open;
    volatile uint32_t* const v_done = &sh->cta_done;
    volatile uint32_t* const v_abort = &sh->abort_flag;
    constexpr int R0 = basis_reach<UC, B, 0>(), R1 = basis_reach<UC, B, (NB > 1 ? 1 : 0)>(), R2 = basis_reach<UC, B, (NB > 2 ? 2 : 0)>();
    real fs[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    auto count_plane = [&](uint32_t zdone) {   // lane 0: my warp's stores of plane zdone are complete
        __threadfence_block();   // my completed stores before my count, the other warps' counts before what follows
        const uint32_t c = atomicAdd(&sh->plane_cnt[zdone % BP_CNT], 1u) + 1u;
        __threadfence_block();
        if (c == n_warps) {
            sh->plane_cnt[zdone % BP_CNT] = 0u;
            atomicMax(&sh->cta_done, zdone + 1u);
            if ((zdone + 1u) % A.pub_every == 0u || zdone + 1u == nz) {
                __threadfence();
                atomicMax(my_prog, A.base + (unsigned long long)(zdone + 1u));
            }
        }
    };
    for (uint32_t z = 0; z < nz; ++z) {
        // ---- may plane z be updated?  (other colours' progress, and not too far ahead of this CTA's slowest warp)
        // lane 0 reads the shared control words and decides for the warp: 1 go, 2 poll (it took the lock), 3 abandon, 0 wait
        {
            unsigned long long t0 = 0;
            uint32_t spins = 0;
            for (;;) {
                uint32_t st = 0;
                if (lane == 0) {
                    const uint32_t gate = *v_gate, done = *v_done;
                    if (*v_abort) st = 3u;
                    else if (z < gate && z <= done + BP_AHEAD) st = 1u;
                    else if (z >= gate && atomicCAS(&sh->gate_lock, 0u, 1u) == 0u) st = 2u;
                }
                st = __shfl_sync(0xffffffffu, st, 0);
                if (st == 1u || st == 3u) break;
                if (st == 2u) {
                    // lanes 3a .. 3a+2: lower colour a on the bands t-1, t, t+1; lane 9: the last colour (lead bound)
                    const uint32_t a = lane / 3u, which = lane - a * 3u;
                    const int reach = a == 0 ? R0 : (a == 1 ? R1 : R2);
                    uint32_t open = nz;
                    if (lane < 9u && (int)a < B && reach >= 0) {
                        const unsigned long long v = ld_acquire_gpu(A.prog + (size_t)a * A.tiles + (which == 0 ? tm : (which == 1 ? band : tp)));
                        const uint32_t pl = v > A.base ? (uint32_t)min(v - A.base, (unsigned long long)nz) : 0u;   // planes published
                        open = pl >= nz ? nz : (pl > (uint32_t)reach ? pl - (uint32_t)reach : 0u);
                    } else if (lane == 9u && B == 0 && NB > 1) {
                        const unsigned long long v = ld_acquire_gpu(A.prog + (size_t)(NB - 1) * A.tiles + band);
                        const uint32_t pl = v > A.base ? (uint32_t)min(v - A.base, (unsigned long long)nz) : 0u;
                        open = min(nz, pl + A.lead);
                    }
                    open = __reduce_min_sync(0xffffffffu, open);
                    if (lane == 0) {
                        if (open > *v_gate) *v_gate = open;
                        __threadfence_block();
                        atomicExch(&sh->gate_lock, 0u);
                    }
                    __syncwarp();
                    if (open > z) continue;          // decided at the top of the loop (the run-ahead bound may still hold the warp)
                }
                __nanosleep(200);
                if (lane == 0 && (++spins & 63u) == 0) {
                    if (t0 == 0) t0 = global_timer();
                    else if (global_timer() - t0 > PIPE_TIMEOUT_NS) { sh->abort_flag = 1u; atomicExch(A.error, (unsigned int)PIPE_ERR_GATE); }
                }
            }
            if (st_abort(sh)) break;
        }
        // ---- my warp's output stage: the store of plane z - 2 has read it
        if (lane == 0 && z >= BP_STAGES) tma_store_wait_read<BP_STAGES - 1>();
        __syncwarp();
        real* out = out_ring + (size_t)(z % BP_STAGES) * 3 * out_comp;
        if (active) basis_item<real, UC, B, FLIP, RECORD>(A, z, iy, x0, out, threadIdx.x * N, out_comp, fs, accepted);
        fence_proxy_async_smem();     // my shared-memory writes before the bulk store reads them
        __syncwarp();
        if (lane == 0) {
            if (w_count) {
                const size_t goff = (size_t)z * plane_elems + w_goff;
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) bulk_store_1d(A.P.s[B][c] + goff, out + (size_t)c * out_comp + (size_t)w_first * N, w_bytes);
            }
            tma_store_commit();
            if (z > 0) { tma_store_wait<1>(); count_plane(z - 1); }   // plane z - 1 of my warp is in global memory
        }
        if (RECORD && (z & 15u) == 15u) heis_flush(fs, s_acc);
    }
    if (!st_abort(sh) && lane == 0) { tma_store_wait<0>(); count_plane(nz - 1); }
    if (RECORD) heis_flush(fs, s_acc);
    const int a = __reduce_add_sync(0xffffffffu, accepted);
    if (lane == 0 && a != 0) atomicAdd(&s_acc[5], (double)a);
}

template <typename real, int UC, bool FLIP, bool RECORD>
__global__ void __launch_bounds__(1024, 1) basis_pipe_kernel(const __grid_constant__ BasisPipeArgs<real> A) {
    constexpr int NB = BasisCell<UC>::NB, N = VecOf<real>::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t b = blockIdx.x / A.tiles, band = blockIdx.x - b * A.tiles;
    const uint32_t total = A.ny * (A.nx / N);                         // items of a plane
    const uint32_t item0 = band * A.items_per_band, n_items = min(A.items_per_band, total - item0);
    const uint32_t out_comp = blockDim.x * N;                         // elements per component of an output stage
    real* out_ring = reinterpret_cast<real*>(smem_raw);
    double* s_acc = reinterpret_cast<double*>(out_ring + (size_t)BP_STAGES * 3 * out_comp);
    BandShared* sh = reinterpret_cast<BandShared*>(s_acc + 6);
    if (threadIdx.x == 0) {
        sh->gate_open = (b == 0 && NB > 1) ? min(A.nz, A.lead) : (b == 0 ? A.nz : 0u);
        sh->gate_lock = 0u; sh->cta_done = 0u; sh->abort_flag = 0u;
        for (uint32_t k = 0; k < BP_CNT; ++k) sh->plane_cnt[k] = 0u;
    }
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
    switch (b) {
        case 0: basis_march<real, UC, 0, FLIP, RECORD>(A, band, item0, n_items, out_ring, out_comp, sh, s_acc); break;
        case 1: basis_march<real, UC, 1, FLIP, RECORD>(A, band, item0, n_items, out_ring, out_comp, sh, s_acc); break;
        case 2: if (NB > 2) basis_march<real, UC, (NB > 2 ? 2 : 0), FLIP, RECORD>(A, band, item0, n_items, out_ring, out_comp, sh, s_acc); break;
        default: if (NB > 3) basis_march<real, UC, (NB > 3 ? 3 : 0), FLIP, RECORD>(A, band, item0, n_items, out_ring, out_comp, sh, s_acc); break;
    }
    __syncthreads();
    // layout of the basis kernels' observable row: [0] = sum_i sum_j J s_i.s_j with every bond twice (here: twice the bonds
    // towards the lower colours), [1..3] = sum s, [4] = sum (s.a)^2, [5] = accepted
    if (threadIdx.x < 6 && s_acc[threadIdx.x] != 0.0)
        atomicAdd(A.obs + threadIdx.x, threadIdx.x == 0 ? 2.0 * (double)A.p.J * s_acc[0] : s_acc[threadIdx.x]);
}

}  // namespace

struct BasisPipeState {
    BasisPipeDesc d;
    uint32_t NB = 0, tiles = 0, items_per_band = 0, threads = 0, lead = 0, pub_every = 0;
    size_t smem = 0;
    unsigned long long* d_prog = nullptr;
    unsigned int* d_error = nullptr;
    unsigned long long launches = 0;
    std::string text;
};

namespace {

template <typename real, int UC>
const void* bp_kernel_ptr(bool flip, bool record) {
    if (flip) return record ? (const void*)basis_pipe_kernel<real, UC, true, true> : (const void*)basis_pipe_kernel<real, UC, true, false>;
    return record ? (const void*)basis_pipe_kernel<real, UC, false, true> : (const void*)basis_pipe_kernel<real, UC, false, false>;
}
template <typename real>
const void* bp_kernel(int uc, bool flip, bool record) {
    return uc == 1 ? bp_kernel_ptr<real, 1>(flip, record) : bp_kernel_ptr<real, 2>(flip, record);
}
const void* bp_kernel_any(bool f64, int uc, bool flip, bool record) {
    return f64 ? bp_kernel<double>(uc, flip, record) : bp_kernel<float>(uc, flip, record);
}

}  // namespace

BasisPipeState* basis_pipe_create(const BasisPipeDesc& d, std::string& why) {
    const size_t sz = d.f64 ? 8 : 4;
    const uint32_t N = (uint32_t)(16 / sz);
    if (d.unitcell != 1 && d.unitcell != 2) { why = "unit cell is neither bcc nor fcc"; return nullptr; }
    const uint32_t NB = d.unitcell == 1 ? 2u : 4u;
    if (d.nx % N || d.nx < N || d.ny < 2 || d.nz < 8) { why = "needs nx a multiple of a 16-byte vector, >= 2 rows and >= 8 planes"; return nullptr; }
    int sms = 0, smem_max = 0, coop = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d.device) != cudaSuccess ||
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.device) != cudaSuccess ||
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d.device) != cudaSuccess || !coop) {
        cudaGetLastError();
        why = "device attributes / cooperative launch unavailable";
        return nullptr;
    }
    BasisPipeState* st = new BasisPipeState();
    st->d = d; st->NB = NB;
    // bands of consecutive work items (16-byte vectors, row-major inside a plane): one per thread, at least one row each
    const uint32_t VX = d.nx / N, total = d.ny * VX;
    uint32_t tiles = d.tiles ? d.tiles : (uint32_t)sms / NB;
    tiles = std::max(1u, std::min(std::min(tiles, d.ny), (uint32_t)sms / NB));
    uint32_t ipb = std::max(VX, (total + tiles - 1) / tiles);
    tiles = (total + ipb - 1) / ipb;
    if (ipb > 1024) { why = "a band has more than 1024 work items per plane (lattice plane too large for one CTA per SM and colour)"; delete st; return nullptr; }
    st->tiles = tiles; st->items_per_band = ipb;
    st->threads = std::max(64u, (ipb + 31u) / 32u * 32u);
    st->smem = (size_t)BP_STAGES * 3 * st->threads * 16 + 6 * 8 + sizeof(BandShared) + 16;
    if (st->smem > (size_t)smem_max) { why = "band does not fit in shared memory"; delete st; return nullptr; }
    for (int v = 0; v < 4; ++v) {
        const void* k = bp_kernel_any(d.f64, d.unitcell, v & 1, v & 2);
        int per_sm = 0;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, (int)st->threads, st->smem) != cudaSuccess || per_sm < 1 ||
            (uint32_t)(per_sm * sms) < NB * tiles) {
            cudaGetLastError();
            why = "kernel cannot be made co-resident (registers / shared memory)";
            delete st;
            return nullptr;
        }
    }
    st->pub_every = std::max(1u, d.pub_every ? d.pub_every : 1u);
    st->lead = std::max(NB * (st->pub_every + 2u), d.lead ? d.lead : NB * (st->pub_every + 3u));
    if (cudaMalloc(&st->d_prog, (size_t)NB * tiles * 8) != cudaSuccess || cudaMalloc(&st->d_error, 4) != cudaSuccess) {
        cudaGetLastError();
        why = "cudaMalloc failed";
        basis_pipe_destroy(st);
        return nullptr;
    }
    cudaMemset(st->d_prog, 0, (size_t)NB * tiles * 8);
    cudaMemset(st->d_error, 0, 4);
    char buf[256];
    snprintf(buf, sizeof buf, "basis_pipe: %u colours x %u bands, %u items/band, %u threads, lead %u planes, publish every %u, %zu B smem",
             NB, tiles, ipb, st->threads, st->lead, st->pub_every, st->smem);
    st->text = buf;
    return st;
}

void basis_pipe_destroy(BasisPipeState* st) {
    if (!st) return;
    cudaFree(st->d_prog); cudaFree(st->d_error);
    delete st;
}

const char* basis_pipe_describe(const BasisPipeState* st) { return st ? st->text.c_str() : ""; }

template <typename real>
int basis_pipe_step(BasisPipeState* st, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                    double* obs_row, cudaStream_t stream, std::string& err) {
    const BasisPipeDesc& d = st->d;
    BasisPipeArgs<real> A;
    memset(&A, 0, sizeof A);
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) A.P.s[b][c] = (real*)d.arr[b][c];
    A.nx = d.nx; A.ny = d.ny; A.nz = d.nz;
    A.tiles = st->tiles; A.items_per_band = st->items_per_band; A.lead = st->lead; A.pub_every = st->pub_every;
    A.prog = st->d_prog;
    A.base = st->launches * (unsigned long long)d.nz;
    A.error = st->d_error;
    A.p = p; A.sweep = sweep; A.pk = pk; A.obs = obs_row;
    const void* k = bp_kernel<real>(d.unitcell, flip, record);
    void* args[] = {&A};
    const cudaError_t e = cudaLaunchCooperativeKernel(k, dim3(st->NB * st->tiles), dim3(st->threads), args, st->smem, stream);
    if (e != cudaSuccess) {
        err = std::string("basis_pipe_kernel launch failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return -1;
    }
    st->launches++;
    return 0;
}
template int basis_pipe_step<float>(BasisPipeState*, const HeisParams<float>&, bool, bool, uint64_t, const PhiloxKey&, double*, cudaStream_t, std::string&);
template int basis_pipe_step<double>(BasisPipeState*, const HeisParams<double>&, bool, bool, uint64_t, const PhiloxKey&, double*, cudaStream_t, std::string&);

int basis_pipe_check(BasisPipeState* st, std::string& err) {
    if (!st) return 0;
    unsigned int e = 0;
    if (cudaMemcpy(&e, st->d_error, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { err = "basis_pipe: cannot read the error flag"; return -1; }
    if (e == 0) return 0;
    err = "basis_pipe_kernel: a dependency or ring wait timed out (results invalid)";
    cudaMemset(st->d_prog, 0, (size_t)st->NB * st->tiles * 8);
    cudaMemset(st->d_error, 0, 4);
    st->launches = 0;
    return -1;
}

}  // namespace vg
