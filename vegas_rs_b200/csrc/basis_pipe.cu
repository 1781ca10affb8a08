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
    uint32_t tiles, rows, tiles_long, n_cw, lead, pub_every;
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

constexpr uint32_t BP_STAGES = 2;   // output tiles in flight: one being filled, one being stored

// One work item: NV consecutive cells of row iy of plane iz.  Neighbour walk copied from heis_basis_vec_kernel (heis_basis.cuh).
template <typename real, int UC, int B, bool FLIP, bool RECORD>
__device__ __forceinline__ void basis_item(const BasisPipeArgs<real>& A, uint32_t iz, uint32_t iy, uint32_t x0, real* out /* [3][rows * nx] */,
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

struct BandCtx {
    uint32_t tile, y0, nr, n_ct;            // band, first row, rows, consumer threads
    uint64_t *done, *freed;                 // per output stage: filled by every consumer warp / read by the bulk store
    volatile uint32_t* abort_flag;
    double* s_acc;
};

// The march of the consumers of CTA (B, band) over the planes.
template <typename real, int UC, int B, bool FLIP, bool RECORD>
__device__ __forceinline__ void basis_march(const BasisPipeArgs<real>& A, const BandCtx& cx, real* out_ring, uint32_t stage_elems) {
    constexpr int NB = BasisCell<UC>::NB, N = VecOf<real>::N;
    const uint32_t nz = A.nz, VX = A.nx / N, items = cx.nr * VX;
    const uint32_t tm = cx.tile == 0 ? A.tiles - 1 : cx.tile - 1, tp = cx.tile + 1 == A.tiles ? 0u : cx.tile + 1;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t out_comp = A.rows * A.nx;
    unsigned long long seen = 0;   // warp 0: last value read of the progress counter this lane watches
    real fs[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    RingPos ps;     // output stage of plane z; parity of its use
    for (uint32_t z = 0; z < nz; ++z) {
        if (threadIdx.x < 32) {
            // warp 0 does the waiting, one lane per counter, so that the (up to eleven) round trips to L2 overlap:
            //   lanes 3a .. 3a+2: lower colour a on the bands t-1, t, t+1 (its planes my bonds reach must be final);
            //   lane 9: the last colour (the first one stays within `lead` planes of it: L2 working set);
            //   lane 10: my output stage (the store of plane z - BP_STAGES has read it)
            constexpr int R0 = basis_reach<UC, B, 0>(), R1 = basis_reach<UC, B, (NB > 1 ? 1 : 0)>(), R2 = basis_reach<UC, B, (NB > 2 ? 2 : 0)>();
            const uint32_t a = lane / 3u, which = lane - a * 3u;
            const int reach = a == 0 ? R0 : (a == 1 ? R1 : R2);
            if (lane < 9u && (int)a < B && reach >= 0) {
                const unsigned long long target = A.base + (unsigned long long)min(z + (uint32_t)reach + 1u, nz);
                if (seen < target)
                    wait_counter<false>(A.prog + (size_t)a * A.tiles + (which == 0 ? tm : (which == 1 ? cx.tile : tp)), target, seen, cx.abort_flag, A.error, PIPE_ERR_GATE);
            } else if (lane == 9u && B == 0 && NB > 1 && z >= A.lead) {
                const unsigned long long target = A.base + (unsigned long long)(z - A.lead) + 1ull;
                if (seen < target) wait_counter<false>(A.prog + (size_t)(NB - 1) * A.tiles + cx.tile, target, seen, cx.abort_flag, A.error, PIPE_ERR_GATE);
            } else if (lane == 10u && z >= BP_STAGES) {
                wait_bar(cx.freed + ps.slot, ps.parity ^ 1u, cx.abort_flag, A.error, PIPE_ERR_EMPTY);
            }
            __syncwarp();
        }
        asm volatile("bar.sync 1, %0;" ::"r"(cx.n_ct) : "memory");
        if (*cx.abort_flag) break;
        real* out = out_ring + (size_t)ps.slot * stage_elems;
        for (uint32_t item = threadIdx.x; item < items; item += cx.n_ct) {
            const uint32_t r = item / VX, x0 = (item - r * VX) * N;
            basis_item<real, UC, B, FLIP, RECORD>(A, z, cx.y0 + r, x0, out, r * A.nx + x0, out_comp, fs, accepted);
        }
        fence_proxy_async_smem();     // my shared-memory writes before the publisher's bulk store reads them
        __syncwarp();
        if (lane == 0) mbar_arrive(cx.done + ps.slot);
        ps.advance(BP_STAGES);
        if (RECORD && (z & 15u) == 15u) heis_flush(fs, cx.s_acc);
    }
    if (RECORD) heis_flush(fs, cx.s_acc);
    const int a = __reduce_add_sync(0xffffffffu, accepted);
    if (lane == 0 && a != 0) atomicAdd(&cx.s_acc[5], (double)a);
}

template <typename real, int UC, bool FLIP, bool RECORD, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) basis_pipe_kernel(const __grid_constant__ BasisPipeArgs<real> A) {
    constexpr int NB = BasisCell<UC>::NB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t b = blockIdx.x / A.tiles, tile = blockIdx.x - b * A.tiles;
    const uint32_t rows = A.rows;
    const uint32_t nr = tile < A.tiles_long ? rows : rows - 1;
    const uint32_t y0 = tile < A.tiles_long ? tile * rows : A.tiles_long * rows + (tile - A.tiles_long) * (rows - 1);
    const uint32_t stage_elems = 3 * rows * A.nx;
    real* out_ring = reinterpret_cast<real*>(smem_raw);
    uint64_t* done = reinterpret_cast<uint64_t*>(out_ring + (size_t)BP_STAGES * stage_elems);
    uint64_t* freed = done + BP_STAGES;
    double* s_acc = reinterpret_cast<double*>(freed + BP_STAGES);
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(s_acc + 6);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, n_cw = A.n_cw;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < BP_STAGES; ++s) { mbar_init(done + s, n_cw); mbar_init(freed + s, 1u); }
        *abort_flag = 0u;
        fence_barrier_init();
        fence_proxy_async();
    }
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;
    __syncthreads();

    if (warp == n_cw) {
        // ===================== publisher: bulk-stores the finished tiles and releases the band's progress =====================
        if (lane == 0) {
            unsigned long long* const my_prog = A.prog + (size_t)b * A.tiles + tile;
            const uint32_t bytes = nr * A.nx * (uint32_t)sizeof(real);
            RingPos pd;
            uint32_t since_pub = 0;
            for (uint32_t z = 0; z < A.nz; ++z) {
                if (!wait_bar(done + pd.slot, pd.parity, abort_flag, A.error, PIPE_ERR_FULL)) break;
                const real* src = out_ring + (size_t)pd.slot * stage_elems;
                const size_t goff = ((size_t)z * A.ny + y0) * A.nx;     // the band's rows of a plane are contiguous
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) bulk_store_1d(A.P.s[b][c] + goff, src + (size_t)c * rows * A.nx, bytes);
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(freed + pd.slot);
                pd.advance(BP_STAGES);
                // a plane takes far longer than its store: publishing it at once keeps the colour fronts (and with them the
                // working set that has to survive in L2) as close together as the dependencies allow
                if (++since_pub == A.pub_every || z + 1 == A.nz) {
                    since_pub = 0;
                    tma_store_wait<0>();
                    fence_proxy_async();
                    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(my_prog), "l"(A.base + (unsigned long long)(z + 1u)) : "memory");
                }
            }
        }
    } else {
        BandCtx cx{tile, y0, nr, n_cw * 32u, done, freed, abort_flag, s_acc};
        switch (b) {
            case 0: basis_march<real, UC, 0, FLIP, RECORD>(A, cx, out_ring, stage_elems); break;
            case 1: basis_march<real, UC, 1, FLIP, RECORD>(A, cx, out_ring, stage_elems); break;
            case 2: if (NB > 2) basis_march<real, UC, (NB > 2 ? 2 : 0), FLIP, RECORD>(A, cx, out_ring, stage_elems); break;
            default: if (NB > 3) basis_march<real, UC, (NB > 3 ? 3 : 0), FLIP, RECORD>(A, cx, out_ring, stage_elems); break;
        }
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
    uint32_t NB = 0, tiles = 0, rows = 0, tiles_long = 0, n_cw = 0, threads = 0, lead = 0, pub_every = 0;
    size_t smem = 0;
    unsigned long long* d_prog = nullptr;
    unsigned int* d_error = nullptr;
    unsigned long long launches = 0;
    std::string text;
};

namespace {

template <typename real, int UC, int MAXT>
const void* bp_kernel_ptr(bool flip, bool record) {
    if (flip) return record ? (const void*)basis_pipe_kernel<real, UC, true, true, MAXT> : (const void*)basis_pipe_kernel<real, UC, true, false, MAXT>;
    return record ? (const void*)basis_pipe_kernel<real, UC, false, true, MAXT> : (const void*)basis_pipe_kernel<real, UC, false, false, MAXT>;
}
template <typename real>
const void* bp_kernel(int uc, bool flip, bool record, uint32_t threads) {
    if (threads <= 576) return uc == 1 ? bp_kernel_ptr<real, 1, 576>(flip, record) : bp_kernel_ptr<real, 2, 576>(flip, record);
    return uc == 1 ? bp_kernel_ptr<real, 1, 1024>(flip, record) : bp_kernel_ptr<real, 2, 1024>(flip, record);
}
const void* bp_kernel_any(bool f64, int uc, bool flip, bool record, uint32_t threads) {
    return f64 ? bp_kernel<double>(uc, flip, record, threads) : bp_kernel<float>(uc, flip, record, threads);
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
    uint32_t tiles = d.tiles ? d.tiles : (uint32_t)sms / NB;
    tiles = std::max(1u, std::min(std::min(tiles, d.ny), (uint32_t)sms / NB));
    const uint32_t rows = (d.ny + tiles - 1) / tiles;
    tiles = (d.ny + rows - 1) / rows;
    st->tiles = tiles; st->rows = rows; st->tiles_long = d.ny - tiles * (rows - 1);
    // consumer threads: the band's items (16-byte vectors of a plane) in as few equal rounds as 992 threads allow
    const uint32_t items = rows * (d.nx / N);
    const uint32_t rounds = (items + 991u) / 992u;
    const uint32_t cthreads = std::max(32u, ((items + rounds - 1) / rounds + 31u) / 32u * 32u);
    st->n_cw = cthreads / 32; st->threads = cthreads + 32;
    st->smem = (size_t)BP_STAGES * 3 * rows * d.nx * sz + 2 * BP_STAGES * 8 + 6 * 8 + 16;
    if (st->smem > (size_t)smem_max) { why = "band does not fit in shared memory"; delete st; return nullptr; }
    for (int v = 0; v < 4; ++v) {
        const void* k = bp_kernel_any(d.f64, d.unitcell, v & 1, v & 2, st->threads);
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
    snprintf(buf, sizeof buf, "basis_pipe: %u colours x %u bands, %u rows/band, %u threads, lead %u planes, publish every %u, %zu B smem",
             NB, tiles, rows, st->threads, st->lead, st->pub_every, st->smem);
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
    A.tiles = st->tiles; A.rows = st->rows; A.tiles_long = st->tiles_long; A.n_cw = st->n_cw; A.lead = st->lead; A.pub_every = st->pub_every;
    A.prog = st->d_prog;
    A.base = st->launches * (unsigned long long)d.nz;
    A.error = st->d_error;
    A.p = p; A.sweep = sweep; A.pk = pk; A.obs = obs_row;
    const void* k = bp_kernel<real>(d.unitcell, flip, record, st->threads);
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
