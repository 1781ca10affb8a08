// basis_wave.cu -- K4w: the 2 (bcc) / 4 (fcc) colour passes of a periodic Heisenberg step as ONE persistent launch whose work
// items run in wave order.
//
// Replaces MetropolisIntegrator::step (src/integrator.rs:66-92; MetropolisFlipIntegrator :109-138 when FLIP) for
// HeisenbergSpin on `Lattice::bcc / fcc (..).expand(x, y, z)` (src/input.rs:296-322), compound energy of
// src/energy.rs:63-257.  Colour = basis index.
//
// Why.  As separate launches every colour pass re-reads its partner sublattices from DRAM: 62 B/attempt measured for fcc
// 384^3 against 24 B/attempt algorithmic (profiles/r01z_heis_basis_vec.*); the kernel already runs at 83 % of the copy
// bandwidth, so only fewer bytes help.  The band-persistent pipeline (basis_pipe.cu) has the ideal traffic but one CTA per
// SM and one work item per thread and plane: latency bound (profiles/r02/README.md).  Here the work item and the occupancy
// are those of heis_basis_vec_kernel (128 threads, 8 CTAs per SM); only the ORDER changes.
//
// A unit is one colour on one cell plane, cut into tiles of 128 * ipt work items.  Units are listed by time slot: in slot t
// colour b works on plane t - lag[b], lag[b] = max over the bonded lower colours a of (lag[a] + reach(b, a)) + L, where
// reach is the largest dz of the bonds b -> a (0 or 1 in the unit-cell tables: towards LOWER colours dz >= 0, towards higher
// ones dz <= 0).  Items (unit, tile) are dealt round robin to the co-resident CTAs (cooperative launch), in order: an item
// only waits for lower-numbered items, so the lowest unfinished item can always run.  Tile (b, z) starts once every
// bonded lower colour a is complete on the planes z .. z + reach(b, a) (completion counters per (colour, plane), release /
// acquire): that one rule covers the true dependencies (b sees the NEW spins of the lower colours, exactly as in
// colour-ordered launches) and, bonds being symmetric, the anti-dependencies (a has finished reading the OLD spins b
// overwrites).  With L slots between the colours the wait is normally over before it starts, and everything between the
// first colour's front and the last colour's back (about lag[NB-1] + 2 planes of every sublattice) stays in L2.
//
// z-slab (one process per GPU): the planes below / above the slab are halo planes the neighbours store into directly
// (basis_vec_item<SLAB>).  The last tile of (a, plane 0) tells the lower neighbour, the last tile of (c, top plane) the
// upper one (st.release.sys on a flag word in peer memory); (b, top plane) waits for the upper neighbour's plane 0 of the
// lower colours of THIS step, (b, plane 0) for the lower neighbour's top planes of the PREVIOUS step.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "basis_wave.hpp"
#include "pipe_ptx.cuh"

namespace vg {

namespace {

template <typename real>
struct BasisWaveArgs {
    BasisPtrs<real> P;
    BasisPeers<real> peers;
    BasisGeom g;
    const uint32_t* units;          // [n_units] colour << 24 | plane, by time slot
    uint32_t n_units, tiles, ipt;
    uint32_t need[4];               // colour b: bit 2a + r set = b at plane z needs colour a complete on plane z + r
    unsigned long long* done;       // [NB][nz] tiles finished, monotone over the launches
    unsigned long long target;      // value of done[.][.] once a unit is complete in THIS launch
    unsigned long long steps;       // launches so far including this one (slab flags count steps)
    unsigned long long* ticket;     // next work item, monotone over the launches (every CTA draws one ticket past the end)
    unsigned long long ticket_base; // ticket of item 0 in THIS launch
    unsigned long long* flags;      // slab: my flag words (basis_wave.hpp)
    unsigned long long* peer_flags[2];
    unsigned int* error;
    HeisParams<real> p;
    uint64_t sweep;
    PhiloxKey pk;
    double* obs;
};

// Polls with RELAXED loads (an acquire load invalidates the SM's L1 every time: 63 M CCTL.IVALL per step in the first version,
// profiles/r02/README.md); the caller fences once after the wait.
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p, bool sys) {
    unsigned long long v;
    if (sys) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <bool SYS>
__device__ __forceinline__ bool wave_wait(const unsigned long long* p, unsigned long long target, const unsigned int* gerr) {
    if (ld_relaxed(p, SYS) >= target) return true;
    const unsigned long long t0 = global_timer();
    uint32_t n = 0;
    while (ld_relaxed(p, SYS) < target) {
        __nanosleep(128);
        if ((++n & 255u) == 0) {
            if (*(volatile const unsigned int*)gerr != 0u) return false;           // another CTA gave up: so do I
            if (global_timer() - t0 > PIPE_TIMEOUT_NS) return false;
        }
    }
    return true;
}

template <typename real, int UC, int B, bool FLIP, bool RECORD, bool SLAB>
__device__ __forceinline__ void wave_tile(const BasisWaveArgs<real>& A, uint32_t z, uint32_t tile, real (&fs)[5], int& accepted) {
    constexpr int N = VecOf<real>::N;
    const uint32_t nz = A.g.nz, VX = A.g.nx / N, items = VX * A.g.ny;
    const int zs[3] = {SLAB ? (int)z - 1 : (int)(z == 0 ? nz - 1 : z - 1), (int)z, SLAB ? (int)z + 1 : (int)(z + 1 == nz ? 0u : z + 1)};
    const uint32_t w0 = tile * (A.ipt * blockDim.x) + threadIdx.x;
    for (uint32_t it = 0; it < A.ipt; ++it) {
        const uint32_t w = w0 + it * blockDim.x;
        if (w >= items) break;
        basis_vec_item<real, UC, B, FLIP, RECORD ? 1 : 0, SLAB>(A.P, A.P, A.peers, A.g, z, zs, w, VX, A.p, A.sweep, A.pk, fs, accepted);
    }
}

template <typename real, int UC, bool FLIP, bool RECORD, bool SLAB>
__global__ void __launch_bounds__(128, BASIS_VEC_MINB) basis_wave_kernel(const __grid_constant__ BasisWaveArgs<real> A) {
    constexpr int NB = BasisCell<UC>::NB;
    __shared__ double s_acc[6];
    __shared__ uint32_t s_abort;
    __shared__ unsigned long long s_item[2];
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) { s_abort = 0u; s_item[0] = atomicAdd(A.ticket, 1ull) - A.ticket_base; }
    __syncthreads();
    real fs[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    uint32_t since_flush = 0;
    const uint32_t nz = A.g.nz, n_items = A.n_units * A.tiles;
    auto flush = [&]() {
        if (RECORD) heis_flush(fs, s_acc);
        const int a = __reduce_add_sync(0xffffffffu, accepted);
        if ((threadIdx.x & 31u) == 0 && a != 0) atomicAdd(&s_acc[5], (double)a);
        accepted = 0;
        since_flush = 0;
    };
    // Items are drawn from a ticket counter, in order: the colour-0 items fetch every sublattice from DRAM and take several
    // times longer than the others (static round robin left three quarters of the CTAs waiting for them).  An item still only
    // depends on lower-numbered items, all of which are held by running CTAs.  The next ticket is drawn while this item works.
    for (uint32_t it = 0;; ++it) {
        const unsigned long long drawn = s_item[it & 1u];
        if (drawn >= (unsigned long long)n_items) break;
        if (threadIdx.x == 0) s_item[(it + 1u) & 1u] = atomicAdd(A.ticket, 1ull) - A.ticket_base;
        const uint32_t item = (uint32_t)drawn;
        const uint32_t u = item / A.tiles, tile = item - u * A.tiles;
        const uint32_t unit = A.units[u];
        const uint32_t b = unit >> 24, z = unit & 0x00FFFFFFu;
        // ---- dependencies: one thread per (lower colour, plane offset); slab: the neighbours' boundary planes
        const bool slab_wait = SLAB && (z == 0 || z + 1 == nz);
        if (b > 0 || slab_wait) {
            bool ok = true;
            if (threadIdx.x < 2u * b) {
                const uint32_t a = threadIdx.x >> 1, r = threadIdx.x & 1u;
                const uint32_t mask = b == 1 ? A.need[1] : (b == 2 ? A.need[2] : A.need[3]);   // static indices: no local copy of A
                if ((mask >> threadIdx.x) & 1u) {
                    if (SLAB && z + r == nz) ok = wave_wait<true>(A.flags + a, A.steps, A.error);   // the upper neighbour's plane 0 of colour a, this step
                    else ok = wave_wait<false>(A.done + (size_t)a * nz + (z + r == nz ? 0u : z + r), A.target, A.error);
                }
            } else if (SLAB && z == 0 && threadIdx.x >= 8u && threadIdx.x < 8u + NB) {
                // the lower neighbour's top planes of the PREVIOUS step are in my lower halo (higher colours read at dz = -1)
                ok = wave_wait<true>(A.flags + 4 + (threadIdx.x - 8u), A.steps - 1ull, A.error);
            }
            if (!ok) { atomicExch(A.error, (unsigned int)PIPE_ERR_GATE); s_abort = 1u; }
            if (threadIdx.x < 12u) { if (SLAB) __threadfence_system(); else __threadfence(); }
            __syncthreads();
            if (s_abort) break;
        }
        switch (b) {
            case 0: wave_tile<real, UC, 0, FLIP, RECORD, SLAB>(A, z, tile, fs, accepted); break;
            case 1: wave_tile<real, UC, 1, FLIP, RECORD, SLAB>(A, z, tile, fs, accepted); break;
            case 2: if (NB > 2) wave_tile<real, UC, (NB > 2 ? 2 : 0), FLIP, RECORD, SLAB>(A, z, tile, fs, accepted); break;
            default: if (NB > 3) wave_tile<real, UC, (NB > 3 ? 3 : 0), FLIP, RECORD, SLAB>(A, z, tile, fs, accepted); break;
        }
        // ---- completion: the last colour has no dependants inside the launch (a slab still tells its neighbours)
        const bool boundary = SLAB && (z == 0 || z + 1 == nz);
        __syncthreads();   // every thread's stores of this tile are issued; thread 0's next ticket is in s_item
        if (b + 1 < (uint32_t)NB || boundary) {
            if (threadIdx.x == 0) {
                if (boundary) __threadfence_system(); else __threadfence();   // boundary tiles also stored into peer memory
                const unsigned long long c = atomicAdd(A.done + (size_t)b * nz + z, 1ull) + 1ull;
                if (boundary && c == A.target) {
                    // last tile of a boundary unit: every tile's peer stores are ordered before its count; tell the neighbour
                    __threadfence_system();
                    unsigned long long* f = z == 0 ? A.peer_flags[0] + b : A.peer_flags[1] + 4 + b;
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(A.steps) : "memory");
                }
            }
        }
        if (++since_flush == 8) flush();
    }
    if (since_flush) flush();
    __syncthreads();
    // layout of the basis kernels' observable row: [0] = sum_i sum_j J s_i.s_j with every bond twice (here: twice the bonds
    // towards the lower colours), [1..3] = sum s, [4] = sum (s.a)^2, [5] = accepted
    if (threadIdx.x < 6 && s_acc[threadIdx.x] != 0.0)
        atomicAdd(A.obs + threadIdx.x, threadIdx.x == 0 ? 2.0 * (double)A.p.J * s_acc[0] : s_acc[threadIdx.x]);
}

// dz of the bonds between colour b and colour a, seen from b, as a bit set: bit (dz + 1)
template <int UC>
uint32_t bond_dz_set(int b, int a) {
    uint32_t m = 0;
    for (int e = 0; e < BasisCell<UC>::NE; ++e) {
        int s = 0, t = 0, dx = 0, dy = 0, dz = 0;
        BasisCell<UC>::edge(e, s, t, dx, dy, dz);
        if (s == b && t == a) m |= 1u << (dz + 1);
        if (t == b && s == a) m |= 1u << (-dz + 1);
    }
    return m;
}
uint32_t bond_dz_set_any(int uc, int b, int a) { return uc == 1 ? bond_dz_set<1>(b, a) : bond_dz_set<2>(b, a); }

template <typename real, int UC>
const void* bw_kernel_ptr(bool flip, bool record, bool slab) {
#define BW(F, R) (slab ? (const void*)basis_wave_kernel<real, UC, F, R, true> : (const void*)basis_wave_kernel<real, UC, F, R, false>)
    if (flip) return record ? BW(true, true) : BW(true, false);
    return record ? BW(false, true) : BW(false, false);
#undef BW
}
template <typename real>
const void* bw_kernel(int uc, bool flip, bool record, bool slab) {
    return uc == 1 ? bw_kernel_ptr<real, 1>(flip, record, slab) : bw_kernel_ptr<real, 2>(flip, record, slab);
}
const void* bw_kernel_any(bool f64, int uc, bool flip, bool record, bool slab) {
    return f64 ? bw_kernel<double>(uc, flip, record, slab) : bw_kernel<float>(uc, flip, record, slab);
}

}  // namespace

// Unit order of one step.  need[b]: bit 2a + r = colour b at plane z waits for colour a on plane z + r.  Empty when the
// unit-cell table does not have the structure the scheme relies on (lower colours at dz >= 0 only).
std::vector<uint32_t> basis_wave_units(int unitcell, uint32_t nz, uint32_t L, uint32_t (&need)[4]) {
    const int NB = unitcell == 1 ? 2 : 4;
    uint32_t lag[4] = {0, 0, 0, 0};
    for (int b = 0; b < 4; ++b) need[b] = 0;
    for (int b = 1; b < NB; ++b) {
        uint32_t lg = lag[b - 1];
        for (int a = 0; a < b; ++a) {
            const uint32_t set = bond_dz_set_any(unitcell, b, a);
            if (set & 1u) return {};                     // a bond towards a lower colour at dz = -1: not this scheme
            if (set & 2u) need[b] |= 1u << (2 * a);
            if (set & 4u) need[b] |= 1u << (2 * a + 1);
            if (set) lg = std::max(lg, lag[a] + ((set & 4u) ? 1u : 0u) + L);
        }
        lag[b] = lg;
    }
    std::vector<uint32_t> units;
    if (nz < lag[NB - 1] + 2) return units;              // the wrap-around dependencies must lie far in the past
    units.reserve((size_t)NB * nz);
    for (uint32_t t = 0; t < nz + lag[NB - 1]; ++t)
        for (int b = 0; b < NB; ++b)
            if (t >= lag[b] && t - lag[b] < nz) units.push_back((uint32_t)b << 24 | (t - lag[b]));
    return units;
}

struct BasisWaveState {
    BasisWaveDesc d;
    uint32_t NB = 0, tiles = 0, ipt = 0, lag = 0, n_units = 0, need[4] = {};
    int grid = 0;
    uint32_t* d_units = nullptr;
    unsigned long long* d_done = nullptr;    // [NB * nz] completion counters, then the ticket counter
    unsigned int* d_error = nullptr;
    unsigned long long launches = 0;
    bool broken = false;
    std::string text;
};

BasisWaveState* basis_wave_create(const BasisWaveDesc& d, std::string& why) {
    const size_t sz = d.f64 ? 8 : 4;
    const uint32_t N = (uint32_t)(16 / sz);
    if (d.unitcell != 1 && d.unitcell != 2) { why = "unit cell is neither bcc nor fcc"; return nullptr; }
    const uint32_t NB = d.unitcell == 1 ? 2u : 4u;
    if (d.nx % N || d.nx < N) { why = "needs nx a multiple of a 16-byte vector"; return nullptr; }
    int sms = 0, coop = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d.device) != cudaSuccess ||
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d.device) != cudaSuccess || !coop) {
        cudaGetLastError();
        why = "device attributes / cooperative launch unavailable";
        return nullptr;
    }
    BasisWaveState* st = new BasisWaveState();
    st->d = d; st->NB = NB;
    st->lag = std::max(1u, d.lag ? d.lag : 2u);
    const std::vector<uint32_t> units = basis_wave_units(d.unitcell, d.nz, st->lag, st->need);
    if (units.empty()) { why = "too few planes for the wave order"; delete st; return nullptr; }
    st->n_units = (uint32_t)units.size();
    // every variant must be co-resident on the grid used for all of them
    int per_sm = 1 << 30;
    for (int v = 0; v < 4; ++v) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bw_kernel_any(d.f64, d.unitcell, v & 1, v & 2, d.slab), 128, 0) != cudaSuccess || n < 1) {
            cudaGetLastError();
            why = "kernel cannot be made co-resident";
            delete st;
            return nullptr;
        }
        per_sm = std::min(per_sm, n);
    }
    st->grid = per_sm * sms;
    if (d.grid) st->grid = (int)std::min<uint32_t>(d.grid, (uint32_t)st->grid);
    // tile = 128 * ipt work items: about one time slot (every colour, one plane) per grid of co-resident CTAs, so that a
    // dependency (`lag` slots back) is complete before its user is drawn
    const uint64_t items = (uint64_t)(d.nx / N) * d.ny;
    uint32_t ipt = d.ipt;
    if (!ipt) ipt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(64 / N, items * NB / ((uint64_t)st->grid * 128)));
    st->ipt = ipt;
    st->tiles = (uint32_t)((items + 128ull * ipt - 1) / (128ull * ipt));
    if ((uint64_t)st->n_units * st->tiles >= (1ull << 32)) { why = "too many work items"; delete st; return nullptr; }
    if (cudaMalloc(&st->d_units, units.size() * 4) != cudaSuccess || cudaMalloc(&st->d_done, ((size_t)NB * d.nz + 1) * 8) != cudaSuccess ||
        cudaMalloc(&st->d_error, 4) != cudaSuccess) {
        cudaGetLastError();
        why = "cudaMalloc failed";
        basis_wave_destroy(st);
        return nullptr;
    }
    cudaMemcpy(st->d_units, units.data(), units.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(st->d_done, 0, ((size_t)NB * d.nz + 1) * 8);
    cudaMemset(st->d_error, 0, 4);
    char buf[256];
    snprintf(buf, sizeof buf, "basis_wave: %u colours x %u planes, %u tiles of %u x 128 items per unit, %u slots between colours, %d CTAs%s",
             NB, d.nz, st->tiles, ipt, st->lag, st->grid, d.slab ? ", slab" : "");
    st->text = buf;
    return st;
}

void basis_wave_destroy(BasisWaveState* st) {
    if (!st) return;
    cudaFree(st->d_units); cudaFree(st->d_done); cudaFree(st->d_error);
    delete st;
}

const char* basis_wave_describe(const BasisWaveState* st) { return st ? st->text.c_str() : ""; }

template <typename real>
int basis_wave_step(BasisWaveState* st, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                    double* obs_row, cudaStream_t stream, std::string& err) {
    const BasisWaveDesc& d = st->d;
    if (st->broken) { err = "basis_wave: disabled after a timed-out wait"; return -1; }
    BasisWaveArgs<real> A;
    memset(&A, 0, sizeof A);
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) A.P.s[b][c] = (real*)d.arr[b][c];
    A.peers.lo = (real*)d.peer_lo; A.peers.hi = (real*)d.peer_hi;
    A.g.nx = d.nx; A.g.ny = d.ny; A.g.nz = d.nz; A.g.ncells = d.nx * d.ny * d.nz;
    A.g.z_offset = d.z_offset; A.g.nz_global = d.nz_global ? d.nz_global : d.nz;
    A.g.ext = (size_t)(d.nz + 2) * d.ny * d.nx;
    A.units = st->d_units; A.n_units = st->n_units; A.tiles = st->tiles; A.ipt = st->ipt;
    for (int b = 0; b < 4; ++b) A.need[b] = st->need[b];
    A.done = st->d_done;
    A.steps = st->launches + 1;
    A.target = (unsigned long long)st->tiles * A.steps;
    const int grid = (int)std::min<uint64_t>((uint64_t)st->grid, (uint64_t)st->n_units * st->tiles);
    A.ticket = st->d_done + (size_t)st->NB * d.nz;
    A.ticket_base = st->launches * ((unsigned long long)st->n_units * st->tiles + (unsigned long long)grid);
    A.flags = d.flags; A.peer_flags[0] = d.peer_flags[0]; A.peer_flags[1] = d.peer_flags[1];
    A.error = st->d_error;
    A.p = p; A.sweep = sweep; A.pk = pk; A.obs = obs_row;
    const void* k = bw_kernel<real>(d.unitcell, flip, record, d.slab);
    void* args[] = {&A};
    const cudaError_t e = cudaLaunchCooperativeKernel(k, dim3(grid), dim3(128), args, 0, stream);
    if (e != cudaSuccess) {
        err = std::string("basis_wave_kernel launch failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return -1;
    }
    st->launches++;
    return 0;
}
template int basis_wave_step<float>(BasisWaveState*, const HeisParams<float>&, bool, bool, uint64_t, const PhiloxKey&, double*, cudaStream_t, std::string&);
template int basis_wave_step<double>(BasisWaveState*, const HeisParams<double>&, bool, bool, uint64_t, const PhiloxKey&, double*, cudaStream_t, std::string&);

unsigned long long basis_wave_steps_done(const BasisWaveState* st) { return st ? st->launches : 0; }

int basis_wave_check(BasisWaveState* st, std::string& err) {
    if (!st) return 0;
    unsigned int e = 0;
    if (cudaMemcpy(&e, st->d_error, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { err = "basis_wave: cannot read the error flag"; return -1; }
    if (e == 0) return 0;
    err = "basis_wave_kernel: a dependency wait timed out (results invalid)";
    cudaMemset(st->d_error, 0, 4);
    st->broken = true;
    return -1;
}

bool basis_wave_usable(const BasisWaveState* st) { return st && !st->broken; }

}  // namespace vg
