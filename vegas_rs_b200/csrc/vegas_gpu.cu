// vegas_gpu.cu -- C ABI (include/vegas_gpu.h) of the B200-native Metropolis sweep.
// Host logic only: handle, layout selection, thermostat tables, launch sequencing.
// Kernels live in ising_msc.cuh (K1), heis.cuh (K3), general.cuh (K2/K4/K5), resident.cuh (K2r), heis_basis.cuh (K4b).
#include "../../include/vegas_gpu.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <type_traits>
#include <vector>

#include "general.cuh"
#include "heis.cuh"
#include "heis_basis.cuh"
#include "heis_fused.cuh"
#include "heis_pipe.hpp"
#include "basis_pipe.hpp"
#include "basis_wave.hpp"
#include "host_pack.hpp"
#include "ising_msc.cuh"
#include "lattice.hpp"
#include "resident.cuh"

using namespace vg;

namespace {

enum Family { FAM_ISING_MSC = 0, FAM_HEIS_STENCIL = 1, FAM_ISING_GEN = 2, FAM_HEIS_GEN = 3, FAM_HEIS_BASIS = 4 };
const char* const FAMILY_NAME[5] = {"ising_msc", "heis_stencil", "ising_general", "heis_general", "heis_basis"};

constexpr uint64_t OBS_CAP = 4096;  // steps of observables kept on the device per batch
constexpr int OBS_W = 8;            // 8 x 8-byte slots per step
constexpr int BWAVE_FLAG_WORD = 16;  // words 16..23: boundary-plane step counters of the wave-ordered bcc / fcc step (basis_wave.cu)
constexpr int PIPE_FLAG_WORD = 8;   // words 8..11 of a slab's flag block: boundary-plane counters of the pipelined kernel (heis_pipe.cu)

std::string g_create_error;

}  // namespace

struct vegas_gpu {
    vegas_model_desc md{};
    int family = 0;
    bool structured = false, csr_input = false;
    vgl::Desc ld{};                 // local lattice (structured)
    uint64_t nz_global = 0, z_offset = 0;
    uint64_t n = 0;                 // local sites
    int n_colours = 0;
    int ndim = 3;                   // stencil dimensionality (2 when nz == 1)
    int n_self = 0;                 // periodic axes of extent 1 (self bond, constant energy)
    // thermostat (src/thermostat.rs:19-79)
    double T = 2.8, fdir[3] = {0, 0, 1}, fmag = 0.0;
    int econv = VEGAS_E_REFERENCE_COMPOUND;
    // device
    int device = 0;
    cudaStream_t stream = nullptr, stream_b = nullptr;   // sweep stream; boundary-plane stream of a connected slab
    cudaEvent_t ev_main = nullptr, ev_bnd = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // --- Ising MSC
    uint32_t* msc[2] = {nullptr, nullptr};  // colour arrays (bit-packed)
    MscThr<MSC_MAX_SLOT> msc_thr{};       // threshold bit-planes (0/1), MSB first; passed to the kernel by value
    MscSlots<MSC_MAX_SLOT> msc_slots{};
    int msc_nslot = 0;
    bool msc_field = false, msc_ferro = false;   // ferro: the slots are exactly count == q for q < Z/2 (J > 0, no field)
    std::vector<uint64_t> ising_thr;      // [2][8]
    std::vector<uint8_t> ising_always;    // [2][8]
    // --- Heisenberg stencil: [colour][component]
    void* hs[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    // --- periodic bcc / fcc Heisenberg (heis_basis.cuh): basis-split SoA [basis][component][cell]
    void* hb[4][3] = {};
    // --- fused two-colour step (heis_fused.cuh): second buffer set, tile geometry; hs <-> hs_alt swap every fused step
    void* hs_alt[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    int fused_enable = -1;                // -1 auto, 0 off, 1 on
    uint32_t fused_ty = 0, fused_cz = 0;  // 0 = auto
    uint32_t wave_c = 0;                  // experiment: interleave the two colour passes in chunks of wave_c planes
    // --- persistent wave kernel (heis_wave_kernel): both colour passes in one launch, L2-friendly order
    int wave_enable = -1;                 // tuning key heis_wave: -1 auto (lattices with >= 32 planes), 0 never, 1 always
    uint32_t wave_planes = 4, wave_lag = 5;
    bool wave_ready = false;
    WaveSched wave_sched{};
    uint32_t wave_k = 1;                  // tuning key heis_wave_steps: steps fused into one launch (<= WAVE_MAX_STEPS)
    uint32_t wave_kmax = 1;               // planned value
    uint32_t* wave_units[WAVE_MAX_STEPS] = {};      // unit order of a launch of k steps at index k - 1
    uint32_t wave_n_units[WAVE_MAX_STEPS] = {};
    unsigned long long* wave_done = nullptr;
    unsigned long long wave_phase_launches[2 * WAVE_MAX_STEPS] = {};  // launches that ran phase p so far
    unsigned int* wave_error = nullptr;
    int wave_grid = 0;
    // --- phase-pipelined TMA kernel (heis_pipe.cu): the default step of big 3-D sc Heisenberg lattices
    int pipe_enable = -1;                 // tuning key heis_pipe: -1 auto (>= 32 planes), 0 never, 1 whenever the lattice fits
    uint32_t pipe_stages_other = 0, pipe_stages_own = 0, pipe_tiles = 0, pipe_vec = 0, pipe_lead = 0, pipe_pub = 0, pipe_backoff_c = 0, pipe_backoff_h = 0, pipe_l2 = 1;   // tuning keys heis_pipe_stages / _own / _tiles / _vec (0 = auto)
    bool pipe_planned = false;
    uint64_t pipe_slab_steps = 0;         // pipelined steps since the slab was connected (the neighbours' boundary counters are relative to it)
    HeisPipeState* pipe = nullptr;
    std::string pipe_why;
    // --- phase-pipelined bcc / fcc step (basis_pipe.cu): all colour passes in one cooperative launch
    int bpipe_enable = -1;                // tuning key basis_pipe: 1 = whenever the lattice fits (single handle); -1 / 0: colour launches
    uint32_t bpipe_lead = 0, bpipe_pub = 0, bpipe_tiles = 0;
    bool bpipe_planned = false;
    BasisPipeState* bpipe = nullptr;
    // --- pair launches of the fcc step (heis_basis_pair_kernel): colours (0,1) and (2,3) in one launch each, S -> D arrays
    int bpair_enable = -1;                // tuning key basis_pair: -1 auto (single-handle fcc lattices larger than L2), 0 never, 1 whenever possible
    uint32_t bpair_rows = 0;              // tuning key basis_pair_rows: rows of a plane per CTA (0 = auto: 48; 24 / 32 / 48 / 96 rows: 2.18 / 2.21 / 2.15 / 2.27 ms)
    uint32_t bpair_chunk = 0;             // tuning key basis_pair_chunk: rows the two colours alternate in (0 = auto)
    void* hb2[4][3] = {};                 // the second set of arrays (allocated at first use); hb / hb2 swap after every pair step
    int bpair_ok = -1;                    // cached: the unit-cell table has the structure the pair kernel needs
    size_t pair_set_bytes = 0;            // fcc slab: its ONE allocation holds two array sets this many bytes apart (0: one set)
    int cur_set = 0;                      // which of the two sets hb points at (all ranks of a slab group swap in lock step)
    // --- wave-ordered bcc / fcc step (basis_wave.cu): the colour passes of a step as one persistent launch, L2-friendly order
    int bwave_enable = -1;                // tuning key basis_wave: 1 whenever possible; -1 / 0: colour launches
    uint32_t bwave_lag = 0, bwave_ipt = 0, bwave_grid = 0;   // tuning keys basis_wave_lag / _ipt / _grid (0 = auto)
    bool bwave_planned = false;
    BasisWaveState* bwave = nullptr;
    bool fused_ready = false;
    FusedGeom fused_geom{};
    size_t fused_smem = 0;
    // --- halos (slab decomposition): per colour lower/upper halo of the *other* ranks' planes
    void* halo = nullptr;                 // one allocation: [colour][lo/hi][comp] planes
    size_t halo_plane_bytes = 0;          // bytes of one colour plane (one component)
    size_t halo_bytes = 0, slab_bytes = 0; // halos, then flags at halo + halo_bytes; slab_bytes = whole allocation
    unsigned long long* flags = nullptr;  // [2]: pass counters written by lower / upper neighbour
    void* peer_halo[2] = {nullptr, nullptr};            // lower / upper neighbour's halo allocation
    unsigned long long* peer_flags[2] = {nullptr, nullptr};
    bool slab = false, connected = false, peer_is_ipc = false;
    unsigned int* slab_error = nullptr;   // set by a slab wait that timed out (neighbour missing)
    bool peers_remote = false;            // both z-neighbours sweep on other devices (other processes, or other GPUs of this one)
    uint64_t pass_counter = 0;
    // --- general family
    int8_t* g_s8 = nullptr;               // Ising natural order
    void* g_s[3] = {nullptr, nullptr, nullptr};  // Heisenberg SoA natural order
    std::vector<uint32_t*> g_sites;       // per colour site lists (device)
    std::vector<uint32_t> g_counts;
    unsigned long long* d_row_ptr = nullptr; uint32_t* d_col = nullptr; double* d_val = nullptr;
    unsigned long long* g_thr = nullptr; uint8_t* g_code = nullptr;
    std::vector<uint64_t> h_row_ptr; std::vector<uint32_t> h_col; std::vector<double> h_val;  // csr input copy
    std::vector<uint8_t> h_colour;
    // --- shared-memory-resident batches of steps for small general-family lattices (resident.cuh)
    // --- host-packed State transfer of big ising_msc lattices (host_pack.cpp): sign bitmap over PCIe instead of one byte per spin
    long host_pack_min = 1l << 22;        // tuning key "host_pack_min": fewest spins that take it (0 = always, -1 = never)
    uint64_t host_pack_chunk = 1ull << 26; // tuning key "host_pack_chunk": spins per pipelined chunk
    uint32_t* hp_host = nullptr;          // pinned staging bitmap (n / 8 bytes)
    uint32_t* hp_dev = nullptr;           // device bitmap
    std::vector<cudaEvent_t> hp_events;
    int msc_full = 1;                     // tuning key "msc_full": the exact-cover variant of the Ising colour pass when the grid allows
    int basis_vec = 1;                    // tuning key "basis_vec": 16-byte accesses in the bcc / fcc colour pass when nx allows
    uint32_t resident_max = 8192;         // tuning key "resident_max": largest site count that takes this path (0: never)
    int resident_cols = -2;               // cached resident_columns(): -2 not planned yet, -1 no, 0 direct, > 0 table columns
    // --- observables
    unsigned long long* obs = nullptr;    // [OBS_CAP + 2][OBS_W]; row OBS_CAP = scratch, OBS_CAP+1 = query
    // --- counters
    uint64_t sweeps = 0, attempts = 0, accepted = 0, launches = 0;
    uint64_t obs_unread = 0;              // rows of the last recorded batch whose accepted counts are not in `accepted` yet
    bool tables_dirty = true;
    std::string err;
};

namespace {

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            if (h) h->err = b__; else g_create_error = b__;                                        \
            return VEGAS_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

int fail(vegas_gpu* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }
inline size_t real_bytes(const vegas_gpu* h) { return h->md.precision == VEGAS_F64 ? 8 : 4; }

double h_abs(const vegas_gpu* h) { return h->md.has_zeeman ? std::fabs(h->fmag) : 0.0; }
// Ising: (orientation . up) * |H|   (IsingSpin::dot src/state.rs:95-101; orientation Up when dir[2] >= 0)
double ising_h_o(const vegas_gpu* h) { return h_abs(h) * (h->fdir[2] >= 0.0 ? 1.0 : -1.0); }

// ---------------------------------------------------------------------------------------
// Acceptance thresholds: accept iff U < thr, U uniform in [0, 2^64)  (src/integrator.rs:82-88, :128-134)
// ---------------------------------------------------------------------------------------
void threshold(double delta, double T, uint64_t& thr, uint8_t& code) {
    if (delta < 0.0) { code = 2; thr = ~0ull; return; }
    const double p = std::exp(-delta / T);
    if (!(p < 1.0)) { code = 2; thr = ~0ull; return; }      // u < 1 always holds
    const double scaled = std::floor(std::ldexp(p, 64));
    thr = scaled >= 18446744073709551615.0 ? ~0ull : (uint64_t)scaled;
    code = thr == 0 ? 0 : 1;
}

// dE of flipping a spin s whose neighbours sum to m (uniform J), reference arithmetic:
// e_old = -J*(s*m) + (s.o)|H|,  e_new = -e_old  (gauge and Ising anisotropy cancel).
double ising_delta(const vegas_gpu* h, int s, int m) {
    const double ex = h->md.has_exchange ? -h->md.exchange * (double)(s * m) : 0.0;
    const double ze = h->md.has_zeeman ? (double)s * ising_h_o(h) : 0.0;
    const double e_old = ex + ze, e_new = (-ex) + (-ze);
    return e_new - e_old;
}

int update_tables(vegas_gpu* h) {
    if (!h->tables_dirty) return VEGAS_OK;
    if (h->family == FAM_ISING_MSC) {
        const int Z = 2 * h->ndim;
        h->ising_thr.assign(16, 0); h->ising_always.assign(16, 0);
        std::vector<uint64_t> slots;
        h->msc_field = h->md.has_zeeman && std::fabs(h->fmag) != 0.0;
        MscSlots<MSC_MAX_SLOT>& ms = h->msc_slots;
        for (int q = 0; q < MSC_MAX_SLOT; ++q) { ms.a0[q] = ms.a1[q] = ms.a2[q] = 0u; ms.sx[q] = 0u; }  // count 7: never matches
        for (int sidx = (h->msc_field ? 0 : 1); sidx < 2; ++sidx)
            for (int c = 0; c <= Z; ++c) {
                const int s = sidx ? 1 : -1;
                const int m = s * (Z - 2 * c);  // s*m = Z - 2c
                uint64_t thr; uint8_t code;
                threshold(ising_delta(h, s, m), h->T, thr, code);
                h->ising_thr[sidx * 8 + c] = thr;
                h->ising_always[sidx * 8 + c] = code == 2;
                if (code == 2) continue;  // dE <= 0: no slot, always accepted
                const size_t q = slots.size();
                if ((int)q >= MSC_MAX_SLOT) return fail(h, VEGAS_ERR_INVALID, "internal: too many threshold slots");
                slots.push_back(thr);  // code 0 (p < 2^-64) is a slot whose threshold is 0: never accepted
                ms.a0[q] = (c & 1) ? 0u : ~0u; ms.a1[q] = (c & 2) ? 0u : ~0u; ms.a2[q] = (c & 4) ? 0u : ~0u;
                ms.sx[q] = sidx ? 0u : ~0u;
            }
        if (!h->msc_field) for (int c = 0; c < 8; ++c) { h->ising_thr[c] = h->ising_thr[8 + c]; h->ising_always[c] = h->ising_always[8 + c]; }
        h->msc_nslot = h->msc_field ? 14 : 3;
        h->msc_ferro = !h->msc_field && (int)slots.size() == Z / 2;
        for (int q = 0; q < Z / 2 && h->msc_ferro; ++q)
            h->msc_ferro = ms.a0[q] == ((q & 1) ? 0u : ~0u) && ms.a1[q] == ((q & 2) ? 0u : ~0u) && ms.a2[q] == ~0u;
        if ((int)slots.size() > h->msc_nslot) return fail(h, VEGAS_ERR_INVALID, "internal: too many threshold slots");
        memset(&h->msc_thr, 0, sizeof h->msc_thr);
        for (size_t q = 0; q < slots.size(); ++q)
            for (int j = 0; j < 64; ++j) h->msc_thr.bit[q][j] = (uint32_t)(slots[q] >> (63 - j) & 1ull);
    } else if (h->family == FAM_ISING_GEN) {
        const int W = 2 * ISING_ZMAX + 1;
        std::vector<uint64_t> thr((size_t)2 * W);
        std::vector<uint8_t> code((size_t)2 * W);
        for (int sidx = 0; sidx < 2; ++sidx)
            for (int m = -ISING_ZMAX; m <= ISING_ZMAX; ++m)
                threshold(ising_delta(h, sidx ? 1 : -1, m), h->T, thr[sidx * W + m + ISING_ZMAX], code[sidx * W + m + ISING_ZMAX]);
        CU(cudaMemcpyAsync(h->g_thr, thr.data(), thr.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->g_code, code.data(), code.size(), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    h->tables_dirty = false;
    return VEGAS_OK;
}

template <typename real>
HeisParams<real> heis_params(const vegas_gpu* h) {
    HeisParams<real> p;
    p.J = (real)(h->md.has_exchange ? h->md.exchange : 0.0);
    const double ha = h_abs(h);
    for (int c = 0; c < 3; ++c) {
        p.h[c] = (real)(ha * h->fdir[c]);
        p.a[c] = (real)(h->md.has_anisotropy ? h->md.anisotropy_axis[c] : 0.0);
    }
    p.k = (real)(h->md.has_anisotropy ? h->md.anisotropy_k : 0.0);
    p.invT = (real)(1.0 / h->T);
    p.invTl = (real)(1.4426950408889634 / h->T);
    return p;
}

EnergyParams energy_params(const vegas_gpu* h) {
    EnergyParams ep{};
    const double ha = h_abs(h);
    const bool ising = h->md.model == VEGAS_ISING;
    for (int c = 0; c < 3; ++c) {
        ep.h[c] = ising ? 0.0 : ha * h->fdir[c];
        ep.a[c] = ising ? 0.0 : h->md.anisotropy_axis[c];
    }
    if (ising) { ep.h[2] = ising_h_o(h); ep.a[2] = h->md.anisotropy_axis[2] >= 0.0 ? 1.0 : -1.0; }
    ep.k = h->md.anisotropy_k; ep.gauge = h->md.gauge;
    ep.has_exchange = h->md.has_exchange; ep.has_zeeman = h->md.has_zeeman;
    ep.has_aniso = h->md.has_anisotropy; ep.has_gauge = h->md.has_gauge;
    return ep;
}

StructuredNb structured_nb(const vegas_gpu* h) {
    StructuredNb nb{};
    nb.nx = (uint32_t)h->ld.nx; nb.ny = (uint32_t)h->ld.ny; nb.nz = (uint32_t)h->ld.nz;
    nb.nb = vgl::basis_count(h->ld.unitcell);
    for (int a = 0; a < 3; ++a) nb.pbc[a] = h->ld.pbc[a];
    nb.literal = h->ld.literal;
    nb.J = h->md.has_exchange ? h->md.exchange : 0.0;
    for (int b = 0; b < 4; ++b) nb.count[b] = 0;
    for (const vgl::UcEdge& e : vgl::unitcell_edges(h->ld.unitcell)) {
        NbEntry f{}; f.tb = (int8_t)e.t; f.dx = (int8_t)e.dx; f.dy = (int8_t)e.dy; f.dz = (int8_t)e.dz; f.fwd = 1;
        nb.e[e.s][nb.count[e.s]++] = f;
    }
    for (const vgl::UcEdge& e : vgl::unitcell_edges(h->ld.unitcell)) {
        NbEntry r{}; r.tb = (int8_t)e.s; r.dx = (int8_t)-e.dx; r.dy = (int8_t)-e.dy; r.dz = (int8_t)-e.dz; r.fwd = 0;
        nb.e[e.t][nb.count[e.t]++] = r;
    }
    return nb;
}

CsrNb csr_nb(const vegas_gpu* h) {
    CsrNb nb{};
    nb.row_ptr = h->d_row_ptr; nb.col = h->d_col; nb.val = h->d_val;
    nb.J = h->md.has_exchange ? h->md.exchange : 0.0;
    return nb;
}

// ---------------------------------------------------------------------------------------
// stencil geometry helpers
// ---------------------------------------------------------------------------------------
MscGeom msc_geom(const vegas_gpu* h) {
    MscGeom g{};
    g.Wx = (uint32_t)(h->ld.nx / 64); g.Ly = (uint32_t)h->ld.ny; g.Lz = (uint32_t)h->ld.nz;
    g.z_offset = (uint32_t)h->z_offset;
    return g;
}
size_t msc_words(const vegas_gpu* h) { return (size_t)(h->ld.nx / 64) * h->ld.ny * h->ld.nz; }

HeisGeom heis_geom(const vegas_gpu* h) {
    HeisGeom g{};
    const int N = h->md.precision == VEGAS_F64 ? 2 : 4;
    g.Hx = (uint32_t)(h->ld.nx / 2); g.Gx = g.Hx / N; g.Ly = (uint32_t)h->ld.ny; g.Lz = (uint32_t)h->ld.nz;
    g.z_offset = (uint32_t)h->z_offset; g.Lx = (uint32_t)h->ld.nx;
    return g;
}
size_t heis_colour_elems(const vegas_gpu* h) { return (size_t)(h->ld.nx / 2) * h->ld.ny * h->ld.nz; }

// halo layout inside h->halo: [colour 0/1][lo 0 / hi 1][component] planes of halo_plane_bytes
size_t halo_offset(const vegas_gpu* h, int colour, int hi, int comp) {
    const int ncomp = h->family == FAM_ISING_MSC ? 1 : 3;
    return (((size_t)colour * 2 + hi) * ncomp + comp) * h->halo_plane_bytes;
}

// ---------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------
template <int NSLOT>
MscThr<NSLOT> thr_prefix(const MscThr<MSC_MAX_SLOT>& a) {
    MscThr<NSLOT> r;
    memcpy(&r, &a, sizeof r);  // bit[q][64] rows are contiguous: the first NSLOT slots
    return r;
}

template <int NSLOT>
MscSlots<NSLOT> slots_prefix(const MscSlots<MSC_MAX_SLOT>& a) {
    MscSlots<NSLOT> r;
    for (int q = 0; q < NSLOT; ++q) { r.a0[q] = a.a0[q]; r.a1[q] = a.a1[q]; r.a2[q] = a.a2[q]; r.sx[q] = a.sx[q]; }
    return r;
}

template <int NDIM, bool FIELD, int NSLOT, bool RP, bool FERRO, bool HALO, bool FULL = false>
void launch_msc_mode(vegas_gpu* h, int mode, dim3 grid, dim3 block, uint32_t* own, const uint32_t* oth, const uint32_t* lo,
                     const uint32_t* hi, uint32_t* plo, uint32_t* phi, int colour, uint32_t zb, uint32_t zstep,
                     unsigned long long* obs, cudaStream_t st) {
    const MscGeom g = msc_geom(h);
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const MscSlots<NSLOT> sl = slots_prefix<NSLOT>(h->msc_slots);
    const MscThr<NSLOT> th = thr_prefix<NSLOT>(h->msc_thr);
    if (mode == 0)
        ising_msc_kernel<NDIM, FIELD, NSLOT, RP, 0, FERRO, HALO, FULL><<<grid, block, 0, st>>>(own, oth, lo, hi, plo, phi, g, colour, zb, zstep,
                                                                                      sl, th, h->sweeps, pk, obs);
    else
        ising_msc_kernel<NDIM, FIELD, NSLOT, RP, 1, FERRO, HALO, FULL><<<grid, block, 0, st>>>(own, oth, lo, hi, plo, phi, g, colour, zb, zstep,
                                                                                      sl, th, h->sweeps, pk, obs);
}

template <int NDIM>
void launch_msc(vegas_gpu* h, int mode, int colour, uint32_t zb, uint32_t zc, uint32_t zstep, const uint32_t* lo, const uint32_t* hi,
                uint32_t* plo, uint32_t* phi, unsigned long long* obs, cudaStream_t st) {
    const MscGeom g = msc_geom(h);
    uint32_t bx = 32;
    while (bx > g.Wx) bx >>= 1;  // block = (bx words) x (256/bx rows); one warp spans 32/bx rows
    const dim3 block(bx, 256 / bx);
    const dim3 grid(cdiv(g.Wx, block.x), cdiv(g.Ly, block.y * msc_rows(NDIM)), zc);
    uint32_t* own = h->msc[colour];
    const uint32_t* oth = h->msc[1 - colour];
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool rp = h->md.proposal == VEGAS_PROPOSE_RANDOM;
    // halo pointers matter only for launches that touch local plane 0 or Lz-1 of a connected slab; everything else
    // (single handle: periodic wrap of `oth` itself; interior planes of a slab) takes the leaner variant
    const bool halo = h->slab && h->connected && (zb == 0 || zb + (zc - 1) * zstep + 1 >= g.Lz);
    const bool full = h->msc_full != 0 && g.Wx % block.x == 0 && g.Ly % (block.y * msc_rows(NDIM)) == 0;
    h->launches++;
    if (mode == 2) {
        ising_msc_kernel<NDIM, false, 3, false, 2><<<grid, block, 0, st>>>(own, oth, lo, hi, plo, phi, g, colour, zb, zstep,
                                                                       slots_prefix<3>(h->msc_slots), thr_prefix<3>(h->msc_thr), h->sweeps, pk, obs);
        return;
    }
#define ML(FIELD, NSLOT, RP, FERRO, HALO) launch_msc_mode<NDIM, FIELD, NSLOT, RP, FERRO, HALO>(h, mode, grid, block, own, oth, lo, hi, plo, phi, colour, zb, zstep, obs, st)
    if (h->msc_field) {
        if (rp) ML(true, 14, true, false, true); else ML(true, 14, false, false, true);
    } else if (h->msc_ferro) {
        if (halo) { if (rp) ML(false, 3, true, true, true); else ML(false, 3, false, true, true); }
        else if (full) {   // the grid covers the lattice exactly: the variant without per-row bounds and predicates
            if (rp) launch_msc_mode<NDIM, false, 3, true, true, false, true>(h, mode, grid, block, own, oth, lo, hi, plo, phi, colour, zb, zstep, obs, st);
            else launch_msc_mode<NDIM, false, 3, false, true, false, true>(h, mode, grid, block, own, oth, lo, hi, plo, phi, colour, zb, zstep, obs, st);
        } else { if (rp) ML(false, 3, true, true, false); else ML(false, 3, false, true, false); }
    } else {
        if (rp) ML(false, 3, true, false, true); else ML(false, 3, false, false, true);
    }
#undef ML
}

template <typename real, int NDIM>
void launch_heis(vegas_gpu* h, int mode, int colour, uint32_t zb, uint32_t zc, uint32_t zstep, const HeisPtrs<real>& P, double* obs,
                 cudaStream_t st) {
    const HeisGeom g = heis_geom(h);
    // grid.x tiles one plane, grid.y splits the z range into chunks a thread marches through; aim at ~16 CTAs per SM
    const uint32_t per_plane = cdiv((uint64_t)g.Ly * g.Gx, 128);
    uint32_t chunks = std::max<uint32_t>(1, (148u * 16u + per_plane - 1) / per_plane);
    chunks = std::min(chunks, zc);
    uint32_t z_chunk = (zc + chunks - 1) / chunks, z_stride = z_chunk;
    dim3 grid(per_plane, (zc + z_chunk - 1) / z_chunk);
    if (zstep > 1) { z_chunk = 1; z_stride = zstep; grid.y = zc; zc = (zc - 1) * zstep + 1; }  // zc single planes, zstep apart
    const HeisParams<real> p = heis_params<real>(h);
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    h->launches++;
    // halo / peer pointers only matter for launches that touch local plane 0 or Lz-1 of a connected slab
    const bool halo = h->slab && h->connected && (zb == 0 || zb + zc >= g.Lz);
#define HL(FLIP, MODE)                                                                                                             \
    do {                                                                                                                           \
        if (halo) heis_stencil_kernel<real, NDIM, FLIP, MODE, true><<<grid, 128, 0, st>>>(P, g, colour, zb, zc, z_chunk, z_stride, p, h->sweeps, pk, obs); \
        else heis_stencil_kernel<real, NDIM, FLIP, MODE, false><<<grid, 128, 0, st>>>(P, g, colour, zb, zc, z_chunk, z_stride, p, h->sweeps, pk, obs);    \
    } while (0)
    if (mode == 2) HL(false, 2);
    else if (mode == 1) { if (flip) HL(true, 1); else HL(false, 1); }
    else if (mode == 3) { if (flip) HL(true, 3); else HL(false, 3); }
    else { if (flip) HL(true, 0); else HL(false, 0); }
#undef HL
}

// One colour pass of a stencil family over local planes [zb, zb+zc).
template <typename real>
void heis_pass(vegas_gpu* h, int mode, int colour, uint32_t zb, uint32_t zc, double* obs, uint32_t zstep = 1, cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    HeisPtrs<real> P{};
    const size_t plane = (size_t)(h->ld.nx / 2) * h->ld.ny;
    for (int c = 0; c < 3; ++c) {
        P.own[c] = (real*)h->hs[colour][c];
        P.oth[c] = (const real*)h->hs[1 - colour][c];
        if (h->slab && h->connected) {
            P.oth_lo[c] = (const real*)((char*)h->halo + halo_offset(h, 1 - colour, 0, c));
            P.oth_hi[c] = (const real*)((char*)h->halo + halo_offset(h, 1 - colour, 1, c));
            // my plane 0 goes to the lower neighbour's *upper* halo of my colour, plane Lz-1 to the upper neighbour's lower halo
            P.peer_lo[c] = mode == 2 ? nullptr : (real*)((char*)h->peer_halo[0] + halo_offset(h, colour, 1, c));
            P.peer_hi[c] = mode == 2 ? nullptr : (real*)((char*)h->peer_halo[1] + halo_offset(h, colour, 0, c));
        } else {
            P.oth_lo[c] = P.oth[c] + (size_t)(h->ld.nz - 1) * plane;
            P.oth_hi[c] = P.oth[c];
            P.peer_lo[c] = nullptr; P.peer_hi[c] = nullptr;
        }
    }
    if (h->ndim == 3) launch_heis<real, 3>(h, mode, colour, zb, zc, zstep, P, obs, st);
    else launch_heis<real, 2>(h, mode, colour, zb, zc, zstep, P, obs, st);
}

void msc_pass(vegas_gpu* h, int mode, int colour, uint32_t zb, uint32_t zc, unsigned long long* obs, uint32_t zstep = 1,
              cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    const size_t plane = (size_t)(h->ld.nx / 64) * h->ld.ny;
    const uint32_t *lo, *hi;
    uint32_t *plo = nullptr, *phi = nullptr;
    if (h->slab && h->connected) {
        lo = (const uint32_t*)((char*)h->halo + halo_offset(h, 1 - colour, 0, 0));
        hi = (const uint32_t*)((char*)h->halo + halo_offset(h, 1 - colour, 1, 0));
        if (mode != 2) {
            plo = (uint32_t*)((char*)h->peer_halo[0] + halo_offset(h, colour, 1, 0));
            phi = (uint32_t*)((char*)h->peer_halo[1] + halo_offset(h, colour, 0, 0));
        }
    } else {
        lo = h->msc[1 - colour] + (size_t)(h->ld.nz - 1) * plane;
        hi = h->msc[1 - colour];
    }
    if (h->ndim == 3) launch_msc<3>(h, mode, colour, zb, zc, zstep, lo, hi, plo, phi, obs, st);
    else launch_msc<2>(h, mode, colour, zb, zc, zstep, lo, hi, plo, phi, obs, st);
}

// ---- slab flags: tiny kernels on the sweep stream (no host sync per colour) -------------
__global__ void signal_kernel(unsigned long long* lower_flag, unsigned long long* upper_flag, unsigned long long value) {
    __threadfence_system();
    if (threadIdx.x == 0) {
        // the lower neighbour reads what its upper neighbour (me) wrote in its flags[1]; vice versa
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(lower_flag + 1), "l"(value) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(upper_flag + 0), "l"(value) : "memory");
    }
}
// Bounded: a neighbour that died or ran a different number of passes must not hang this GPU for ever (the sweep then
// continues on stale halos and `*error` tells every entry point that hands out results).
__global__ void wait_kernel(const unsigned long long* flags, unsigned long long value, unsigned int* error) {
    if (threadIdx.x < 2) {
        unsigned long long v;
        uint32_t spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= value) break;
            if (++spins > (1u << 24)) { atomicExch(error, 1u); break; }   // ~10 s
            if (spins > 1024u) __nanosleep(512);
        }
    }
    __threadfence_system();
}

// CUDA loads kernels lazily, and loading one while a spinning wait_kernel is resident can block the launching host
// thread -- fatal when that thread is also the one that must enqueue the neighbour slab's work (all slabs in one
// process).  A connected slab therefore loads every kernel variant it can launch up front.
template <typename K>
void preload(K kernel) {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, kernel);
}
template <typename real>
void preload_heis_slab() {
    preload(heis_stencil_kernel<real, 3, false, 0, true>); preload(heis_stencil_kernel<real, 3, false, 0, false>);
    preload(heis_stencil_kernel<real, 3, false, 1, true>); preload(heis_stencil_kernel<real, 3, false, 1, false>);
    preload(heis_stencil_kernel<real, 3, false, 3, true>); preload(heis_stencil_kernel<real, 3, false, 3, false>);
    preload(heis_stencil_kernel<real, 3, true, 0, true>); preload(heis_stencil_kernel<real, 3, true, 0, false>);
    preload(heis_stencil_kernel<real, 3, true, 1, true>); preload(heis_stencil_kernel<real, 3, true, 1, false>);
    preload(heis_stencil_kernel<real, 3, true, 3, true>); preload(heis_stencil_kernel<real, 3, true, 3, false>);
    preload(heis_stencil_kernel<real, 3, false, 2, true>); preload(heis_stencil_kernel<real, 3, false, 2, false>);
}
template <typename real, int UC, int B>
void preload_basis_one() {
    preload(heis_basis_kernel<real, UC, B, false, 0, true>); preload(heis_basis_kernel<real, UC, B, false, 1, true>);
    preload(heis_basis_kernel<real, UC, B, true, 0, true>); preload(heis_basis_kernel<real, UC, B, true, 1, true>);
    preload(heis_basis_kernel<real, UC, B, false, 2, true>);
    preload(heis_basis_vec_kernel<real, UC, B, false, 0, true>); preload(heis_basis_vec_kernel<real, UC, B, false, 1, true>);
    preload(heis_basis_vec_kernel<real, UC, B, true, 0, true>); preload(heis_basis_vec_kernel<real, UC, B, true, 1, true>);
    preload(heis_basis_vec_kernel<real, UC, B, false, 2, true>);
}
template <typename real>
void preload_pair_slab() {
    preload(heis_basis_pair_kernel<real, 2, 0, false, 0, true>); preload(heis_basis_pair_kernel<real, 2, 0, false, 1, true>);
    preload(heis_basis_pair_kernel<real, 2, 0, true, 0, true>); preload(heis_basis_pair_kernel<real, 2, 0, true, 1, true>);
    preload(heis_basis_pair_kernel<real, 2, 2, false, 0, true>); preload(heis_basis_pair_kernel<real, 2, 2, false, 1, true>);
    preload(heis_basis_pair_kernel<real, 2, 2, true, 0, true>); preload(heis_basis_pair_kernel<real, 2, 2, true, 1, true>);
}
template <typename real>
void preload_basis_slab() {
    preload_pair_slab<real>();
    preload_basis_one<real, 1, 0>(); preload_basis_one<real, 1, 1>();
    preload_basis_one<real, 2, 0>(); preload_basis_one<real, 2, 1>(); preload_basis_one<real, 2, 2>(); preload_basis_one<real, 2, 3>();
}
template <bool RP>
void preload_msc_slab() {
    preload(ising_msc_kernel<3, true, 14, RP, 0, false, true>); preload(ising_msc_kernel<3, true, 14, RP, 1, false, true>);
    preload(ising_msc_kernel<3, false, 3, RP, 0, true, true>); preload(ising_msc_kernel<3, false, 3, RP, 1, true, true>);
    preload(ising_msc_kernel<3, false, 3, RP, 0, true, false>); preload(ising_msc_kernel<3, false, 3, RP, 1, true, false>);
    preload(ising_msc_kernel<3, false, 3, RP, 0, false, true>); preload(ising_msc_kernel<3, false, 3, RP, 1, false, true>);
}
__global__ void signal_kernel(unsigned long long*, unsigned long long*, unsigned long long);
__global__ void wait_kernel(const unsigned long long*, unsigned long long, unsigned int*);
__global__ void copy_plane_kernel(uint32_t*, const uint32_t*, size_t);
void preload_slab_kernels(vegas_gpu* h) {
    preload(signal_kernel); preload(wait_kernel); preload(copy_plane_kernel);
    if (h->family == FAM_ISING_MSC) {
        preload_msc_slab<false>(); preload_msc_slab<true>();
        preload(ising_msc_kernel<3, false, 3, false, 2>);
    } else if (h->family == FAM_HEIS_BASIS) {
        if (h->md.precision == VEGAS_F64) preload_basis_slab<double>(); else preload_basis_slab<float>();
    } else if (h->md.precision == VEGAS_F64) preload_heis_slab<double>();
    else preload_heis_slab<float>();
}

void stencil_colour_pass(vegas_gpu* h, int mode, int colour, void* obs_row) {
    const uint32_t Lz = (uint32_t)h->ld.nz;
    auto run = [&](uint32_t zb, uint32_t zc, uint32_t zstep, cudaStream_t st) {
        if (zc == 0) return;
        if (h->family == FAM_ISING_MSC) msc_pass(h, mode, colour, zb, zc, (unsigned long long*)obs_row, zstep, st);
        else if (h->md.precision == VEGAS_F64) heis_pass<double>(h, mode, colour, zb, zc, (double*)obs_row, zstep, st);
        else heis_pass<float>(h, mode, colour, zb, zc, (double*)obs_row, zstep, st);
    };
    if (h->slab && h->connected && mode != 2 && !h->peer_is_ipc) {
        // all slabs in one process (tests, or one process driving several GPUs): one stream, boundary planes first.
        // (A second stream per handle could alias hardware queues with another handle's spinning wait kernel.)
        h->pass_counter++;
        if (h->pass_counter > 1) { wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, h->pass_counter - 1, h->slab_error); h->launches++; }
        if (Lz >= 2) run(0, 2, Lz - 1, h->stream); else run(0, 1, 1, h->stream);
        signal_kernel<<<1, 32, 0, h->stream>>>(h->peer_flags[0], h->peer_flags[1], h->pass_counter);
        h->launches++;
        if (Lz > 2) run(1, Lz - 2, 1, h->stream);
    } else if (h->slab && h->connected && mode != 2) {
        // One process per GPU.  The two boundary planes (ONE launch) feed the neighbours: they run on the boundary stream -- wait for the
        // neighbours' previous pass, update, store into the peers' halos, signal -- while the interior planes run
        // on the main stream.  Both streams first wait for the whole previous pass (it wrote the other colour).
        h->pass_counter++;
        cudaEventRecord(h->ev_main, h->stream);
        cudaStreamWaitEvent(h->stream_b, h->ev_main, 0);
        cudaStreamWaitEvent(h->stream, h->ev_bnd, 0);
        if (h->pass_counter > 1) { wait_kernel<<<1, 32, 0, h->stream_b>>>(h->flags, h->pass_counter - 1, h->slab_error); h->launches++; }
        if (Lz >= 2) run(0, 2, Lz - 1, h->stream_b); else run(0, 1, 1, h->stream_b);
        signal_kernel<<<1, 32, 0, h->stream_b>>>(h->peer_flags[0], h->peer_flags[1], h->pass_counter);
        h->launches++;
        cudaEventRecord(h->ev_bnd, h->stream_b);
        if (Lz > 2) run(1, Lz - 2, 1, h->stream);
    } else {
        run(0, Lz, 1, h->stream);
    }
}

// the main stream joins the boundary stream (slabs): later main-stream work sees both
void join_boundary_stream(vegas_gpu* h) {
    if (h->slab && h->connected && h->peer_is_ipc && h->stream_b) cudaStreamWaitEvent(h->stream, h->ev_bnd, 0);
}

// ---- general family ---------------------------------------------------------------------
// bcc / fcc Heisenberg: the colour passes of a recorded step reduce the observables themselves
bool basis_fused_obs(const vegas_gpu* h) {
    return h->family == FAM_HEIS_GEN && !h->csr_input && vgl::basis_count(h->ld.unitcell) > 1;
}

void basis_obs_pass(vegas_gpu* h, int colour, double* obs_row) {
    const uint32_t count = h->g_counts[colour];
    if (count == 0) return;
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const StructuredNb nb = structured_nb(h);
    const dim3 grid(cdiv(count, 128));
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    h->launches++;
    if (h->md.precision == VEGAS_F64) {
        const HeisParams<double> p = heis_params<double>(h);
        if (flip) heis_basis_sweep_obs_kernel<double, true><<<grid, 128, 0, h->stream>>>((double*)h->g_s[0], (double*)h->g_s[1], (double*)h->g_s[2], nb, h->g_sites[colour], count, colour, p, 0, h->sweeps, pk, obs_row);
        else heis_basis_sweep_obs_kernel<double, false><<<grid, 128, 0, h->stream>>>((double*)h->g_s[0], (double*)h->g_s[1], (double*)h->g_s[2], nb, h->g_sites[colour], count, colour, p, 0, h->sweeps, pk, obs_row);
    } else {
        const HeisParams<float> p = heis_params<float>(h);
        if (flip) heis_basis_sweep_obs_kernel<float, true><<<grid, 128, 0, h->stream>>>((float*)h->g_s[0], (float*)h->g_s[1], (float*)h->g_s[2], nb, h->g_sites[colour], count, colour, p, 0, h->sweeps, pk, obs_row);
        else heis_basis_sweep_obs_kernel<float, false><<<grid, 128, 0, h->stream>>>((float*)h->g_s[0], (float*)h->g_s[1], (float*)h->g_s[2], nb, h->g_sites[colour], count, colour, p, 0, h->sweeps, pk, obs_row);
    }
}

template <typename NB>
void general_colour_pass(vegas_gpu* h, const NB& nb, int colour, unsigned long long* obs_row) {
    const uint32_t count = h->g_counts[colour];
    if (count == 0) return;
    const PhiloxKey pk = make_philox_key(h->md.seed);
    h->launches++;
    if (h->family == FAM_ISING_GEN) {
        IsingGeneralParams p{};
        p.thr = h->g_thr; p.code = h->g_code;
        p.uniform = (h->csr_input && h->d_val) ? 0 : 1;
        p.h_o = ising_h_o(h); p.invT = 1.0 / h->T;
        const dim3 grid(cdiv(count, 256));
        if (h->md.proposal == VEGAS_PROPOSE_RANDOM)
            ising_general_sweep_kernel<NB, true><<<grid, 256, 0, h->stream>>>(h->g_s8, nb, h->g_sites[colour], count, p, 0, h->sweeps, pk, obs_row + 4);
        else
            ising_general_sweep_kernel<NB, false><<<grid, 256, 0, h->stream>>>(h->g_s8, nb, h->g_sites[colour], count, p, 0, h->sweeps, pk, obs_row + 4);
    } else {
        const dim3 grid(cdiv(count, 128));
        const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
        if (h->md.precision == VEGAS_F64) {
            const HeisParams<double> p = heis_params<double>(h);
            if (flip) heis_general_sweep_kernel<NB, double, true><<<grid, 128, 0, h->stream>>>((double*)h->g_s[0], (double*)h->g_s[1], (double*)h->g_s[2], nb, h->g_sites[colour], count, p, 0, h->sweeps, pk, (double*)obs_row);
            else heis_general_sweep_kernel<NB, double, false><<<grid, 128, 0, h->stream>>>((double*)h->g_s[0], (double*)h->g_s[1], (double*)h->g_s[2], nb, h->g_sites[colour], count, p, 0, h->sweeps, pk, (double*)obs_row);
        } else {
            const HeisParams<float> p = heis_params<float>(h);
            if (flip) heis_general_sweep_kernel<NB, float, true><<<grid, 128, 0, h->stream>>>((float*)h->g_s[0], (float*)h->g_s[1], (float*)h->g_s[2], nb, h->g_sites[colour], count, p, 0, h->sweeps, pk, (double*)obs_row);
            else heis_general_sweep_kernel<NB, float, false><<<grid, 128, 0, h->stream>>>((float*)h->g_s[0], (float*)h->g_s[1], (float*)h->g_s[2], nb, h->g_sites[colour], count, p, 0, h->sweeps, pk, (double*)obs_row);
        }
    }
}

template <typename NB>
void general_reduce(vegas_gpu* h, const NB& nb, double* obs_row) {
    const uint32_t n = (uint32_t)h->n;
    const dim3 grid(std::min<uint32_t>(cdiv(n, 256), 148 * 8));
    double ax = 0, ay = 0, az = 1;
    if (h->md.model == VEGAS_HEISENBERG) { ax = h->md.anisotropy_axis[0]; ay = h->md.anisotropy_axis[1]; az = h->md.anisotropy_axis[2]; }
    h->launches++;
    if (h->family == FAM_ISING_GEN) {
        IsingSpins sp{h->g_s8};
        general_reduce_kernel<NB, IsingSpins><<<grid, 256, 0, h->stream>>>(nb, sp, n, ax, ay, az, obs_row);
    } else if (h->md.precision == VEGAS_F64) {
        HeisSpins<double> sp{(const double*)h->g_s[0], (const double*)h->g_s[1], (const double*)h->g_s[2]};
        general_reduce_kernel<NB, HeisSpins<double>><<<grid, 256, 0, h->stream>>>(nb, sp, n, ax, ay, az, obs_row);
    } else {
        HeisSpins<float> sp{(const float*)h->g_s[0], (const float*)h->g_s[1], (const float*)h->g_s[2]};
        general_reduce_kernel<NB, HeisSpins<float>><<<grid, 256, 0, h->stream>>>(nb, sp, n, ax, ay, az, obs_row);
    }
}

// ---- host-only self check: the compile-time neighbour tables of heis_basis.cuh against lattice.hpp -------------
template <int UC, int B>
int check_basis_table() {
    StructuredNb nb{};
    for (int b = 0; b < 4; ++b) nb.count[b] = 0;
    for (const vgl::UcEdge& e : vgl::unitcell_edges(UC)) { NbEntry f{}; f.tb = (int8_t)e.t; f.dx = (int8_t)e.dx; f.dy = (int8_t)e.dy; f.dz = (int8_t)e.dz; nb.e[e.s][nb.count[e.s]++] = f; }
    for (const vgl::UcEdge& e : vgl::unitcell_edges(UC)) { NbEntry r{}; r.tb = (int8_t)e.s; r.dx = (int8_t)-e.dx; r.dy = (int8_t)-e.dy; r.dz = (int8_t)-e.dz; nb.e[e.t][nb.count[e.t]++] = r; }
    if (nb.count[B] != BasisCell<UC>::Z) return 1;
    for (int q = 0; q < BasisCell<UC>::Z; ++q) {
        const BasisNb t = basis_neighbour<UC, B>(q);
        const NbEntry& w = nb.e[B][q];
        if (t.tb != w.tb || t.dx != w.dx || t.dy != w.dy || t.dz != w.dz) return 1;
    }
    return 0;
}

// ---- periodic bcc / fcc Heisenberg (heis_basis.cuh) ----------------------------------------------
template <typename real>
BasisPtrs<real> basis_ptrs(const vegas_gpu* h) {
    BasisPtrs<real> P{};
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) P.s[b][c] = (real*)h->hb[b][c];
    return P;
}
BasisGeom basis_geom(const vegas_gpu* h) {
    BasisGeom g{};
    g.nx = (uint32_t)h->ld.nx; g.ny = (uint32_t)h->ld.ny; g.nz = (uint32_t)h->ld.nz; g.ncells = g.nx * g.ny * g.nz;
    g.z_offset = (uint32_t)h->z_offset; g.nz_global = (uint32_t)h->nz_global;
    g.ext = (size_t)(g.nz + 2) * g.ny * g.nx;
    return g;
}

// base of array set `set` in the lower (0) / upper (1) neighbour's slab allocation
char* peer_set(const vegas_gpu* h, int which, int set) { return (char*)h->peer_halo[which] + (size_t)set * h->pair_set_bytes; }

template <typename real, int UC, int B>
void basis_launch(vegas_gpu* h, int mode, double* obs, uint32_t zb, uint32_t zc, uint32_t zstep, cudaStream_t st) {
    const BasisGeom g = basis_geom(h);
    // rows per CTA: enough CTAs to fill the GPU (~16 per SM), at most 64 rows (fp32 partial sums stay short)
    const uint64_t ctas_per_row_set = (uint64_t)cdiv(g.nx, 128) * zc;
    uint32_t rows = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, (uint64_t)g.ny * ctas_per_row_set / (148u * 16u)));
    const dim3 grid(cdiv(g.nx, 128), cdiv(g.ny, rows), zc);
    const HeisParams<real> p = heis_params<real>(h);
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const BasisPtrs<real> P = basis_ptrs<real>(h);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    const bool slab = h->slab && h->connected;
    BasisPeers<real> peers{};
    if (slab && mode != 2) { peers.lo = (real*)peer_set(h, 0, h->cur_set); peers.hi = (real*)peer_set(h, 1, h->cur_set); }
    h->launches++;
    constexpr uint32_t NV = (uint32_t)VecOf<real>::N;
    if (h->basis_vec != 0 && g.nx % NV == 0) {
        // K4v: 16-byte accesses; a thread walks `ipt` work items (NV cells each) of one plane: ~16 CTAs per SM, at most
        // 64 cells per thread (the fp32 partial sums stay short)
        const uint64_t items = (uint64_t)(g.nx / NV) * g.ny;
        const uint32_t ipt = (uint32_t)std::min<uint64_t>(64 / NV, std::max<uint64_t>(1, items * zc / 128 / (148u * 16u)));
        const dim3 vgrid(cdiv(items, 128 * ipt), 1, zc);
#define BV(FLIP, MODE)                                                                                                          \
    do {                                                                                                                        \
        if (slab) heis_basis_vec_kernel<real, UC, B, FLIP, MODE, true><<<vgrid, 128, 0, st>>>(P, peers, g, ipt, zb, zstep, p, h->sweeps, pk, obs); \
        else heis_basis_vec_kernel<real, UC, B, FLIP, MODE, false><<<vgrid, 128, 0, st>>>(P, peers, g, ipt, zb, zstep, p, h->sweeps, pk, obs);    \
    } while (0)
        if (mode == 2) BV(false, 2);
        else if (mode == 1) { if (flip) BV(true, 1); else BV(false, 1); }
        else { if (flip) BV(true, 0); else BV(false, 0); }
#undef BV
        return;
    }
#define BL(FLIP, MODE)                                                                                                          \
    do {                                                                                                                        \
        if (slab) heis_basis_kernel<real, UC, B, FLIP, MODE, true><<<grid, 128, 0, st>>>(P, peers, g, rows, zb, zstep, p, h->sweeps, pk, obs); \
        else heis_basis_kernel<real, UC, B, FLIP, MODE, false><<<grid, 128, 0, st>>>(P, peers, g, rows, zb, zstep, p, h->sweeps, pk, obs);    \
    } while (0)
    if (mode == 2) BL(false, 2);
    else if (mode == 1) { if (flip) BL(true, 1); else BL(false, 1); }
    else { if (flip) BL(true, 0); else BL(false, 0); }
#undef BL
}

// one colour (= basis) pass over planes zb, zb + zstep, ... (zc of them); mode as heis_basis_kernel
template <typename real>
void basis_pass(vegas_gpu* h, int mode, int b, double* obs, uint32_t zb, uint32_t zc, uint32_t zstep, cudaStream_t st) {
    if (zc == 0) return;
    if (h->ld.unitcell == VEGAS_BCC) {
        if (b == 0) basis_launch<real, 1, 0>(h, mode, obs, zb, zc, zstep, st); else basis_launch<real, 1, 1>(h, mode, obs, zb, zc, zstep, st);
    } else {
        switch (b) {
            case 0: basis_launch<real, 2, 0>(h, mode, obs, zb, zc, zstep, st); break;
            case 1: basis_launch<real, 2, 1>(h, mode, obs, zb, zc, zstep, st); break;
            case 2: basis_launch<real, 2, 2>(h, mode, obs, zb, zc, zstep, st); break;
            default: basis_launch<real, 2, 3>(h, mode, obs, zb, zc, zstep, st); break;
        }
    }
}
__global__ void signal_kernel(unsigned long long*, unsigned long long*, unsigned long long);
__global__ void wait_kernel(const unsigned long long*, unsigned long long, unsigned int*);
void basis_pass_any(vegas_gpu* h, int mode, int b, double* obs) {
    const uint32_t nz = (uint32_t)h->ld.nz;
    auto run = [&](uint32_t zb, uint32_t zc, uint32_t zstep, cudaStream_t st) {
        if (h->md.precision == VEGAS_F64) basis_pass<double>(h, mode, b, obs, zb, zc, zstep, st);
        else basis_pass<float>(h, mode, b, obs, zb, zc, zstep, st);
    };
    // connected slab: the pass reads halo planes the z-neighbours wrote in their previous pass and writes its own
    // boundary planes into theirs (same flag protocol and stream structure as the sc stencil kernels)
    const bool exchange = h->slab && h->connected && mode != 2;
    if (!exchange) { run(0, nz, 1, h->stream); return; }
    h->pass_counter++;
    if (!h->peer_is_ipc) {  // all slabs in one process: one stream (see stencil_colour_pass)
        if (h->pass_counter > 1) { wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, h->pass_counter - 1, h->slab_error); h->launches++; }
        if (nz >= 2) run(0, 2, nz - 1, h->stream); else run(0, 1, 1, h->stream);
        signal_kernel<<<1, 32, 0, h->stream>>>(h->peer_flags[0], h->peer_flags[1], h->pass_counter);
        h->launches++;
        if (nz > 2) run(1, nz - 2, 1, h->stream);
        return;
    }
    // one process per GPU: boundary planes (wait, update, peer stores, signal) on the boundary stream, interior on the main one
    cudaEventRecord(h->ev_main, h->stream);
    cudaStreamWaitEvent(h->stream_b, h->ev_main, 0);
    cudaStreamWaitEvent(h->stream, h->ev_bnd, 0);
    if (h->pass_counter > 1) { wait_kernel<<<1, 32, 0, h->stream_b>>>(h->flags, h->pass_counter - 1, h->slab_error); h->launches++; }
    if (nz >= 2) run(0, 2, nz - 1, h->stream_b); else run(0, 1, 1, h->stream_b);
    signal_kernel<<<1, 32, 0, h->stream_b>>>(h->peer_flags[0], h->peer_flags[1], h->pass_counter);
    h->launches++;
    cudaEventRecord(h->ev_bnd, h->stream_b);
    if (nz > 2) run(1, nz - 2, 1, h->stream);
}

// ---- fused two-colour Heisenberg step -------------------------------------------------------
// Chooses the tile (TY interior rows, full x) and the z-chunking; returns false when the lattice does not fit the
// fused kernel (the two-pass kernels then run).  Limits: 3-D, single handle (no slab), TY + 4 <= Ly, Ly % TY == 0,
// (TY + 4) * Lx/8 threads <= 768 and six plane slots of (TY + 4) rows in 227 KB of shared memory.
bool fused_plan(vegas_gpu* h) {
    if (h->fused_ready) return true;
    if (h->family != FAM_HEIS_STENCIL || h->ndim != 3 || h->slab || h->fused_enable != 1) return false;  // opt-in until it beats two passes
    const size_t rb = real_bytes(h);
    const uint32_t N = (uint32_t)(16 / rb);
    FusedGeom g{};
    g.Lx = (uint32_t)h->ld.nx; g.Ly = (uint32_t)h->ld.ny; g.Lz = (uint32_t)h->ld.nz;
    g.Hx = g.Lx / 2; g.Gx = g.Hx / N;
    g.z_offset = (uint32_t)h->z_offset; g.nz_global = (uint32_t)h->nz_global;
    if (g.Gx == 0 || g.Gx > HEIS_FUSED_THREADS / 5 || g.Lz < 2) return false;
    int smem_max = 0;
    if (cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess) return false;
    const size_t row_bytes = (size_t)6 * 3 * g.Hx * rb;  // six slots x three components
    uint32_t rows = std::min<uint32_t>(HEIS_FUSED_THREADS / g.Gx + 2, (uint32_t)(((size_t)smem_max - 1024) / row_bytes));
    rows = std::min(rows, g.Ly);
    if (rows < 5) return false;
    uint32_t ty = rows - 4;
    if (h->fused_ty) { if (h->fused_ty > ty) return false; ty = h->fused_ty; }
    while (ty > 1 && g.Ly % ty != 0) --ty;
    if (g.Ly % ty != 0) return false;
    if (ty < 4 && h->fused_enable != 1) return false;    // too much redundant halo work to pay off
    g.TY = ty; g.ROWS = ty + 4; g.tiles = g.Ly / ty;
    uint32_t chunks = cdiv(g.Lz, h->fused_cz ? h->fused_cz : g.Lz);
    if (!h->fused_cz)  // auto: about seven waves of one CTA per SM, chunks of at least four planes (two warm-up planes each)
        chunks = std::min(std::max<uint32_t>(1, (148u * 7u + g.tiles / 2) / g.tiles), std::max<uint32_t>(1, g.Lz / 4));
    g.CZ = cdiv(g.Lz, chunks); g.chunks = cdiv(g.Lz, g.CZ);
    const size_t bytes = heis_colour_elems(h) * rb;
    for (int c = 0; c < 2; ++c)
        for (int k = 0; k < 3; ++k)
            if (!h->hs_alt[c][k] && cudaMalloc(&h->hs_alt[c][k], bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    h->fused_geom = g;
    h->fused_smem = row_bytes * g.ROWS;
    h->fused_ready = true;
    return true;
}

template <typename real, int HX_T>
int fused_launch(vegas_gpu* h, const FusedPtrs<real>& P, double* obs_row, bool record) {
    const FusedGeom& g = h->fused_geom;
    const HeisParams<real> p = heis_params<real>(h);
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    const uint32_t threads = (g.Gx * (g.ROWS - 2) + 31u) / 32u * 32u;
    const dim3 grid(g.tiles * g.chunks);
#define FL(FLIP, REC)                                                                                                   \
    do {                                                                                                                \
        CU(cudaFuncSetAttribute(heis_fused_kernel<real, HX_T, FLIP, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fused_smem)); \
        heis_fused_kernel<real, HX_T, FLIP, REC><<<grid, threads, h->fused_smem, h->stream>>>(P, g, p, h->sweeps, pk, obs_row); \
    } while (0)
    if (record) { if (flip) FL(true, true); else FL(false, true); }
    else { if (flip) FL(true, false); else FL(false, false); }
#undef FL
    return VEGAS_OK;
}

template <typename real>
int fused_step_t(vegas_gpu* h, double* obs_row, bool record) {
    const FusedGeom& g = h->fused_geom;
    FusedPtrs<real> P{};
    for (int col = 0; col < 2; ++col)
        for (int c = 0; c < 3; ++c) { P.src[col][c] = (const real*)h->hs[col][c]; P.dst[col][c] = (real*)h->hs_alt[col][c]; }
    h->launches++;
    int rc;
    // row lengths with a specialised kernel (immediate shared-memory offsets); anything else takes the generic one
    if (sizeof(real) == 4 && g.Hx == 256) rc = fused_launch<real, 256>(h, P, obs_row, record);
    else if (g.Hx == 128) rc = fused_launch<real, 128>(h, P, obs_row, record);
    else if (g.Hx == 64) rc = fused_launch<real, 64>(h, P, obs_row, record);
    else rc = fused_launch<real, 0>(h, P, obs_row, record);
    if (rc) return rc;
    for (int col = 0; col < 2; ++col)
        for (int c = 0; c < 3; ++c) std::swap(h->hs[col][c], h->hs_alt[col][c]);
    return VEGAS_OK;
}

// ---- persistent wave step (heis_wave_kernel) ------------------------------------------------------
template <typename real>
const void* wave_kernel_ptr_t(bool flip, bool rec, bool multi) {
#define WK(F, R) (multi ? (const void*)heis_wave_kernel<real, F, R, true> : (const void*)heis_wave_kernel<real, F, R, false>)
    if (flip) return rec ? WK(true, true) : WK(true, false);
    return rec ? WK(false, true) : WK(false, false);
#undef WK
}
const void* wave_kernel_ptr(bool f64, bool flip, bool rec, bool multi) {
    return f64 ? wave_kernel_ptr_t<double>(flip, rec, multi) : wave_kernel_ptr_t<float>(flip, rec, multi);
}

// Unit order of a launch of `k` steps (2k phases) over n chunks: phase p visits chunk (p + pos) % n at time slot
// pos + p * lag; slots ascending, phases ascending inside a slot (see heis.cuh, K3w).
std::vector<uint32_t> wave_units_for(uint32_t n, uint32_t lag, uint32_t k) {
    std::vector<uint32_t> units;
    const uint32_t phases = 2 * k;
    for (uint32_t t = 0; t < n + (phases - 1) * lag; ++t)
        for (uint32_t p = 0; p < phases; ++p) {
            if (t < p * lag || t - p * lag >= n) continue;
            units.push_back((p << 24) | ((p + (t - p * lag)) % n));
        }
    return units;
}

bool wave_plan(vegas_gpu* h) {
    if (h->wave_ready) return true;
    if (h->family != FAM_HEIS_STENCIL || h->ndim != 3 || h->slab || h->wave_enable == 0) return false;
    if (h->wave_enable < 0 && (h->ld.nz < 32 || h->fused_enable == 1 || h->wave_c > 0)) return false;  // auto: big lattices only
    const uint32_t Lz = (uint32_t)h->ld.nz, C = std::max<uint32_t>(1, h->wave_planes);
    const uint32_t n = cdiv(Lz, C);
    if (n < 2 || n >= (1u << 24)) return false;
    const uint32_t D = std::max<uint32_t>(3, h->wave_lag);   // lag >= 3: dependencies precede their users (lag >= n: no overlap)
    const uint32_t kmax = std::max<uint32_t>(1, std::min<uint32_t>(h->wave_k, WAVE_MAX_STEPS));
    for (uint32_t k = 0; k < WAVE_MAX_STEPS; ++k) { cudaFree(h->wave_units[k]); h->wave_units[k] = nullptr; h->wave_n_units[k] = 0; }
    cudaFree(h->wave_done); cudaFree(h->wave_error);
    h->wave_done = nullptr; h->wave_error = nullptr;
    for (uint32_t k = 1; k <= kmax; ++k) {
        const std::vector<uint32_t> units = wave_units_for(n, D, k);
        if (units.size() != (size_t)2 * k * n) return false;
        if (cudaMalloc(&h->wave_units[k - 1], units.size() * 4) != cudaSuccess) { cudaGetLastError(); return false; }
        cudaMemcpy(h->wave_units[k - 1], units.data(), units.size() * 4, cudaMemcpyHostToDevice);
        h->wave_n_units[k - 1] = (uint32_t)units.size();
    }
    const size_t done_bytes = (size_t)2 * WAVE_MAX_STEPS * n * sizeof(unsigned long long);
    if (cudaMalloc(&h->wave_done, done_bytes) != cudaSuccess || cudaMalloc(&h->wave_error, 4) != cudaSuccess) {
        cudaGetLastError(); return false;
    }
    cudaMemset(h->wave_done, 0, done_bytes);
    for (auto& c : h->wave_phase_launches) c = 0;
    cudaMemset(h->wave_error, 0, 4);
    const HeisGeom g = heis_geom(h);
    WaveSched ws{};
    ws.tiles = cdiv((uint64_t)g.Ly * g.Gx, 128);
    ws.C = C; ws.n_chunks = n; ws.done = h->wave_done; ws.error = h->wave_error;
    h->wave_sched = ws;
    h->wave_kmax = kmax;
    // every CTA of the grid must be resident at once (static round-robin over the items)
    // (the grid must fit EVERY instantiation that can be launched, and the launch is cooperative so that
    // co-residency is guaranteed by the driver or the launch fails -- never a silent dependency time-out)
    int per_sm = 1 << 30, sms = 0, coop = 0;
    const bool f64 = h->md.precision == VEGAS_F64;
    for (int v = 0; v < 8; ++v) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wave_kernel_ptr(f64, v & 1, v & 2, v & 4), 128, 0) != cudaSuccess || n < 1) { cudaGetLastError(); return false; }
        per_sm = std::min(per_sm, n);
    }
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) != cudaSuccess ||
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device) != cudaSuccess || !coop) { cudaGetLastError(); return false; }
    h->wave_grid = per_sm * sms;
    h->wave_ready = true;
    return true;
}

template <typename real>
void wave_ptrs(vegas_gpu* h, int colour, HeisPtrs<real>& P) {
    const size_t plane = (size_t)(h->ld.nx / 2) * h->ld.ny;
    for (int c = 0; c < 3; ++c) {
        P.own[c] = (real*)h->hs[colour][c];
        P.oth[c] = (const real*)h->hs[1 - colour][c];
        P.oth_lo[c] = P.oth[c] + (size_t)(h->ld.nz - 1) * plane;   // periodic wrap of the single handle
        P.oth_hi[c] = P.oth[c];
        P.peer_lo[c] = nullptr; P.peer_hi[c] = nullptr;
    }
}

// k consecutive steps (1 <= k <= wave_kmax) in one launch; obs_row = row of the first step (rows OBS_W apart) when
// recording, else the scratch row (all steps add into it)
template <typename real>
void wave_steps_t(vegas_gpu* h, uint32_t k, double* obs_row, bool record) {
    HeisPtrs<real> P0{}, P1{};
    wave_ptrs<real>(h, 0, P0); wave_ptrs<real>(h, 1, P1);
    WaveSched ws = h->wave_sched;
    ws.units = h->wave_units[k - 1]; ws.n_units = h->wave_n_units[k - 1]; ws.n_phases = 2 * k;
    // the last phase of a launch has no dependants and is not counted (heis_wave_kernel)
    for (uint32_t ph = 0; ph + 1 < 2 * k; ++ph) ws.target[ph] = (unsigned long long)ws.tiles * (++h->wave_phase_launches[ph]);
    const HeisGeom g = heis_geom(h);
    const HeisParams<real> p = heis_params<real>(h);
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    const int grid = (int)std::min<uint64_t>((uint64_t)h->wave_grid, (uint64_t)ws.n_units * ws.tiles);
    int stride = record ? OBS_W : 0;
    uint64_t sweep = h->sweeps;
    PhiloxKey key = pk;
    HeisGeom geom = g;
    HeisParams<real> par = p;
    h->launches++;
    void* args[] = {&P0, &P1, &geom, &ws, &par, &sweep, &key, &obs_row, &stride};
    const cudaError_t e = cudaLaunchCooperativeKernel(wave_kernel_ptr_t<real>(flip, record, k > 1), dim3(grid), dim3(128), args, 0, h->stream);
    if (e != cudaSuccess) {   // co-residency refused: the caller's cudaGetLastError check reports it; two-pass kernels from now on
        h->err = std::string("heis_wave_kernel cooperative launch failed: ") + cudaGetErrorString(e);
        h->wave_ready = false; h->wave_enable = 0;
    }
}

// ---- phase-pipelined TMA step (heis_pipe.cu) ---------------------------------------------------------
// Both colour passes of a step in ONE cooperative launch: dedicated CTAs per colour march the planes in lock step, the
// second colour a few planes behind the first (per-plane progress counters), tiles staged by TMA into mbarrier rings.
bool pipe_plan(vegas_gpu* h) {
    if (h->pipe_planned) return h->pipe != nullptr;
    if (h->family != FAM_HEIS_STENCIL || h->ndim != 3 || h->pipe_enable == 0) return false;
    if (h->pipe_enable < 0 && (h->ld.nz < 32 || h->fused_enable == 1 || h->wave_c > 0 || h->wave_enable == 1 || h->wave_k > 1)) return false;
    // a slab runs it once it is connected to neighbours that sweep on OTHER devices: the kernel occupies every SM of its
    // device and waits for the neighbours' boundary planes inside the launch (slabs sharing a device would deadlock)
    if (h->slab && !(h->connected && h->peers_remote)) return false;
    h->pipe_planned = true;
    HeisPipeDesc d;
    d.device = h->device; d.f64 = h->md.precision == VEGAS_F64;
    d.Lx = (uint32_t)h->ld.nx; d.Ly = (uint32_t)h->ld.ny; d.Lz = (uint32_t)h->ld.nz; d.z_offset = (uint32_t)h->z_offset;
    for (int col = 0; col < 2; ++col) for (int c = 0; c < 3; ++c) d.arr[col][c] = h->hs[col][c];
    if (h->slab) {
        d.slab = true;
        for (int col = 0; col < 2; ++col)
            for (int c = 0; c < 3; ++c) {
                d.halo[col][0][c] = (char*)h->halo + halo_offset(h, col, 0, c);
                d.halo[col][1][c] = (char*)h->halo + halo_offset(h, col, 1, c);
                // my plane 0 goes to the lower neighbour's UPPER halo of my colour, my plane Lz-1 to the upper neighbour's LOWER halo
                d.peer[col][0][c] = (char*)h->peer_halo[0] + halo_offset(h, col, 1, c);
                d.peer[col][1][c] = (char*)h->peer_halo[1] + halo_offset(h, col, 0, c);
            }
        d.flags = h->flags + PIPE_FLAG_WORD;
        d.peer_flags[0] = h->peer_flags[0] + PIPE_FLAG_WORD; d.peer_flags[1] = h->peer_flags[1] + PIPE_FLAG_WORD;
    }
    d.stages_other = h->pipe_stages_other; d.stages_own = h->pipe_stages_own; d.tiles = h->pipe_tiles; d.vec = h->pipe_vec; d.lead = h->pipe_lead; d.pub_every = h->pipe_pub; d.backoff_consumer = h->pipe_backoff_c; d.backoff_helper = h->pipe_backoff_h; d.l2_hints = h->pipe_l2;
    h->pipe = heis_pipe_create(d, h->pipe_why);
    return h->pipe != nullptr;
}

bool bpipe_plan(vegas_gpu* h) {
    if (h->bpipe_planned) return h->bpipe != nullptr;
    // opt-in (tuning key basis_pipe=1): it moves the compulsory 24 B/attempt through DRAM (5.6 GB per fcc 384^3 step against
    // 14 GB for four colour launches) but is latency bound at 17 warps per SM: 3.3 ms per step against 2.5 (profiles/r02/README.md)
    if (h->family != FAM_HEIS_BASIS || h->slab || h->bpipe_enable != 1) return false;
    h->bpipe_planned = true;
    BasisPipeDesc d;
    d.device = h->device; d.f64 = h->md.precision == VEGAS_F64;
    d.unitcell = h->ld.unitcell == VEGAS_BCC ? 1 : 2;
    d.nx = (uint32_t)h->ld.nx; d.ny = (uint32_t)h->ld.ny; d.nz = (uint32_t)h->ld.nz;
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) d.arr[b][c] = h->hb[b][c];
    d.tiles = h->bpipe_tiles; d.lead = h->bpipe_lead; d.pub_every = h->bpipe_pub;
    h->bpipe = basis_pipe_create(d, h->pipe_why);
    return h->bpipe != nullptr;
}

// ---- fcc pair launches (heis_basis_pair_kernel) -----------------------------------------------------
// bonds between colours b0 and b0 + 1, seen from b0 + 1: dz must be 0 and dy in {0, +1}
template <int UC>
bool pair_structure_ok() {
    constexpr int NB = BasisCell<UC>::NB;
    if (NB % 2) return false;
    for (int e = 0; e < BasisCell<UC>::NE; ++e) {
        int s = 0, t = 0, dx = 0, dy = 0, dz = 0;
        BasisCell<UC>::edge(e, s, t, dx, dy, dz);
        int lo = s, hi = t, ddy = dy, ddz = dz;              // neighbour of `lo` at +d is `hi`
        if (lo > hi) { std::swap(lo, hi); ddy = -dy; ddz = -dz; }
        if (hi != lo + 1 || (lo & 1)) continue;              // not a bond inside a pair
        // seen from hi, the neighbour lo sits at -d
        if (ddz != 0 || !(-ddy == 0 || -ddy == 1)) return false;
    }
    return true;
}

bool bpair_plan(vegas_gpu* h) {
    // Bit-identical to the colour launches and synchronisation-free; with 3 fat CTAs per SM the window between a CTA's two
    // colours fits L2: 2.22 ms per fcc 384^3 step against 2.45 for four colour launches (profiles/r02/README.md section 11).
    // Default (-1) for single-handle lattices beyond L2 size; it needs the State twice in HBM.  Slabs keep the colour launches
    // (the second array set would have to live in the slab's IPC allocation with its own halo planes).
    if (h->family != FAM_HEIS_BASIS || h->bpair_enable == 0 || h->basis_vec == 0) return false;
    // a slab steps from one array set of its allocation into the other, in lock step with its neighbours; it needs the
    // connection, three planes, and no wave state (which captured the set it was created on)
    if (h->slab && !(h->pair_set_bytes && h->connected && h->ld.nz >= 3 && !h->bwave)) return false;
    if (h->bpair_enable < 0 && h->n * 3 * real_bytes(h) < (96ull << 20)) return false;   // auto: the State does not fit in L2
    if (h->ld.unitcell != VEGAS_FCC) return false;
    const uint32_t NV = h->md.precision == VEGAS_F64 ? 2u : 4u;
    if (h->ld.nx % NV) return false;
    if (h->bpair_ok < 0) h->bpair_ok = pair_structure_ok<2>() ? 1 : 0;
    if (!h->bpair_ok) return false;
    if (!h->hb2[0][0]) {   // second set of arrays, once (a slab's second set lives in its one allocation)
        const size_t bytes = (size_t)(h->n / h->n_colours) * real_bytes(h);
        for (int b = 0; b < h->n_colours; ++b)
            for (int k = 0; k < 3; ++k)
                if (cudaMalloc(&h->hb2[b][k], bytes) != cudaSuccess) {
                    cudaGetLastError();
                    for (int bb = 0; bb < 4; ++bb) for (int kk = 0; kk < 3; ++kk) { cudaFree(h->hb2[bb][kk]); h->hb2[bb][kk] = nullptr; }
                    h->bpair_enable = 0;   // not enough memory for the second set: colour launches
                    return false;
                }
    }
    return true;
}

template <typename real>
void bpair_step_t(vegas_gpu* h, double* obs_row, bool record) {
    const BasisGeom g = basis_geom(h);
    const HeisParams<real> p = heis_params<real>(h);
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    BasisPtrs<real> S = basis_ptrs<real>(h), D{};
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) D.s[b][c] = (real*)h->hb2[b][c];
    const uint32_t rows = std::max<uint32_t>(1, std::min<uint32_t>(h->bpair_rows ? h->bpair_rows : 48u, g.ny));
    // rows after which a CTA switches colours: as few as keep all 128 threads busy (rows * vectors per row a multiple of 128);
    // fcc 384^3 (96 vectors per row): 4 rows 2.21 ms, 2 / 3 / 5 / 6 rows 2.57 / 2.52 / 2.38 / 2.51, 8 rows 2.58 (window beyond L2)
    uint32_t auto_chunk = 128u / std::gcd(g.nx / (uint32_t)VecOf<real>::N, 128u);
    if (auto_chunk > 8u) auto_chunk = 4u;
    const uint32_t chunk = std::max<uint32_t>(1, h->bpair_chunk ? h->bpair_chunk : auto_chunk);
    const bool slab = h->slab;
    BasisPeers<real> peers{nullptr, nullptr};
    if (slab) { peers.lo = (real*)peer_set(h, 0, h->cur_set ^ 1); peers.hi = (real*)peer_set(h, 1, h->cur_set ^ 1); }   // the neighbours' D sets
    const uint32_t nz = g.nz;
    // planes zb, zb + zstep, ... (zc of them) of pair B0 on stream st
    auto launch = [&](int B0, uint32_t zb, uint32_t zc, uint32_t zstep, cudaStream_t st) {
        if (zc == 0) return;
        const dim3 grid(cdiv(g.ny, rows), 1, zc);
#define BPK(B0v, F, M, SL) heis_basis_pair_kernel<real, 2, B0v, F, M, SL><<<grid, 128, 0, st>>>(S, D, peers, g, rows, chunk, zb, zstep, p, h->sweeps, pk, obs_row)
#define BPM(B0v, SL)                                                                              \
    do {                                                                                          \
        if (record) { if (flip) BPK(B0v, true, 1, SL); else BPK(B0v, false, 1, SL); }             \
        else { if (flip) BPK(B0v, true, 0, SL); else BPK(B0v, false, 0, SL); }                    \
    } while (0)
        if (B0 == 0) { if (slab) BPM(0, true); else BPM(0, false); }
        else { if (slab) BPM(2, true); else BPM(2, false); }
#undef BPM
#undef BPK
        h->launches++;
    };
    for (int B0 = 0; B0 < 4; B0 += 2) {
        if (!slab) { launch(B0, 0, nz, 1, h->stream); continue; }
        // connected slab: the same flag protocol and stream structure as the colour launches (basis_pass_any), one exchange
        // per PAIR launch: the boundary planes wait for the neighbours' previous launch, store into their halos, signal
        h->pass_counter++;
        if (!h->peer_is_ipc) {   // all slabs in one process: one stream
            if (h->pass_counter > 1) { wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, h->pass_counter - 1, h->slab_error); h->launches++; }
            launch(B0, 0, 2, nz - 1, h->stream);
            signal_kernel<<<1, 32, 0, h->stream>>>(h->peer_flags[0], h->peer_flags[1], h->pass_counter);
            h->launches++;
            launch(B0, 1, nz - 2, 1, h->stream);
            continue;
        }
        cudaEventRecord(h->ev_main, h->stream);
        cudaStreamWaitEvent(h->stream_b, h->ev_main, 0);
        cudaStreamWaitEvent(h->stream, h->ev_bnd, 0);
        if (h->pass_counter > 1) { wait_kernel<<<1, 32, 0, h->stream_b>>>(h->flags, h->pass_counter - 1, h->slab_error); h->launches++; }
        launch(B0, 0, 2, nz - 1, h->stream_b);
        signal_kernel<<<1, 32, 0, h->stream_b>>>(h->peer_flags[0], h->peer_flags[1], h->pass_counter);
        h->launches++;
        cudaEventRecord(h->ev_bnd, h->stream_b);
        launch(B0, 1, nz - 2, 1, h->stream);
    }
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) std::swap(h->hb[b][c], h->hb2[b][c]);   // D is the State now
    h->cur_set ^= 1;
}

bool bwave_plan(vegas_gpu* h) {
    if (h->bwave_planned) return basis_wave_usable(h->bwave);
    // opt-in (tuning key basis_wave=1): 9.2 GB of DRAM traffic per fcc 384^3 step against 14 GB for four colour launches, but
    // every item pays a dependency poll, two fences and two barriers and tiles of 128 items lose the L1 reuse between rows:
    // 2.98 ms per step against 2.51 (profiles/r02/README.md section 2)
    if (h->family != FAM_HEIS_BASIS || h->bwave_enable != 1 || h->basis_vec == 0) return false;
    // a slab runs it once it is connected to neighbours that sweep on OTHER devices: the kernel occupies every SM of its
    // device and waits for the neighbours' boundary planes inside the launch (slabs sharing a device would deadlock)
    if (h->slab && !(h->connected && h->peers_remote)) return false;
    h->bwave_planned = true;
    BasisWaveDesc d;
    d.device = h->device; d.f64 = h->md.precision == VEGAS_F64;
    d.unitcell = h->ld.unitcell == VEGAS_BCC ? 1 : 2;
    d.nx = (uint32_t)h->ld.nx; d.ny = (uint32_t)h->ld.ny; d.nz = (uint32_t)h->ld.nz;
    d.z_offset = (uint32_t)h->z_offset; d.nz_global = (uint32_t)h->nz_global;
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) d.arr[b][c] = h->hb[b][c];
    d.lag = h->bwave_lag; d.ipt = h->bwave_ipt; d.grid = h->bwave_grid;
    if (h->slab) {
        d.slab = true;
        d.peer_lo = peer_set(h, 0, h->cur_set); d.peer_hi = peer_set(h, 1, h->cur_set);
        d.flags = h->flags + BWAVE_FLAG_WORD;
        d.peer_flags[0] = h->peer_flags[0] + BWAVE_FLAG_WORD; d.peer_flags[1] = h->peer_flags[1] + BWAVE_FLAG_WORD;
    }
    h->bwave = basis_wave_create(d, h->pipe_why);
    return h->bwave != nullptr;
}

int bwave_step(vegas_gpu* h, double* obs_row, bool record) {
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    h->launches++;
    std::string err;
    const int rc = h->md.precision == VEGAS_F64 ? basis_wave_step<double>(h->bwave, heis_params<double>(h), flip, record, h->sweeps, pk, obs_row, h->stream, err)
                                               : basis_wave_step<float>(h->bwave, heis_params<float>(h), flip, record, h->sweeps, pk, obs_row, h->stream, err);
    if (rc) h->err = err;
    return rc;
}

int bpipe_step(vegas_gpu* h, double* obs_row, bool record) {
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    h->launches++;
    std::string err;
    const int rc = h->md.precision == VEGAS_F64 ? basis_pipe_step<double>(h->bpipe, heis_params<double>(h), flip, record, h->sweeps, pk, obs_row, h->stream, err)
                                               : basis_pipe_step<float>(h->bpipe, heis_params<float>(h), flip, record, h->sweeps, pk, obs_row, h->stream, err);
    if (rc) h->err = err;
    return rc;
}

int pipe_step(vegas_gpu* h, double* obs_row, bool record) {
    const PhiloxKey pk = make_philox_key(h->md.seed);
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    h->launches++;
    std::string err;
    const int rc = h->md.precision == VEGAS_F64 ? heis_pipe_step<double>(h->pipe, heis_params<double>(h), flip, record, h->sweeps, pk, obs_row, h->pipe_slab_steps, h->stream, err)
                                               : heis_pipe_step<float>(h->pipe, heis_params<float>(h), flip, record, h->sweeps, pk, obs_row, h->pipe_slab_steps, h->stream, err);
    if (rc) h->err = err;
    else h->pipe_slab_steps++;
    return rc;
}

// Persistent kernels report a dependency wait that timed out through a device flag; every entry point that hands
// results to the caller checks it after synchronising (a stale-plane sweep must never look like a valid one).
int check_async_errors(vegas_gpu* h) {
    if (h->slab && h->connected && h->slab_error) {
        unsigned int serr = 0;
        CU(cudaMemcpy(&serr, h->slab_error, 4, cudaMemcpyDeviceToHost));
        if (serr) {
            cudaMemset(h->slab_error, 0, 4);
            return fail(h, VEGAS_ERR_CUDA, "slab halo exchange: a z-neighbour never signalled its pass (results invalid)");
        }
    }
    if (h->wave_error) {
        unsigned int werr = 0;
        CU(cudaMemcpy(&werr, h->wave_error, 4, cudaMemcpyDeviceToHost));
        if (werr) {
            cudaMemset(h->wave_error, 0, 4);
            h->wave_ready = false; h->wave_enable = 0;   // co-residency cannot be relied on here: two-pass kernels from now on
            return fail(h, VEGAS_ERR_CUDA, "heis_wave_kernel: a dependency wait timed out (results invalid)");
        }
    }
    if (h->pipe) {
        std::string e;
        if (heis_pipe_check(h->pipe, e)) return fail(h, VEGAS_ERR_CUDA, e);
    }
    if (h->bpipe) {
        std::string e;
        if (basis_pipe_check(h->bpipe, e)) return fail(h, VEGAS_ERR_CUDA, e);
    }
    if (h->bwave) {
        std::string e;
        if (basis_wave_check(h->bwave, e)) return fail(h, VEGAS_ERR_CUDA, e);
    }
    return VEGAS_OK;
}

// ---- K2r: a batch of steps of a small general-family lattice in ONE launch (state resident in shared memory) ----
constexpr size_t RES_SMEM_LIMIT = 226 * 1024;  // of the 227 KB a CTA may opt in to
constexpr uint32_t RES_DIRECT_MAX = 2048;      // the direct variant (per-bond values) only pays off on tiny lattices

// widest neighbour row (self entries included: an upper bound for the table columns); 0 when the rows are unknown
int resident_row_width(const vegas_gpu* h) {
    if (h->csr_input) {
        uint64_t w = 0;
        for (size_t i = 0; i + 1 < h->h_row_ptr.size(); ++i) w = std::max<uint64_t>(w, h->h_row_ptr[i + 1] - h->h_row_ptr[i]);
        return (int)std::min<uint64_t>(w, 1u << 20);
    }
    const StructuredNb nb = structured_nb(h);
    int w = 0;
    for (int b = 0; b < nb.nb; ++b) w = std::max(w, nb.count[b]);
    return w;
}

// table columns of the resident launch: > 0 table variant, 0 direct variant, -1 no resident path for this handle
int resident_columns(const vegas_gpu* h) {
    if (h->family != FAM_ISING_GEN && h->family != FAM_HEIS_GEN) return -1;
    if (h->slab || h->n == 0 || h->n > h->resident_max || h->n > 16384 || h->n_colours > RES_MAX_COLOURS) return -1;
    const size_t per_site = h->family == FAM_ISING_GEN ? 1 : 3 * real_bytes(h);
    const bool uniform = !(h->csr_input && h->d_val);
    if (uniform) {
        const int z = std::max(1, resident_row_width(h));
        if (z <= RES_MAX_Z && resident_smem_bytes((uint32_t)h->n, z, per_site) <= RES_SMEM_LIMIT) return z;
    }
    if (h->n <= RES_DIRECT_MAX && resident_smem_bytes((uint32_t)h->n, 0, per_site) <= RES_SMEM_LIMIT) return 0;
    return -1;
}

bool resident_plan(vegas_gpu* h) {
    if (h->resident_cols == -2) h->resident_cols = resident_columns(h);
    return h->resident_cols >= 0;
}

template <typename K, typename... Args>
int resident_launch(vegas_gpu* h, K kernel, uint32_t threads, size_t smem, Args... args) {
    // the attribute is per kernel instantiation and cheap to set; the variants of one handle never change
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<1, threads, smem, h->stream>>>(args...);
    h->launches++;
    return VEGAS_OK;
}

template <typename NB, bool TABLE>
int resident_steps_t(vegas_gpu* h, const NB& nb, uint32_t n_steps, bool record) {
    ResidentPlan rp{};
    rp.n_colours = h->n_colours;
    rp.n = (uint32_t)h->n;
    rp.zmax = TABLE ? h->resident_cols : 0;
    uint32_t widest = 1;
    for (int c = 0; c < h->n_colours; ++c) {
        rp.sites[c] = h->g_sites[c];
        rp.counts[c] = h->g_counts[c];
        widest = std::max(widest, h->g_counts[c]);
    }
    const uint32_t threads = std::min<uint32_t>(1024, std::max<uint32_t>(128, (widest + 31) / 32 * 32));
    const size_t smem = resident_smem_bytes(rp.n, rp.zmax, h->family == FAM_ISING_GEN ? 1 : 3 * real_bytes(h));
    const PhiloxKey pk = make_philox_key(h->md.seed);
    unsigned long long* rows = record ? h->obs : nullptr;
    unsigned long long* scratch = h->obs + OBS_CAP * OBS_W;
    const double J = nb.J;
    if (h->family == FAM_ISING_GEN) {
        IsingGeneralParams p{};
        p.thr = h->g_thr; p.code = h->g_code;
        p.uniform = (h->csr_input && h->d_val) ? 0 : 1;
        p.h_o = ising_h_o(h); p.invT = 1.0 / h->T;
        constexpr bool SYMM = std::is_same<NB, StructuredNb>::value;  // a user CSR may be asymmetric: full reduction per step
        if (h->md.proposal == VEGAS_PROPOSE_RANDOM)
            return resident_launch(h, ising_resident_kernel<NB, true, TABLE, SYMM>, threads, smem, h->g_s8, nb, rp, p, J, h->sweeps, n_steps, pk, rows, OBS_W, scratch);
        return resident_launch(h, ising_resident_kernel<NB, false, TABLE, SYMM>, threads, smem, h->g_s8, nb, rp, p, J, h->sweeps, n_steps, pk, rows, OBS_W, scratch);
    }
    const bool flip = h->md.proposal == VEGAS_PROPOSE_FLIP;
    const double ax = h->md.anisotropy_axis[0], ay = h->md.anisotropy_axis[1], az = h->md.anisotropy_axis[2];
    if (h->md.precision == VEGAS_F64) {
        const HeisParams<double> p = heis_params<double>(h);
        double *x = (double*)h->g_s[0], *y = (double*)h->g_s[1], *z = (double*)h->g_s[2];
        if (flip) return resident_launch(h, heis_resident_kernel<NB, double, true, TABLE>, threads, smem, x, y, z, nb, rp, p, J, ax, ay, az, h->sweeps, n_steps, pk, (double*)rows, OBS_W, (double*)scratch);
        return resident_launch(h, heis_resident_kernel<NB, double, false, TABLE>, threads, smem, x, y, z, nb, rp, p, J, ax, ay, az, h->sweeps, n_steps, pk, (double*)rows, OBS_W, (double*)scratch);
    }
    const HeisParams<float> p = heis_params<float>(h);
    float *x = (float*)h->g_s[0], *y = (float*)h->g_s[1], *z = (float*)h->g_s[2];
    if (flip) return resident_launch(h, heis_resident_kernel<NB, float, true, TABLE>, threads, smem, x, y, z, nb, rp, p, J, ax, ay, az, h->sweeps, n_steps, pk, (double*)rows, OBS_W, (double*)scratch);
    return resident_launch(h, heis_resident_kernel<NB, float, false, TABLE>, threads, smem, x, y, z, nb, rp, p, J, ax, ay, az, h->sweeps, n_steps, pk, (double*)rows, OBS_W, (double*)scratch);
}

template <typename NB>
int resident_steps(vegas_gpu* h, const NB& nb, uint32_t n_steps, bool record) {
    return h->resident_cols > 0 ? resident_steps_t<NB, true>(h, nb, n_steps, record) : resident_steps_t<NB, false>(h, nb, n_steps, record);
}

// One Monte Carlo step (= N attempts): every colour once.  obs_row != null records observables.
void do_step(vegas_gpu* h, void* obs_row, void* scratch_row) {
    const bool rec = obs_row != nullptr;
    if (h->family == FAM_HEIS_BASIS && bpipe_plan(h)) {
        bpipe_step(h, (double*)(rec ? obs_row : scratch_row), rec);
    } else if (h->family == FAM_HEIS_BASIS && bwave_plan(h)) {
        bwave_step(h, (double*)(rec ? obs_row : scratch_row), rec);   // a failed launch surfaces through cudaGetLastError / h->err
    } else if (h->family == FAM_HEIS_BASIS && bpair_plan(h)) {
        if (h->md.precision == VEGAS_F64) bpair_step_t<double>(h, (double*)(rec ? obs_row : scratch_row), rec);
        else bpair_step_t<float>(h, (double*)(rec ? obs_row : scratch_row), rec);
    } else if (h->family == FAM_HEIS_BASIS) {
        for (int b = 0; b < h->n_colours; ++b) basis_pass_any(h, rec ? 1 : 0, b, (double*)(rec ? obs_row : scratch_row));
    } else if (h->family == FAM_HEIS_STENCIL && pipe_plan(h)) {
        pipe_step(h, (double*)(rec ? obs_row : scratch_row), rec);   // a failed launch surfaces through cudaGetLastError / h->err
    } else if (h->family == FAM_HEIS_STENCIL && wave_plan(h)) {
        double* row = (double*)(rec ? obs_row : scratch_row);
        if (h->md.precision == VEGAS_F64) wave_steps_t<double>(h, 1, row, rec); else wave_steps_t<float>(h, 1, row, rec);
    } else if (h->family == FAM_HEIS_STENCIL && fused_plan(h)) {
        double* row = (double*)(rec ? obs_row : scratch_row);
        if (h->md.precision == VEGAS_F64) fused_step_t<double>(h, row, rec); else fused_step_t<float>(h, row, rec);
    } else if (h->family == FAM_HEIS_STENCIL && h->wave_c > 0 && h->ndim == 3 && !h->slab && h->ld.nz >= 2 * h->wave_c) {
        // EXPERIMENT (tuning key "heis_wave_c"): the two colour passes interleaved in chunks of wave_c planes so that the
        // second pass finds the planes of the first in L2.  Colour 1 on planes [z, z+C) needs colour 0 on [z-1, z+C],
        // and colour 0 on plane Lz-1 reads the OLD colour 1 on plane 0, so colour-1 plane 0 goes last.
        const uint32_t Lz = (uint32_t)h->ld.nz, C = h->wave_c;
        void* row = rec ? obs_row : scratch_row;
        auto pass = [&](int colour, uint32_t zb, uint32_t zc) {
            if (zc == 0) return;
            const int mode = !rec ? 0 : (colour == 1 ? 1 : 3);
            if (h->md.precision == VEGAS_F64) heis_pass<double>(h, mode, colour, zb, zc, (double*)row);
            else heis_pass<float>(h, mode, colour, zb, zc, (double*)row);
        };
        uint32_t a_done = 0, b_done = 1;                 // colour 0 done on [0, a_done); colour 1 done on [1, b_done)
        while (a_done < Lz) {
            const uint32_t n = std::min(C, Lz - a_done);
            pass(0, a_done, n); a_done += n;
            const uint32_t b_to = a_done == Lz ? Lz : a_done - 1;   // colour 1 may advance to plane a_done - 2
            if (b_to > b_done) { pass(1, b_done, b_to - b_done); b_done = b_to; }
        }
        pass(1, 0, 1);
    } else if (h->family == FAM_ISING_MSC || h->family == FAM_HEIS_STENCIL) {
        for (int c = 0; c < 2; ++c) {
            // recorded step: the Ising kernel reduces both colours in the last pass; the Heisenberg kernel reduces the
            // own colour in each pass (mode 3 / 1) and the exchange energy in the last one
            const int mode = !rec ? 0 : (c == 1 ? 1 : (h->family == FAM_HEIS_STENCIL ? 3 : 0));
            stencil_colour_pass(h, mode, c, rec ? obs_row : scratch_row);
        }
    } else {
        unsigned long long* row = (unsigned long long*)(rec ? obs_row : scratch_row);
        const bool fused_obs = rec && basis_fused_obs(h);
        for (int c = 0; c < h->n_colours; ++c) {
            if (fused_obs) basis_obs_pass(h, c, (double*)row);
            else if (h->csr_input) general_colour_pass(h, csr_nb(h), c, row);
            else general_colour_pass(h, structured_nb(h), c, row);
        }
        if (rec && !fused_obs) {
            if (h->csr_input) general_reduce(h, csr_nb(h), (double*)obs_row);
            else general_reduce(h, structured_nb(h), (double*)obs_row);
        }
    }
    join_boundary_stream(h);
    h->sweeps++;
    h->attempts += h->n;
}

// ---------------------------------------------------------------------------------------
// observables: device row -> canonical (E_exchange physical, M, A, accepted) -> convention
// ---------------------------------------------------------------------------------------
struct Canon { double eex, m[3], a; uint64_t accepted; };

Canon canon_of(const vegas_gpu* h, const unsigned long long* row) {
    Canon c{};
    const double J = h->md.has_exchange ? h->md.exchange : 0.0;
    const double N = (double)h->n;
    if (h->family == FAM_ISING_MSC) {
        const long long bonds = (long long)row[0], M = (long long)row[1];
        c.eex = -J * (double)bonds; c.m[2] = (double)M; c.a = N; c.accepted = row[2];
        c.eex += -J * N * h->n_self;
    } else if (h->family == FAM_HEIS_STENCIL) {
        const double* d = (const double*)row;
        c.eex = d[0] + (-J * N * h->n_self); c.m[0] = d[1]; c.m[1] = d[2]; c.m[2] = d[3]; c.a = d[4];
        c.accepted = (uint64_t)llround(d[5]);
    } else {
        const double* d = (const double*)row;
        c.eex = -d[0] / 2.0; c.m[0] = d[1]; c.m[1] = d[2]; c.m[2] = d[3];
        c.a = h->md.model == VEGAS_ISING ? N : d[4];
        c.accepted = h->family == FAM_ISING_GEN ? row[6] : (uint64_t)llround(d[5]);
    }
    return c;
}

double energy_of(const vegas_gpu* h, const Canon& c) {
    const double N = (double)h->n;
    const double ha = h_abs(h);
    double zs;  // sum_i s_i . orientation
    if (h->md.model == VEGAS_ISING) zs = c.m[2] * (h->fdir[2] >= 0.0 ? 1.0 : -1.0);
    else zs = c.m[0] * h->fdir[0] + c.m[1] * h->fdir[1] + c.m[2] * h->fdir[2];
    const double k = h->md.has_anisotropy ? h->md.anisotropy_k : 0.0;
    const double g = h->md.has_gauge ? h->md.gauge : 0.0;
    const double eex = h->md.has_exchange ? c.eex : 0.0;
    switch (h->econv) {
        case VEGAS_E_PHYSICAL: return eex - ha * zs + k * c.a + g * N;
        case VEGAS_E_REFERENCE_EXCHANGE: return eex;
        default: return 2.0 * eex + ha * zs + k * c.a + g * N;  // src/energy.rs:55-59 over the compound
    }
}

// Before anything reads the halos of a slab that steps with the pipelined kernel: the neighbours' boundary planes of the
// last step may still be on their way (the kernel only waits for what it needs itself).  Bounded spin, never hangs.
__global__ void pipe_halo_wait_kernel(const unsigned long long* flags, unsigned long long target) {
    if (threadIdx.x < 4) {
        unsigned long long v;
        for (uint32_t spins = 0; spins < (1u << 22); ++spins) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= target) break;
            __nanosleep(256);
        }
    }
    __threadfence_system();
}

// The same for a slab that steps with the wave-ordered bcc / fcc kernel: words [a] / [4 + c] count the steps whose plane 0 /
// top plane of colour a / c the upper / lower neighbour has stored into my halos.
__global__ void bwave_halo_wait_kernel(const unsigned long long* flags, unsigned long long target, uint32_t nb) {
    if (threadIdx.x < 8 && (threadIdx.x & 3u) < nb) {
        unsigned long long v;
        for (uint32_t spins = 0; spins < (1u << 22); ++spins) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= target) break;
            __nanosleep(256);
        }
    }
    __threadfence_system();
}

int measure_now(vegas_gpu* h, Canon& out) {
    unsigned long long* row = h->obs + (OBS_CAP + 1) * OBS_W;
    CU(cudaMemsetAsync(row, 0, OBS_W * 8, h->stream));
    if (h->slab && h->pipe && h->pipe_slab_steps > 0) {
        pipe_halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags + PIPE_FLAG_WORD, h->pipe_slab_steps * (unsigned long long)heis_pipe_tiles(h->pipe));
        h->launches++;
    }
    if (h->slab && h->bwave && basis_wave_steps_done(h->bwave) > 0) {
        bwave_halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags + BWAVE_FLAG_WORD, basis_wave_steps_done(h->bwave), (uint32_t)h->n_colours);
        h->launches++;
    }
    if (h->family == FAM_ISING_MSC || h->family == FAM_HEIS_STENCIL) stencil_colour_pass(h, 2, 1, row);
    else if (h->family == FAM_HEIS_BASIS) { for (int b = 0; b < h->n_colours; ++b) basis_pass_any(h, 2, b, (double*)row); }
    else if (h->csr_input) general_reduce(h, csr_nb(h), (double*)row);
    else general_reduce(h, structured_nb(h), (double*)row);
    unsigned long long host[OBS_W];
    CU(cudaMemcpyAsync(host, row, sizeof host, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    { const int rc = check_async_errors(h); if (rc) return rc; }
    out = canon_of(h, host);
    return VEGAS_OK;
}

// ---------------------------------------------------------------------------------------
// creation
// ---------------------------------------------------------------------------------------
int common_init(vegas_gpu* h) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(h, VEGAS_ERR_CUDA, "no CUDA device: the vegas_gpu sweep has no CPU fallback");
    if (h->md.device < 0 || h->md.device >= ndev) return fail(h, VEGAS_ERR_INVALID, "device ordinal out of range");
    h->device = h->md.device;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
        int lo_prio = 0, hi_prio = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CU(cudaStreamCreateWithPriority(&h->stream_b, cudaStreamNonBlocking, hi_prio));
        CU(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_bnd, cudaEventDisableTiming));
    }
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    {   // staging buffers of the host-layout entry points come from the stream-ordered pool: keep freed blocks cached,
        // otherwise every Integrator::step-style call pays a fresh 1-3 GB driver allocation
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, h->device));
        unsigned long long keep = ~0ull;
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CU(cudaMalloc(&h->slab_error, 4));
    CU(cudaMemsetAsync(h->slab_error, 0, 4, h->stream));
    CU(cudaMalloc(&h->obs, (OBS_CAP + 2) * OBS_W * 8));
    CU(cudaMemsetAsync(h->obs, 0, (OBS_CAP + 2) * OBS_W * 8, h->stream));
    return VEGAS_OK;
}

int alloc_general(vegas_gpu* h, const std::vector<uint8_t>& colour) {
    const uint64_t n = h->n;
    if (n >= (1ull << 32)) return fail(h, VEGAS_ERR_INVALID, "general-adjacency path supports < 2^32 sites");
    if (h->md.model == VEGAS_ISING) {
        CU(cudaMalloc(&h->g_s8, n));
        CU(cudaMalloc(&h->g_thr, (size_t)2 * (2 * ISING_ZMAX + 1) * 8));
        CU(cudaMalloc(&h->g_code, (size_t)2 * (2 * ISING_ZMAX + 1)));
    } else {
        for (int c = 0; c < 3; ++c) CU(cudaMalloc(&h->g_s[c], n * real_bytes(h)));
    }
    std::vector<std::vector<uint32_t>> lists(h->n_colours);
    for (uint64_t i = 0; i < n; ++i) lists[colour[i]].push_back((uint32_t)i);
    h->g_sites.assign(h->n_colours, nullptr);
    h->g_counts.assign(h->n_colours, 0);
    for (int c = 0; c < h->n_colours; ++c) {
        h->g_counts[c] = (uint32_t)lists[c].size();
        if (lists[c].empty()) continue;
        CU(cudaMalloc(&h->g_sites[c], lists[c].size() * 4));
        CU(cudaMemcpy(h->g_sites[c], lists[c].data(), lists[c].size() * 4, cudaMemcpyHostToDevice));
    }
    return VEGAS_OK;
}

int check_model(vegas_gpu* h, const vegas_model_desc* md) {
    if (!md) return fail(h, VEGAS_ERR_INVALID, "null model descriptor");
    if (md->model != VEGAS_ISING && md->model != VEGAS_HEISENBERG) return fail(h, VEGAS_ERR_INVALID, "unknown model");
    if (md->proposal != VEGAS_PROPOSE_FLIP && md->proposal != VEGAS_PROPOSE_RANDOM) return fail(h, VEGAS_ERR_INVALID, "unknown proposal");
    if (md->precision != VEGAS_F32 && md->precision != VEGAS_F64) return fail(h, VEGAS_ERR_INVALID, "unknown precision");
    return VEGAS_OK;
}

int run_fill(vegas_gpu* h, int up);
int push_boundaries(vegas_gpu* h);

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* vegas_gpu_version(void) { return "vegas_gpu 0.1 (sm_100a)"; }

const char* vegas_gpu_last_error(vegas_gpu_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vegas_gpu_create_lattice(const vegas_model_desc* md, const vegas_lattice_desc* ldesc, vegas_gpu_t* out) {
    vegas_gpu* h = nullptr;
    if (!out) return fail(h, VEGAS_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int rc = check_model(h, md);
    if (rc) return rc;
    if (!ldesc || ldesc->unitcell < 0 || ldesc->unitcell > 2) return fail(h, VEGAS_ERR_INVALID, "unknown unit cell");
    if (ldesc->nx == 0 || ldesc->ny == 0 || ldesc->nz == 0) return fail(h, VEGAS_ERR_INVALID, "empty lattice");
    vegas_gpu* hh = new vegas_gpu();
    hh->md = *md;
    hh->structured = true;
    hh->ld.unitcell = ldesc->unitcell;
    hh->ld.nx = ldesc->nx; hh->ld.ny = ldesc->ny; hh->ld.nz = ldesc->nz;
    hh->ld.pbc[0] = ldesc->pbc_x != 0; hh->ld.pbc[1] = ldesc->pbc_y != 0; hh->ld.pbc[2] = ldesc->pbc_z != 0;
    hh->ld.literal = ldesc->literal_from_lattice_filter != 0;
    hh->nz_global = ldesc->nz_global ? ldesc->nz_global : ldesc->nz;
    hh->z_offset = ldesc->z_offset;
    hh->slab = hh->nz_global != ldesc->nz;
    hh->n = ldesc->nx * ldesc->ny * ldesc->nz * (uint64_t)vgl::basis_count(ldesc->unitcell);
    h = hh;
    auto bail = [&](int code) { std::string e = h->err; vegas_gpu_destroy(h); g_create_error = e; return code; };
    if ((rc = common_init(h))) return bail(rc);

    // ---- family selection: structured stencil needs sc, periodic even extents, no literal filter
    const uint64_t L[3] = {hh->ld.nx, hh->ld.ny, hh->nz_global};
    bool stencil = ldesc->unitcell == VEGAS_SC && !hh->ld.literal && !md->force_general;
    for (int a = 0; a < 2 && stencil; ++a) stencil = hh->ld.pbc[a] && L[a] >= 2 && (L[a] % 2 == 0);
    if (stencil) stencil = (L[2] == 1) || (hh->ld.pbc[2] && L[2] % 2 == 0);
    if (stencil && md->model == VEGAS_ISING) stencil = L[0] % 64 == 0;
    if (stencil && md->model == VEGAS_HEISENBERG) stencil = L[0] % 8 == 0;
    if (stencil && hh->ld.nx * hh->ld.ny * hh->ld.nz >= (1ull << 32) * (md->model == VEGAS_ISING ? 64 : 2)) stencil = false;  // 32-bit offsets per colour array
    const bool basis_family = md->model == VEGAS_HEISENBERG && ldesc->unitcell != VEGAS_SC && !hh->ld.literal && !md->force_general &&
                              hh->ld.pbc[0] && hh->ld.pbc[1] && hh->ld.pbc[2] && h->n < (1ull << 31);
    if (hh->slab) {
        if (!(basis_family || (stencil && L[2] != 1)))
            return bail(fail(h, VEGAS_ERR_INVALID, "z-slab decomposition needs the sc stencil path (periodic, even extents) or periodic Heisenberg bcc/fcc"));
        if (hh->z_offset + hh->ld.nz > hh->nz_global) return bail(fail(h, VEGAS_ERR_INVALID, "slab outside the global lattice"));
    }
    if (stencil) {
        h->family = md->model == VEGAS_ISING ? FAM_ISING_MSC : FAM_HEIS_STENCIL;
        h->ndim = L[2] == 1 ? 2 : 3;
        h->n_self = (L[2] == 1 && hh->ld.pbc[2]) ? 1 : 0;
        h->n_colours = 2;
        if (h->family == FAM_ISING_MSC) {
            const size_t bytes = msc_words(h) * sizeof(uint32_t);
            for (int c = 0; c < 2; ++c) { if (cudaMalloc(&h->msc[c], bytes) != cudaSuccess) return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (spins) failed")); }
            h->halo_plane_bytes = (size_t)(h->ld.nx / 64) * h->ld.ny * sizeof(uint32_t);
        } else {
            const size_t bytes = heis_colour_elems(h) * real_bytes(h);
            for (int c = 0; c < 2; ++c)
                for (int k = 0; k < 3; ++k)
                    if (cudaMalloc(&h->hs[c][k], bytes) != cudaSuccess) return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (spins) failed"));
            h->halo_plane_bytes = (size_t)(h->ld.nx / 2) * h->ld.ny * real_bytes(h);
        }
        if (h->slab) {
            const int ncomp = h->family == FAM_ISING_MSC ? 1 : 3;
            // ONE allocation holds the halos and the flags; it is sized in whole 2 MiB pages so that it is not
            // sub-allocated from a shared page and its CUDA IPC handle maps exactly [halo, halo + slab_bytes).
            h->halo_bytes = (h->halo_plane_bytes * 4 * ncomp + 255) / 256 * 256;
            h->slab_bytes = (h->halo_bytes + 256 + (2u << 20) - 1) / (2u << 20) * (2u << 20);
            if (cudaMalloc(&h->halo, h->slab_bytes) != cudaSuccess)
                return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (halo) failed"));
            h->flags = (unsigned long long*)((char*)h->halo + h->halo_bytes);
            cudaMemsetAsync(h->halo, 0, h->slab_bytes, h->stream);
            const unsigned long long magic = 0x76656761735f6770ull ^ h->z_offset;  // checked by the peer after mapping
            cudaMemcpyAsync(h->flags + 4, &magic, 8, cudaMemcpyHostToDevice, h->stream);
            cudaStreamSynchronize(h->stream);
        }
    } else if (basis_family) {
        // periodic bcc / fcc: basis-split arrays, colour = basis, compile-time neighbour tables (heis_basis.cuh)
        h->family = FAM_HEIS_BASIS;
        h->n_colours = vgl::basis_count(ldesc->unitcell);
        h->h_colour.resize(h->n);
        for (uint64_t i = 0; i < h->n; ++i) h->h_colour[i] = (uint8_t)(i % (uint64_t)h->n_colours);
        const size_t cells = (size_t)(h->n / h->n_colours), pl = (size_t)h->ld.nx * h->ld.ny;
        if (!h->slab) {
            for (int b = 0; b < h->n_colours; ++b)
                for (int k = 0; k < 3; ++k)
                    if (cudaMalloc(&h->hb[b][k], cells * real_bytes(h)) != cudaSuccess) return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (spins) failed"));
        } else {
            // ONE 2 MiB-granular allocation (its CUDA IPC handle maps exactly this range in the neighbour): the arrays
            // [basis][component], each with a halo plane below and above its nz local planes, then the flags
            // fcc: TWO such array sets, for the pair launches that step from one set into the other (heis_basis_pair_kernel)
            const size_t ext = cells + 2 * pl;
            h->halo_plane_bytes = pl * real_bytes(h);
            const size_t set_bytes = ((size_t)h->n_colours * 3 * ext * real_bytes(h) + 255) / 256 * 256;
            const int n_sets = ldesc->unitcell == VEGAS_FCC ? 2 : 1;
            h->pair_set_bytes = n_sets == 2 ? set_bytes : 0;
            h->halo_bytes = set_bytes * n_sets;
            h->slab_bytes = (h->halo_bytes + 256 + (2u << 20) - 1) / (2u << 20) * (2u << 20);
            if (cudaMalloc(&h->halo, h->slab_bytes) != cudaSuccess) return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (spins + halos) failed"));
            for (int b = 0; b < h->n_colours; ++b)
                for (int k = 0; k < 3; ++k) {
                    h->hb[b][k] = (char*)h->halo + ((size_t)(b * 3 + k) * ext + pl) * real_bytes(h);
                    if (n_sets == 2) h->hb2[b][k] = (char*)h->hb[b][k] + set_bytes;
                }
            h->flags = (unsigned long long*)((char*)h->halo + h->halo_bytes);
            cudaMemsetAsync(h->halo, 0, h->slab_bytes, h->stream);
            const unsigned long long magic = 0x76656761735f6770ull ^ h->z_offset;  // checked by the peer after mapping
            cudaMemcpyAsync(h->flags + 4, &magic, 8, cudaMemcpyHostToDevice, h->stream);
            cudaStreamSynchronize(h->stream);
        }
    } else {
        h->family = md->model == VEGAS_ISING ? FAM_ISING_GEN : FAM_HEIS_GEN;
        if (h->n >= (1ull << 32)) return bail(fail(h, VEGAS_ERR_INVALID, "general-adjacency path supports < 2^32 sites"));
        vgl::Colouring col(h->ld);
        h->n_colours = col.n_colours;
        std::vector<uint8_t> colour(h->n);
        for (uint64_t i = 0; i < h->n; ++i) colour[i] = (uint8_t)col.colour(i);
        h->h_colour = colour;
        if ((rc = alloc_general(h, colour))) return bail(rc);
    }
    if ((rc = run_fill(h, 1))) return bail(rc);
    *out = h;
    return VEGAS_OK;
}

int vegas_gpu_create_csr(const vegas_model_desc* md, const vegas_csr_desc* cd, vegas_gpu_t* out) {
    vegas_gpu* h = nullptr;
    if (!out) return fail(h, VEGAS_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int rc = check_model(h, md);
    if (rc) return rc;
    if (!cd || cd->n == 0 || !cd->row_ptr || (!cd->col_idx && cd->row_ptr[cd->n] > 0)) return fail(h, VEGAS_ERR_INVALID, "bad CSR descriptor");
    if (cd->n >= (1ull << 32)) return fail(h, VEGAS_ERR_INVALID, "CSR path supports < 2^32 sites");
    const uint64_t nnz = cd->row_ptr[cd->n];
    for (uint64_t i = 0; i < cd->n; ++i) {
        if (cd->row_ptr[i + 1] < cd->row_ptr[i]) return fail(h, VEGAS_ERR_INVALID, "row_ptr not monotone");
        if (!cd->values && cd->row_ptr[i + 1] - cd->row_ptr[i] > (uint64_t)ISING_ZMAX && md->model == VEGAS_ISING)
            return fail(h, VEGAS_ERR_INVALID, "uniform-J Ising rows longer than 32 are not supported");
    }
    for (uint64_t p = 0; p < nnz; ++p)
        if (cd->col_idx[p] >= cd->n) return fail(h, VEGAS_ERR_INVALID, "column index out of range");
    vegas_gpu* hh = new vegas_gpu();
    hh->md = *md;
    hh->csr_input = true;
    hh->n = cd->n;
    hh->family = md->model == VEGAS_ISING ? FAM_ISING_GEN : FAM_HEIS_GEN;
    hh->h_row_ptr.assign(cd->row_ptr, cd->row_ptr + cd->n + 1);
    hh->h_col.assign(cd->col_idx, cd->col_idx + nnz);
    if (cd->values) hh->h_val.assign(cd->values, cd->values + nnz);
    h = hh;
    auto bail = [&](int code) { std::string e = h->err; vegas_gpu_destroy(h); g_create_error = e; return code; };
    if ((rc = common_init(h))) return bail(rc);
    std::vector<uint8_t> colour;
    const int nc = vgl::greedy_colour(h->n, h->h_row_ptr.data(), h->h_col.data(), colour);
    if (nc < 0) return bail(fail(h, VEGAS_ERR_INVALID, "graph needs more than 64 colours"));
    h->n_colours = nc;
    h->h_colour = colour;
    if (cudaMalloc(&h->d_row_ptr, (h->n + 1) * 8) != cudaSuccess || cudaMalloc(&h->d_col, (nnz + 1) * 4) != cudaSuccess)
        return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (csr) failed"));
    cudaMemcpy(h->d_row_ptr, h->h_row_ptr.data(), (h->n + 1) * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_col, h->h_col.data(), nnz * 4, cudaMemcpyHostToDevice);
    if (cd->values) {
        if (cudaMalloc(&h->d_val, (nnz + 1) * 8) != cudaSuccess) return bail(fail(h, VEGAS_ERR_ALLOC, "cudaMalloc (csr) failed"));
        cudaMemcpy(h->d_val, h->h_val.data(), nnz * 8, cudaMemcpyHostToDevice);
    }
    if ((rc = alloc_general(h, colour))) return bail(rc);
    if ((rc = run_fill(h, 1))) return bail(rc);
    *out = h;
    return VEGAS_OK;
}

void vegas_gpu_destroy(vegas_gpu_t h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int c = 0; c < 2; ++c) {
        cudaFree(h->msc[c]);
        for (int k = 0; k < 3; ++k) { cudaFree(h->hs[c][k]); cudaFree(h->hs_alt[c][k]); }
    }
    if (h->peer_is_ipc) {
        if (h->peer_halo[0]) cudaIpcCloseMemHandle(h->peer_halo[0]);
        if (h->peer_halo[1] && h->peer_halo[1] != h->peer_halo[0]) cudaIpcCloseMemHandle(h->peer_halo[1]);
    }
    cudaFree(h->halo);
    if (!(h->family == FAM_HEIS_BASIS && h->slab)) for (int b = 0; b < 4; ++b) for (int k = 0; k < 3; ++k) cudaFree(h->hb[b][k]);
    cudaFree(h->g_s8);
    for (int k = 0; k < 3; ++k) cudaFree(h->g_s[k]);
    for (uint32_t* p : h->g_sites) cudaFree(p);
    cudaFree(h->d_row_ptr); cudaFree(h->d_col); cudaFree(h->d_val);
    cudaFree(h->g_thr); cudaFree(h->g_code);
    cudaFree(h->obs); cudaFree(h->slab_error);
    for (int k = 0; k < WAVE_MAX_STEPS; ++k) cudaFree(h->wave_units[k]);
    cudaFree(h->wave_done); cudaFree(h->wave_error);
    heis_pipe_destroy(h->pipe);
    basis_pipe_destroy(h->bpipe);
    if (h->hp_host) cudaFreeHost(h->hp_host);
    cudaFree(h->hp_dev);
    for (cudaEvent_t ev : h->hp_events) cudaEventDestroy(ev);
    basis_wave_destroy(h->bwave);
    if (!h->slab) for (int b = 0; b < 4; ++b) for (int k = 0; k < 3; ++k) cudaFree(h->hb2[b][k]);
    if (h->stream_b) { cudaStreamSynchronize(h->stream_b); cudaStreamDestroy(h->stream_b); }
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    if (h->ev_bnd) cudaEventDestroy(h->ev_bnd);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

uint64_t vegas_gpu_n_sites(vegas_gpu_t h) { return h ? h->n : 0; }
int vegas_gpu_n_colours(vegas_gpu_t h) { return h ? h->n_colours : 0; }
const char* vegas_gpu_kernel_family(vegas_gpu_t h) { return h ? FAMILY_NAME[h->family] : ""; }
void* vegas_gpu_stream(vegas_gpu_t h) { return h ? (void*)h->stream : nullptr; }
uint64_t vegas_gpu_launch_count(vegas_gpu_t h) { return h ? h->launches : 0; }

int vegas_gpu_adjacency(vegas_gpu_t h, uint64_t* n, uint64_t* nnz, uint64_t* row_ptr, uint32_t* col_idx, double* values) {
    if (!h) return VEGAS_ERR_INVALID;
    if (h->csr_input) {
        if (n) *n = h->n;
        if (nnz) *nnz = h->h_col.size();
        if (row_ptr) memcpy(row_ptr, h->h_row_ptr.data(), (h->n + 1) * 8);
        if (col_idx) memcpy(col_idx, h->h_col.data(), h->h_col.size() * 4);
        if (values) for (size_t p = 0; p < h->h_col.size(); ++p) values[p] = h->h_val.empty() ? h->md.exchange : h->h_val[p];
        return VEGAS_OK;
    }
    if (h->slab) return fail(h, VEGAS_ERR_STATE, "adjacency export is not available for a slab");
    if (h->n > (1ull << 27)) return fail(h, VEGAS_ERR_STATE, "adjacency export limited to 2^27 sites");
    std::vector<uint64_t> rp; std::vector<uint32_t> col; std::vector<double> val;
    vgl::build_csr(h->ld, h->md.exchange, rp, col, val);
    if (n) *n = h->n;
    if (nnz) *nnz = col.size();
    if (row_ptr) memcpy(row_ptr, rp.data(), rp.size() * 8);
    if (col_idx) memcpy(col_idx, col.data(), col.size() * 4);
    if (values) memcpy(values, val.data(), val.size() * 8);
    return VEGAS_OK;
}

static int desc_of(const vegas_lattice_desc* ld, vgl::Desc& d) {
    if (!ld || ld->unitcell < 0 || ld->unitcell > 2 || ld->nx == 0 || ld->ny == 0 || ld->nz == 0) return VEGAS_ERR_INVALID;
    d.unitcell = ld->unitcell; d.nx = ld->nx; d.ny = ld->ny; d.nz = ld->nz;
    d.pbc[0] = ld->pbc_x != 0; d.pbc[1] = ld->pbc_y != 0; d.pbc[2] = ld->pbc_z != 0;
    d.literal = ld->literal_from_lattice_filter != 0;
    return VEGAS_OK;
}

int vegas_gpu_lattice_adjacency(const vegas_lattice_desc* ld, double exchange, uint64_t* n, uint64_t* nnz,
                                uint64_t* row_ptr, uint32_t* col_idx, double* values) {
    vgl::Desc d{};
    if (desc_of(ld, d)) return fail(nullptr, VEGAS_ERR_INVALID, "bad lattice descriptor");
    const uint64_t sites = d.nx * d.ny * d.nz * (uint64_t)vgl::basis_count(d.unitcell);
    if (sites > (1ull << 27)) return fail(nullptr, VEGAS_ERR_INVALID, "adjacency export limited to 2^27 sites");
    std::vector<uint64_t> rp; std::vector<uint32_t> col; std::vector<double> val;
    vgl::build_csr(d, exchange, rp, col, val);
    if (n) *n = sites;
    if (nnz) *nnz = col.size();
    if (row_ptr) memcpy(row_ptr, rp.data(), rp.size() * 8);
    if (col_idx) memcpy(col_idx, col.data(), col.size() * 4);
    if (values) memcpy(values, val.data(), val.size() * 8);
    return VEGAS_OK;
}

int vegas_gpu_lattice_colours(const vegas_lattice_desc* ld, int* n_colours, uint8_t* colour_of_site) {
    vgl::Desc d{};
    if (desc_of(ld, d)) return fail(nullptr, VEGAS_ERR_INVALID, "bad lattice descriptor");
    vgl::Colouring col(d);
    if (n_colours) *n_colours = col.n_colours;
    const uint64_t sites = d.nx * d.ny * d.nz * (uint64_t)vgl::basis_count(d.unitcell);
    if (colour_of_site) for (uint64_t i = 0; i < sites; ++i) colour_of_site[i] = (uint8_t)col.colour(i);
    return VEGAS_OK;
}

int vegas_gpu_colours(vegas_gpu_t h, uint8_t* colour_of_site) {
    if (!h || !colour_of_site) return VEGAS_ERR_INVALID;
    if (h->family == FAM_ISING_MSC || h->family == FAM_HEIS_STENCIL) {
        if (h->n > (1ull << 30)) return fail(h, VEGAS_ERR_STATE, "colour export limited to 2^30 sites");
        for (uint64_t i = 0; i < h->n; ++i) {
            const uint64_t x = i % h->ld.nx, y = (i / h->ld.nx) % h->ld.ny, z = i / (h->ld.nx * h->ld.ny) + h->z_offset;
            colour_of_site[i] = (uint8_t)((x + y + z) & 1);
        }
        return VEGAS_OK;
    }
    memcpy(colour_of_site, h->h_colour.data(), h->n);
    return VEGAS_OK;
}

// ---- state I/O --------------------------------------------------------------------------
namespace {
// staging buffers of the host-packed transfer, allocated at first use; false: this handle takes the byte-per-spin path
bool host_pack_ready(vegas_gpu* h) {
    if (h->family != FAM_ISING_MSC || h->host_pack_min < 0 || (long)std::min<uint64_t>(h->n, 1ull << 62) < h->host_pack_min || h->n % 64) return false;
    if (h->hp_host && h->hp_dev) return true;
    const size_t bytes = (size_t)(h->n / 8);
    if (!h->hp_host && cudaHostAlloc((void**)&h->hp_host, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); h->hp_host = nullptr; return false; }
    if (!h->hp_dev && cudaMalloc((void**)&h->hp_dev, bytes) != cudaSuccess) { cudaGetLastError(); h->hp_dev = nullptr; return false; }
    return true;
}
}  // namespace

int vegas_gpu_upload_ising(vegas_gpu_t h, const int8_t* s, uint64_t n) {
    if (!h || !s) return VEGAS_ERR_INVALID;
    if (h->md.model != VEGAS_ISING || n != h->n) return fail(h, VEGAS_ERR_INVALID, "upload_ising: wrong model or size");
    CU(cudaSetDevice(h->device));
    if (h->family == FAM_ISING_GEN) {
        CU(cudaMemcpyAsync(h->g_s8, s, n, cudaMemcpyHostToDevice, h->stream));
    } else if (host_pack_ready(h)) {
        // big lattice: the host threads turn the State into a sign bitmap chunk by chunk (1/8 of the PCIe bytes); the copy of
        // a finished chunk overlaps the packing of the next ones; one kernel splits the bitmap into the colour arrays
        const size_t words = (size_t)(n / 32);
        int err = 0;
        try {
            host_chunked(words, (size_t)(h->host_pack_chunk / 32), host_pack_threads(),
                         [&](size_t first, size_t nw) { host_pack_signs(s + 32 * first, h->hp_host + first, nw); },
                         [&](size_t, size_t first, size_t nw) {
                             if (cudaMemcpyAsync(h->hp_dev + first, h->hp_host + first, nw * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) err = 1;
                         },
                         nullptr);
        } catch (const std::exception& e) {   // worker threads could not be started: nothing crosses the C ABI as an exception
            return fail(h, VEGAS_ERR_ALLOC, std::string("upload_ising: host packing failed: ") + e.what());
        }
        if (err) CU(cudaGetLastError());
        const uint32_t Wx = (uint32_t)(h->ld.nx / 64);
        const size_t total = (size_t)Wx * h->ld.ny * h->ld.nz;
        ising_msc_from_bitmap_kernel<<<cdiv(total, 256), 256, 0, h->stream>>>((const uint2*)h->hp_dev, h->msc[0], h->msc[1], Wx,
                                                                             (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
        h->launches++;
    } else {
        int8_t* tmp = nullptr;
        CU(cudaMallocAsync(&tmp, n, h->stream));
        CU(cudaMemcpyAsync(tmp, s, n, cudaMemcpyHostToDevice, h->stream));
        const uint32_t Wx = (uint32_t)(h->ld.nx / 64);
        const size_t total = (size_t)Wx * h->ld.ny * h->ld.nz;
        ising_msc_pack_kernel<<<cdiv(total, 256), 256, 0, h->stream>>>(tmp, h->msc[0], h->msc[1], Wx,
                                                                      (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
        h->launches++;
        CU(cudaFreeAsync(tmp, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return push_boundaries(h);
}

int vegas_gpu_download_ising(vegas_gpu_t h, int8_t* s, uint64_t n) {
    if (!h || !s) return VEGAS_ERR_INVALID;
    if (h->md.model != VEGAS_ISING || n != h->n) return fail(h, VEGAS_ERR_INVALID, "download_ising: wrong model or size");
    CU(cudaSetDevice(h->device));
    if (h->family == FAM_ISING_GEN) {
        CU(cudaMemcpyAsync(s, h->g_s8, n, cudaMemcpyDeviceToHost, h->stream));
    } else if (host_pack_ready(h)) {
        // the reverse: colour arrays -> bitmap on the device, chunked copies, the host threads write the int8 State of a chunk
        // as soon as its copy has landed (one event per chunk)
        const uint32_t Wx = (uint32_t)(h->ld.nx / 64);
        const size_t total = (size_t)Wx * h->ld.ny * h->ld.nz;
        ising_msc_to_bitmap_kernel<<<cdiv(total, 256), 256, 0, h->stream>>>((uint2*)h->hp_dev, h->msc[0], h->msc[1], Wx,
                                                                           (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
        h->launches++;
        const size_t words = (size_t)(n / 32), cw = (size_t)(h->host_pack_chunk / 32), n_chunks = (words + cw - 1) / cw;
        while (h->hp_events.size() < n_chunks) {
            cudaEvent_t ev;
            CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            h->hp_events.push_back(ev);
        }
        for (size_t c = 0; c < n_chunks; ++c) {
            const size_t first = c * cw, nw = std::min(cw, words - first);
            CU(cudaMemcpyAsync(h->hp_host + first, h->hp_dev + first, nw * 4, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaEventRecord(h->hp_events[c], h->stream));
        }
        int err = 0;
        try {
            host_chunked(words, cw, host_pack_threads(),
                         [&](size_t first, size_t nw) { host_unpack_signs(h->hp_host + first, s + 32 * first, nw); },
                         nullptr,
                         [&](size_t c) { if (cudaEventSynchronize(h->hp_events[c]) != cudaSuccess) err = 1; });
        } catch (const std::exception& e) {
            cudaStreamSynchronize(h->stream);
            return fail(h, VEGAS_ERR_ALLOC, std::string("download_ising: host unpacking failed: ") + e.what());
        }
        if (err) CU(cudaGetLastError());
    } else {
        int8_t* tmp = nullptr;
        CU(cudaMallocAsync(&tmp, n, h->stream));
        const uint32_t Wx = (uint32_t)(h->ld.nx / 64);
        const size_t total = (size_t)Wx * h->ld.ny * h->ld.nz;
        ising_msc_unpack_kernel<<<cdiv(total, 256), 256, 0, h->stream>>>(tmp, h->msc[0], h->msc[1], Wx,
                                                                        (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
        h->launches++;
        CU(cudaMemcpyAsync(s, tmp, n, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaFreeAsync(tmp, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return VEGAS_OK;
}

}  // extern "C"

namespace {

template <typename real>
int heis_upload_t(vegas_gpu* h, const double* dev_aos) {
    const uint32_t n = (uint32_t)h->n;
    if (h->family == FAM_HEIS_BASIS) {
        basis_pack_kernel<real><<<cdiv(h->n, 256), 256, 0, h->stream>>>(dev_aos, basis_ptrs<real>(h), (uint32_t)h->n_colours, (size_t)h->n);
    } else if (h->family == FAM_HEIS_GEN) {
        aos_to_soa_kernel<real><<<cdiv(n, 256), 256, 0, h->stream>>>(dev_aos, (real*)h->g_s[0], (real*)h->g_s[1], (real*)h->g_s[2], n);
    } else {
        heis_pack_kernel<real><<<cdiv(h->n, 256), 256, 0, h->stream>>>(dev_aos, (real*)h->hs[0][0], (real*)h->hs[0][1], (real*)h->hs[0][2],
                                                                      (real*)h->hs[1][0], (real*)h->hs[1][1], (real*)h->hs[1][2],
                                                                      (uint32_t)h->ld.nx, (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
    }
    h->launches++;
    return VEGAS_OK;
}

template <typename real>
int heis_download_t(vegas_gpu* h, double* dev_aos) {
    const uint32_t n = (uint32_t)h->n;
    if (h->family == FAM_HEIS_BASIS) {
        basis_unpack_kernel<real, double><<<cdiv(h->n, 256), 256, 0, h->stream>>>(dev_aos, dev_aos + 1, dev_aos + 2, 3, basis_ptrs<real>(h),
                                                                                 (uint32_t)h->n_colours, (size_t)h->n);
    } else if (h->family == FAM_HEIS_GEN) {
        soa_to_aos_kernel<real><<<cdiv(n, 256), 256, 0, h->stream>>>(dev_aos, (const real*)h->g_s[0], (const real*)h->g_s[1], (const real*)h->g_s[2], n);
    } else {
        heis_unpack_kernel<real, double><<<cdiv(h->n, 256), 256, 0, h->stream>>>(
            dev_aos, dev_aos + 1, dev_aos + 2, 3, (const real*)h->hs[0][0], (const real*)h->hs[0][1], (const real*)h->hs[0][2],
            (const real*)h->hs[1][0], (const real*)h->hs[1][1], (const real*)h->hs[1][2], (uint32_t)h->ld.nx, (uint32_t)h->ld.ny,
            (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
    }
    h->launches++;
    return VEGAS_OK;
}

int run_fill(vegas_gpu* h, int up) {
    CU(cudaSetDevice(h->device));
    if (h->family == FAM_ISING_MSC) {
        for (int c = 0; c < 2; ++c) CU(cudaMemsetAsync(h->msc[c], up ? 0xFF : 0x00, msc_words(h) * sizeof(uint32_t), h->stream));
    } else if (h->family == FAM_ISING_GEN) {
        CU(cudaMemsetAsync(h->g_s8, up ? 0x01 : 0xFF, h->n, h->stream));
    } else {
        const bool f64 = h->md.precision == VEGAS_F64;
        auto fill = [&](void* p, size_t cnt, double v) {
            if (f64) fill_kernel<double><<<cdiv(cnt, 256), 256, 0, h->stream>>>((double*)p, cnt, v);
            else fill_kernel<float><<<cdiv(cnt, 256), 256, 0, h->stream>>>((float*)p, cnt, (float)v);
            h->launches++;
        };
        if (h->family == FAM_HEIS_GEN) {
            fill(h->g_s[0], h->n, 0.0); fill(h->g_s[1], h->n, 0.0); fill(h->g_s[2], h->n, up ? 1.0 : -1.0);
        } else if (h->family == FAM_HEIS_BASIS) {
            const size_t cells = h->n / h->n_colours;
            for (int b = 0; b < h->n_colours; ++b) { fill(h->hb[b][0], cells, 0.0); fill(h->hb[b][1], cells, 0.0); fill(h->hb[b][2], cells, up ? 1.0 : -1.0); }
        } else {
            for (int c = 0; c < 2; ++c) {
                fill(h->hs[c][0], heis_colour_elems(h), 0.0); fill(h->hs[c][1], heis_colour_elems(h), 0.0);
                fill(h->hs[c][2], heis_colour_elems(h), up ? 1.0 : -1.0);
            }
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return VEGAS_OK;
}

// Fill the halos of a connected slab from the neighbours' current boundary planes is done by the
// neighbours themselves: each rank pushes its own boundary planes (both colours) to its peers.
__global__ void copy_plane_kernel(uint32_t* dst, const uint32_t* src, size_t n4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) dst[i] = src[i];
}

int push_boundaries(vegas_gpu* h) {
    if (!(h->slab && h->connected)) return VEGAS_OK;
    if (h->family == FAM_HEIS_BASIS) {
        // my first / last local plane of every array -> the lower neighbour's plane "nz" / the upper neighbour's plane "-1"
        const size_t n4 = h->halo_plane_bytes / 4, pl = (size_t)h->ld.nx * h->ld.ny, rb = real_bytes(h);
        const size_t ext = (size_t)(h->n / h->n_colours) + 2 * pl;
        for (int a = 0; a < h->n_colours * 3; ++a) {
            const char* mine = (const char*)h->hb[a / 3][a % 3];
            char* lo_dst = peer_set(h, 0, h->cur_set) + ((size_t)a * ext + (size_t)(h->ld.nz + 1) * pl) * rb;
            char* hi_dst = peer_set(h, 1, h->cur_set) + ((size_t)a * ext) * rb;
            copy_plane_kernel<<<cdiv(n4, 256), 256, 0, h->stream>>>((uint32_t*)lo_dst, (const uint32_t*)mine, n4);
            copy_plane_kernel<<<cdiv(n4, 256), 256, 0, h->stream>>>((uint32_t*)hi_dst, (const uint32_t*)(mine + (size_t)(h->ld.nz - 1) * pl * rb), n4);
            h->launches += 2;
        }
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
        return VEGAS_OK;
    }
    const int ncomp = h->family == FAM_ISING_MSC ? 1 : 3;
    const size_t n4 = h->halo_plane_bytes / 4;
    for (int colour = 0; colour < 2; ++colour)
        for (int c = 0; c < ncomp; ++c) {
            const char* base = h->family == FAM_ISING_MSC ? (const char*)h->msc[colour] : (const char*)h->hs[colour][c];
            const char* first = base;
            const char* last = base + (size_t)(h->ld.nz - 1) * h->halo_plane_bytes;
            copy_plane_kernel<<<cdiv(n4, 256), 256, 0, h->stream>>>((uint32_t*)((char*)h->peer_halo[0] + halo_offset(h, colour, 1, c)), (const uint32_t*)first, n4);
            copy_plane_kernel<<<cdiv(n4, 256), 256, 0, h->stream>>>((uint32_t*)((char*)h->peer_halo[1] + halo_offset(h, colour, 0, c)), (const uint32_t*)last, n4);
            h->launches += 2;
        }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return VEGAS_OK;
}

}  // namespace

extern "C" {

int vegas_gpu_upload_heisenberg(vegas_gpu_t h, const double* sxyz, uint64_t n) {
    if (!h || !sxyz) return VEGAS_ERR_INVALID;
    if (h->md.model != VEGAS_HEISENBERG || n != h->n) return fail(h, VEGAS_ERR_INVALID, "upload_heisenberg: wrong model or size");
    CU(cudaSetDevice(h->device));
    double* tmp = nullptr;
    CU(cudaMallocAsync(&tmp, n * 24, h->stream));
    CU(cudaMemcpyAsync(tmp, sxyz, n * 24, cudaMemcpyHostToDevice, h->stream));
    if (h->md.precision == VEGAS_F64) heis_upload_t<double>(h, tmp); else heis_upload_t<float>(h, tmp);
    CU(cudaFreeAsync(tmp, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return push_boundaries(h);
}

int vegas_gpu_download_heisenberg(vegas_gpu_t h, double* sxyz, uint64_t n) {
    if (!h || !sxyz) return VEGAS_ERR_INVALID;
    if (h->md.model != VEGAS_HEISENBERG || n != h->n) return fail(h, VEGAS_ERR_INVALID, "download_heisenberg: wrong model or size");
    CU(cudaSetDevice(h->device));
    double* tmp = nullptr;
    CU(cudaMallocAsync(&tmp, n * 24, h->stream));
    if (h->md.precision == VEGAS_F64) heis_download_t<double>(h, tmp); else heis_download_t<float>(h, tmp);
    CU(cudaMemcpyAsync(sxyz, tmp, n * 24, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaFreeAsync(tmp, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return check_async_errors(h);
}

int vegas_gpu_randomize(vegas_gpu_t h) {
    if (!h) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    const PhiloxKey pk = make_philox_key(h->md.seed);
    h->launches++;
    if (h->family == FAM_ISING_MSC) {
        const size_t words = msc_words(h);
        const uint64_t woff = (uint64_t)h->z_offset * h->ld.ny * (h->ld.nx / 64);
        ising_msc_randomize_kernel<<<cdiv(words, 256), 256, 0, h->stream>>>(h->msc[0], h->msc[1], words, woff, pk);
    } else if (h->family == FAM_ISING_GEN) {
        ising_general_randomize_kernel<<<cdiv(h->n, 256), 256, 0, h->stream>>>(h->g_s8, (uint32_t)h->n, 0, pk);
    } else if (h->family == FAM_HEIS_BASIS) {
        if (h->md.precision == VEGAS_F64) basis_randomize_kernel<double><<<cdiv(h->n, 256), 256, 0, h->stream>>>(basis_ptrs<double>(h), (uint32_t)h->n_colours, (size_t)h->n, (uint64_t)h->z_offset * h->ld.ny * h->ld.nx * h->n_colours, pk);
        else basis_randomize_kernel<float><<<cdiv(h->n, 256), 256, 0, h->stream>>>(basis_ptrs<float>(h), (uint32_t)h->n_colours, (size_t)h->n, (uint64_t)h->z_offset * h->ld.ny * h->ld.nx * h->n_colours, pk);
    } else if (h->family == FAM_HEIS_GEN) {
        if (h->md.precision == VEGAS_F64) heis_general_randomize_kernel<double><<<cdiv(h->n, 256), 256, 0, h->stream>>>((double*)h->g_s[0], (double*)h->g_s[1], (double*)h->g_s[2], (uint32_t)h->n, 0, pk);
        else heis_general_randomize_kernel<float><<<cdiv(h->n, 256), 256, 0, h->stream>>>((float*)h->g_s[0], (float*)h->g_s[1], (float*)h->g_s[2], (uint32_t)h->n, 0, pk);
    } else {
        const uint32_t Lx = (uint32_t)h->ld.nx, Ly = (uint32_t)h->ld.ny, Lz = (uint32_t)h->ld.nz, zo = (uint32_t)h->z_offset;
        if (h->md.precision == VEGAS_F64)
            heis_stencil_randomize_kernel<double><<<cdiv(h->n, 256), 256, 0, h->stream>>>((double*)h->hs[0][0], (double*)h->hs[0][1], (double*)h->hs[0][2], (double*)h->hs[1][0], (double*)h->hs[1][1], (double*)h->hs[1][2], Lx, Ly, Lz, zo, pk);
        else
            heis_stencil_randomize_kernel<float><<<cdiv(h->n, 256), 256, 0, h->stream>>>((float*)h->hs[0][0], (float*)h->hs[0][1], (float*)h->hs[0][2], (float*)h->hs[1][0], (float*)h->hs[1][1], (float*)h->hs[1][2], Lx, Ly, Lz, zo, pk);
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return push_boundaries(h);
}

int vegas_gpu_fill(vegas_gpu_t h, int up) {
    if (!h) return VEGAS_ERR_INVALID;
    int rc = run_fill(h, up);
    if (rc) return rc;
    return push_boundaries(h);
}

// ---- thermostat -------------------------------------------------------------------------
int vegas_gpu_set_thermostat(vegas_gpu_t h, double temperature, const double field_dir[3], double field_mag) {
    if (!h) return VEGAS_ERR_INVALID;
    if (temperature != temperature) return fail(h, VEGAS_ERR_INVALID, "temperature is NaN");
    h->T = temperature < DBL_EPSILON ? DBL_EPSILON : temperature;  // src/thermostat.rs:30-34
    if (field_dir) { h->fdir[0] = field_dir[0]; h->fdir[1] = field_dir[1]; h->fdir[2] = field_dir[2]; }
    h->fmag = field_mag;
    h->tables_dirty = true;
    return VEGAS_OK;
}

int vegas_gpu_set_energy_convention(vegas_gpu_t h, int conv) {
    if (!h || conv < 0 || conv > 2) return VEGAS_ERR_INVALID;
    h->econv = conv;
    return VEGAS_OK;
}

// ---- the hot path -----------------------------------------------------------------------
int vegas_gpu_step_async(vegas_gpu_t h, uint64_t n_steps, int record) {
    if (!h) return VEGAS_ERR_INVALID;
    if (record && n_steps > OBS_CAP) return fail(h, VEGAS_ERR_INVALID, "step_async records at most 4096 steps per call");
    if (h->slab && !h->connected) return fail(h, VEGAS_ERR_STATE, "slab handle is not connected to its z-neighbours (vegas_gpu_slab_connect)");
    CU(cudaSetDevice(h->device));
    int rc = update_tables(h);
    if (rc) return rc;
    unsigned long long* scratch = h->obs + OBS_CAP * OBS_W;
    if (record) { CU(cudaMemsetAsync(h->obs, 0, n_steps * OBS_W * 8, h->stream)); h->obs_unread = n_steps; }
    if (resident_plan(h)) {  // small lattice: the whole batch is one launch with the State in shared memory
        uint64_t done = 0;
        while (done < n_steps) {
            const uint32_t batch = (uint32_t)std::min<uint64_t>(n_steps - done, OBS_CAP);
            rc = h->csr_input ? resident_steps(h, csr_nb(h), batch, record != 0) : resident_steps(h, structured_nb(h), batch, record != 0);
            if (rc) return rc;
            h->sweeps += batch;
            h->attempts += h->n * batch;
            done += batch;
        }
        CU(cudaGetLastError());
        return VEGAS_OK;
    }
    for (uint64_t s = 0; s < n_steps;) {
        if (h->family == FAM_HEIS_STENCIL && h->wave_k > 1 && !pipe_plan(h) && wave_plan(h) && h->wave_kmax > 1 && n_steps - s > 1) {
            // several steps per persistent launch: the later colour passes find the earlier ones' planes in L2
            const uint32_t k = (uint32_t)std::min<uint64_t>(h->wave_kmax, n_steps - s);
            double* row = (double*)(record ? h->obs + s * OBS_W : scratch);
            if (h->md.precision == VEGAS_F64) wave_steps_t<double>(h, k, row, record != 0); else wave_steps_t<float>(h, k, row, record != 0);
            h->sweeps += k;
            h->attempts += h->n * k;
            s += k;
            continue;
        }
        do_step(h, record ? (void*)(h->obs + s * OBS_W) : nullptr, scratch);
        ++s;
    }
    CU(cudaGetLastError());
    return VEGAS_OK;
}

int vegas_gpu_read_observables(vegas_gpu_t h, uint64_t n_steps, double* energy, double* mag_xyz) {
    if (!h || n_steps > OBS_CAP) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    std::vector<unsigned long long> host(n_steps * OBS_W);
    CU(cudaMemcpyAsync(host.data(), h->obs, host.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    { const int rc = check_async_errors(h); if (rc) return rc; }
    for (uint64_t s = 0; s < n_steps; ++s) {
        const Canon c = canon_of(h, host.data() + s * OBS_W);
        if (s < h->obs_unread) h->accepted += c.accepted;   // each recorded batch is folded into the counter once
        if (energy) energy[s] = energy_of(h, c);
        if (mag_xyz) { mag_xyz[3 * s] = c.m[0]; mag_xyz[3 * s + 1] = c.m[1]; mag_xyz[3 * s + 2] = c.m[2]; }
    }
    h->obs_unread = 0;
    return VEGAS_OK;
}

int vegas_gpu_step(vegas_gpu_t h, uint64_t n_steps, double* energy, double* mag_xyz) {
    if (!h) return VEGAS_ERR_INVALID;
    const bool rec = energy != nullptr || mag_xyz != nullptr;
    uint64_t done = 0;
    while (done < n_steps) {
        const uint64_t batch = std::min<uint64_t>(n_steps - done, rec ? OBS_CAP : 1024);
        int rc = vegas_gpu_step_async(h, batch, rec);
        if (rc) return rc;
        if (rec) {
            rc = vegas_gpu_read_observables(h, batch, energy ? energy + done : nullptr, mag_xyz ? mag_xyz + 3 * done : nullptr);
            if (rc) return rc;
        }
        done += batch;
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return check_async_errors(h);
}

int vegas_gpu_synchronize(vegas_gpu_t h) {
    if (!h) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return check_async_errors(h);
}

int vegas_gpu_step_host_ising(vegas_gpu_t h, int8_t* state, uint64_t n, double* energy, double* mag_xyz) {
    int rc = vegas_gpu_upload_ising(h, state, n);
    if (rc) return rc;
    double e, m[3];
    if ((rc = vegas_gpu_step(h, 1, &e, m))) return rc;
    if (energy) *energy = e;
    if (mag_xyz) { mag_xyz[0] = m[0]; mag_xyz[1] = m[1]; mag_xyz[2] = m[2]; }
    return vegas_gpu_download_ising(h, state, n);
}

int vegas_gpu_step_host_heisenberg(vegas_gpu_t h, double* sxyz, uint64_t n, double* energy, double* mag_xyz) {
    int rc = vegas_gpu_upload_heisenberg(h, sxyz, n);
    if (rc) return rc;
    double e, m[3];
    if ((rc = vegas_gpu_step(h, 1, &e, m))) return rc;
    if (energy) *energy = e;
    if (mag_xyz) { mag_xyz[0] = m[0]; mag_xyz[1] = m[1]; mag_xyz[2] = m[2]; }
    return vegas_gpu_download_heisenberg(h, sxyz, n);
}

// ---- deterministic parity entry points ----------------------------------------------------
int vegas_gpu_total_energy(vegas_gpu_t h, double* out) {
    if (!h || !out) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    Canon c;
    int rc = measure_now(h, c);
    if (rc) return rc;
    *out = energy_of(h, c);
    return VEGAS_OK;
}

int vegas_gpu_magnetization(vegas_gpu_t h, double out_xyz[3]) {
    if (!h || !out_xyz) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    Canon c;
    int rc = measure_now(h, c);
    if (rc) return rc;
    out_xyz[0] = c.m[0]; out_xyz[1] = c.m[1]; out_xyz[2] = c.m[2];
    return VEGAS_OK;
}

}  // extern "C"

namespace {

// Per-site energies run on natural-order copies of the state with the implicit / CSR adjacency.
template <typename NB>
int site_energy_run(vegas_gpu* h, const NB& nb, const void* proposal, int want_delta, double* out_host) {
    const uint32_t n = (uint32_t)h->n;
    double *d_out = nullptr, *d_prop = nullptr;
    CU(cudaMallocAsync(&d_out, (size_t)n * 8, h->stream));
    const EnergyParams ep = energy_params(h);
    const int flip = proposal == nullptr;
    if (want_delta && proposal) {
        CU(cudaMallocAsync(&d_prop, (size_t)n * 24, h->stream));
        if (h->md.model == VEGAS_ISING) {
            std::vector<double> p3((size_t)n * 3, 0.0);
            const int8_t* p8 = (const int8_t*)proposal;
            for (uint32_t i = 0; i < n; ++i) p3[3 * (size_t)i + 2] = p8[i] > 0 ? 1.0 : -1.0;
            CU(cudaMemcpyAsync(d_prop, p3.data(), (size_t)n * 24, cudaMemcpyHostToDevice, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        } else {
            CU(cudaMemcpyAsync(d_prop, proposal, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream));
        }
    }
    double* oe = want_delta ? nullptr : d_out;
    double* ode = want_delta ? d_out : nullptr;
    const dim3 grid(cdiv(n, 256));
    void* tmp[3] = {nullptr, nullptr, nullptr};
    h->launches++;
    if (h->md.model == VEGAS_ISING) {
        const int8_t* s8 = h->g_s8;
        if (h->family == FAM_ISING_MSC) {
            CU(cudaMallocAsync(&tmp[0], n, h->stream));
            const uint32_t Wx = (uint32_t)(h->ld.nx / 64);
            const size_t total = (size_t)Wx * h->ld.ny * h->ld.nz;
            ising_msc_unpack_kernel<<<cdiv(total, 256), 256, 0, h->stream>>>((int8_t*)tmp[0], h->msc[0], h->msc[1], Wx, (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
            s8 = (const int8_t*)tmp[0];
        }
        IsingSpins sp{s8};
        site_energy_kernel<NB, IsingSpins><<<grid, 256, 0, h->stream>>>(nb, sp, n, ep, d_prop, flip, oe, ode);
    } else if (h->md.precision == VEGAS_F64) {
        const double* s[3] = {(const double*)h->g_s[0], (const double*)h->g_s[1], (const double*)h->g_s[2]};
        if (h->family == FAM_HEIS_BASIS) {
            for (int c = 0; c < 3; ++c) CU(cudaMallocAsync(&tmp[c], (size_t)n * 8, h->stream));
            basis_unpack_kernel<double, double><<<cdiv(n, 256), 256, 0, h->stream>>>((double*)tmp[0], (double*)tmp[1], (double*)tmp[2], 1,
                                                                                    basis_ptrs<double>(h), (uint32_t)h->n_colours, (size_t)n);
            for (int c = 0; c < 3; ++c) s[c] = (const double*)tmp[c];
        }
        if (h->family == FAM_HEIS_STENCIL) {
            for (int c = 0; c < 3; ++c) CU(cudaMallocAsync(&tmp[c], (size_t)n * 8, h->stream));
            heis_unpack_kernel<double, double><<<cdiv(n, 256), 256, 0, h->stream>>>((double*)tmp[0], (double*)tmp[1], (double*)tmp[2], 1,
                (const double*)h->hs[0][0], (const double*)h->hs[0][1], (const double*)h->hs[0][2], (const double*)h->hs[1][0], (const double*)h->hs[1][1], (const double*)h->hs[1][2],
                (uint32_t)h->ld.nx, (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
            for (int c = 0; c < 3; ++c) s[c] = (const double*)tmp[c];
        }
        HeisSpins<double> sp{s[0], s[1], s[2]};
        site_energy_kernel<NB, HeisSpins<double>><<<grid, 256, 0, h->stream>>>(nb, sp, n, ep, d_prop, flip, oe, ode);
    } else {
        const float* s[3] = {(const float*)h->g_s[0], (const float*)h->g_s[1], (const float*)h->g_s[2]};
        if (h->family == FAM_HEIS_BASIS) {
            for (int c = 0; c < 3; ++c) CU(cudaMallocAsync(&tmp[c], (size_t)n * 4, h->stream));
            basis_unpack_kernel<float, float><<<cdiv(n, 256), 256, 0, h->stream>>>((float*)tmp[0], (float*)tmp[1], (float*)tmp[2], 1,
                                                                                  basis_ptrs<float>(h), (uint32_t)h->n_colours, (size_t)n);
            for (int c = 0; c < 3; ++c) s[c] = (const float*)tmp[c];
        }
        if (h->family == FAM_HEIS_STENCIL) {
            for (int c = 0; c < 3; ++c) CU(cudaMallocAsync(&tmp[c], (size_t)n * 4, h->stream));
            heis_unpack_kernel<float, float><<<cdiv(n, 256), 256, 0, h->stream>>>((float*)tmp[0], (float*)tmp[1], (float*)tmp[2], 1,
                (const float*)h->hs[0][0], (const float*)h->hs[0][1], (const float*)h->hs[0][2], (const float*)h->hs[1][0], (const float*)h->hs[1][1], (const float*)h->hs[1][2],
                (uint32_t)h->ld.nx, (uint32_t)h->ld.ny, (uint32_t)h->ld.nz, (uint32_t)h->z_offset);
            for (int c = 0; c < 3; ++c) s[c] = (const float*)tmp[c];
        }
        HeisSpins<float> sp{s[0], s[1], s[2]};
        site_energy_kernel<NB, HeisSpins<float>><<<grid, 256, 0, h->stream>>>(nb, sp, n, ep, d_prop, flip, oe, ode);
    }
    CU(cudaMemcpyAsync(out_host, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    for (int c = 0; c < 3; ++c) if (tmp[c]) CU(cudaFreeAsync(tmp[c], h->stream));
    if (d_prop) CU(cudaFreeAsync(d_prop, h->stream));
    CU(cudaFreeAsync(d_out, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return VEGAS_OK;
}

int site_energy_dispatch(vegas_gpu* h, const void* proposal, int want_delta, double* out) {
    if (h->slab) return fail(h, VEGAS_ERR_STATE, "per-site energies are not available for a slab");
    if (h->n >= (1ull << 32)) return fail(h, VEGAS_ERR_STATE, "per-site energies support < 2^32 sites");
    CU(cudaSetDevice(h->device));
    if (h->csr_input) return site_energy_run(h, csr_nb(h), proposal, want_delta, out);
    return site_energy_run(h, structured_nb(h), proposal, want_delta, out);
}

}  // namespace

extern "C" {

int vegas_gpu_site_energies(vegas_gpu_t h, double* out_n) {
    if (!h || !out_n) return VEGAS_ERR_INVALID;
    return site_energy_dispatch(h, nullptr, 0, out_n);
}

int vegas_gpu_delta_energies(vegas_gpu_t h, const void* proposal, double* out_n) {
    if (!h || !out_n) return VEGAS_ERR_INVALID;
    return site_energy_dispatch(h, proposal, 1, out_n);
}

int vegas_gpu_attempt_count(vegas_gpu_t h, uint64_t* attempts, uint64_t* accepted) {
    if (!h) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    unsigned long long row[OBS_W];
    CU(cudaMemcpyAsync(row, h->obs + OBS_CAP * OBS_W, sizeof row, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const Canon c = canon_of(h, row);
    if (attempts) *attempts = h->attempts;
    if (accepted) *accepted = h->accepted + c.accepted;
    return VEGAS_OK;
}

int vegas_gpu_sweep_count(vegas_gpu_t h, uint64_t* sweeps) {
    if (!h || !sweeps) return VEGAS_ERR_INVALID;
    *sweeps = h->sweeps;
    return VEGAS_OK;
}

int vegas_gpu_set_sweep_count(vegas_gpu_t h, uint64_t sweeps) {
    if (!h) return VEGAS_ERR_INVALID;
    h->sweeps = sweeps;
    return VEGAS_OK;
}

int vegas_gpu_ising_thresholds(vegas_gpu_t h, int* n_classes, uint64_t* thr, uint8_t* always) {
    if (!h) return VEGAS_ERR_INVALID;
    if (h->family != FAM_ISING_MSC) return fail(h, VEGAS_ERR_STATE, "thresholds are exposed for the ising_msc family only");
    CU(cudaSetDevice(h->device));
    int rc = update_tables(h);
    if (rc) return rc;
    if (n_classes) *n_classes = 2 * h->ndim + 1;
    if (thr) memcpy(thr, h->ising_thr.data(), 16 * 8);
    if (always) memcpy(always, h->ising_always.data(), 16);
    return VEGAS_OK;
}

// ---- slab decomposition -------------------------------------------------------------------
struct SlabBlob {
    cudaIpcMemHandle_t mem;      // the single halo+flags allocation
    int device;
    uint64_t plane_bytes, halo_bytes, slab_bytes, z_offset;
};
static_assert(sizeof(SlabBlob) <= VEGAS_IPC_BYTES, "blob size");

int vegas_gpu_slab_export(vegas_gpu_t h, void* blob) {
    if (!h || !blob) return VEGAS_ERR_INVALID;
    if (!h->slab) return fail(h, VEGAS_ERR_STATE, "handle is not a slab");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    SlabBlob b;
    memset(&b, 0, sizeof b);
    CU(cudaIpcGetMemHandle(&b.mem, h->halo));
    b.device = h->device;
    b.plane_bytes = h->halo_plane_bytes; b.halo_bytes = h->halo_bytes; b.slab_bytes = h->slab_bytes; b.z_offset = h->z_offset;
    memset(blob, 0, VEGAS_IPC_BYTES);
    memcpy(blob, &b, sizeof b);
    return VEGAS_OK;
}

int vegas_gpu_slab_connect(vegas_gpu_t h, const void* lower, const void* upper) {
    if (!h || !lower || !upper) return VEGAS_ERR_INVALID;
    if (!h->slab) return fail(h, VEGAS_ERR_STATE, "handle is not a slab");
    CU(cudaSetDevice(h->device));
    const void* blobs[2] = {lower, upper};
    const bool same = memcmp(lower, upper, VEGAS_IPC_BYTES) == 0;  // two ranks: both neighbours are the same peer
    for (int d = 0; d < 2; ++d) {
        if (d == 1 && same) { h->peer_halo[1] = h->peer_halo[0]; h->peer_flags[1] = h->peer_flags[0]; break; }
        SlabBlob b;
        memcpy(&b, blobs[d], sizeof b);
        if (b.plane_bytes != h->halo_plane_bytes || b.halo_bytes != h->halo_bytes)
            return fail(h, VEGAS_ERR_INVALID, "neighbour slab has a different plane size");
        CU(cudaIpcOpenMemHandle(&h->peer_halo[d], b.mem, cudaIpcMemLazyEnablePeerAccess));
        h->peer_flags[d] = (unsigned long long*)((char*)h->peer_halo[d] + b.halo_bytes);
        unsigned long long magic = 0;  // the mapping must start exactly at the neighbour's allocation
        CU(cudaMemcpy(&magic, h->peer_flags[d] + 4, 8, cudaMemcpyDeviceToHost));
        if (magic != (0x76656761735f6770ull ^ b.z_offset))
            return fail(h, VEGAS_ERR_CUDA, "CUDA IPC mapping of the neighbour's halo buffer is not where expected");
    }
    h->peer_is_ipc = true;
    h->peers_remote = true;     // one process per GPU (SlabBlob::device is the neighbour's ordinal in ITS process: not comparable)
    h->connected = true;
    h->pipe_slab_steps = 0;
    if (h->pipe_planned) { heis_pipe_destroy(h->pipe); h->pipe = nullptr; h->pipe_planned = false; }
    if (h->bwave_planned) { basis_wave_destroy(h->bwave); h->bwave = nullptr; h->bwave_planned = false; }
    CU(cudaMemset(h->flags + BWAVE_FLAG_WORD, 0, 8 * sizeof(unsigned long long)));   // the wave step counts from the connection on
    preload_slab_kernels(h);
    return push_boundaries(h);
}

int vegas_gpu_slab_connect_local(vegas_gpu_t h, vegas_gpu_t lower, vegas_gpu_t upper) {
    if (!h || !lower || !upper) return VEGAS_ERR_INVALID;
    if (!h->slab || !lower->slab || !upper->slab) return fail(h, VEGAS_ERR_STATE, "handle is not a slab");
    CU(cudaSetDevice(h->device));
    vegas_gpu* nbr[2] = {lower, upper};
    for (int d = 0; d < 2; ++d) {
        if (nbr[d]->halo_plane_bytes != h->halo_plane_bytes) return fail(h, VEGAS_ERR_INVALID, "neighbour slab has a different plane size");
        if (nbr[d]->device != h->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, h->device, nbr[d]->device));
            if (!can) return fail(h, VEGAS_ERR_CUDA, "no peer access between slab devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(nbr[d]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
        }
        h->peer_halo[d] = nbr[d]->halo;
        h->peer_flags[d] = nbr[d]->flags;
    }
    h->peer_is_ipc = false;
    h->peers_remote = lower->device != h->device && upper->device != h->device;
    h->connected = true;
    h->pipe_slab_steps = 0;
    if (h->pipe_planned) { heis_pipe_destroy(h->pipe); h->pipe = nullptr; h->pipe_planned = false; }
    if (h->bwave_planned) { basis_wave_destroy(h->bwave); h->bwave = nullptr; h->bwave_planned = false; }
    CU(cudaMemset(h->flags + BWAVE_FLAG_WORD, 0, 8 * sizeof(unsigned long long)));   // the wave step counts from the connection on
    preload_slab_kernels(h);
    return push_boundaries(h);
}

int vegas_gpu_check_basis_tables(void) {
    return check_basis_table<1, 0>() + check_basis_table<1, 1>() + check_basis_table<2, 0>() + check_basis_table<2, 1>() +
           check_basis_table<2, 2>() + check_basis_table<2, 3>();
}

int vegas_gpu_wave_schedule(uint32_t n_chunks, uint32_t lag, uint32_t steps, uint32_t* units, uint64_t capacity, uint64_t* count) {
    if (n_chunks == 0 || n_chunks >= (1u << 24) || steps == 0 || steps > (uint32_t)WAVE_MAX_STEPS || !count) return VEGAS_ERR_INVALID;
    const std::vector<uint32_t> u = wave_units_for(n_chunks, std::max<uint32_t>(3, lag), steps);
    *count = u.size();
    if (units) {
        if (capacity < u.size()) return VEGAS_ERR_INVALID;
        std::memcpy(units, u.data(), u.size() * sizeof(uint32_t));
    }
    return VEGAS_OK;
}

int vegas_gpu_basis_pair_structure(int unitcell) {
    if (unitcell == VEGAS_FCC) return pair_structure_ok<2>() ? 1 : 0;
    if (unitcell == VEGAS_BCC) return pair_structure_ok<1>() ? 1 : 0;
    return 0;
}

int vegas_gpu_basis_wave_schedule(int unitcell, uint32_t nz, uint32_t lag, uint32_t* units, uint64_t capacity, uint64_t* count, uint32_t need[4]) {
    if ((unitcell != VEGAS_BCC && unitcell != VEGAS_FCC) || nz == 0 || nz >= (1u << 24) || lag == 0 || !count || !need) return VEGAS_ERR_INVALID;
    uint32_t nd[4];
    const std::vector<uint32_t> u = basis_wave_units(unitcell == VEGAS_BCC ? 1 : 2, nz, lag, nd);
    for (int b = 0; b < 4; ++b) need[b] = nd[b];
    *count = u.size();
    if (units) {
        if (capacity < u.size()) return VEGAS_ERR_INVALID;
        std::memcpy(units, u.data(), u.size() * sizeof(uint32_t));
    }
    return VEGAS_OK;
}

uint64_t vegas_gpu_state_transfer_bytes(vegas_gpu_t h) {
    if (!h) return 0;
    if (h->md.model == VEGAS_HEISENBERG) return h->n * 24;
    if (cudaSetDevice(h->device) == cudaSuccess && host_pack_ready(h)) return h->n / 8;
    return h->n;
}

int vegas_gpu_host_pack(const int8_t* s, uint32_t* words, uint64_t n_words, int threads, uint64_t chunk_words) {
    if (!s || !words || threads < 1) return VEGAS_ERR_INVALID;
    host_chunked((size_t)n_words, (size_t)std::max<uint64_t>(1, chunk_words), (unsigned)threads,
                 [&](size_t first, size_t nw) { host_pack_signs(s + 32 * first, words + first, nw); }, nullptr, nullptr);
    return VEGAS_OK;
}
int vegas_gpu_host_unpack(const uint32_t* words, int8_t* s, uint64_t n_words, int threads, uint64_t chunk_words) {
    if (!s || !words || threads < 1) return VEGAS_ERR_INVALID;
    host_chunked((size_t)n_words, (size_t)std::max<uint64_t>(1, chunk_words), (unsigned)threads,
                 [&](size_t first, size_t nw) { host_unpack_signs(words + first, s + 32 * first, nw); }, nullptr, nullptr);
    return VEGAS_OK;
}

// ---- tuning knobs ---------------------------------------------------------------------------
int vegas_gpu_set_tuning(vegas_gpu_t h, const char* key, long value) {
    if (!h || !key) return VEGAS_ERR_INVALID;
    const std::string k(key);
    if (k == "heis_fused") h->fused_enable = (int)value;
    else if (k == "heis_fused_ty") h->fused_ty = (uint32_t)value;
    else if (k == "heis_fused_cz") h->fused_cz = (uint32_t)value;
    else if (k == "heis_wave_c") h->wave_c = (uint32_t)value;
    else if (k == "heis_wave") h->wave_enable = (int)value;
    else if (k == "heis_wave_planes") h->wave_planes = (uint32_t)value;
    else if (k == "heis_wave_lag") h->wave_lag = (uint32_t)value;
    else if (k == "heis_wave_steps") h->wave_k = (uint32_t)std::max<long>(1, std::min<long>(value, WAVE_MAX_STEPS));
    else if (k == "heis_pipe") h->pipe_enable = (int)value;
    else if (k == "heis_pipe_stages") h->pipe_stages_other = (uint32_t)value;
    else if (k == "heis_pipe_own") h->pipe_stages_own = (uint32_t)value;
    else if (k == "heis_pipe_tiles") h->pipe_tiles = (uint32_t)value;
    else if (k == "heis_pipe_vec") h->pipe_vec = (uint32_t)value;
    else if (k == "heis_pipe_lead") h->pipe_lead = (uint32_t)value;
    else if (k == "heis_pipe_pub") h->pipe_pub = (uint32_t)value;
    else if (k == "heis_pipe_l2") h->pipe_l2 = (uint32_t)value;
    else if (k == "heis_pipe_backoff") h->pipe_backoff_c = (uint32_t)value;
    else if (k == "heis_pipe_backoff_helper") h->pipe_backoff_h = (uint32_t)value;
    else if (k == "basis_pipe") h->bpipe_enable = (int)value;
    else if (k == "basis_pipe_lead") h->bpipe_lead = (uint32_t)value;
    else if (k == "basis_pipe_pub") h->bpipe_pub = (uint32_t)value;
    else if (k == "basis_pipe_tiles") h->bpipe_tiles = (uint32_t)value;
    else if (k == "basis_pair") h->bpair_enable = (int)value;
    else if (k == "basis_pair_rows") h->bpair_rows = (uint32_t)value;
    else if (k == "basis_pair_chunk") h->bpair_chunk = (uint32_t)value;
    else if (k == "basis_wave") h->bwave_enable = (int)value;
    else if (k == "basis_wave_lag") h->bwave_lag = (uint32_t)value;
    else if (k == "basis_wave_ipt") h->bwave_ipt = (uint32_t)value;
    else if (k == "basis_wave_grid") h->bwave_grid = (uint32_t)value;
    else if (k == "basis_vec") h->basis_vec = (int)value;
    else if (k == "msc_full") h->msc_full = (int)value;
    else if (k == "host_pack_min") h->host_pack_min = value;
    else if (k == "host_pack_chunk") h->host_pack_chunk = (uint64_t)std::max<long>(64, value) / 64 * 64;
    else if (k == "resident_max") { h->resident_max = (uint32_t)value; h->resident_cols = -2; }
    else return fail(h, VEGAS_ERR_INVALID, "unknown tuning key: " + k);
    h->fused_ready = false;  // re-plan at the next step
    h->wave_ready = false;
    if (h->pipe_planned) { cudaStreamSynchronize(h->stream); heis_pipe_destroy(h->pipe); h->pipe = nullptr; h->pipe_planned = false; }
    if (h->bpipe_planned) { cudaStreamSynchronize(h->stream); basis_pipe_destroy(h->bpipe); h->bpipe = nullptr; h->bpipe_planned = false; }
    // a connected slab keeps its wave state: the neighbours' flag words count its steps
    if (h->bwave_planned && !(h->slab && h->bwave)) { cudaStreamSynchronize(h->stream); basis_wave_destroy(h->bwave); h->bwave = nullptr; h->bwave_planned = false; }
    return VEGAS_OK;
}

const char* vegas_gpu_step_kernel(vegas_gpu_t h) {
    if (!h) return "";
    if (h->family == FAM_HEIS_STENCIL && cudaSetDevice(h->device) == cudaSuccess && pipe_plan(h)) return "heis_pipe";
    if (h->family == FAM_HEIS_BASIS && cudaSetDevice(h->device) == cudaSuccess && bpipe_plan(h)) return "basis_pipe";
    if (h->family == FAM_HEIS_BASIS && cudaSetDevice(h->device) == cudaSuccess && bwave_plan(h)) return "basis_wave";
    if (h->family == FAM_HEIS_BASIS && cudaSetDevice(h->device) == cudaSuccess && bpair_plan(h)) return "basis_pair";
    if (h->family == FAM_HEIS_STENCIL && cudaSetDevice(h->device) == cudaSuccess && wave_plan(h)) return "heis_wave";
    if (h->family == FAM_HEIS_STENCIL && cudaSetDevice(h->device) == cudaSuccess && fused_plan(h)) return "heis_fused";
    if (resident_plan(h)) return h->family == FAM_ISING_GEN ? "ising_resident" : "heis_resident";
    return FAMILY_NAME[h->family];
}

// ---- timing hooks -------------------------------------------------------------------------
int vegas_gpu_timer_start(vegas_gpu_t h) {
    if (!h) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev0, h->stream));
    return VEGAS_OK;
}

int vegas_gpu_timer_stop(vegas_gpu_t h, float* ms) {
    if (!h || !ms) return VEGAS_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return VEGAS_OK;
}

}  // extern "C"
