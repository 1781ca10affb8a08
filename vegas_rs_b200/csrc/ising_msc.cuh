// ising_msc.cuh -- K1: multi-spin-coded Ising checkerboard half-sweep for sc lattices (2D/3D, pbc).
//
// Replaces MetropolisFlipIntegrator::step / MetropolisIntegrator::step for IsingSpin
// (src/integrator.rs:109-138 / :66-92) with Exchange::energy + Zeeman::energy
// (src/energy.rs:194-206, :147-151) folded into an integer class table.
//
// Layout: one bit per spin (1 = Up), colours stored apart.  Colour c holds the sites with
// (x+y+z) % 2 == c; in row (y,z) the compact index xc maps to x = 2*xc + ((y+z+c)&1).
// A 32-bit word holds 32 consecutive xc; a thread owns one uint4 (128 spins of one colour).
// All neighbours of a word's spins live in 6 words of the other colour: same index (x-nb 1),
// same index shifted by one bit with a carry from the adjacent word (x-nb 2), rows y+-1 and
// planes z+-1 at the same index.
//
// Acceptance: every spin needs "U < thr(class)" for a 64-bit uniform U.  U is never built:
// random bit-planes (one Philox word = bit j of U for 32 spins) are compared MSB-first with
// the class thresholds until every spin of the word is decided (about 8-12 planes instead of
// 64 bits per spin).  Exact: the decision equals a full 64-bit compare.
#pragma once
#include "common.cuh"

namespace vg {

constexpr int MSC_MAX_SLOT = 14;
constexpr uint8_t MSC_ALWAYS = 0xFE;
constexpr uint8_t MSC_NEVER = 0xFF;

// thresholds grouped into "slots"; class (spin s, antiparallel count c) -> slot / always / never
struct MscTable {
    uint8_t slot_of[2][8];
};

struct MscGeom {
    uint32_t Gx;        // uint4 groups per row per colour = Lx / 256
    uint32_t Ly, Lz;    // local rows / planes
    uint32_t z_offset;  // global z of local plane 0
    uint32_t Ly_g;      // == Ly (rows are never decomposed)
};

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

template <int C>
__device__ __forceinline__ uint32_t class_mask(uint32_t n0, uint32_t n1, uint32_t n2) {
    return ((C & 1) ? n0 : ~n0) & ((C & 2) ? n1 : ~n1) & ((C & 4) ? n2 : ~n2);
}

// MODE 0: update; 1: update + fused energy/magnetisation reduction; 2: reduction only.
// obs[0] += sum over own sites of s_i * (sum_nb s_j)   (every bond once, bipartite)
// obs[1] += sum of s over both colours (own word after update + partner word)
// obs[2] += accepted moves
template <int NDIM, bool FIELD, int NSLOT, bool RANDPROP, int MODE>
__global__ void __launch_bounds__(256)
ising_msc_kernel(uint4* __restrict__ own, const uint4* __restrict__ oth, const uint4* __restrict__ oth_lo,
                 const uint4* __restrict__ oth_hi, uint4* __restrict__ peer_lo, uint4* __restrict__ peer_hi,
                 MscGeom g, int colour, uint32_t z_begin, uint32_t z_count, MscTable tab,
                 const uint4* __restrict__ thr_bits /* [NSLOT][16] uint4: 0/~0 masks of threshold bit-planes */,
                 uint64_t sweep, uint32_t k0, uint32_t k1, unsigned long long* __restrict__ obs) {
    constexpr int Z = 2 * NDIM;
    __shared__ uint4 s_bits[(MODE == 2) ? 1 : NSLOT * 16];
    __shared__ unsigned long long s_red[3 * 32];
    if (MODE != 2) {
        for (int i = threadIdx.x; i < NSLOT * 16; i += blockDim.x) s_bits[i] = thr_bits[i];
        __syncthreads();
    }

    const uint32_t rows = g.Ly * g.Gx;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < z_count * rows;
    unsigned long long acc[3] = {0ull, 0ull, 0ull};

    if (active) {
        const uint32_t zl = z_begin + t / rows;
        const uint32_t rem = t % rows;
        const uint32_t y = rem / g.Gx, gx = rem % g.Gx;
        const uint32_t zg = zl + g.z_offset;
        const uint32_t rp = (y + zg + (uint32_t)colour) & 1u;
        const size_t row = ((size_t)zl * g.Ly + y) * g.Gx;

        const uint4 s4 = own[row + gx];
        const uint4 n4 = oth[row + gx];
        const uint32_t gxc = rp ? (gx + 1 == g.Gx ? 0u : gx + 1) : (gx == 0 ? g.Gx - 1 : gx - 1);
        const uint32_t cw = reinterpret_cast<const uint32_t*>(oth)[(row + gxc) * 4 + (rp ? 0 : 3)];
        const uint32_t ym = y == 0 ? g.Ly - 1 : y - 1, yp = y + 1 == g.Ly ? 0 : y + 1;
        const uint4 a4 = oth[((size_t)zl * g.Ly + ym) * g.Gx + gx];
        const uint4 b4 = oth[((size_t)zl * g.Ly + yp) * g.Gx + gx];
        uint4 c4 = make_uint4(0, 0, 0, 0), d4 = make_uint4(0, 0, 0, 0);
        if (NDIM == 3) {
            c4 = zl == 0 ? oth_lo[(size_t)y * g.Gx + gx] : oth[row - rows + gx];
            d4 = zl + 1 == g.Lz ? oth_hi[(size_t)y * g.Gx + gx] : oth[row + rows + gx];
        }
        const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w}, nw[4] = {n4.x, n4.y, n4.z, n4.w};
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
        const uint32_t cwz[4] = {c4.x, c4.y, c4.z, c4.w}, dwz[4] = {d4.x, d4.y, d4.z, d4.w};
        uint32_t out[4];

        // global index of this thread's first 32-bit word inside the colour array (RNG key)
        const uint64_t wbase = ((((uint64_t)zg * g.Ly_g + y) * g.Gx + gx) << 2) | ((uint64_t)colour << 62);

#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t s = sw[k], N0 = nw[k];
            const uint32_t prev = k > 0 ? nw[k > 0 ? k - 1 : 0] : cw;
            const uint32_t next = k < 3 ? nw[k < 3 ? k + 1 : 3] : cw;
            const uint32_t Nsh = rp ? __funnelshift_r(N0, next, 1) : __funnelshift_l(prev, N0, 1);
            // antiparallel indicators and their bit-sliced count n2 n1 n0
            const uint32_t a1 = s ^ N0, a2 = s ^ Nsh, a3 = s ^ aw[k], a4b = s ^ bw[k];
            uint32_t n0, n1, n2;
            if (NDIM == 3) {
                const uint32_t a5 = s ^ cwz[k], a6 = s ^ dwz[k];
                const uint32_t s1 = a1 ^ a2 ^ a3, c1 = maj3(a1, a2, a3);
                const uint32_t s2 = a4b ^ a5 ^ a6, c2 = maj3(a4b, a5, a6);
                n0 = s1 ^ s2;
                const uint32_t c3 = s1 & s2;
                n1 = c1 ^ c2 ^ c3;
                n2 = maj3(c1, c2, c3);
            } else {
                const uint32_t s1 = a1 ^ a2 ^ a3, c1 = maj3(a1, a2, a3);
                n0 = s1 ^ a4b;
                const uint32_t c3 = s1 & a4b;
                n1 = c1 ^ c3;
                n2 = c1 & c3;
            }
            uint32_t flip = 0;
            if (MODE != 2) {
                uint32_t accept = 0, pm[NSLOT];
#pragma unroll
                for (int q = 0; q < NSLOT; ++q) pm[q] = 0;
                auto assign = [&](uint32_t m, uint32_t slot) {
                    if (slot == MSC_ALWAYS) accept |= m;
#pragma unroll
                    for (int q = 0; q < NSLOT; ++q)
                        if (slot == (uint32_t)q) pm[q] |= m;
                };
                auto per_class = [&](uint32_t m, int c) {
                    if (FIELD) {
                        assign(m & ~s, tab.slot_of[0][c]);
                        assign(m & s, tab.slot_of[1][c]);
                    } else {
                        assign(m, tab.slot_of[0][c]);
                    }
                };
                per_class(class_mask<0>(n0, n1, n2), 0);
                per_class(class_mask<1>(n0, n1, n2), 1);
                per_class(class_mask<2>(n0, n1, n2), 2);
                per_class(class_mask<3>(n0, n1, n2), 3);
                per_class(class_mask<4>(n0, n1, n2), 4);
                if (NDIM == 3) {
                    per_class(class_mask<5>(n0, n1, n2), 5);
                    per_class(class_mask<6>(n0, n1, n2), 6);
                }
                uint32_t cand = 0xFFFFFFFFu;
                if (RANDPROP) {  // IsingSpin::rand (src/state.rs:76-84): proposed spin is a fair coin
                    uint32_t r[4];
                    philox_at(wbase + k, sweep, 0xFFu, k0, k1, r);
                    cand = r[0] ^ s;  // proposal differs from the current spin
                }
                uint32_t eq = 0, lt = 0;
#pragma unroll
                for (int q = 0; q < NSLOT; ++q) eq |= pm[q];
                eq &= cand;
                for (uint32_t ch = 0; ch < 16u && eq != 0u; ++ch) {
                    uint32_t r[4];
                    philox_at(wbase + k, sweep, ch, k0, k1, r);
                    uint32_t tb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int q = 0; q < NSLOT; ++q) {
                        const uint4 b = s_bits[q * 16 + ch];
                        tb[0] |= pm[q] & b.x; tb[1] |= pm[q] & b.y; tb[2] |= pm[q] & b.z; tb[3] |= pm[q] & b.w;
                    }
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const uint32_t tt = eq & (tb[p] ^ r[p]);  // undecided spins whose U bit differs from thr bit
                        lt |= tt & tb[p];                        // thr bit 1, U bit 0  ->  U < thr
                        eq ^= tt;
                    }
                }
                flip = (accept | lt) & cand;
                acc[2] += (unsigned long long)__popc(RANDPROP ? ((accept | lt) | ~cand) : flip);
            }
            const uint32_t snew = s ^ flip;
            out[k] = snew;
            if (MODE != 0) {
                // sum of final antiparallel counts: c' = flip ? Z - c : c
                const int s_all = __popc(n0) + 2 * __popc(n1) + 4 * __popc(n2);
                const int s_f = __popc(n0 & flip) + 2 * __popc(n1 & flip) + 4 * __popc(n2 & flip);
                const int cfin = s_all - 2 * s_f + Z * __popc(flip);
                acc[0] += (unsigned long long)(long long)(Z * 32 - 2 * cfin);
                acc[1] += (unsigned long long)(long long)(2 * __popc(snew) - 32 + 2 * __popc(N0) - 32);
            }
        }
        if (MODE != 2) {
            const uint4 o4 = make_uint4(out[0], out[1], out[2], out[3]);
            own[row + gx] = o4;
            if (NDIM == 3) {
                if (peer_lo != nullptr && zl == 0) peer_lo[(size_t)y * g.Gx + gx] = o4;
                if (peer_hi != nullptr && zl + 1 == g.Lz) peer_hi[(size_t)y * g.Gx + gx] = o4;
            }
        }
    }
    block_atomic_add<unsigned long long, 3>(acc, s_red, obs);
}

// ---------------------------------------------------------------------------------------
// K7: reference host layout (int8 +1/-1 per site, natural order) <-> bit-packed colours.
// One thread owns the pair (colour-0 word, colour-1 word) at the same index = 64 consecutive x.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ising_msc_pack_kernel(const int8_t* __restrict__ src, uint32_t* __restrict__ c0, uint32_t* __restrict__ c1,
                      uint32_t Wx, uint32_t Ly, uint32_t Lz, uint32_t z_offset) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Wx * Ly * Lz;
    if (t >= total) return;
    const uint32_t y = (uint32_t)((t / Wx) % Ly), z = (uint32_t)(t / ((size_t)Wx * Ly));
    const uint32_t par = (y + z + z_offset) & 1u;  // colour of x = 0 in this row
    const uint4* p = reinterpret_cast<const uint4*>(src + t * 64);
    uint32_t even = 0, odd = 0;  // bits of even-x / odd-x sites
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const uint4 q = p[v];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = v * 16 + j * 4 + b;
                const uint32_t up = ((int8_t)(w[j] >> (8 * b)) > 0) ? 1u : 0u;
                if (x & 1) odd |= up << (x >> 1); else even |= up << (x >> 1);
            }
    }
    // colour 0 holds x with (x+y+z) even: even x when par==0
    c0[t] = par ? odd : even;
    c1[t] = par ? even : odd;
}

__global__ void __launch_bounds__(256)
ising_msc_unpack_kernel(int8_t* __restrict__ dst, const uint32_t* __restrict__ c0, const uint32_t* __restrict__ c1,
                        uint32_t Wx, uint32_t Ly, uint32_t Lz, uint32_t z_offset) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Wx * Ly * Lz;
    if (t >= total) return;
    const uint32_t y = (uint32_t)((t / Wx) % Ly), z = (uint32_t)(t / ((size_t)Wx * Ly));
    const uint32_t par = (y + z + z_offset) & 1u;
    const uint32_t even = par ? c1[t] : c0[t], odd = par ? c0[t] : c1[t];
    uint4* p = reinterpret_cast<uint4*>(dst + t * 64);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = v * 16 + j * 4 + b;
                const uint32_t up = (((x & 1) ? odd : even) >> (x >> 1)) & 1u;
                word |= (up ? 0x01u : 0xFFu) << (8 * b);
            }
            w[j] = word;
        }
        p[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// State::rand_with_size on device (src/state.rs:260-262): one fair bit per spin.
__global__ void __launch_bounds__(256)
ising_msc_randomize_kernel(uint4* __restrict__ c0, uint4* __restrict__ c1, size_t groups_local, uint64_t group_offset,
                           uint32_t k0, uint32_t k1) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= groups_local) return;
    uint32_t r[4];
    philox_at((t + group_offset) | (1ull << 61), ~0ull, 0xFEu, k0, k1, r);
    c0[t] = make_uint4(r[0], r[1], r[2], r[3]);
    philox_at((t + group_offset) | (1ull << 61) | (1ull << 62), ~0ull, 0xFEu, k0, k1, r);
    c1[t] = make_uint4(r[0], r[1], r[2], r[3]);
}

}  // namespace vg
