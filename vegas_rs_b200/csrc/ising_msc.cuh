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

// Every (spin, antiparallel-count) class with dE > 0 owns one "slot" = one 64-bit threshold.
// A slot is selected bit-parallel as (n0^a0) & (n1^a1) & (n2^a2) [& (s^sx)]; unused slots ask for
// the impossible count 7.  Classes without a slot have dE <= 0 and are always accepted.
template <int NSLOT>
struct MscSlots {
    uint32_t a0[NSLOT], a1[NSLOT], a2[NSLOT], sx[NSLOT];
};

// Bit-planes of the slot thresholds, MSB first: bit[q][j] = bit (63 - j) of slot q's 64-bit threshold, as 0/1.
// Passed by value (constant bank): with a compile-time plane index the bits are immediate-like operands.
template <int NSLOT>
struct MscThr {
    uint32_t bit[NSLOT][64];
};

struct MscGeom {
    uint32_t Wx;        // 32-bit words per row per colour = Lx / 64
    uint32_t Ly, Lz;    // local rows / planes
    uint32_t z_offset;  // global z of local plane 0
};

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

#ifndef MSC_MINB
#define MSC_MINB 4   // resident CTAs of 256 threads per SM the sweep kernel is compiled for (register cap 64).  Measured on
#endif               // 1024^3 / 8192^2 (profiles/r02o*): 2 -> 2.319e12 / 1.888e12, 3 -> 2.337e12 / 1.852e12, 4 -> 2.423e12 / 1.867e12;
                     // final kernel (profiles/r02/i6.sh): 4 -> 3.065e12 / 2.43e12, 5 (48 registers, 16-40 B spilled) -> 2.995e12 / 2.32e12
// rows (words along y) per thread: amortises addressing, shares the y-neighbour loads.  2-D lattices are small
// (8192^2 = 1 Mi words per colour): 1, 2, 4 rows per thread measured 1.42e12, 1.76e12, 1.91e12 attempts/s there.
// of the 8 threshold bit-planes of the main compare, how many are built with logic operations instead of multiply-adds.
// ncu of the all-IMAD kernel: FMA-heavy pipe 67 % busy (IMAD 2, IMAD.WIDE 4 cycles per warp instruction), ALU pipe 53 % --
// yet moving planes to the ALU pipe loses: 0 / 2 / 4 / 8 planes measured 3.06e12 / 3.04e12 / 2.99e12 / 2.86e12 attempts/s
// on 1024^3 (profiles/r02/i5.sh): the and-or form needs the negated bits and one more operation per plane
#ifndef MSC_LOP_PLANES
#define MSC_LOP_PLANES 0
#endif
#ifndef MSC_ROWS_2D
#define MSC_ROWS_2D 4
#endif
__host__ __device__ constexpr int msc_rows(int ndim) { return ndim == 2 ? MSC_ROWS_2D : 4; }

// One thread owns K = msc_rows(NDIM) 32-bit words: the same word column w in rows y0 .. y0+K-1.
// MODE 0: update; 1: update + fused energy/magnetisation reduction; 2: reduction only.
// obs[0] += sum over own sites of s_i * (sum_nb s_j)   (every bond once, bipartite)
// obs[1] += sum of s over both colours (own word after update + partner word)
// obs[2] += accepted moves
// FERRO: the slots are exactly "count == q" for q < Z/2 (uniform J > 0, no field): the class masks need no operands.
// HALO:  planes z-1 / z+1 outside the local range come from oth_lo / oth_hi and boundary words are also stored into the
//        neighbours' halos (connected slab); otherwise the periodic wrap of `oth` itself is used and those four
//        pointers are never touched (fewer uniform registers: the Philox round keys and threshold bits stay resident).
// FULL:  the grid covers the lattice exactly (Wx % blockDim.x == 0, Ly % (blockDim.y * K) == 0; never with HALO): no thread or
//        row is ever inactive and every address is one base offset plus compile-time multiples of the row pitch -- the
//        generic prologue (per-row bounds, wraps and predicates) was ~50 of the ~275 instructions per word.
template <int NDIM, bool FIELD, int NSLOT, bool RANDPROP, int MODE, bool FERRO = false, bool HALO = true, bool FULL = false>
__global__ void __launch_bounds__(256, MSC_MINB)
ising_msc_kernel(uint32_t* __restrict__ own, const uint32_t* __restrict__ oth, const uint32_t* __restrict__ oth_lo,
                 const uint32_t* __restrict__ oth_hi, uint32_t* __restrict__ peer_lo, uint32_t* __restrict__ peer_hi,
                 MscGeom g, int colour, uint32_t z_begin, uint32_t z_step, MscSlots<NSLOT> slots, MscThr<NSLOT> thr,
                 uint64_t sweep, PhiloxKey pk, unsigned long long* __restrict__ obs) {
    constexpr int Z = 2 * NDIM;
    constexpr int K = msc_rows(NDIM);
    __shared__ int s_acc[3];
    __shared__ unsigned int s_cnt;
    // straggler records of the warp-compacted tail (NSLOT <= 3): one per word that still holds an undecided spin after
    // 8 bit-planes: {eq, n0, n1, n2, word offset, cand, pm0, pm1, pm2}; at most K * 32 per warp.  FERRO: the slot masks
    // follow from the counts and are not stored; the word index (RNG key) always follows from the word offset.
    constexpr int REC_W = FERRO ? (RANDPROP ? 6 : 5) : 9;
    __shared__ uint32_t s_rec[NSLOT <= 3 ? 8 : 1][NSLOT <= 3 ? K * 32 : 1][NSLOT <= 3 ? REC_W : 1];
    uint32_t n_rec = 0;  // warp-uniform
    if (threadIdx.x == 0 && threadIdx.y == 0) { s_acc[0] = 0; s_acc[1] = 0; s_acc[2] = 0; s_cnt = 0; }
    __syncthreads();

    // block = (BX words of a row) x (BY threads, K rows each); grid = (word tiles, row tiles, planes)
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y0 = (blockIdx.y * blockDim.y + threadIdx.y) * K;
    const uint32_t zl = z_begin + blockIdx.z * z_step;   // z_step > 1: the two boundary planes of a slab in one launch
    const uint32_t Wx = g.Wx, Ly = g.Ly;
    const uint32_t plane = Ly * Wx;            // 32-bit word offsets: a colour array has < 2^32 words
    static_assert(!(FULL && HALO), "FULL is an interior / single-handle variant");
    const bool active = FULL || (w < Wx && y0 < Ly);
    int acc[3] = {0, 0, 0};
    const uint32_t lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31u, warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;

    const uint32_t zg = zl + g.z_offset;
    const uint32_t rp0 = (y0 + zg + (uint32_t)colour) & 1u;
    const uint32_t base = zl * plane + w;      // + y * Wx
    const uint32_t wl = w == 0 ? Wx - 1 : w - 1, wr = w + 1 == Wx ? 0u : w + 1;

    // other-colour rows y0-1 .. y0+K at column w (periodic in y); rows[k+1] is the partner word of own row k.
    // Inactive lanes (ragged grids) load nothing and carry zero masks through the warp-collective code below.
    uint32_t rows[K + 2];
    uint32_t sv[K], cwv[K], Cv[K], Dv[K];
    if constexpr (FULL) {
        // 32-bit word offsets relative to (zl, y0, w); the additions wrap modulo 2^32 and land inside the array
        const uint32_t row0 = base + y0 * Wx;
        rows[0] = oth[base + (y0 == 0 ? Ly - 1 : y0 - 1) * Wx];
        rows[K + 1] = oth[base + (y0 + K == Ly ? 0u : y0 + K) * Wx];
        const uint32_t dw[2] = {(rp0 ? wr : wl) - w, (rp0 ? wl : wr) - w};
        const uint32_t dzm = zl == 0 ? (g.Lz - 1) * plane : 0u - plane, dzp = zl + 1 == g.Lz ? 0u - zl * plane : plane;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t o = row0 + k * Wx;
            rows[k + 1] = oth[o];
            sv[k] = own[o];
            cwv[k] = oth[o + dw[k & 1]];
            if (NDIM == 3) { Cv[k] = oth[o + dzm]; Dv[k] = oth[o + dzp]; } else { Cv[k] = 0; Dv[k] = 0; }
        }
    } else {
#pragma unroll
    for (int k = -1; k <= K; ++k) {
        uint32_t y = y0 + k;
        if (k < 0) y = y0 == 0 ? Ly - 1 : y0 - 1;
        if (k > 0 && y >= Ly) y -= Ly;
        rows[k + 1] = active ? oth[base + y * Wx] : 0u;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t y = y0 + k;
        const bool in = active && y < Ly;
        const uint32_t yy = y < Ly ? y : y0;
        const uint32_t rp = (rp0 + k) & 1u;
        sv[k] = 0; cwv[k] = 0; Cv[k] = 0; Dv[k] = 0;
        if (in) {
            sv[k] = own[base + yy * Wx];
            cwv[k] = oth[zl * plane + yy * Wx + (rp ? wr : wl)];
            if (NDIM == 3) {
                if (HALO) {
                    Cv[k] = zl == 0 ? oth_lo[yy * Wx + w] : oth[base - plane + yy * Wx];
                    Dv[k] = zl + 1 == g.Lz ? oth_hi[yy * Wx + w] : oth[base + plane + yy * Wx];
                } else {
                    Cv[k] = oth[(zl == 0 ? (g.Lz - 1) * plane : (zl - 1) * plane) + yy * Wx + w];
                    Dv[k] = oth[(zl + 1 == g.Lz ? 0u : (zl + 1) * plane) + yy * Wx + w];
                }
            }
        }
    }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t y = y0 + k;
        const bool in = FULL || (active && y < Ly);
        const uint32_t rp = (rp0 + k) & 1u;
        const uint32_t s = sv[k], N0 = rows[k + 1];
        // x-neighbour 2 sits one compact index to the left (row parity 0) or right (1): shift with carry
        const uint32_t Nsh = rp ? __funnelshift_r(N0, cwv[k], 1) : __funnelshift_l(cwv[k], N0, 1);
        // antiparallel indicators and their bit-sliced count n2 n1 n0
        const uint32_t a1 = s ^ N0, a2 = s ^ Nsh, a3 = s ^ rows[k], a4 = s ^ rows[k + 2];
        uint32_t n0, n1, n2;
        if (NDIM == 3) {
            const uint32_t a5 = s ^ Cv[k], a6 = s ^ Dv[k];
            const uint32_t s1 = a1 ^ a2 ^ a3, c1 = maj3(a1, a2, a3);
            const uint32_t s2 = a4 ^ a5 ^ a6, c2 = maj3(a4, a5, a6);
            n0 = s1 ^ s2;
            const uint32_t c3 = s1 & s2;
            n1 = c1 ^ c2 ^ c3;
            n2 = maj3(c1, c2, c3);
        } else {
            const uint32_t s1 = a1 ^ a2 ^ a3, c1 = maj3(a1, a2, a3);
            n0 = s1 ^ a4;
            const uint32_t c3 = s1 & a4;
            n1 = c1 ^ c3;
            n2 = c1 & c3;
        }
        // global index of this 32-bit word inside the colour array (RNG key)
        const uint64_t widx = (((uint64_t)zg * Ly + y) * Wx + w) | ((uint64_t)colour << 62);
        uint32_t flip = 0;
        if (MODE != 2) {
            uint32_t pm[NSLOT], eq = 0, lt = 0;
#pragma unroll
            for (int q = 0; q < NSLOT; ++q) {
                uint32_t m;
                if (FERRO) m = q >= Z / 2 ? 0u : ((q & 1 ? n0 : ~n0) & (q & 2 ? n1 : ~n1) & ~n2);
                else m = (n0 ^ slots.a0[q]) & (n1 ^ slots.a1[q]) & (n2 ^ slots.a2[q]);
                if (FIELD) m &= s ^ slots.sx[q];
                pm[q] = m;
                eq |= m;
            }
            const uint32_t always = ~eq;
            uint32_t cand = in ? 0xFFFFFFFFu : 0u;
            if (RANDPROP) {  // IsingSpin::rand (src/state.rs:76-84): the proposed spin is a fair coin
                uint32_t r[4];
                philox_at(widx, sweep, 0xFFu, pk, r);
                cand &= r[0] ^ s;  // proposal differs from the current spin
            }
            eq &= cand;
            // One chunk = 4 bit-planes of U from one Philox call.  The slot masks are disjoint, so the per-spin
            // threshold bit-plane is sum_q pm[q] * bit_q: integer multiply-adds (FMA pipe) instead of and/or (ALU pipe),
            // which balances the two issue pipes; with a compile-time chunk the bits are constant-bank operands.
            auto planes = [&](uint32_t ch, const uint32_t (&r)[4], const uint32_t (&m)[NSLOT], uint32_t& e, uint32_t& l) {
                uint32_t tb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int q = 0; q < NSLOT; ++q) {
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) tb[pl] += m[q] * thr.bit[q][4 * ch + pl];
                }
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const uint32_t tt = e & (tb[pl] ^ r[pl]);  // undecided spins whose U bit differs from the threshold bit
                    l |= tt & tb[pl];                          // threshold bit 1, U bit 0  ->  U < thr
                    e ^= tt;
                }
            };
            auto chunk = [&](uint32_t ch) {
                uint32_t r[4];
                philox_at(widx, sweep, ch, pk, r);
                planes(ch, r, pm, eq, lt);
            };
            if constexpr (NSLOT <= 3) {
                // 8 planes are needed by practically every word: both calls unconditionally, and the compare as ONE
                // borrow chain from the least significant of the 8 planes up: l <- (t & ~u) | (~(t ^ u) & l) leaves the
                // verdict of the MOST significant plane where U and the threshold differ, ea the spins where none does
                // (2 logic operations per plane instead of the 3 of the MSB-first form; same decisions).
                uint32_t r0[4], r1[4], t0[4] = {0u, 0u, 0u, 0u}, t1[4] = {0u, 0u, 0u, 0u};
                philox_at(widx, sweep, 0u, pk, r0);
                philox_at(widx, sweep, 1u, pk, r1);
                // threshold bit-planes: sum_q pm[q] * bit (IMAD, FMA-heavy pipe) for the first 8 - MSC_LOP_PLANES planes,
                // or_q pm[q] & -bit (LOP3, ALU pipe) for the rest: the split balances the two pipes
#pragma unroll
                for (int pl = 0; pl < 8; ++pl) {
                    uint32_t t = 0u;
#pragma unroll
                    for (int q = 0; q < NSLOT; ++q) {
                        if (pl < 8 - MSC_LOP_PLANES) t += pm[q] * thr.bit[q][pl];
                        else t |= pm[q] & (0u - thr.bit[q][pl]);
                    }
                    if (pl < 4) t0[pl] = t; else t1[pl - 4] = t;
                }
                uint32_t l8 = 0u, ea = 0xFFFFFFFFu;
#pragma unroll
                for (int pl = 7; pl >= 0; --pl) {
                    const uint32_t t = pl < 4 ? t0[pl] : t1[pl - 4], u = pl < 4 ? r0[pl] : r1[pl - 4];
                    l8 = (t & ~u) | (~(t ^ u) & l8);
                    ea &= ~(t ^ u);
                }
                lt = eq & l8;
                eq &= ea;
            } else {
                chunk(0);
                chunk(1);
            }
            if constexpr (NSLOT <= 3) {
                // Tail.  After 8 planes about 6 % of the words still hold an undecided spin.  Running every lane through
                // more chunks would cost the whole warp a Philox call per word; instead the word is finished
                // provisionally (undecided = rejected) and a record goes to the warp's list in shared memory.  After the
                // K words the list is compacted over the lanes (one record per lane) and resolved there, with the SAME
                // random numbers (keyed by the word index), so the decisions are exactly those of the plain loop.
                const uint32_t need = __ballot_sync(0xffffffffu, eq != 0u);
                if (eq != 0u) {
                    uint32_t* q = s_rec[warp][n_rec + __popc(need & ((1u << lane) - 1u))];
                    q[0] = eq; q[1] = n0; q[2] = n1; q[3] = n2; q[4] = base + y * Wx;
                    if (!FERRO || RANDPROP) q[5] = cand;
                    if (!FERRO) {
#pragma unroll
                        for (int i = 0; i < NSLOT; ++i) q[6 + i] = pm[i];
                    }
                }
                n_rec += __popc(need);
            } else {
                for (uint32_t ch = 2; ch < 16u && eq != 0u; ++ch) chunk(ch);
            }
            const uint32_t ok = always | lt;
            flip = ok & cand;
            acc[2] += __popc(RANDPROP ? ((ok | ~cand) & (in ? 0xFFFFFFFFu : 0u)) : flip);
        }
        const uint32_t snew = s ^ flip;
        if (MODE != 0 && in) {
            // sum of final antiparallel counts: c' = flip ? Z - c : c
            const int s_all = __popc(n0) + 2 * __popc(n1) + 4 * __popc(n2);
            const int s_f = __popc(n0 & flip) + 2 * __popc(n1 & flip) + 4 * __popc(n2 & flip);
            const int cfin = s_all - 2 * s_f + Z * __popc(flip);
            acc[0] += Z * 32 - 2 * cfin;
            acc[1] += 2 * __popc(snew) - 32 + 2 * __popc(N0) - 32;
        }
        if (MODE != 2 && in) {
            own[base + y * Wx] = snew;
            if (NDIM == 3 && HALO) {
                if (peer_lo != nullptr && zl == 0) peer_lo[y * Wx + w] = snew;
                if (peer_hi != nullptr && zl + 1 == g.Lz) peer_hi[y * Wx + w] = snew;
            }
        }
    }
    if constexpr (NSLOT <= 3) {
        if (MODE != 2) {
            __syncwarp();
            for (uint32_t b0 = 0; b0 < n_rec; b0 += 32u) {
                const uint32_t idx = b0 + lane;
                if (idx < n_rec) {
                    const uint32_t* q = s_rec[warp][idx];
                    uint32_t e = q[0], l = 0u, m[NSLOT];
                    const uint32_t r0 = q[1], r1 = q[2], r2 = q[3], off = q[4];
#pragma unroll
                    for (int i = 0; i < NSLOT; ++i) {
                        if (FERRO) m[i] = i >= Z / 2 ? 0u : ((i & 1 ? r0 : ~r0) & (i & 2 ? r1 : ~r1) & ~r2);
                        else m[i] = q[6 + i];
                    }
                    const uint32_t rcand = (!FERRO || RANDPROP) ? q[5] : 0xFFFFFFFFu;
                    // word offset -> global word index of the colour array (the key the main loop drew with)
                    const uint64_t wi = ((uint64_t)off + (uint64_t)g.z_offset * plane) | ((uint64_t)colour << 62);
                    for (uint32_t ch = 2; e != 0u && ch < 16u; ++ch) {
                        uint32_t r[4], tb[4] = {0u, 0u, 0u, 0u};
                        philox_at(wi, sweep, ch, pk, r);
#pragma unroll
                        for (int qq = 0; qq < NSLOT; ++qq) {
#pragma unroll
                            for (int pl = 0; pl < 4; ++pl) tb[pl] += m[qq] * thr.bit[qq][4 * ch + pl];
                        }
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl) {
                            const uint32_t tt = e & (tb[pl] ^ r[pl]);
                            l |= tt & tb[pl];
                            e ^= tt;
                        }
                    }
                    const uint32_t f = l & rcand;  // spins accepted after all: flip them on top of the provisional word
                    if (f != 0u) {
                        const uint32_t prov = own[off], fin = prov ^ f;
                        own[off] = fin;
                        if (NDIM == 3 && HALO) {
                            const uint32_t in_plane = off - zl * plane;
                            if (peer_lo != nullptr && zl == 0) peer_lo[in_plane] = fin;
                            if (peer_hi != nullptr && zl + 1 == g.Lz) peer_hi[in_plane] = fin;
                        }
                        acc[2] += __popc(f);
                        if (MODE != 0) {
                            // a flipped spin with antiparallel count c ends with Z - c: the sum of final counts moves by Z - 2c
                            const int dc = Z * __popc(f) - 2 * (__popc(r0 & f) + 2 * __popc(r1 & f) + 4 * __popc(r2 & f));
                            acc[0] -= 2 * dc;
                            acc[1] += 2 * (__popc(f & ~prov) - __popc(f & prov));
                        }
                    }
                }
            }
        }
    }
    if (MODE == 0) {
        const int a1[1] = {acc[2]};
        block_flush_int<1>(a1, s_acc, &s_cnt, obs + 2);
    } else {
        block_flush_int<3>(acc, s_acc, &s_cnt, obs);
    }
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

// The same conversions from / to a sign BITMAP in natural site order (bit x & 31 of word x >> 5 of a row of Lx spins: 1 = Up),
// what the host packs for big lattices (host_pack.cpp): 1/8 of the PCIe bytes.  One thread owns 64 consecutive x.
__device__ __forceinline__ uint32_t compress_even_bits(uint64_t v) {   // bits 0, 2, 4, ... of v -> bits 0, 1, 2, ...
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
    v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
    v = (v | (v >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)v;
}
__device__ __forceinline__ uint64_t spread_to_even_bits(uint32_t w) {  // bits 0, 1, 2, ... of w -> bits 0, 2, 4, ...
    uint64_t v = w;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
__global__ void __launch_bounds__(256)
ising_msc_from_bitmap_kernel(const uint2* __restrict__ bitmap, uint32_t* __restrict__ c0, uint32_t* __restrict__ c1,
                             uint32_t Wx, uint32_t Ly, uint32_t Lz, uint32_t z_offset) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Wx * Ly * Lz;
    if (t >= total) return;
    const uint32_t y = (uint32_t)((t / Wx) % Ly), z = (uint32_t)(t / ((size_t)Wx * Ly));
    const uint32_t par = (y + z + z_offset) & 1u;  // colour of x = 0 in this row
    const uint2 q = bitmap[t];
    const uint64_t v = (uint64_t)q.x | ((uint64_t)q.y << 32);
    const uint32_t even = compress_even_bits(v), odd = compress_even_bits(v >> 1);
    c0[t] = par ? odd : even;
    c1[t] = par ? even : odd;
}
__global__ void __launch_bounds__(256)
ising_msc_to_bitmap_kernel(uint2* __restrict__ bitmap, const uint32_t* __restrict__ c0, const uint32_t* __restrict__ c1,
                           uint32_t Wx, uint32_t Ly, uint32_t Lz, uint32_t z_offset) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Wx * Ly * Lz;
    if (t >= total) return;
    const uint32_t y = (uint32_t)((t / Wx) % Ly), z = (uint32_t)(t / ((size_t)Wx * Ly));
    const uint32_t par = (y + z + z_offset) & 1u;
    const uint32_t even = par ? c1[t] : c0[t], odd = par ? c0[t] : c1[t];
    const uint64_t v = spread_to_even_bits(even) | (spread_to_even_bits(odd) << 1);
    bitmap[t] = make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}

// State::rand_with_size on device (src/state.rs:260-262): one fair bit per spin.
__global__ void __launch_bounds__(256)
ising_msc_randomize_kernel(uint32_t* __restrict__ c0, uint32_t* __restrict__ c1, size_t words_local, uint64_t word_offset,
                           PhiloxKey pk) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= words_local) return;
    uint32_t r[4];
    philox_at((t + word_offset) | (1ull << 61), ~0ull, 0xFEu, pk, r);
    c0[t] = r[0];
    c1[t] = r[1];
}

}  // namespace vg
