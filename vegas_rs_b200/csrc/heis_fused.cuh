// heis_fused.cuh -- K3f: one launch = one whole Monte Carlo step (both checkerboard colours) of the
// Heisenberg sc stencil, with the spins read from HBM once and written once (24 B/attempt in fp32).
//
// Same arithmetic, same Philox keys and therefore the same trajectory as two heis_stencil_kernel
// colour passes (heis.cuh); it replaces MetropolisIntegrator::step / MetropolisFlipIntegrator::step
// (src/integrator.rs:66-92, :109-138) over the compound Hamiltonian (src/energy.rs:63-257).
//
// Scheme (2.5-D temporal blocking).  A CTA owns TY full-x rows and a chunk of z-planes and marches
// along z.  At march step k it has the old colour-1 plane k ("B[k]") in shared memory and
//   * updates colour 0 on plane k-1 from B[k-2] (registers), B[k-1] (shared, in-plane neighbours)
//     and B[k] (shared, own column)                                           -> Anew[k-1]
//   * updates colour 1 on plane k-2 from Anew[k-3] (registers), Anew[k-2] (shared) and Anew[k-1].
// Rows: the CTA loads TY+4 rows of colour 1 and TY+2 rows of colour 0, recomputes the colour-0 update
// of its two halo rows (site-keyed random numbers make the redundant updates identical to the
// neighbour CTA's) and writes only its TY interior rows.  Because neighbouring CTAs read each
// other's rows of the OLD state, the step reads `src` arrays and writes `dst` arrays (ping-pong).
// Planes are staged with cp.async (16 B per thread, L2 only), one plane ahead of the arithmetic;
// one CTA barrier per march step.
//
// Shared memory: 3 ring slots per colour, slot = [row][component][HX] so that with a compile-time row
// length (HX_T != 0) every neighbour access is one LDS with an immediate offset from the thread's base.
#pragma once
#include <type_traits>

#include "heis.cuh"

namespace vg {

template <typename real>
struct FusedPtrs {
    const real* src[2][3];      // [colour][component] old state, local planes [0, Lz)
    real* dst[2][3];            // new state
};

struct FusedGeom {
    uint32_t Hx, Gx, Lx, Ly, Lz;   // compact row length, 16-byte groups per row, extents (local Lz)
    uint32_t z_offset, nz_global;  // global plane of local plane 0
    uint32_t TY, ROWS;             // interior rows per CTA, ROWS = TY + 4
    uint32_t tiles;                // Ly / TY
    uint32_t CZ, chunks;           // planes per z-chunk
};

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef HEIS_FUSED_THREADS
#define HEIS_FUSED_THREADS 640   // (ROWS - 2) rows x up to 64 sixteen-byte groups: every thread owns a column that colour 0 updates
#endif

// obs layout as heis_stencil_kernel: [0] -sum_{colour 1} s.n  [1..3] sum s  [4] sum (s.a)^2  [5] accepted
template <typename real, int HX_T, bool FLIP, bool RECORD>
__global__ void __launch_bounds__(HEIS_FUSED_THREADS, 1)
heis_fused_kernel(FusedPtrs<real> P, FusedGeom g, HeisParams<real> p, uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    constexpr int N = VecOf<real>::N;
    constexpr uint32_t RB = sizeof(real);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double s_acc[6];
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;           // ordered before the first flush by the loop's barriers

    const uint32_t HX = HX_T ? (uint32_t)HX_T : g.Hx;        // compile-time row length when specialised
    const uint32_t GX = HX / N;
    const uint32_t CB = HX * RB;                             // byte stride between components of a row
    const uint32_t ROWB = 3u * CB;                           // byte stride between rows
    const uint32_t SLOTB = g.ROWS * ROWB;                    // bytes of one plane slot

    const uint32_t tile = blockIdx.x % g.tiles, chunk = blockIdx.x / g.tiles;
    const int z0 = (int)(chunk * g.CZ), z1 = min((int)g.Lz, z0 + (int)g.CZ);
    // threads cover tile rows 1 .. ROWS-2 (the rows whose colour-0 update this CTA computes); the threads of rows 1 and
    // ROWS-2 also stage the outermost colour-1 rows 0 and ROWS-1, which are only read
    const uint32_t r = 1u + threadIdx.x / GX, gx = threadIdx.x % GX;
    const bool a_row = r + 1 < g.ROWS;
    const bool b_row = a_row && r >= 2 && r + 2 < g.ROWS;    // interior rows: written by this CTA
    const int outer = !a_row ? 0 : (r == 1 ? -1 : (r + 2 == g.ROWS ? 1 : 0));   // also stages the colour-1 row r + outer
    const uint32_t y = (tile * g.TY + g.Ly - 2u + r) % g.Ly; // lattice row of tile row r
    const size_t plane = (size_t)g.Ly * HX;
    const size_t e_row = (size_t)y * HX + (size_t)gx * N;    // element offset inside a plane
    const size_t e_outer = (size_t)((y + g.Ly + outer) % g.Ly) * HX + (size_t)gx * N;
    const uint32_t own = r * ROWB + gx * (N * RB);           // byte offset of the own 16 bytes inside a slot (component 0)
    // x-neighbour 2 of element 0 / N-1 lives in the adjacent group of the same row (periodic in x)
    const int dl = (gx == 0 ? (int)(CB - RB) : -(int)RB);                          // byte delta to the left carry
    const int dr = (gx + 1 == GX ? -(int)(gx * N * RB) : (int)(N * RB));           // ... to the right carry
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);

    // ring slots as byte offsets (+ own): oB[0..2] = planes k-1, k, k+1;  oA[0..2] = planes k-2, k-1, k
    uint32_t oB0 = own, oB1 = own + SLOTB, oB2 = own + 2 * SLOTB;
    uint32_t oA0 = own + 3 * SLOTB, oA1 = own + 4 * SLOTB, oA2 = own + 5 * SLOTB;

    auto lds = [&](uint32_t off, real (&v)[N]) { vec_load(reinterpret_cast<const real*>(smem + off), v); };
    auto lds1 = [&](uint32_t off) { return *reinterpret_cast<const real*>(smem + off); };
    auto sts = [&](uint32_t off, const real (&v)[N]) { vec_store(reinterpret_cast<real*>(smem + off), v); };
    auto wrapz = [&](int zl) { return (size_t)(zl < 0 ? zl + (int)g.Lz : (zl >= (int)g.Lz ? zl - (int)g.Lz : zl)); };
    auto stageB = [&](uint32_t off, int zl) {                // async copy of this thread's 16 bytes x 3 components of plane zl
        const size_t zo = wrapz(zl) * plane;
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async16(smem_base + off + c * CB, P.src[1][c] + zo + e_row);
        if (outer != 0) {
            const uint32_t off2 = outer < 0 ? off - ROWB : off + ROWB;
#pragma unroll
            for (int c = 0; c < 3; ++c) cp_async16(smem_base + off2 + c * CB, P.src[1][c] + zo + e_outer);
        }
    };
    auto stageA = [&](uint32_t off, int zl) {
        const size_t zo = wrapz(zl) * plane + e_row;
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async16(smem_base + off + c * CB, P.src[0][c] + zo);
    };

    real Bm2[3][N], Am3[3][N];                               // B[k-2], Anew[k-3] of the own column
    real facc[5] = {0, 0, 0, 0, 0};                          // per-thread partial sums, flushed every 16 planes
    int accepted = 0;
    auto flush = [&]() {                                     // all threads of the CTA call this together
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double v = warp_sum((double)facc[i]);
            if ((threadIdx.x & 31u) == 0 && v != 0.0) atomicAdd(&s_acc[i], v);
            facc[i] = 0;
        }
    };
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < N; ++e) { Bm2[c][e] = 0; Am3[c][e] = 0; }

    // prologue: B[z0-2] -> registers; B[z0-1], B[z0], A[z0-1] -> shared
    if (a_row) {
        const size_t e = wrapz(z0 - 2) * plane + e_row;
#pragma unroll
        for (int c = 0; c < 3; ++c) vec_load(P.src[1][c] + e, Bm2[c]);
        stageB(oB0, z0 - 1);
        stageB(oB1, z0);
        stageA(oA1, z0 - 1);
    }
    cp_async_commit();

    // N attempts of one row segment: s <- updated spins.  nsum = raw neighbour sums.
    auto update = [&](real (&s)[3][N], const real (&nsum)[3][N], uint32_t zg, uint32_t rp, bool count) {
        const uint64_t site0 = (uint64_t)(zg * g.Ly + y) * g.Lx + (2u * N) * gx + rp;  // element e is site0 + 2e
        const uint32_t s_lo = (uint32_t)site0, s_hi = (uint32_t)(site0 >> 32);       // site0 % 8 == rp: adding 2e never carries
        const uint32_t c2 = (uint32_t)sweep, c3 = (uint32_t)(sweep >> 32) & 0x00FFFFFFu;
        if (sizeof(real) == 4) {
#pragma unroll
            for (int e = 0; e < N; e += 2) {
                uint32_t w[4];
                philox4x32(s_lo + 2u * e, s_hi, c2, c3, pk, w);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    HeisRand<real> rnd;
                    reinterpret_cast<HeisRand<float>&>(rnd) = heis_rand_words(w[2 * h2], w[2 * h2 + 1]);
                    const bool ok = heis_attempt<real, FLIP>(s[0][e + h2], s[1][e + h2], s[2][e + h2],
                                                             heis_field(p.J, nsum[0][e + h2], p.h[0]), heis_field(p.J, nsum[1][e + h2], p.h[1]),
                                                             heis_field(p.J, nsum[2][e + h2], p.h[2]), p, rnd);
                    accepted += (ok && count) ? 1 : 0;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < N; ++e) {
                HeisRand<real> rnd;
                heis_rand(site0 + 2u * e, sweep, pk, rnd);
                const bool ok = heis_attempt<real, FLIP>(s[0][e], s[1][e], s[2][e], heis_field(p.J, nsum[0][e], p.h[0]),
                                                         heis_field(p.J, nsum[1][e], p.h[1]), heis_field(p.J, nsum[2][e], p.h[2]), p, rnd);
                accepted += (ok && count) ? 1 : 0;
            }
        }
    };

    // The arithmetic of one march step for a row whose x-neighbour 2 sits to the right (RP = 1) or left (RP = 0).
    // INTERIOR rows run colour 0 on plane k-1 and colour 1 on plane k-2 as ONE straight-line block (the warm-up steps
    // compute on not-yet-valid planes and simply do not store), so that the scheduler can overlap the two.
    auto body = [&](auto rp_tag, auto interior_tag, int k, uint32_t zgA, uint32_t zgB) {
        constexpr int RP = decltype(rp_tag)::value;
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        const int dc = RP ? dr : dl;
        real nA[3][N], s[3][N], bk1[3][N];
        // ---- colour 0 on plane k-1
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            real up[N], dn[N], zp[N];
            lds(oB0 + c * CB, bk1[c]);
            lds(oB0 + c * CB - ROWB, up);
            lds(oB0 + c * CB + ROWB, dn);
            lds(oB1 + c * CB, zp);
            lds(oA1 + c * CB, s[c]);
            const real carry = lds1(oB0 + c * CB + dc);
#pragma unroll
            for (int e = 0; e < N; ++e) {
                const real sh = RP ? (e + 1 < N ? bk1[c][(e + 1) % N] : carry) : (e > 0 ? bk1[c][(e + N - 1) % N] : carry);
                nA[c][e] = ((bk1[c][e] + (up[e] + dn[e])) + sh) + (Bm2[c][e] + zp[e]);  // same order as heis_march
            }
        }
        const bool mineA = INTERIOR && k - 1 >= z0 && k - 1 < z1;  // not a redundant halo update
        update(s, nA, zgA, RP, mineA);
#pragma unroll
        for (int c = 0; c < 3; ++c) sts(oA1 + c * CB, s[c]);
        if (mineA) {
            const size_t e = (size_t)(k - 1) * plane + e_row;
#pragma unroll
            for (int c = 0; c < 3; ++c) vec_store(P.dst[0][c] + e, s[c]);
            if (RECORD) {
#pragma unroll
                for (int e2 = 0; e2 < N; ++e2) {
                    facc[1] += s[0][e2]; facc[2] += s[1][e2]; facc[3] += s[2][e2];
                    const real d1 = s[0][e2] * p.a[0] + s[1][e2] * p.a[1] + s[2][e2] * p.a[2];
                    facc[4] += d1 * d1;
                }
            }
        }
        // ---- colour 1 on plane k-2 (its x-neighbour 2 sits on the same side: (y + z + colour) has the same parity)
        if (INTERIOR) {
            real nB[3][N];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                real n0[N], up[N], dn[N];
                lds(oA0 + c * CB, n0);
                lds(oA0 + c * CB - ROWB, up);
                lds(oA0 + c * CB + ROWB, dn);
                const real carry = lds1(oA0 + c * CB + dc);
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    const real sh = RP ? (e + 1 < N ? n0[(e + 1) % N] : carry) : (e > 0 ? n0[(e + N - 1) % N] : carry);
                    nB[c][e] = ((n0[e] + (up[e] + dn[e])) + sh) + (Am3[c][e] + s[c][e]);   // s = Anew[k-1] of the own column
                    Am3[c][e] = n0[e];                       // Anew[k-2] is next step's Anew[k-3]
                }
            }
            const bool mineB = k - 2 >= z0;
            update(Bm2, nB, zgB, RP, mineB);
            if (mineB) {
                const size_t e = (size_t)(k - 2) * plane + e_row;
#pragma unroll
                for (int c = 0; c < 3; ++c) vec_store(P.dst[1][c] + e, Bm2[c]);
                if (RECORD) {
#pragma unroll
                    for (int e2 = 0; e2 < N; ++e2) {
                        facc[0] -= p.J * (Bm2[0][e2] * nB[0][e2] + Bm2[1][e2] * nB[1][e2] + Bm2[2][e2] * nB[2][e2]);
                        facc[1] += Bm2[0][e2]; facc[2] += Bm2[1][e2]; facc[3] += Bm2[2][e2];
                        const real d1 = Bm2[0][e2] * p.a[0] + Bm2[1][e2] * p.a[1] + Bm2[2][e2] * p.a[2];
                        facc[4] += d1 * d1;
                    }
                }
            }
        }
        // ---- B[k-1] becomes next step's B[k-2] (own column)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int e = 0; e < N; ++e) Bm2[c][e] = bk1[c][e];
    };

    const uint32_t nzg = g.nz_global;
    uint32_t zgA = (uint32_t)(((int)g.z_offset + z0 - 1 + (int)nzg) % (int)nzg);    // global plane of k-1
    for (int k = z0; k <= z1 + 1; ++k) {
        cp_async_wait_all();
        __syncthreads();                                     // planes B[k], A[k-1] have landed; the previous step is complete
        if (a_row) {                                         // prefetch for step k+1
            if (k + 1 <= z1 + 1) stageB(oB2, k + 1);
            if (k <= z1) stageA(oA2, k);
        }
        cp_async_commit();
        const uint32_t zgB = zgA == 0 ? nzg - 1 : zgA - 1;
        if (a_row) {
            const bool rp = (y + zgA) & 1u;
            if (b_row) {
                if (rp) body(std::integral_constant<int, 1>{}, std::true_type{}, k, zgA, zgB);
                else body(std::integral_constant<int, 0>{}, std::true_type{}, k, zgA, zgB);
            } else {
                if (rp) body(std::integral_constant<int, 1>{}, std::false_type{}, k, zgA, zgB);
                else body(std::integral_constant<int, 0>{}, std::false_type{}, k, zgA, zgB);
            }
        }
        if (RECORD && ((k - z0) & 15) == 15) flush();        // keep the fp32 partial sums short
        // rotate the rings
        { const uint32_t t = oB0; oB0 = oB1; oB1 = oB2; oB2 = t; }
        { const uint32_t t = oA0; oA0 = oA1; oA1 = oA2; oA2 = t; }
        zgA = zgA + 1 == nzg ? 0u : zgA + 1;
    }
    cp_async_wait_all();
    if (RECORD) flush();
    {
        const int a = __reduce_add_sync(0xffffffffu, accepted);
        if ((threadIdx.x & 31u) == 0 && a != 0) atomicAdd(&s_acc[5], (double)a);
    }
    __syncthreads();
    if (threadIdx.x < 6 && (RECORD || threadIdx.x == 5) && s_acc[threadIdx.x] != 0.0) atomicAdd(obs + threadIdx.x, s_acc[threadIdx.x]);
}

}  // namespace vg
