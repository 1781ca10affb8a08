// heis.cuh -- K3: Heisenberg checkerboard half-sweep for sc lattices (pbc), SoA colour-split storage,
// plus the single-site update shared with the general-adjacency kernels.
//
// Replaces MetropolisIntegrator::step (src/integrator.rs:66-92; MetropolisFlipIntegrator :109-138
// when FLIP) for HeisenbergSpin with the compound Gauge+Exchange+Anisotropy+Zeeman energy
// (src/energy.rs:63-214) in the closed form of SURVEY App. B:
//   dE = -(s'-s).n + (s'-s).h + k[(s'.a)^2 - (s.a)^2],   n = sum_j J_ij s_j,  h = |H| * orientation
// (reference sign: Zeeman::energy = +|H| s.o, src/energy.rs:147-151).
#pragma once
#include "common.cuh"

namespace vg {

template <typename real>
struct HeisParams {
    real J;        // uniform exchange (multiplies the neighbour sum)
    real h[3];     // |H| * field orientation
    real k;        // anisotropy strength (0 when absent)
    real a[3];     // anisotropy reference spin
    real invT;
    real invTl;    // log2(e) / T: the Boltzmann factor is evaluated as 2^(-dE * invTl)
};

// ---------------------------------------------------------------------------------------
// fp32 lane arithmetic.  The attempt is written ONCE over a lane type L: float (one site per instruction) or float2 (two
// sites per instruction: Blackwell's packed FADD2 / FMUL2 / FFMA2, half the issue slots).  Every operation is an explicit
// round-to-nearest intrinsic -- never contracted, never reassociated -- so the scalar kernels and the packed one produce
// the same bits per site.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float l_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float l_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float l_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float l_neg(float a) { return -a; }
__device__ __forceinline__ float2 l_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 l_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 l_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 l_neg(float2 a) { return make_float2(-a.x, -a.y); }   // folds into the operand modifiers
template <typename L> __device__ __forceinline__ L l_bc(float x);
template <> __device__ __forceinline__ float l_bc<float>(float x) { return x; }
template <> __device__ __forceinline__ float2 l_bc<float2>(float x) { return make_float2(x, x); }
// per-lane special functions (MUFU): hardware rsqrt / cos / sin / ex2
__device__ __forceinline__ float l_rsqrt(float a) { return rsqrtf(a); }
__device__ __forceinline__ float l_cos(float a) { return __cosf(a); }
__device__ __forceinline__ float l_sin(float a) { return __sinf(a); }
__device__ __forceinline__ float2 l_rsqrt(float2 a) { return make_float2(rsqrtf(a.x), rsqrtf(a.y)); }
__device__ __forceinline__ float2 l_cos(float2 a) { return make_float2(__cosf(a.x), __cosf(a.y)); }
__device__ __forceinline__ float2 l_sin(float2 a) { return make_float2(__sinf(a.x), __sinf(a.y)); }

// Uniform point on the sphere from two uniforms (Archimedes' hat-box); same distribution as
// util.rs:21-34 (Marsaglia) without a rejection loop.
// fp32: the uniforms arrive as integer-valued floats f0, f1 in [0, 2^21) so that the scalings fold into
// the multiply-adds; hardware sin/cos/rsqrt (abs. error ~2^-21, spin norm 1 +- 1e-6, inside the fp32 bar).
//   z = 1 - 2 u0 with u0 = (f0 + 1/2) 2^-21 in (0,1);   azimuth = 2 pi (f1 2^-21 - 1/2)
template <typename L>
__device__ __forceinline__ void sphere_point_f32(L f0, L f1, L& x, L& y, L& z) {
    z = l_fma(f0, l_bc<L>(-0x1.0p-20f), l_bc<L>(1.0f - 0x1.0p-21f));
    const L t = l_fma(l_neg(z), z, l_bc<L>(1.0f));   // = 4 u0 (1 - u0) > 0: one rounding of the exact 1 - z^2
    const L rxy = l_mul(t, l_rsqrt(t));
    const L ang = l_fma(f1, l_bc<L>(6.283185307179586f * 0x1.0p-21f), l_bc<L>(-3.141592653589793f));
    x = l_mul(rxy, l_cos(ang)); y = l_mul(rxy, l_sin(ang));
}
__device__ __forceinline__ void sphere_point(float f0, float f1, float& x, float& y, float& z) { sphere_point_f32<float>(f0, f1, x, y, z); }
// fp64: u0 in [0,1), u1 in [0,1)
__device__ __forceinline__ void sphere_point(double u0, double u1, double& x, double& y, double& z) {
    z = 1.0 - 2.0 * u0;
    const double rxy = sqrt(fmax(0.0, (1.0 - z) * (1.0 + z)));
    double sn, cs;
    sincospi(2.0 * u1, &sn, &cs);
    x = rxy * cs; y = rxy * sn;
}

// 2^x; fp32 uses the hardware approximation (rel. error 2^-22)
__device__ __forceinline__ float fast_exp2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double fast_exp2(double x) { return exp2(x); }

// The three random numbers of one attempt: direction (u0, u1) and acceptance (ua).
// fp32: u0, u1 integer-valued floats in [0, 2^21); ua = k + 1/2 with k in [0, 2^22), i.e. the acceptance uniform is the
//       CENTRE (k + 1/2) 2^-22 of its cell (64 random bits per attempt).  With the cell's lower edge instead, k = 0 would
//       accept any move whose Boltzmann factor is merely positive: a floor of 2^-22 per attempt that breaks balance at low T;
// fp64: uniforms in [0, 1) with 53 bits.
template <typename real>
struct HeisRand { real u0, u1, ua; };
template <typename real> struct HeisAcceptShift;
template <> struct HeisAcceptShift<float> { static constexpr float value = 22.0f; };   // (k + 1/2) < 2^22 * exp(-dE/T)
template <> struct HeisAcceptShift<double> { static constexpr double value = 0.0; };

__device__ __forceinline__ HeisRand<float> heis_rand_words(uint32_t w0, uint32_t w1) {
    HeisRand<float> o;
    o.u0 = (float)(w0 >> 11);
    o.u1 = (float)(w1 >> 11);
    o.ua = (float)(((w0 & 0x7FFu) << 11) | (w1 & 0x7FFu)) + 0.5f;   // exact: 23 significant bits
    return o;
}

// Random numbers of site `site` in sweep `sweep`.
// fp32: ONE Philox call serves the two sites that differ in bit 1 of the index (same colour, x and x+2
//       on an sc row): counter index = site & ~2, words (0,1) for bit 1 clear, (2,3) for bit 1 set.
// fp64: call 0 -> 53-bit u0, u1; call 1 -> 53-bit ua (the fp64 path is the 1e-12 parity path).
__device__ __forceinline__ void heis_rand(uint64_t site, uint64_t sweep, const PhiloxKey& pk, HeisRand<float>& o) {
    uint32_t r[4];
    philox_at(site & ~2ull, sweep, 0u, pk, r);
    o = (site & 2ull) ? heis_rand_words(r[2], r[3]) : heis_rand_words(r[0], r[1]);
}
__device__ __forceinline__ void heis_rand(uint64_t site, uint64_t sweep, const PhiloxKey& pk, HeisRand<double>& o) {
    uint32_t r[4], q[4];
    philox_at(site, sweep, 0u, pk, r);
    philox_at(site, sweep, 1u, pk, q);
    o.u0 = u53(r[0], r[1]); o.u1 = u53(r[2], r[3]); o.ua = u53(q[0], q[1]);
}

// Effective field of a site from its neighbour sum: f = J n - h, one fused multiply-add (all stencil / basis kernels
// go through here so that they agree bit for bit).
__device__ __forceinline__ float heis_field(float J, float n, float h) { return __fmaf_rn(J, n, -h); }
__device__ __forceinline__ float2 heis_field(float J, float2 n, float h) { return __ffma2_rn(make_float2(J, J), n, make_float2(-h, -h)); }
__device__ __forceinline__ double heis_field(double J, double n, double h) { return __fma_rn(J, n, -h); }

// One Metropolis attempt on a site whose effective field f = sum_j J_ij s_j - |H| o (energy units) and random
// numbers are known:  -dE = (s' - s).f - k[(s'.a)^2 - (s.a)^2]   (reference signs).
// src/integrator.rs:82-88 accepts if dE < 0, else if u < exp(-dE/T); since u < 1 both cases are
// u < exp(-dE/T), evaluated as 2^(-dE log2(e)/T).
// AXZ: the caller guarantees a = (0, 0, a_z); s.a = s_z a_z is then bit-identical to the general dot product.
//
// fp32, any lane type: the proposal (px, py, pz) and the exponent x = -dE log2(e)/T + 22 of the acceptance bound 2^x
template <typename L, bool FLIP, bool AXZ>
__device__ __forceinline__ L heis_propose_f32(L sx, L sy, L sz, L fx, L fy, L fz, const HeisParams<float>& p, L u0, L u1, L& px, L& py, L& pz) {
    if (FLIP) { px = l_neg(sx); py = l_neg(sy); pz = l_neg(sz); }
    else sphere_point_f32<L>(u0, u1, px, py, pz);
    const L dx = l_add(px, l_neg(sx)), dy = l_add(py, l_neg(sy)), dz = l_add(pz, l_neg(sz));
    L mdE = l_fma(dz, fz, l_fma(dy, fy, l_mul(dx, fx)));
    if (!FLIP) {  // (s.a)^2 is invariant under a flip
        const L da_new = AXZ ? l_mul(pz, l_bc<L>(p.a[2])) : l_fma(pz, l_bc<L>(p.a[2]), l_fma(py, l_bc<L>(p.a[1]), l_mul(px, l_bc<L>(p.a[0]))));
        const L da_old = AXZ ? l_mul(sz, l_bc<L>(p.a[2])) : l_fma(sz, l_bc<L>(p.a[2]), l_fma(sy, l_bc<L>(p.a[1]), l_mul(sx, l_bc<L>(p.a[0]))));
        mdE = l_fma(l_bc<L>(-p.k), l_mul(l_add(da_new, l_neg(da_old)), l_add(da_new, da_old)), mdE);
    }
    return l_fma(mdE, l_bc<L>(p.invTl), l_bc<L>(HeisAcceptShift<float>::value));
}

// Returns true when accepted (the spin is then replaced by the proposal).
template <typename real, bool FLIP, bool AXZ = false>
__device__ __forceinline__ bool heis_attempt(real& sx, real& sy, real& sz, real fx, real fy, real fz,
                                             const HeisParams<real>& p, const HeisRand<real>& rnd) {
    if constexpr (sizeof(real) == 4) {
        float px, py, pz;
        const float x = heis_propose_f32<float, FLIP, AXZ>(sx, sy, sz, fx, fy, fz, p, rnd.u0, rnd.u1, px, py, pz);
        const bool acc = rnd.ua < fast_exp2(x);
        if (acc) { sx = px; sy = py; sz = pz; }
        return acc;
    } else {
        real px, py, pz;
        if (FLIP) { px = -sx; py = -sy; pz = -sz; }
        else sphere_point(rnd.u0, rnd.u1, px, py, pz);
        const real dx = px - sx, dy = py - sy, dz = pz - sz;
        real mdE = dx * fx + dy * fy + dz * fz;
        if (!FLIP) {
            const real da_new = AXZ ? pz * p.a[2] : px * p.a[0] + py * p.a[1] + pz * p.a[2];
            const real da_old = AXZ ? sz * p.a[2] : sx * p.a[0] + sy * p.a[1] + sz * p.a[2];
            mdE -= p.k * ((da_new - da_old) * (da_new + da_old));
        }
        const bool acc = rnd.ua < fast_exp2(mdE * p.invTl + HeisAcceptShift<real>::value);
        if (acc) { sx = px; sy = py; sz = pz; }
        return acc;
    }
}

// Two sites per instruction (fp32): lanes .x / .y with their own random numbers; a0 / a1 = accepted.
template <bool FLIP, bool AXZ>
__device__ __forceinline__ void heis_attempt2(float2& sx, float2& sy, float2& sz, float2 fx, float2 fy, float2 fz, const HeisParams<float>& p,
                                              const HeisRand<float>& r0, const HeisRand<float>& r1, bool& a0, bool& a1) {
    float2 px, py, pz;
    const float2 x = heis_propose_f32<float2, FLIP, AXZ>(sx, sy, sz, fx, fy, fz, p, make_float2(r0.u0, r1.u0), make_float2(r0.u1, r1.u1), px, py, pz);
    a0 = r0.ua < fast_exp2(x.x);
    a1 = r1.ua < fast_exp2(x.y);
    if (a0) { sx.x = px.x; sy.x = py.x; sz.x = pz.x; }
    if (a1) { sx.y = px.y; sy.y = py.y; sz.y = pz.y; }
}

struct HeisGeom {
    uint32_t Gx;        // vector groups per row per colour
    uint32_t Hx;        // elements per row per colour = Lx / 2
    uint32_t Ly, Lz;
    uint32_t z_offset;
    uint32_t Lx;
};

template <typename real> struct VecOf;
template <> struct VecOf<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct VecOf<double> { typedef double2 type; static constexpr int N = 2; };

__device__ __forceinline__ void vec_load(const float* p, float (&v)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void vec_load(const double* p, double (&v)[2]) {
    const double2 q = *reinterpret_cast<const double2*>(p);
    v[0] = q.x; v[1] = q.y;
}
__device__ __forceinline__ void vec_store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void vec_store(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

template <typename real>
struct HeisPtrs {
    real* own[3];
    const real* oth[3];
    const real* oth_lo[3];  // plane below local z = 0 (halo or periodic wrap)
    const real* oth_hi[3];  // plane above local z = Lz-1
    real* peer_lo[3];       // neighbour GPU's halo that receives my plane 0 (or null)
    real* peer_hi[3];
};

// One thread owns one 16-byte vector (4 floats / 2 doubles) of a row and marches over `z_chunk` planes,
// so index arithmetic and the block reduction are paid once per thread, not once per site.  All element offsets are
// 32-bit (a colour/component array holds < 2^32 elements: checked when the family is chosen), so that every address
// is one IMAD.WIDE from a base pointer.
// MODE 0: update.  1: update + observables of the OWN colour + exchange energy (last colour pass of a recorded step).
//      3: update + observables of the own colour only (first colour pass of a recorded step).
//      2: no update; energy and the observables of BOTH colours (measure-only entry points).
// obs[0] += -sum_own s.n (exchange energy, each bond once)   obs[1..3] += sum s
// obs[4] += sum (s.a)^2                                       obs[5] += accepted (as double)
#ifndef HEIS_MINB
#define HEIS_MINB 6   // resident CTAs of 128 threads per SM the stencil kernel is compiled for (register cap)
#endif
// The march of one thread: vector (y, gx) = t2 of colour `colour`, planes [z0, z1).
//   UPDATE  attempt the moves and store (false: measure only)
//   OBS     accumulate observables into facc: [1..3] sum s, [4] sum (s.a)^2 of the own colour, and, when `energy`,
//           [0] -sum s.n (exchange energy, each bond once);  BOTH adds the other colour's s and (s.a)^2 as well
// `every16` is called (by all threads of the CTA, uniformly) after every 16th plane so that the caller can shorten the
// fp32 partial sums.
// HALO: planes outside the local z range come from P.oth_lo / P.oth_hi and boundary planes are also stored into the
// neighbours' halos (connected slab); otherwise the periodic wrap of P.oth itself is used and those twelve pointers
// are never touched (they would crowd the Philox round keys out of the uniform registers).
template <typename real, int NDIM, bool FLIP, bool UPDATE, bool OBS, bool BOTH, bool HALO, typename F>
__device__ __forceinline__ void heis_march(const HeisPtrs<real>& P, const HeisGeom& g, int colour, uint32_t t2, uint32_t z0,
                                           uint32_t z1, bool energy, const HeisParams<real>& p, uint64_t sweep,
                                           const PhiloxKey& pk, real (&facc)[5], int& accepted, F&& every16) {
    constexpr int N = VecOf<real>::N;
    const bool active = t2 < g.Ly * g.Gx;
    const uint32_t y = active ? t2 / g.Gx : 0u, gx = active ? t2 % g.Gx : 0u;
    const uint32_t ym = y == 0 ? g.Ly - 1 : y - 1, yp = y + 1 == g.Ly ? 0 : y + 1;
    const uint32_t plane = g.Ly * g.Hx;
    const uint32_t el = y * g.Hx + gx * N;                          // offsets inside a plane
    const uint32_t ela = ym * g.Hx + gx * N, elb = yp * g.Hx + gx * N;
    const uint32_t row0 = y * g.Hx;
    // carry element of the x-neighbour that lives in the adjacent group (periodic in x)
    const uint32_t c_right = row0 + ((gx + 1 == g.Gx) ? 0u : (gx + 1) * N), c_left = row0 + (gx == 0 ? g.Gx : gx) * N - 1;
    for (uint32_t zl = z0; zl < z1; ++zl) {
        if (active) {
            const uint32_t zg = zl + g.z_offset;
            const uint32_t rp = (y + zg + (uint32_t)colour) & 1u;
            const uint32_t zb = zl * plane;
            const uint32_t e0 = zb + el;
            const uint32_t e_carry = zb + (rp ? c_right : c_left);

            // planes z-1 / z+1: the base pointer choice is uniform over the CTA, the offset is one select per thread
            const bool at_lo = NDIM == 3 && zl == 0, at_hi = NDIM == 3 && zl + 1 == g.Lz;
            const uint32_t e_lo = at_lo ? (HALO ? el : (g.Lz - 1) * plane + el) : e0 - plane, e_hi = at_hi ? el : e0 + plane;
            real s[3][N], nsum[3][N], partner[3][N];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                real n0[N], a[N], b[N];
                vec_load(P.own[c] + e0, s[c]);
                vec_load(P.oth[c] + e0, n0);
                vec_load(P.oth[c] + (zb + ela), a);
                vec_load(P.oth[c] + (zb + elb), b);
                const real carry = P.oth[c][e_carry];
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    nsum[c][e] = n0[e] + (a[e] + b[e]);
                    if (BOTH) partner[c][e] = n0[e];
                }
                // x-neighbour 2: the row shifted by one element towards the carry side (a real branch, warp-uniform when a
                // warp lies inside one row: both sides are plain adds, no per-element selects)
                if (rp) {
#pragma unroll
                    for (int e = 0; e < N; ++e) nsum[c][e] += e + 1 < N ? n0[(e + 1) % N] : carry;
                } else {
#pragma unroll
                    for (int e = 0; e < N; ++e) nsum[c][e] += e > 0 ? n0[(e + N - 1) % N] : carry;
                }
                if (NDIM == 3) {
                    real lo[N], hi[N];
                    vec_load((HALO && at_lo ? P.oth_lo[c] : P.oth[c]) + e_lo, lo);
                    vec_load((HALO && at_hi ? P.oth_hi[c] : P.oth[c]) + e_hi, hi);
#pragma unroll
                    for (int e = 0; e < N; ++e) nsum[c][e] += lo[e] + hi[e];
                }
            }
            HeisRand<real> rnd[N];
            if (UPDATE) {
                const uint64_t site0 = (uint64_t)(zg * g.Ly + y) * g.Lx + 2u * (gx * N) + rp;  // element e: site0 + 2e
                if (sizeof(real) == 4) {
#pragma unroll
                    for (int e = 0; e < N; e += 2) {  // bit 1 of site0 is clear (Lx % 8 == 0): elements e, e+1 share a call
                        uint32_t r[4];
                        philox_at(site0 + 2u * e, sweep, 0u, pk, r);
                        reinterpret_cast<HeisRand<float>&>(rnd[e]) = heis_rand_words(r[0], r[1]);
                        reinterpret_cast<HeisRand<float>&>(rnd[e + 1]) = heis_rand_words(r[2], r[3]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < N; ++e) heis_rand(site0 + 2u * e, sweep, pk, rnd[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < N; ++e) {
                if (UPDATE) {
                    const bool ok = heis_attempt<real, FLIP>(s[0][e], s[1][e], s[2][e], heis_field(p.J, nsum[0][e], p.h[0]),
                                                             heis_field(p.J, nsum[1][e], p.h[1]), heis_field(p.J, nsum[2][e], p.h[2]), p, rnd[e]);
                    accepted += ok ? 1 : 0;
                }
                if (OBS) {
                    if (energy) facc[0] -= p.J * (s[0][e] * nsum[0][e] + s[1][e] * nsum[1][e] + s[2][e] * nsum[2][e]);
                    facc[1] += s[0][e]; facc[2] += s[1][e]; facc[3] += s[2][e];
                    const real d1 = s[0][e] * p.a[0] + s[1][e] * p.a[1] + s[2][e] * p.a[2];
                    facc[4] += d1 * d1;
                    if (BOTH) {
                        facc[1] += partner[0][e]; facc[2] += partner[1][e]; facc[3] += partner[2][e];
                        const real d2 = partner[0][e] * p.a[0] + partner[1][e] * p.a[1] + partner[2][e] * p.a[2];
                        facc[4] += d2 * d2;
                    }
                }
            }
            if (UPDATE) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    vec_store(P.own[c] + e0, s[c]);
                    if (NDIM == 3 && HALO) {
                        if (P.peer_lo[c] != nullptr && zl == 0) vec_store(P.peer_lo[c] + el, s[c]);
                        if (P.peer_hi[c] != nullptr && zl + 1 == g.Lz) vec_store(P.peer_hi[c] + el, s[c]);
                    }
                }
            }
        }
        if (OBS && ((zl - z0) & 15u) == 15u) every16();  // uniform branch
    }
}

// fp32 per-thread partial sums -> f64 sums of the CTA in shared memory (every thread of the CTA calls this)
template <typename real>
__device__ __forceinline__ void heis_flush(real (&facc)[5], double* s_acc) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = warp_sum((double)facc[i]);
        if ((threadIdx.x & 31u) == 0 && v != 0.0) atomicAdd(&s_acc[i], v);
        facc[i] = 0;
    }
}

template <typename real, int NDIM, bool FLIP, int MODE, bool HALO = true>
__global__ void __launch_bounds__(128, HEIS_MINB)
heis_stencil_kernel(HeisPtrs<real> P, HeisGeom g, int colour, uint32_t z_begin, uint32_t z_count, uint32_t z_chunk,
                    uint32_t z_stride /* distance between the chunk starts of consecutive blockIdx.y */,
                    HeisParams<real> p, uint64_t sweep, PhiloxKey pk, double* __restrict__ obs) {
    constexpr bool OBS = MODE != 0, ENERGY = MODE == 1 || MODE == 2, BOTH = MODE == 2, UPDATE = MODE != 2;
    __shared__ double s_acc[6];
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
    const uint32_t t2 = blockIdx.x * blockDim.x + threadIdx.x;  // (y, gx) inside a plane
    real facc[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    const uint32_t z0 = z_begin + blockIdx.y * z_stride;
    const uint32_t z1 = min(z0 + z_chunk, z_begin + z_count);
    heis_march<real, NDIM, FLIP, UPDATE, OBS, BOTH, HALO>(P, g, colour, t2, z0, z1, ENERGY, p, sweep, pk, facc, accepted,
                                                    [&]() { heis_flush(facc, s_acc); });
    if (OBS) heis_flush(facc, s_acc);
    if (UPDATE) {
        const int a = __reduce_add_sync(0xffffffffu, accepted);
        if ((threadIdx.x & 31u) == 0 && a != 0) atomicAdd(&s_acc[5], (double)a);
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_acc[threadIdx.x] != 0.0) atomicAdd(obs + threadIdx.x, s_acc[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------
// K3w: the colour passes of ONE OR SEVERAL steps as one persistent launch in wave order, so that every pass finds the
// planes the previous pass wrote in L2 (two separate passes move 36 B/attempt through DRAM; one step in wave order
// ~33 B measured; k steps per launch divide the compulsory 24 B by k).
// A phase is one colour pass: phase = 2 * step + colour.  Work items = (unit, tile); a unit is one phase on a chunk of
// C planes.  Phase p visits the chunks in the rotated order p, p+1, ..., n-1, 0, ..., p-1 (its chunk p-1 needs phase
// p-1 on chunk p-2, which that phase visits last) and trails phase p-1 by `lag` >= 3 positions.  Unit (p, j) may run once
// phase p-1 is complete on chunks j-1, j, j+1: that single rule covers the true dependencies (the neighbours' new
// values) and the anti-dependencies (phase p-1 on j+-1 has read the old values (p, j) overwrites).  `units` lists the
// units by time slot; items are dealt round-robin to the co-resident CTAs (static, in order: an item only waits for
// lower-numbered items, so the lowest unfinished item can always run).  Completion is counted per (phase, chunk) with
// release/acquire fences; the counters are monotone over the launches (each launch passes its own targets).
// ---------------------------------------------------------------------------------------
constexpr int WAVE_MAX_STEPS = 4;
struct WaveSched {
    const uint32_t* units;       // [n_units]: phase << 24 | chunk
    uint32_t n_units, tiles, C, n_chunks, n_phases;
    unsigned long long* done;    // [2 * WAVE_MAX_STEPS][n_chunks] tiles finished, monotone over the launches
    unsigned long long target[2 * WAVE_MAX_STEPS];  // value of done[p][.] once phase p is complete on a chunk in THIS launch
    unsigned int* error;         // != 0: a dependency wait timed out (results invalid)
};

// MULTI = false: one step per launch (phases 0 and 1 only; the step bookkeeping below folds away)
template <typename real, bool FLIP, bool RECORD, bool MULTI>
__global__ void __launch_bounds__(128, HEIS_MINB)
heis_wave_kernel(HeisPtrs<real> P0, HeisPtrs<real> P1, HeisGeom g, WaveSched ws, HeisParams<real> p, uint64_t sweep,
                 PhiloxKey pk, double* __restrict__ obs, int obs_stride /* doubles between the rows of consecutive steps */) {
    constexpr int KS = MULTI ? WAVE_MAX_STEPS : 1;
    __shared__ double s_acc[KS][6];
    if (threadIdx.x < KS * 6) (&s_acc[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
    real facc[5] = {0, 0, 0, 0, 0};
    int accepted = 0;
    const uint32_t n_items = ws.n_units * ws.tiles;
    uint32_t since_flush = 0, acc_step = 0;  // facc / accepted hold sums of step acc_step only
    auto flush = [&]() {
        if (RECORD) heis_flush(facc, s_acc[acc_step]);
        const int a = __reduce_add_sync(0xffffffffu, accepted);
        if ((threadIdx.x & 31u) == 0 && a != 0) atomicAdd(&s_acc[acc_step][5], (double)a);
        accepted = 0;
        since_flush = 0;
    };
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t u = item / ws.tiles, tile = item - u * ws.tiles;
        const uint32_t unit = ws.units[u];
        const uint32_t phase = unit >> 24, chunk = unit & 0x00FFFFFFu;
        const int colour = (int)(phase & 1u);
        const uint32_t step = MULTI ? phase >> 1 : 0u;
        const uint32_t z0 = chunk * ws.C, z1 = min(z0 + ws.C, g.Lz);
        if (MULTI && step != acc_step) { if (since_flush) flush(); acc_step = step; }
        if (phase > 0) {
            if (threadIdx.x < 3) {
                const uint32_t dep = threadIdx.x == 0 ? (chunk == 0 ? ws.n_chunks - 1 : chunk - 1)
                                   : threadIdx.x == 1 ? chunk : (chunk + 1 == ws.n_chunks ? 0u : chunk + 1);
                const volatile unsigned long long* d = ws.done + (size_t)(phase - 1) * ws.n_chunks + dep;
                unsigned long long target = ws.target[0];  // static indices only: a dynamic one would copy ws to local memory
                if (MULTI) {
#pragma unroll
                    for (uint32_t q = 1; q + 1 < 2 * WAVE_MAX_STEPS; ++q) if (phase - 1 == q) target = ws.target[q];
                }
                uint32_t spins = 0;
                while (*d < target) {
                    __nanosleep(256);
                    if (++spins > (1u << 22)) { atomicExch(ws.error, 1u); break; }  // never hang the GPU
                }
                __threadfence();
            }
            __syncthreads();
        }
        const uint32_t t2 = tile * blockDim.x + threadIdx.x;
        if (colour == 0)
            heis_march<real, 3, FLIP, true, RECORD, false, false>(P0, g, 0, t2, z0, z1, false, p, sweep + step, pk, facc, accepted, [] {});
        else
            heis_march<real, 3, FLIP, true, RECORD, false, false>(P1, g, 1, t2, z0, z1, true, p, sweep + step, pk, facc, accepted, [] {});
        if (MULTI ? phase + 1 < ws.n_phases : phase == 0) {
            __syncthreads();  // every thread's stores of this tile are issued
            if (threadIdx.x == 0) { __threadfence(); atomicAdd(ws.done + (size_t)phase * ws.n_chunks + chunk, 1ull); }
        }
        if (++since_flush == 8) flush();
    }
    if (since_flush) flush();
    __syncthreads();
    const uint32_t n_steps = MULTI ? ws.n_phases >> 1 : 1u;
    if (threadIdx.x < 6 * n_steps) {
        const uint32_t st = threadIdx.x / 6, k = threadIdx.x - st * 6;
        if (s_acc[st][k] != 0.0) atomicAdd(obs + (size_t)st * obs_stride + k, s_acc[st][k]);
    }
}

// ---------------------------------------------------------------------------------------
// K7: reference host layout (AoS double[3] per site, natural order) <-> colour-split SoA.
// One thread per site; reads/writes of the AoS side are 24 B strided (one-off cost).
// ---------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256)
heis_pack_kernel(const double* __restrict__ aos, real* c0x, real* c0y, real* c0z, real* c1x, real* c1y, real* c1z,
                 uint32_t Lx, uint32_t Ly, uint32_t Lz, uint32_t z_offset) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Lx * Ly * Lz;
    if (t >= total) return;
    const uint32_t x = (uint32_t)(t % Lx), y = (uint32_t)((t / Lx) % Ly), z = (uint32_t)(t / ((size_t)Lx * Ly));
    const uint32_t col = (x + y + z + z_offset) & 1u;
    const size_t e = ((size_t)z * Ly + y) * (Lx / 2) + (x >> 1);
    const double sx = aos[3 * t], sy = aos[3 * t + 1], sz = aos[3 * t + 2];
    if (col == 0) { c0x[e] = (real)sx; c0y[e] = (real)sy; c0z[e] = (real)sz; }
    else { c1x[e] = (real)sx; c1y[e] = (real)sy; c1z[e] = (real)sz; }
}

template <typename real, typename outT>
__global__ void __launch_bounds__(256)
heis_unpack_kernel(outT* __restrict__ ox, outT* __restrict__ oy, outT* __restrict__ oz, size_t stride,
                   const real* c0x, const real* c0y, const real* c0z, const real* c1x, const real* c1y,
                   const real* c1z, uint32_t Lx, uint32_t Ly, uint32_t Lz, uint32_t z_offset) {
    // stride 3 with ox=aos, oy=aos+1, oz=aos+2 writes AoS; stride 1 writes natural-order SoA
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Lx * Ly * Lz;
    if (t >= total) return;
    const uint32_t x = (uint32_t)(t % Lx), y = (uint32_t)((t / Lx) % Ly), z = (uint32_t)(t / ((size_t)Lx * Ly));
    const uint32_t col = (x + y + z + z_offset) & 1u;
    const size_t e = ((size_t)z * Ly + y) * (Lx / 2) + (x >> 1);
    ox[t * stride] = (outT)(col ? c1x[e] : c0x[e]);
    oy[t * stride] = (outT)(col ? c1y[e] : c0y[e]);
    oz[t * stride] = (outT)(col ? c1z[e] : c0z[e]);
}

// State::rand_with_size for Heisenberg spins on device, natural site index keyed.
template <typename real>
__device__ __forceinline__ void heis_random_spin(uint64_t site, PhiloxKey pk, real& x, real& y, real& z) {
    uint32_t r[4];
    philox_at(site | (1ull << 61), ~0ull, 0xFEu, pk, r);
    double dx, dy, dz;
    sphere_point(u53(r[0], r[1]), u53(r[2], r[3]), dx, dy, dz);
    x = (real)dx; y = (real)dy; z = (real)dz;
}

template <typename real>
__global__ void __launch_bounds__(256)
heis_stencil_randomize_kernel(real* c0x, real* c0y, real* c0z, real* c1x, real* c1y, real* c1z, uint32_t Lx, uint32_t Ly,
                              uint32_t Lz, uint32_t z_offset, PhiloxKey pk) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)Lx * Ly * Lz;
    if (t >= total) return;
    const uint32_t x = (uint32_t)(t % Lx), y = (uint32_t)((t / Lx) % Ly), z = (uint32_t)(t / ((size_t)Lx * Ly));
    const uint32_t col = (x + y + z + z_offset) & 1u;
    const size_t e = ((size_t)z * Ly + y) * (Lx / 2) + (x >> 1);
    const uint64_t site = ((uint64_t)(z + z_offset) * Ly + y) * Lx + x;
    real sx, sy, sz;
    heis_random_spin<real>(site, pk, sx, sy, sz);
    if (col == 0) { c0x[e] = sx; c0y[e] = sy; c0z[e] = sz; }
    else { c1x[e] = sx; c1y[e] = sy; c1z[e] = sz; }
}

}  // namespace vg
