/*
 * vegas_replay.c -- CPU replay of ONE colour-ordered sweep with the GPU's counter-based random numbers.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference draws sites at random from a PCG stream
 * (src/integrator.rs:74-76), so its trajectories cannot be compared with a checkerboard sweep.
 * What CAN be checked exactly is that every single-site decision the GPU takes is the
 * reference's Metropolis rule (src/integrator.rs:77-90 / :123-136): this file visits the sites
 * in the GPU's colour order, re-derives the GPU's Philox numbers for each site, and applies
 *     delta = H.energy(after) - H.energy(before);  accept iff delta < 0 or u < exp(-delta/T)
 * with delta evaluated by the ORACLE's restatement of Hamiltonian::energy (vegas_oracle.c),
 * i.e. by the reference arithmetic, not by the GPU's closed forms.  The documented mapping of
 * random numbers to sites is part of the product's contract (DESIGN.md "RNG keying").
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vegas_oracle.h"

static void philox_at(uint64_t index, uint64_t sweep, uint32_t call, uint64_t seed, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)sweep,
                       ((uint32_t)(sweep >> 32) & 0x00FFFFFFu) | (call << 24)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    vo_philox4x32(ctr, key, VO_PHILOX_ROUNDS, out);
}

/* accept iff U < floor(exp(-delta/T) * 2^64); delta < 0 or p >= 1 always accepts (u < 1). */
static int accept_u64(double delta, double T, uint64_t U) {
    if (delta < 0.0) return 1;
    double p = exp(-delta / T);
    if (!(p < 1.0)) return 1;
    double scaled = floor(ldexp(p, 64));
    uint64_t thr = scaled >= 18446744073709551615.0 ? ~0ull : (uint64_t)scaled;
    return U < thr;
}

/* Ising, bit-plane keyed randoms of the multi-spin-coded sc kernel (ising_msc.cuh):
 * site (x,y,z) has colour (x+y+z)&1, compact index xc = x>>1, word W = ((z*Ly+y)*(Lx/64) + xc>>5),
 * bit b = xc&31; plane j of word W is Philox(W | colour<<62, sweep, call=j>>2)[j&3]; U bit (63-j) = its bit b.
 * proposal != 0: IsingSpin::rand proposal from Philox(..., call=0xFF)[0] bit b (1 = Up).
 * Returns the number of accepted moves; state is updated in place. */
uint64_t vo_replay_ising_msc(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, uint64_t seed,
                             uint64_t sweep, uint64_t Lx, uint64_t Ly, uint64_t Lz, int8_t* state) {
    const uint64_t n = Lx * Ly * Lz, Wx = Lx / 64;
    uint64_t accepted = 0;
    for (int colour = 0; colour < 2; ++colour)
        for (uint64_t z = 0; z < Lz; ++z)
            for (uint64_t y = 0; y < Ly; ++y)
                for (uint64_t w = 0; w < Wx; ++w) {
                    const uint64_t index = ((z * Ly + y) * Wx + w) | ((uint64_t)colour << 62);
                    uint32_t planes[64], prop[4];
                    for (uint32_t ch = 0; ch < 16; ++ch) philox_at(index, sweep, ch, seed, planes + 4 * ch);
                    philox_at(index, sweep, 0xFFu, seed, prop);
                    for (uint64_t b = 0; b < 32; ++b) {
                        const uint64_t x = 2 * (32 * w + b) + ((y + z + (uint64_t)colour) & 1);
                        const uint64_t i = (z * Ly + y) * Lx + x;
                        uint64_t U = 0;
                        for (int j = 0; j < 64; ++j) U |= (uint64_t)((planes[j] >> b) & 1u) << (63 - j);
                        const int8_t old = state[i];
                        int8_t cand = (int8_t)-old;
                        if (proposal == VO_PROPOSE_RANDOM) cand = ((prop[0] >> b) & 1u) ? 1 : -1;
                        const double e_old = vo_energy(h, th, state, n, i);
                        state[i] = cand;
                        const double e_new = vo_energy(h, th, state, n, i);
                        if (accept_u64(e_new - e_old, th->temperature, U)) ++accepted;
                        else state[i] = old;
                    }
                }
    return accepted;
}

/* Ising, site keyed randoms of the general kernel (general.cuh): U = r0<<32|r1 of Philox(site, sweep, 0);
 * random proposal: Up when r2&1.  Sites are visited colour by colour. */
uint64_t vo_replay_ising_sites(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, uint64_t seed,
                               uint64_t sweep, uint64_t n, const uint8_t* colour, int n_colours, int8_t* state) {
    uint64_t accepted = 0;
    for (int c = 0; c < n_colours; ++c)
        for (uint64_t i = 0; i < n; ++i) {
            if (colour[i] != c) continue;
            uint32_t r[4];
            philox_at(i, sweep, 0u, seed, r);
            const uint64_t U = ((uint64_t)r[0] << 32) | r[1];
            const int8_t old = state[i];
            int8_t cand = (int8_t)-old;
            if (proposal == VO_PROPOSE_RANDOM) cand = (r[2] & 1u) ? 1 : -1;
            const double e_old = vo_energy(h, th, state, n, i);
            state[i] = cand;
            const double e_new = vo_energy(h, th, state, n, i);
            if (accept_u64(e_new - e_old, th->temperature, U)) ++accepted;
            else state[i] = old;
        }
    return accepted;
}

static double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }

/* Heisenberg, site keyed (heis.cuh heis_attempt).  f32 != 0 mirrors the float kernel's number formats
 * (24-bit uniforms, float proposal), the energy difference is always the oracle's f64 evaluation. */
uint64_t vo_replay_heisenberg(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, int f32, uint64_t seed,
                              uint64_t sweep, uint64_t n, const uint8_t* colour, int n_colours, double* state) {
    uint64_t accepted = 0;
    for (int c = 0; c < n_colours; ++c)
        for (uint64_t i = 0; i < n; ++i) {
            if (colour[i] != c) continue;
            uint32_t r[4], q[4];
            double old[3] = {state[3 * i], state[3 * i + 1], state[3 * i + 2]}, p[3], u;
            if (proposal == VO_PROPOSE_FLIP) { p[0] = -old[0]; p[1] = -old[1]; p[2] = -old[2]; }
            if (f32) {
                /* one call serves sites i and i^2: index i & ~2, words (0,1) / (2,3); 21+21+22 bits */
                philox_at(i & ~2ull, sweep, 0u, seed, r);
                const uint32_t w0 = (i & 2ull) ? r[2] : r[0], w1 = (i & 2ull) ? r[3] : r[1];
                if (proposal != VO_PROPOSE_FLIP) {
                    float u0 = ((float)(w0 >> 11) + 0.5f) * 0x1.0p-21f, u1 = (float)(w1 >> 11) * 0x1.0p-21f;
                    float z = 1.0f - 2.0f * u0;
                    float rxy = sqrtf(4.0f * u0 * (1.0f - u0));
                    float ang = 6.283185307179586f * (u1 - 0.5f);
                    p[0] = (double)(rxy * (float)cos((double)ang));
                    p[1] = (double)(rxy * (float)sin((double)ang));
                    p[2] = (double)z;
                }
                /* the centre of the 2^-22 cell (heis.cuh heis_rand_words): never 0, so a vanishing Boltzmann factor is never accepted */
                u = (double)(((float)(((w0 & 0x7FFu) << 11) | (w1 & 0x7FFu)) + 0.5f) * 0x1.0p-22f);
            } else {
                philox_at(i, sweep, 0u, seed, r);
                if (proposal != VO_PROPOSE_FLIP) {
                    double z = 1.0 - 2.0 * u53(r[0], r[1]);
                    double rxy = sqrt(fmax(0.0, (1.0 - z) * (1.0 + z)));
                    double ang = 2.0 * u53(r[2], r[3]);
                    p[0] = rxy * cos(M_PI * ang); p[1] = rxy * sin(M_PI * ang); p[2] = z;
                }
                philox_at(i, sweep, 1u, seed, q);
                u = u53(q[0], q[1]);
            }
            const double e_old = vo_energy(h, th, state, n, i);
            state[3 * i] = p[0]; state[3 * i + 1] = p[1]; state[3 * i + 2] = p[2];
            const double e_new = vo_energy(h, th, state, n, i);
            const double delta = e_new - e_old;
            if (delta < 0.0 || u < exp(-delta / th->temperature)) { ++accepted; continue; }
            state[3 * i] = old[0]; state[3 * i + 1] = old[1]; state[3 * i + 2] = old[2];
        }
    return accepted;
}
