/*
 * vegas_oracle.c -- CPU restatement of the vegas-rs 0.9.0 Metropolis path (plain C11).
 *
 * TEST INFRASTRUCTURE ONLY (see vegas_oracle.h).  Written from the behaviour of the
 * reference, one function per reference function, each with its file:line.  It keeps
 * the reference's data layouts on purpose (1-byte Ising spins, 24-byte AoS f64 Heisenberg
 * spins, CSR with 8-byte indices + f64 values, random site selection with replacement, two
 * energy() calls per attempt, a state clone per step, per-sensor total_energy per step) so
 * that timing it is a fair CPU baseline of the reference's algorithm ("kind": "port").
 *
 * parity unpinned: lattice adjacency (vegas-lattice 0.13), CSR assembly (sprs 0.11) and the
 * RNG bit stream (rand 0.9 / rand_pcg 0.9) are external crates without reference vectors.
 */
#include "vegas_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== RNG */
/* rand_pcg::Pcg64 = Lcg128Xsl64 (PCG XSL-RR 128/64).  main.rs:27-30 seeds it with
 * seed_from_u64; the exact expansion is a rand_core detail (unpinned) -> splitmix64 here. */
static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static const unsigned __int128 PCG_MULT =
    ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;

void vo_rng_seed(vo_rng* r, uint64_t seed) {
    uint64_t s = seed;
    uint64_t a = splitmix64(&s), b = splitmix64(&s), c = splitmix64(&s), d = splitmix64(&s);
    r->inc = ((((unsigned __int128)a) << 64) | b) | 1;
    r->state = (((unsigned __int128)c) << 64) | d;
    r->state = r->state * PCG_MULT + r->inc;
}

uint64_t vo_rng_u64(vo_rng* r) {
    r->state = r->state * PCG_MULT + r->inc;
    uint64_t hi = (uint64_t)(r->state >> 64), lo = (uint64_t)r->state;
    unsigned rot = (unsigned)(r->state >> 122);
    uint64_t x = hi ^ lo;
    return (x >> rot) | (x << ((64 - rot) & 63));
}

/* rand 0.9 StandardUniform for f64: 53 random bits scaled by 2^-53, in [0,1). */
double vo_rng_f64(vo_rng* r) { return (double)(vo_rng_u64(r) >> 11) * 0x1.0p-53; }

/* Uniform<usize>::sample: unbiased widening-multiply rejection (Lemire). */
uint64_t vo_rng_below(vo_rng* r, uint64_t n) {
    unsigned __int128 m = (unsigned __int128)vo_rng_u64(r) * n;
    uint64_t l = (uint64_t)m;
    if (l < n) {
        uint64_t t = (0 - n) % n;
        while (l < t) {
            m = (unsigned __int128)vo_rng_u64(r) * n;
            l = (uint64_t)m;
        }
    }
    return (uint64_t)(m >> 64);
}

/* Uniform<f64>::new(lo,hi).sample: lo + (hi-lo)*u with u in [0,1). */
double vo_rng_range(vo_rng* r, double lo, double hi) { return lo + (hi - lo) * vo_rng_f64(r); }

/* ======================================================================== lattice */
/* Unit cells as vegas-lattice 0.13 is understood to define them (unpinned):
 * sc: 1 site, bonds +x,+y,+z; bcc: A(0,0,0),B(1/2,1/2,1/2), 8 A->B bonds with cell deltas in
 * {0,-1}^3; fcc: A(0,0,0),B(1/2,1/2,0),C(1/2,0,1/2),D(0,1/2,1/2), 24 bonds per cell. */
typedef struct { int s, t, dx, dy, dz; } uc_edge;

static const uc_edge SC_EDGES[3] = {{0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}};
static const uc_edge BCC_EDGES[8] = {
    {0, 1, 0, 0, 0},  {0, 1, -1, 0, 0},  {0, 1, 0, -1, 0},  {0, 1, -1, -1, 0},
    {0, 1, 0, 0, -1}, {0, 1, -1, 0, -1}, {0, 1, 0, -1, -1}, {0, 1, -1, -1, -1}};
static const uc_edge FCC_EDGES[24] = {
    /* A-B */ {0, 1, 0, 0, 0}, {0, 1, -1, 0, 0}, {0, 1, 0, -1, 0}, {0, 1, -1, -1, 0},
    /* A-C */ {0, 2, 0, 0, 0}, {0, 2, -1, 0, 0}, {0, 2, 0, 0, -1}, {0, 2, -1, 0, -1},
    /* A-D */ {0, 3, 0, 0, 0}, {0, 3, 0, -1, 0}, {0, 3, 0, 0, -1}, {0, 3, 0, -1, -1},
    /* B-C */ {1, 2, 0, 0, 0}, {1, 2, 0, 1, 0}, {1, 2, 0, 0, -1}, {1, 2, 0, 1, -1},
    /* B-D */ {1, 3, 0, 0, 0}, {1, 3, 1, 0, 0}, {1, 3, 0, 0, -1}, {1, 3, 1, 0, -1},
    /* C-D */ {2, 3, 0, 0, 0}, {2, 3, 1, 0, 0}, {2, 3, 0, -1, 0}, {2, 3, 1, -1, 0}};

/* input.rs:296-322: unit cell -> expand(x,y,z) -> drop_{x,y,z} for non-periodic axes
 * (drop removes the bonds that cross the supercell boundary along that axis). */
int vo_lattice_build(int unitcell, uint64_t nx, uint64_t ny, uint64_t nz, int pbc_x, int pbc_y, int pbc_z,
                     vo_lattice* out) {
    const uc_edge* tab;
    int ne, nb;
    switch (unitcell) {
        case VO_SC: tab = SC_EDGES; ne = 3; nb = 1; break;
        case VO_BCC: tab = BCC_EDGES; ne = 8; nb = 2; break;
        case VO_FCC: tab = FCC_EDGES; ne = 24; nb = 4; break;
        default: return -1;
    }
    uint64_t cells = nx * ny * nz;
    out->n_sites = cells * (uint64_t)nb;
    out->src = (uint64_t*)malloc(sizeof(uint64_t) * cells * ne + 8);
    out->dst = (uint64_t*)malloc(sizeof(uint64_t) * cells * ne + 8);
    if (!out->src || !out->dst) return -2;
    uint64_t k = 0;
    for (uint64_t iz = 0; iz < nz; ++iz)
        for (uint64_t iy = 0; iy < ny; ++iy)
            for (uint64_t ix = 0; ix < nx; ++ix)
                for (int e = 0; e < ne; ++e) {
                    int64_t tx = (int64_t)ix + tab[e].dx, ty = (int64_t)iy + tab[e].dy, tz = (int64_t)iz + tab[e].dz;
                    if (tx < 0 || tx >= (int64_t)nx) { if (!pbc_x) continue; tx = (tx + (int64_t)nx) % (int64_t)nx; }
                    if (ty < 0 || ty >= (int64_t)ny) { if (!pbc_y) continue; ty = (ty + (int64_t)ny) % (int64_t)ny; }
                    if (tz < 0 || tz >= (int64_t)nz) { if (!pbc_z) continue; tz = (tz + (int64_t)nz) % (int64_t)nz; }
                    out->src[k] = ((iz * ny + iy) * nx + ix) * nb + tab[e].s;
                    out->dst[k] = (((uint64_t)tz * ny + (uint64_t)ty) * nx + (uint64_t)tx) * nb + tab[e].t;
                    ++k;
                }
    out->n_edges = k;
    return 0;
}

void vo_lattice_free(vo_lattice* l) {
    free(l->src); free(l->dst);
    l->src = l->dst = NULL; l->n_edges = l->n_sites = 0;
}

/* ======================================================================== CSR */
/* sprs TriMat::to_csr: duplicates summed, column indices sorted within each row. */
int vo_csr_from_triplets(uint64_t n, uint64_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals,
                         vo_csr* out) {
    uint64_t* cnt = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    uint64_t* tc = (uint64_t*)malloc(sizeof(uint64_t) * (nnz + 1));
    double* tv = (double*)malloc(sizeof(double) * (nnz + 1));
    if (!cnt || !tc || !tv) return -2;
    for (uint64_t k = 0; k < nnz; ++k) cnt[rows[k] + 1]++;
    for (uint64_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    uint64_t* pos = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
    if (!pos) return -2;
    memcpy(pos, cnt, sizeof(uint64_t) * (n + 1));
    for (uint64_t k = 0; k < nnz; ++k) {
        uint64_t p = pos[rows[k]]++;
        tc[p] = cols[k];
        tv[p] = vals[k];
    }
    out->n = n;
    out->row_ptr = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
    out->col_idx = (uint64_t*)malloc(sizeof(uint64_t) * (nnz + 1));
    out->values = (double*)malloc(sizeof(double) * (nnz + 1));
    if (!out->row_ptr || !out->col_idx || !out->values) return -2;
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; ++i) {
        out->row_ptr[i] = w;
        uint64_t a = cnt[i], b = cnt[i + 1];
        for (uint64_t p = a + 1; p < b; ++p) { /* stable insertion sort: rows are short */
            uint64_t c = tc[p]; double v = tv[p]; uint64_t q = p;
            while (q > a && tc[q - 1] > c) { tc[q] = tc[q - 1]; tv[q] = tv[q - 1]; --q; }
            tc[q] = c; tv[q] = v;
        }
        for (uint64_t p = a; p < b; ++p) {
            if (w > out->row_ptr[i] && out->col_idx[w - 1] == tc[p]) out->values[w - 1] += tv[p];
            else { out->col_idx[w] = tc[p]; out->values[w] = tv[p]; ++w; }
        }
    }
    out->row_ptr[n] = w;
    free(cnt); free(pos); free(tc); free(tv);
    return 0;
}

/* Exchange::from_lattice, energy.rs:176-187. */
int vo_csr_from_lattice(const vo_lattice* l, double exchange, int literal_filter, vo_csr* out) {
    uint64_t cap = 2 * l->n_edges + 1, k = 0;
    uint64_t* r = (uint64_t*)malloc(sizeof(uint64_t) * cap);
    uint64_t* c = (uint64_t*)malloc(sizeof(uint64_t) * cap);
    double* v = (double*)malloc(sizeof(double) * cap);
    if (!r || !c || !v) return -2;
    for (uint64_t e = 0; e < l->n_edges; ++e) {
        if (literal_filter && !(l->src[e] <= l->dst[e])) continue; /* energy.rs:180 */
        r[k] = l->src[e]; c[k] = l->dst[e]; v[k] = exchange; ++k;  /* :181 */
        r[k] = l->dst[e]; c[k] = l->src[e]; v[k] = exchange; ++k;  /* :182 */
    }
    int rc = vo_csr_from_triplets(l->n_sites, k, r, c, v, out);
    free(r); free(c); free(v);
    return rc;
}

void vo_csr_free(vo_csr* m) {
    free(m->row_ptr); free(m->col_idx); free(m->values);
    m->row_ptr = m->col_idx = NULL; m->values = NULL; m->n = 0;
}

/* ======================================================================== spins */
/* IsingSpin::dot state.rs:95-101; HeisenbergSpin::dot state.rs:163-171 (fold from 0). */
static inline double dot_ising(int8_t a, int8_t b) { return a == b ? 1.0 : -1.0; }
static inline double dot_heis(const double* a, const double* b) {
    double s = 0.0;
    s = s + a[0] * b[0];
    s = s + a[1] * b[1];
    s = s + a[2] * b[2];
    return s;
}
static inline int8_t ising_of(double z) { return z >= 0.0 ? 1 : -1; }

/* util.rs:21-34 */
void vo_marsaglia(vo_rng* r, double out[3]) {
    for (;;) {
        double x1 = vo_rng_range(r, -1.0, 1.0);
        double x2 = vo_rng_range(r, -1.0, 1.0);
        if (x1 * x1 + x2 * x2 >= 1.0) continue;
        out[0] = 2.0 * x1 * sqrt(1.0 - x1 * x1 - x2 * x2);
        out[1] = 2.0 * x2 * sqrt(1.0 - x1 * x1 - x2 * x2);
        out[2] = 1.0 - 2.0 * (x1 * x1 + x2 * x2);
        return;
    }
}

/* IsingSpin::rand state.rs:76-84: Uniform(0,1) < 0.5 -> Up. */
static inline int8_t ising_rand(vo_rng* r) { return vo_rng_range(r, 0.0, 1.0) < 0.5 ? 1 : -1; }

/* State::rand_with_size state.rs:260-262 */
void vo_state_rand(int model, vo_rng* r, void* state, uint64_t n) {
    if (model == VO_ISING) {
        int8_t* s = (int8_t*)state;
        for (uint64_t i = 0; i < n; ++i) s[i] = ising_rand(r);
    } else {
        double* s = (double*)state;
        for (uint64_t i = 0; i < n; ++i) vo_marsaglia(r, s + 3 * i);
    }
}

/* ======================================================================== thermostat */
/* Thermostat::new / with_temperature thermostat.rs:29-60: T clamped to >= f64::EPSILON. */
vo_thermostat_t vo_thermostat(double temperature, const double dir[3], double mag) {
    vo_thermostat_t t;
    t.temperature = temperature < DBL_EPSILON ? DBL_EPSILON : temperature;
    t.field_dir[0] = dir ? dir[0] : 0.0;
    t.field_dir[1] = dir ? dir[1] : 0.0;
    t.field_dir[2] = dir ? dir[2] : 1.0;
    t.field_mag = mag;
    return t;
}

/* ======================================================================== Hamiltonian terms */
static inline double term_energy_ising(const vo_hamiltonian* h, const vo_thermostat_t* th, const int8_t* s, uint64_t i,
                                       int term) {
    switch (term) {
        case VO_TERM_GAUGE: return h->gauge;                                          /* energy.rs:75-79 */
        case VO_TERM_ANISOTROPY: {                                                    /* energy.rs:108-112 */
            double d = dot_ising(s[i], ising_of(h->aniso_axis[2]));
            return d * d * h->aniso_k;
        }
        case VO_TERM_ZEEMAN:                                                          /* energy.rs:147-151 */
            return dot_ising(s[i], ising_of(th->field_dir[2])) * fabs(th->field_mag);
        default: {                                                                    /* energy.rs:194-206 */
            const vo_csr* m = h->exchange;
            if (!m || i >= m->n) return 0.0;
            double acc = 0.0;
            for (uint64_t p = m->row_ptr[i]; p < m->row_ptr[i + 1]; ++p)
                acc = acc + (-m->values[p] * dot_ising(s[i], s[m->col_idx[p]]));
            return acc;
        }
    }
}

static inline double term_energy_heis(const vo_hamiltonian* h, const vo_thermostat_t* th, const double* s, uint64_t i,
                                      int term) {
    switch (term) {
        case VO_TERM_GAUGE: return h->gauge;
        case VO_TERM_ANISOTROPY: {
            double d = dot_heis(s + 3 * i, h->aniso_axis);
            return d * d * h->aniso_k;
        }
        case VO_TERM_ZEEMAN: return dot_heis(s + 3 * i, th->field_dir) * fabs(th->field_mag);
        default: {
            const vo_csr* m = h->exchange;
            if (!m || i >= m->n) return 0.0;
            double acc = 0.0;
            for (uint64_t p = m->row_ptr[i]; p < m->row_ptr[i + 1]; ++p)
                acc = acc + (-m->values[p] * dot_heis(s + 3 * i, s + 3 * m->col_idx[p]));
            return acc;
        }
    }
}

/* Compound::energy energy.rs:254-256, left-nested by the hamiltonian! macro (:273-290). */
double vo_energy(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n, uint64_t i) {
    (void)n;
    double e;
    if (h->model == VO_ISING) {
        e = term_energy_ising(h, th, (const int8_t*)state, i, h->terms[0]);
        for (int t = 1; t < h->n_terms; ++t) e = e + term_energy_ising(h, th, (const int8_t*)state, i, h->terms[t]);
    } else {
        e = term_energy_heis(h, th, (const double*)state, i, h->terms[0]);
        for (int t = 1; t < h->n_terms; ++t) e = e + term_energy_heis(h, th, (const double*)state, i, h->terms[t]);
    }
    return e;
}

double vo_total_energy(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n) {
    if (h->n_terms == 1) {
        int term = h->terms[0];
        if (term == VO_TERM_ANISOTROPY) { /* energy.rs:114-120 -- strength is NOT applied */
            double acc = 0.0;
            for (uint64_t i = 0; i < n; ++i) {
                double d = h->model == VO_ISING ? dot_ising(((const int8_t*)state)[i], ising_of(h->aniso_axis[2]))
                                                : dot_heis((const double*)state + 3 * i, h->aniso_axis);
                acc += d * d;
            }
            return acc;
        }
        if (term == VO_TERM_ZEEMAN) { /* energy.rs:153-160 -- opposite sign to energy() */
            double acc = 0.0;
            for (uint64_t i = 0; i < n; ++i)
                acc += h->model == VO_ISING ? dot_ising(((const int8_t*)state)[i], ising_of(th->field_dir[2]))
                                            : dot_heis((const double*)state + 3 * i, th->field_dir);
            return -fabs(th->field_mag) * acc;
        }
        if (term == VO_TERM_EXCHANGE) { /* energy.rs:208-213 */
            double acc = 0.0;
            for (uint64_t i = 0; i < n; ++i) acc = acc + vo_energy(h, th, state, n, i);
            return acc / 2.0;
        }
    }
    double acc = 0.0; /* trait default energy.rs:55-59 */
    for (uint64_t i = 0; i < n; ++i) acc += vo_energy(h, th, state, n, i);
    return acc;
}

void vo_site_energies(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n, double* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = vo_energy(h, th, state, n, i);
}

/* integrator.rs:77-81 / :123-127 for every site of a fixed state. */
void vo_delta_energies(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n,
                       const void* proposal, double* out) {
    if (h->model == VO_ISING) {
        int8_t* s = (int8_t*)malloc(n ? n : 1);
        memcpy(s, state, n);
        for (uint64_t i = 0; i < n; ++i) {
            double e_old = vo_energy(h, th, s, n, i);
            int8_t old = s[i];
            s[i] = proposal ? ((const int8_t*)proposal)[i] : (int8_t)-old;
            double e_new = vo_energy(h, th, s, n, i);
            s[i] = old;
            out[i] = e_new - e_old;
        }
        free(s);
    } else {
        double* s = (double*)malloc(sizeof(double) * 3 * (n ? n : 1));
        memcpy(s, state, sizeof(double) * 3 * n);
        for (uint64_t i = 0; i < n; ++i) {
            double e_old = vo_energy(h, th, s, n, i);
            double old[3] = {s[3 * i], s[3 * i + 1], s[3 * i + 2]};
            for (int c = 0; c < 3; ++c) s[3 * i + c] = proposal ? ((const double*)proposal)[3 * i + c] : -old[c];
            double e_new = vo_energy(h, th, s, n, i);
            for (int c = 0; c < 3; ++c) s[3 * i + c] = old[c];
            out[i] = e_new - e_old;
        }
        free(s);
    }
}

/* State::magnetization state.rs:291-296 -> Sum state.rs:235-242 -> from_projections :86-92,:150-160;
 * Field::magnitude state.rs:219-221. */
double vo_magnetization(int model, const void* state, uint64_t n, double out_xyz[3]) {
    double px = 0.0, py = 0.0, pz = 0.0;
    if (model == VO_ISING) {
        const int8_t* s = (const int8_t*)state;
        for (uint64_t i = 0; i < n; ++i) pz = pz + (s[i] > 0 ? 1.0 : -1.0);
        if (out_xyz) { out_xyz[0] = 0.0; out_xyz[1] = 0.0; out_xyz[2] = pz; }
        return fabs(fabs(pz));
    }
    const double* s = (const double*)state;
    for (uint64_t i = 0; i < n; ++i) { px = px + s[3 * i]; py = py + s[3 * i + 1]; pz = pz + s[3 * i + 2]; }
    if (out_xyz) { out_xyz[0] = px; out_xyz[1] = py; out_xyz[2] = pz; }
    double mag = sqrt(px * px + py * py + pz * pz);
    if (fabs(mag) < DBL_EPSILON) return 0.0;
    return fabs(mag);
}

/* ======================================================================== integrators */
/* MetropolisIntegrator::step integrator.rs:66-92 and MetropolisFlipIntegrator::step :109-138. */
uint64_t vo_metropolis_step(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, vo_rng* r, void* state,
                            uint64_t n) {
    uint64_t accepted = 0;
    if (h->model == VO_ISING) {
        int8_t* s = (int8_t*)state;
        for (uint64_t a = 0; a < n; ++a) {
            uint64_t i = vo_rng_below(r, n);
            double e_old = vo_energy(h, th, s, n, i);
            int8_t old = s[i];
            s[i] = proposal == VO_PROPOSE_FLIP ? (int8_t)-old : ising_rand(r);
            double e_new = vo_energy(h, th, s, n, i);
            double delta = e_new - e_old;
            if (delta < 0.0) { ++accepted; continue; }
            if (vo_rng_f64(r) < exp(-delta / th->temperature)) { ++accepted; continue; }
            s[i] = old;
        }
    } else {
        double* s = (double*)state;
        for (uint64_t a = 0; a < n; ++a) {
            uint64_t i = vo_rng_below(r, n);
            double e_old = vo_energy(h, th, s, n, i);
            double old[3] = {s[3 * i], s[3 * i + 1], s[3 * i + 2]};
            if (proposal == VO_PROPOSE_FLIP) { s[3 * i] = -old[0]; s[3 * i + 1] = -old[1]; s[3 * i + 2] = -old[2]; }
            else vo_marsaglia(r, s + 3 * i);
            double e_new = vo_energy(h, th, s, n, i);
            double delta = e_new - e_old;
            if (delta < 0.0) { ++accepted; continue; }
            if (vo_rng_f64(r) < exp(-delta / th->temperature)) { ++accepted; continue; }
            s[3 * i] = old[0]; s[3 * i + 1] = old[1]; s[3 * i + 2] = old[2];
        }
    }
    return accepted;
}

/* ======================================================================== accumulator.rs:23-64 */
void vo_acc_reset(vo_acc* a) { a->sum = a->sum_sq = a->sum_fourth = 0.0; a->count = 0; }
void vo_acc_collect(vo_acc* a, double v) {
    a->sum += v; a->sum_sq += v * v; a->sum_fourth += v * v * v * v; a->count += 1;
}
double vo_acc_mean(const vo_acc* a) { return a->sum / (double)a->count; }
double vo_acc_variance(const vo_acc* a) { double m = vo_acc_mean(a); return a->sum_sq / (double)a->count - m * m; }
double vo_acc_binder(const vo_acc* a) {
    double m2 = a->sum_sq / (double)a->count;
    return 1.0 - (a->sum_fourth / (double)a->count) / (3.0 * (m2 * m2));
}

/* ======================================================================== machine.rs:91-125 + sensors */
int vo_machine_init(vo_machine* m, const vo_hamiltonian* h, int proposal, vo_rng* rng, void* state, uint64_t n,
                    int n_sensors) {
    memset(m, 0, sizeof(*m));
    m->h = h; m->proposal = proposal; m->rng = rng; m->state = state; m->n = n; m->n_sensors = n_sensors;
    m->th = vo_thermostat(2.8, NULL, 0.0); /* input.rs:273-279 */
    size_t bytes = (h->model == VO_ISING ? 1 : 24) * (size_t)(n ? n : 1);
    m->scratch = malloc(bytes);
    return m->scratch ? VO_OK : VO_ERR_ALLOC;
}

void vo_machine_free(vo_machine* m) {
    free(m->scratch); free(m->rows); free(m->obs_energy); free(m->obs_mag);
    memset(m, 0, sizeof(*m));
}

static int push_obs(vo_machine* m, double e, double mag) {
    if (m->obs_len == m->obs_cap) {
        uint64_t cap = m->obs_cap ? 2 * m->obs_cap : 1024;
        double* a = (double*)realloc(m->obs_energy, sizeof(double) * cap);
        double* b = (double*)realloc(m->obs_mag, sizeof(double) * cap);
        if (a) m->obs_energy = a;
        if (b) m->obs_mag = b;
        if (!a || !b) return VO_ERR_ALLOC;
        m->obs_cap = cap;
    }
    m->obs_energy[m->obs_len] = e; m->obs_mag[m->obs_len] = mag; m->obs_len++;
    return VO_OK;
}

/* Machine::run machine.rs:91-101.  measuring!=0 between on_measure_start and on_measure_end. */
static int machine_run(vo_machine* m, uint64_t steps, int measuring, vo_acc* e_acc, vo_acc* m_acc) {
    size_t bytes = (m->h->model == VO_ISING ? 1 : 24) * (size_t)m->n;
    for (uint64_t s = 0; s < steps; ++s) {
        memcpy(m->scratch, m->state, bytes); /* state.clone() machine.rs:95 */
        vo_metropolis_step(m->h, &m->th, m->proposal, m->rng, m->state, m->n);
        m->attempts += m->n;
        if (m->n_sensors >= 1 && measuring) { /* StatSensor::after_step instrument.rs:133-141 */
            double e = vo_total_energy(m->h, &m->th, m->state, m->n);
            double mg = vo_magnetization(m->h->model, m->state, m->n, NULL);
            vo_acc_collect(e_acc, e); vo_acc_collect(m_acc, mg);
        }
        if (m->n_sensors >= 2) { /* ObservableSensor::after_step instrument.rs:254-262 (relax AND measure) */
            double e = vo_total_energy(m->h, &m->th, m->state, m->n);
            double mg = vo_magnetization(m->h->model, m->state, m->n, NULL);
            int rc = push_obs(m, e, mg);
            if (rc) return rc;
        }
    }
    return VO_OK;
}

int vo_relax_for(vo_machine* m, uint64_t steps) { return machine_run(m, steps, 0, NULL, NULL); }

int vo_measure_for(vo_machine* m, uint64_t steps) {
    vo_acc ea, ma;
    vo_acc_reset(&ea); vo_acc_reset(&ma);
    int rc = machine_run(m, steps, 1, &ea, &ma);
    if (rc) return rc;
    if (m->n_sensors >= 1) { /* StatSensor::on_measure_end instrument.rs:110-131 */
        if (m->rows_len == m->rows_cap) {
            uint64_t cap = m->rows_cap ? 2 * m->rows_cap : 64;
            vo_stat_row* r = (vo_stat_row*)realloc(m->rows, sizeof(vo_stat_row) * cap);
            if (!r) return VO_ERR_ALLOC;
            m->rows = r; m->rows_cap = cap;
        }
        double T = m->th.temperature, nn = (double)m->n;
        vo_stat_row* row = &m->rows[m->rows_len++];
        row->temperature = T;
        row->field = fabs(m->th.field_mag);
        row->mean_e = vo_acc_mean(&ea);
        row->cv = vo_acc_variance(&ea) / (nn * (T * T));
        row->mean_m = vo_acc_mean(&ma);
        row->chi = vo_acc_variance(&ma) / (nn * T);
        row->binder = vo_acc_binder(&ma);
    }
    return VO_OK;
}

/* ======================================================================== program.rs */
static void set_temperature(vo_machine* m, double t) { m->th.temperature = t < DBL_EPSILON ? DBL_EPSILON : t; }

int vo_program_relax(vo_machine* m, uint64_t steps, double temperature) { /* program.rs:97-115 */
    if (steps == 0) return VO_ERR_NO_STEPS;
    if (temperature < DBL_EPSILON) return VO_ERR_ZERO_TEMPERATURE;
    set_temperature(m, temperature);
    return vo_relax_for(m, steps);
}

int vo_program_cooldown(vo_machine* m, double tmax, double tmin, double rate, uint64_t relax, uint64_t steps) {
    if (tmax < tmin) return VO_ERR_TMAX_LT_TMIN; /* program.rs:190-201 */
    if (steps == 0) return VO_ERR_NO_STEPS;
    if (tmin < DBL_EPSILON) return VO_ERR_ZERO_TEMPERATURE;
    if (rate < DBL_EPSILON) return VO_ERR_ZERO_COOL_RATE;
    double t = tmax;
    for (;;) { /* program.rs:202-211 */
        set_temperature(m, t);
        int rc = vo_relax_for(m, relax);
        if (rc) return rc;
        rc = vo_measure_for(m, steps);
        if (rc) return rc;
        t -= rate;
        if (t < tmin) break;
    }
    return VO_OK;
}

uint64_t vo_cooldown_points(double tmax, double tmin, double rate, double* out, uint64_t cap) {
    uint64_t k = 0;
    double t = tmax;
    for (;;) {
        if (out && k < cap) out[k] = t < DBL_EPSILON ? DBL_EPSILON : t;
        ++k;
        t -= rate;
        if (t < tmin) break;
    }
    return k;
}

static int hyst_point(vo_machine* m, double magnitude, uint64_t relax, uint64_t steps) {
    m->th.field_dir[0] = 0.0; m->th.field_dir[1] = 0.0; m->th.field_dir[2] = 1.0; /* Field::new(S::up(), magnitude) */
    m->th.field_mag = magnitude;
    int rc = vo_relax_for(m, relax);
    if (rc) return rc;
    return vo_measure_for(m, steps);
}

int vo_program_hysteresis(vo_machine* m, uint64_t steps, uint64_t relax, double temperature, double max_field,
                          double field_step) { /* program.rs:281-336 */
    if (steps == 0) return VO_ERR_NO_STEPS;
    if (temperature < DBL_EPSILON) return VO_ERR_ZERO_TEMPERATURE;
    if (max_field < DBL_EPSILON) return VO_ERR_ZERO_FIELD;
    if (field_step < DBL_EPSILON) return VO_ERR_ZERO_FIELD_STEP;
    set_temperature(m, temperature);
    double mag = 0.0;
    int rc;
    for (;;) { if ((rc = hyst_point(m, mag, relax, steps))) return rc; mag += field_step; if (mag > max_field) break; }
    for (;;) { if ((rc = hyst_point(m, mag, relax, steps))) return rc; mag -= field_step; if (mag < -max_field) break; }
    for (;;) { if ((rc = hyst_point(m, mag, relax, steps))) return rc; mag += field_step; if (mag > max_field) break; }
    return VO_OK;
}

uint64_t vo_hysteresis_points(double max_field, double field_step, double* out, uint64_t cap) {
    uint64_t k = 0;
    double mag = 0.0;
    for (;;) { if (out && k < cap) out[k] = mag; ++k; mag += field_step; if (mag > max_field) break; }
    for (;;) { if (out && k < cap) out[k] = mag; ++k; mag -= field_step; if (mag < -max_field) break; }
    for (;;) { if (out && k < cap) out[k] = mag; ++k; mag += field_step; if (mag > max_field) break; }
    return k;
}

/* instrument.rs:113-123: seven "{:.16}" fields separated by single spaces. */
int vo_stat_line(const vo_stat_row* r, char* buf, size_t cap) {
    return snprintf(buf, cap, "%.16f %.16f %.16f %.16f %.16f %.16f %.16f", r->temperature, r->field, r->mean_e, r->cv,
                    r->mean_m, r->chi, r->binder);
}

/* ======================================================================== Philox4x32-R */
/* Random123 Philox4x32 with R rounds; constants as in curand_philox4x32_x.h:88-91.  Used only by the
 * replay checker (oracle/vegas_replay.c) to re-derive the GPU's random numbers on the CPU. */
void vo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void vo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { vo_philox4x32(ctr, key, 10, out); }
