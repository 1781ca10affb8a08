/*
 * vegas_oracle.h -- CPU restatement of the vegas-rs 0.9.0 Metropolis path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path
 * (vegas_rs_b200/) never links, imports or executes anything in oracle/.
 *
 * The reference is Rust and cannot be built in this image (no cargo/rustc), so there is
 * no oracle/_ref.  Parity status:
 *   - PINNED by the reference's own unit tests (true values of src/energy.rs:301-375 and
 *     src/state.rs:331-376; see tests/golden/reference_known_answers.json).
 *   - UNPINNED ("parity unpinned"): adjacency produced by the external crate
 *     vegas-lattice 0.13 (not in /root/reference), CSR assembly by sprs 0.11, and the
 *     bit stream of rand 0.9 / rand_pcg 0.9.  No reference test holds a vector for them.
 *     They are restated from their published behaviour and documented in DESIGN.md.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#ifndef VEGAS_ORACLE_H
#define VEGAS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- RNG (rand_pcg::Pcg64 shape) */
typedef struct {
    unsigned __int128 state;
    unsigned __int128 inc;
} vo_rng;

void vo_rng_seed(vo_rng* r, uint64_t seed);          /* main.rs:27-30 Pcg64::seed_from_u64 (stream unpinned) */
uint64_t vo_rng_u64(vo_rng* r);
double vo_rng_f64(vo_rng* r);                        /* rng.random::<f64>(): 53-bit [0,1) integrator.rs:85 */
uint64_t vo_rng_below(vo_rng* r, uint64_t n);        /* Uniform::new(0,n).sample integrator.rs:74-76 */
double vo_rng_range(vo_rng* r, double lo, double hi);/* Uniform::new(lo,hi).sample util.rs:23-25 */

/* ---------------------------------------------------------------- lattice (vegas-lattice 0.13 restated) */
enum { VO_SC = 0, VO_BCC = 1, VO_FCC = 2 };

typedef struct {
    uint64_t n_sites;
    uint64_t n_edges;
    uint64_t* src;  /* edge source site */
    uint64_t* dst;  /* edge target site */
} vo_lattice;

/* Lattice::{sc,bcc,fcc}(1.0).expand(x,y,z) then drop_{x,y,z} when !pbc (input.rs:296-322).
 * Site index = ((iz*ny+iy)*nx+ix)*n_basis + b.  Each bond listed once. */
int vo_lattice_build(int unitcell, uint64_t nx, uint64_t ny, uint64_t nz,
                     int pbc_x, int pbc_y, int pbc_z, vo_lattice* out);
void vo_lattice_free(vo_lattice* l);

/* ---------------------------------------------------------------- CSR (sprs TriMat::to_csr restated) */
typedef struct {
    uint64_t n;
    uint64_t* row_ptr;  /* n+1 */
    uint64_t* col_idx;  /* usize indices, 8 B as in the reference */
    double* values;
} vo_csr;

/* Exchange::from_lattice (energy.rs:176-187).  literal_filter=1 keeps the reference's
 * `source <= target` test (energy.rs:180); 0 inserts every edge in both directions. */
int vo_csr_from_lattice(const vo_lattice* l, double exchange, int literal_filter, vo_csr* out);
/* TriMat -> CSR from raw triplets: duplicates summed, columns sorted ascending. */
int vo_csr_from_triplets(uint64_t n, uint64_t nnz, const uint64_t* rows, const uint64_t* cols,
                         const double* vals, vo_csr* out);
void vo_csr_free(vo_csr* m);

/* ---------------------------------------------------------------- Hamiltonian */
enum { VO_ISING = 0, VO_HEISENBERG = 1 };
enum { VO_TERM_GAUGE = 0, VO_TERM_ANISOTROPY = 1, VO_TERM_ZEEMAN = 2, VO_TERM_EXCHANGE = 3 };

typedef struct {
    int model;            /* VO_ISING: spins are int8 +1/-1; VO_HEISENBERG: double[3] AoS */
    int n_terms;          /* 1..4, evaluated left-nested in this order (energy.rs:254-256, :273-290) */
    int terms[4];
    double gauge;         /* energy.rs:63-80 */
    double aniso_k;       /* energy.rs:83-121 */
    double aniso_axis[3]; /* Heisenberg reference spin; Ising uses aniso_axis[2] sign (+1 up / -1 down) */
    const vo_csr* exchange; /* energy.rs:164-214 */
} vo_hamiltonian;

typedef struct {
    double temperature;   /* already clamped by vo_thermostat() */
    double field_dir[3];  /* orientation spin; Ising uses field_dir[2] sign */
    double field_mag;     /* raw magnitude as given to Field::new; magnitude() takes abs (state.rs:219-221) */
} vo_thermostat_t;

vo_thermostat_t vo_thermostat(double temperature, const double dir[3], double mag); /* thermostat.rs:29-60 */

/* Hamiltonian::energy(thermostat,state,index) for the configured compound. */
double vo_energy(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n, uint64_t i);
/* Hamiltonian::total_energy: n_terms==1 uses that term's override (energy.rs:114-120,153-160,208-213),
 * n_terms>1 uses the trait default sum_i energy(i) (energy.rs:55-59). */
double vo_total_energy(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n);
void vo_site_energies(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n, double* out);
/* e_new - e_old for the flip proposal (proposal==NULL) or the given per-site proposal
 * (Heisenberg: double[3] per site; Ising: int8 per site), exactly as integrator.rs:77-81 / :123-127. */
void vo_delta_energies(const vo_hamiltonian* h, const vo_thermostat_t* th, const void* state, uint64_t n,
                       const void* proposal, double* out);
/* State::magnetization (state.rs:291-296): raw projections and Field magnitude. */
double vo_magnetization(int model, const void* state, uint64_t n, double out_xyz[3]);

/* ---------------------------------------------------------------- spins / state */
void vo_marsaglia(vo_rng* r, double out[3]);                       /* util.rs:21-34 */
void vo_state_rand(int model, vo_rng* r, void* state, uint64_t n); /* state.rs:76-84,145-148,260-262 */

/* ---------------------------------------------------------------- integrators */
enum { VO_PROPOSE_FLIP = 0, VO_PROPOSE_RANDOM = 1 };
/* One Integrator::step (= n attempts) in place.  integrator.rs:66-92 (random) / :109-138 (flip).
 * Returns accepted count.  clone_state!=0 reproduces machine.rs:95's per-step state.clone(). */
uint64_t vo_metropolis_step(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, vo_rng* r,
                            void* state, uint64_t n);

/* ---------------------------------------------------------------- accumulator (accumulator.rs:23-64) */
typedef struct { double sum, sum_sq, sum_fourth; uint64_t count; } vo_acc;
void vo_acc_reset(vo_acc* a);
void vo_acc_collect(vo_acc* a, double v);
double vo_acc_mean(const vo_acc* a);
double vo_acc_variance(const vo_acc* a);
double vo_acc_binder(const vo_acc* a);

/* ---------------------------------------------------------------- machine + programs */
typedef struct {
    double temperature, field, mean_e, cv, mean_m, chi, binder;   /* instrument.rs:110-131 */
} vo_stat_row;

typedef struct {
    const vo_hamiltonian* h;
    vo_thermostat_t th;
    int proposal;
    vo_rng* rng;
    void* state;
    uint64_t n;
    int n_sensors;          /* 0: bare sweep; 1: StatSensor; 2: StatSensor+ObservableSensor (instrument.rs:133-141,254-262) */
    /* outputs */
    vo_stat_row* rows; uint64_t rows_cap, rows_len;
    double* obs_energy; double* obs_mag; uint64_t obs_cap, obs_len;  /* per-step series (relax+measure) */
    uint64_t attempts;
    void* scratch;          /* per-step clone target (machine.rs:95) */
} vo_machine;

int vo_machine_init(vo_machine* m, const vo_hamiltonian* h, int proposal, vo_rng* rng, void* state, uint64_t n, int n_sensors);
void vo_machine_free(vo_machine* m);
int vo_relax_for(vo_machine* m, uint64_t steps);    /* machine.rs:104-113 */
int vo_measure_for(vo_machine* m, uint64_t steps);  /* machine.rs:116-125 */

/* error codes mirror ProgramError (error.rs:31-46) */
enum { VO_OK = 0, VO_ERR_NO_STEPS = 1, VO_ERR_ZERO_TEMPERATURE = 2, VO_ERR_TMAX_LT_TMIN = 3,
       VO_ERR_ZERO_COOL_RATE = 4, VO_ERR_ZERO_FIELD = 5, VO_ERR_ZERO_FIELD_STEP = 6, VO_ERR_ALLOC = 7 };

int vo_program_relax(vo_machine* m, uint64_t steps, double temperature);                 /* program.rs:97-115 */
int vo_program_cooldown(vo_machine* m, double tmax, double tmin, double rate,
                        uint64_t relax, uint64_t steps);                                  /* program.rs:182-214 */
int vo_program_hysteresis(vo_machine* m, uint64_t steps, uint64_t relax, double temperature,
                          double max_field, double field_step);                           /* program.rs:281-336 */
/* schedule-only helpers (no dynamics): the (T) / (H) point lists the programs visit. */
uint64_t vo_cooldown_points(double tmax, double tmin, double rate, double* out, uint64_t cap);
uint64_t vo_hysteresis_points(double max_field, double field_step, double* out, uint64_t cap);
/* StatSensor line, "{:.16} x7" (instrument.rs:113-123). Returns chars written. */
int vo_stat_line(const vo_stat_row* row, char* buf, size_t cap);

/* ---------------------------------------------------------------- Philox4x32-R (Random123 KAT-checked for R = 7 and 10) */
#define VO_PHILOX_ROUNDS 7   /* rounds the GPU kernels draw with (PHILOX_ROUNDS, vegas_rs_b200/csrc/common.cuh): the replay uses the same */
void vo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
void vo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---------------------------------------------------------------- replay of one GPU-ordered sweep (vegas_replay.c) */
uint64_t vo_replay_ising_msc(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, uint64_t seed,
                             uint64_t sweep, uint64_t Lx, uint64_t Ly, uint64_t Lz, int8_t* state);
uint64_t vo_replay_ising_sites(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, uint64_t seed,
                               uint64_t sweep, uint64_t n, const uint8_t* colour, int n_colours, int8_t* state);
uint64_t vo_replay_heisenberg(const vo_hamiltonian* h, const vo_thermostat_t* th, int proposal, int f32, uint64_t seed,
                              uint64_t sweep, uint64_t n, const uint8_t* colour, int n_colours, double* state);

#ifdef __cplusplus
}
#endif
#endif
