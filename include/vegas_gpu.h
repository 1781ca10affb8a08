/*
 * vegas_gpu.h -- C ABI of the B200-native Metropolis sweep for vegas-rs 0.9.0.
 *
 * This is the boundary a Rust `GpuMetropolis` shim (extern "C" + build.rs/cc, see
 * INTEGRATION.md) binds.  The reference has no FFI layer; its extension points are the
 * generic traits Integrator (src/integrator.rs:40-49), Hamiltonian (src/energy.rs:45-60),
 * Instrument (src/instrument.rs:19-58) and Program (src/program.rs:55-63).  Each entry
 * point below names the reference item it replaces.  All citations are relative to
 * /root/reference.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every call returns 0 on
 * success or a negative vegas_status_t and never throws or aborts (the reference bubbles
 * Result<> up to main, src/error.rs:13-88); vegas_gpu_last_error() gives the message.
 * A handle is bound to ONE CUDA device and is not thread-safe (the reference Machine is
 * single-threaded, src/machine.rs:44-55).  There is no CPU fallback: without a CUDA
 * device every create call fails with VEGAS_ERR_CUDA.
 */
#ifndef VEGAS_GPU_H
#define VEGAS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vegas_gpu* vegas_gpu_t;

typedef enum {
    VEGAS_OK = 0,
    VEGAS_ERR_INVALID = -1,     /* bad argument / unsupported combination */
    VEGAS_ERR_CUDA = -2,        /* CUDA runtime error (message in last_error) */
    VEGAS_ERR_ALLOC = -3,
    VEGAS_ERR_STATE = -4,       /* call not valid in the handle's current state */
    VEGAS_ERR_NO_STEPS = -10,   /* ProgramError::NoSteps             src/error.rs:31-46 */
    VEGAS_ERR_ZERO_TEMPERATURE = -11,
    VEGAS_ERR_TMAX_LT_TMIN = -12,
    VEGAS_ERR_ZERO_COOL_RATE = -13,
    VEGAS_ERR_ZERO_FIELD = -14,
    VEGAS_ERR_ZERO_FIELD_STEP = -15
} vegas_status_t;

typedef enum { VEGAS_ISING = 0, VEGAS_HEISENBERG = 1 } vegas_model_t;              /* src/input.rs:19-26 */
typedef enum { VEGAS_PROPOSE_FLIP = 0, VEGAS_PROPOSE_RANDOM = 1 } vegas_proposal_t; /* src/integrator.rs:125 vs :79 */
typedef enum { VEGAS_F32 = 0, VEGAS_F64 = 1 } vegas_precision_t;                   /* Heisenberg device storage */
typedef enum { VEGAS_SC = 0, VEGAS_BCC = 1, VEGAS_FCC = 2 } vegas_unitcell_t;      /* src/input.rs:38-48 */

/* Which total_energy the reference would report for this Hamiltonian:
 *  PHYSICAL            every bond once, -|H| sum s.o, k sum (s.n)^2, g N
 *  REFERENCE_COMPOUND  trait default sum_i energy(i) of hamiltonian!(...) (src/energy.rs:55-59,254-256):
 *                      exchange double counted, Zeeman with '+', anisotropy with k
 *  REFERENCE_EXCHANGE  Exchange::total_energy alone (src/energy.rs:208-213), as `vegas bench` uses */
typedef enum { VEGAS_E_PHYSICAL = 0, VEGAS_E_REFERENCE_COMPOUND = 1, VEGAS_E_REFERENCE_EXCHANGE = 2 } vegas_energy_conv_t;

/* Structured lattice = unit cell x expansion, src/input.rs:296-322.  Site index is
 * ((iz*ny+iy)*nx+ix)*n_basis+b.  The z range may be a slab of a larger lattice (multi-GPU):
 * this handle then owns planes/cells [z_offset, z_offset+nz) of nz_global. */
typedef struct {
    int unitcell;                 /* vegas_unitcell_t */
    uint64_t nx, ny, nz;          /* local extent in unit cells */
    int pbc_x, pbc_y, pbc_z;      /* src/input.rs:79-96 */
    int literal_from_lattice_filter; /* 1: apply `source <= target` (src/energy.rs:180) to the generated edge list */
    uint64_t nz_global;           /* 0 or ==nz: not decomposed */
    uint64_t z_offset;
} vegas_lattice_desc;

/* General adjacency = Exchange::new(CsMat<f64>), src/energy.rs:171-173.  Columns sorted
 * ascending per row (sprs CSR); values==NULL means the uniform `exchange` of the model. */
typedef struct {
    uint64_t n;
    const uint64_t* row_ptr;      /* n+1 */
    const uint32_t* col_idx;
    const double* values;
} vegas_csr_desc;

typedef struct {
    int model;                    /* vegas_model_t */
    int proposal;                 /* vegas_proposal_t */
    int precision;                /* vegas_precision_t (Heisenberg only) */
    double exchange;              /* uniform J, src/input.rs:163 (default 1.0, :352) */
    int has_exchange;             /* Exchange term present */
    int has_zeeman;               /* Zeeman term present, src/input.rs:271 */
    int has_anisotropy; double anisotropy_k; double anisotropy_axis[3]; /* src/energy.rs:96-101 */
    int has_gauge; double gauge;  /* src/energy.rs:70-72 */
    uint64_t seed;                /* Philox key; replaces --seed, src/main.rs:27-30 */
    int device;                   /* CUDA device ordinal */
    int force_general;            /* 1: never use the structured-stencil kernels (testing) */
} vegas_model_desc;

/* ---- lifetime ------------------------------------------------------------------------- */
int vegas_gpu_create_lattice(const vegas_model_desc*, const vegas_lattice_desc*, vegas_gpu_t* out);
int vegas_gpu_create_csr(const vegas_model_desc*, const vegas_csr_desc*, vegas_gpu_t* out);
void vegas_gpu_destroy(vegas_gpu_t);
const char* vegas_gpu_last_error(vegas_gpu_t);      /* NULL handle: error of the last failed create */
const char* vegas_gpu_version(void);

/* ---- introspection -------------------------------------------------------------------- */
uint64_t vegas_gpu_n_sites(vegas_gpu_t);            /* State::len, src/state.rs:279-281 (local sites) */
int vegas_gpu_n_colours(vegas_gpu_t);
/* kernel family in use: "ising_msc", "heis_stencil", "heis_basis" (periodic bcc / fcc), "ising_general", "heis_general" */
const char* vegas_gpu_kernel_family(vegas_gpu_t);
/* the host-side adjacency this handle was built with, in the reference's CSR form (tests/oracle parity).
 * Pass NULL arrays to query sizes. Not available (VEGAS_ERR_STATE) for stencil handles above 2^27 sites. */
int vegas_gpu_adjacency(vegas_gpu_t, uint64_t* n, uint64_t* nnz, uint64_t* row_ptr, uint32_t* col_idx, double* values);
int vegas_gpu_colours(vegas_gpu_t, uint8_t* colour_of_site);

/* pure host-side helpers (no CUDA device needed): the adjacency Exchange::from_lattice would build for a
 * lattice descriptor (src/energy.rs:176-187) and the colouring the general-adjacency sweep uses for it.
 * Pass NULL arrays to query sizes. */
int vegas_gpu_lattice_adjacency(const vegas_lattice_desc*, double exchange, uint64_t* n, uint64_t* nnz,
                                uint64_t* row_ptr, uint32_t* col_idx, double* values);
int vegas_gpu_lattice_colours(const vegas_lattice_desc*, int* n_colours, uint8_t* colour_of_site);
/* host-only self check: the compile-time bcc / fcc neighbour tables of the heis_basis kernel equal the unit-cell edge
 * list the adjacency export above is built from (0 = identical) */
int vegas_gpu_check_basis_tables(void);

/* ---- state I/O in the REFERENCE's host layouts (src/state.rs:60-63,133-134,245-246) ---- */
int vegas_gpu_upload_ising(vegas_gpu_t, const int8_t* s, uint64_t n);          /* +1 Up / -1 Down per site */
int vegas_gpu_upload_heisenberg(vegas_gpu_t, const double* sxyz, uint64_t n);  /* AoS [f64;3] per site */
int vegas_gpu_download_ising(vegas_gpu_t, int8_t* s, uint64_t n);
int vegas_gpu_download_heisenberg(vegas_gpu_t, double* sxyz, uint64_t n);
int vegas_gpu_randomize(vegas_gpu_t);               /* State::rand_with_size on device, src/state.rs:260-262 */
int vegas_gpu_fill(vegas_gpu_t, int up);            /* State::{up,down}_with_size, src/state.rs:250-258 */

/* ---- thermostat, src/thermostat.rs:19-79: T clamped to >= DBL_EPSILON; the field is an
 * orientation spin (Ising: sign of dir[2]) and a magnitude of which |.| is used (src/state.rs:219-221) */
int vegas_gpu_set_thermostat(vegas_gpu_t, double temperature, const double field_dir[3], double field_mag);
int vegas_gpu_set_energy_convention(vegas_gpu_t, int conv /* vegas_energy_conv_t */);

/* ---- the hot path: Integrator::step x n_steps (src/integrator.rs:66-92,109-138) with the
 * per-step observers of src/instrument.rs:133-141 fused in.  1 step = N attempts.  When
 * energy / mag_xyz are non-NULL they receive, per step, Hamiltonian::total_energy in the
 * selected convention and the raw magnetisation projections (sum sx, sum sy, sum sz;
 * |M| = Field magnitude is their norm, src/state.rs:235-242).  Host pointers. */
int vegas_gpu_step(vegas_gpu_t, uint64_t n_steps, double* energy, double* mag_xyz);
/* Same, asynchronous: nothing is copied back; observables of the last `n` steps stay on the device
 * until vegas_gpu_read_observables.  Used by the multi-GPU driver and the benchmark. */
int vegas_gpu_step_async(vegas_gpu_t, uint64_t n_steps, int record_observables);
int vegas_gpu_read_observables(vegas_gpu_t, uint64_t n_steps, double* energy, double* mag_xyz);
int vegas_gpu_synchronize(vegas_gpu_t);
/* Literal drop-in of Integrator::step's signature: host State in, host State out (one step),
 * host<->device copies included.  Ising: int8 per site; Heisenberg: double[3] per site. */
int vegas_gpu_step_host_ising(vegas_gpu_t, int8_t* state_inout, uint64_t n, double* energy, double* mag_xyz);
int vegas_gpu_step_host_heisenberg(vegas_gpu_t, double* sxyz_inout, uint64_t n, double* energy, double* mag_xyz);

/* ---- deterministic parity entry points ----------------------------------------------- */
int vegas_gpu_total_energy(vegas_gpu_t, double* out);                    /* Hamiltonian::total_energy, selected convention */
int vegas_gpu_magnetization(vegas_gpu_t, double out_xyz[3]);             /* State::magnetization projections */
int vegas_gpu_site_energies(vegas_gpu_t, double* out_n);                 /* Hamiltonian::energy(i) for all i (compound) */
/* e_new - e_old of src/integrator.rs:77-81/:123-127 for every site: flip proposal when
 * proposal==NULL, else the given spins (Ising int8[n], Heisenberg double[3n]). */
int vegas_gpu_delta_energies(vegas_gpu_t, const void* proposal, double* out_n);
int vegas_gpu_attempt_count(vegas_gpu_t, uint64_t* attempts, uint64_t* accepted);
int vegas_gpu_sweep_count(vegas_gpu_t, uint64_t* sweeps);                /* Philox sweep counter */
int vegas_gpu_set_sweep_count(vegas_gpu_t, uint64_t sweeps);
/* integer acceptance thresholds in use (Ising, uniform J): thr[2][n_classes], always[2][n_classes] */
int vegas_gpu_ising_thresholds(vegas_gpu_t, int* n_classes, uint64_t* thr, uint8_t* always);

/* ---- z-slab decomposition over one process per GPU (no reference counterpart) ---------
 * Each rank exports CUDA IPC handles of its halo buffers and flags; the launcher (torch.distributed,
 * any transport) delivers them to the z-neighbours, which connect.  After connect the sweep kernels
 * store boundary planes straight into the neighbour's halo over NVLink and signal with flags. */
#define VEGAS_IPC_BYTES 256
int vegas_gpu_slab_export(vegas_gpu_t, void* blob /* VEGAS_IPC_BYTES */);
int vegas_gpu_slab_connect(vegas_gpu_t, const void* blob_lower_neighbour, const void* blob_upper_neighbour);
/* single-process variant: both handles live in this process (tests, or one process driving several GPUs) */
int vegas_gpu_slab_connect_local(vegas_gpu_t self, vegas_gpu_t lower, vegas_gpu_t upper);

/* ---- kernel selection knobs (no reference counterpart; tests and tuning) ----------------
 * key "heis_fused"    : -1 auto (default), 0 never, 1 whenever the lattice fits  -- the one-launch-per-step
 *                       two-colour Heisenberg kernel (heis_fused.cuh) instead of two colour passes
 *     "heis_fused_ty" : interior rows per CTA tile (0 = auto), "heis_fused_cz": planes per z-chunk (0 = auto)
 *     "heis_pipe"     : -1 auto (default: 3-D lattices with >= 32 planes whose rows fit), 0 never, 1 whenever the lattice
 *                       fits -- both colour passes of a Heisenberg step as ONE cooperative, phase-pipelined launch whose
 *                       plane tiles are staged by TMA into mbarrier-guarded shared-memory rings (heis_pipe.cu);
 *                       "heis_pipe_stages" / "heis_pipe_own" (ring depths of the other / own colour, 0 = auto),
 *                       "heis_pipe_tiles" (bands of rows per colour, 0 = auto: half the SM count), "heis_pipe_vec" (sites per
 *                       consumer thread: a whole or half 16-byte vector, 0 = auto), "heis_pipe_lead" (planes the first
 *                       colour may run ahead of the second, 0 = auto; bounds the working set kept in L2), "heis_pipe_pub"
 *                       (planes per published progress update = per gpu-scope release, 0 = auto), "heis_pipe_backoff" /
 *                       "heis_pipe_backoff_helper" (ns a consumer / helper warp sleeps between failed mbarrier polls),
 *                       "heis_pipe_l2" (1, default: L2 eviction-priority hints on the TMA loads / stores -- what the next colour is about
 *                       to read is kept, what was used for the last time in this step goes first; 0: none)
 *     "heis_wave"     : -1 auto (default: lattices with >= 32 planes), 0 never, 1 always -- both colour passes of a
 *                       Heisenberg step as ONE persistent launch in wave order (second pass finds the first in L2);
 *                       "heis_wave_planes" (planes per chunk, default 4), "heis_wave_lag" (positions a colour pass
 *                       trails the previous one by, >= 3, default 5), "heis_wave_steps" (steps fused into one launch,
 *                       1..4, default 1: more was measured slower)
 *     "heis_wave_c"   : experiment: the two passes as separate launches interleaved in chunks of C planes
 *     "basis_pipe"    : 1 whenever the lattice fits (single handle); default -1 / 0: one launch per colour -- all 2 / 4
 *                       colour passes of a periodic bcc / fcc Heisenberg step as ONE cooperative, phase-pipelined launch
 *                       (basis_pipe.cu): colour b trails colour b-1 by a few planes so that the partner sublattices are read
 *                       from L2; "basis_pipe_lead" (planes the first colour may lead the last, 0 = auto), "basis_pipe_pub"
 *                       (planes per published progress update, 0 = auto: 1), "basis_pipe_tiles" (bands per colour, 0 = auto).
 *                       Opt-in: compulsory DRAM traffic only, but latency bound and slower than the colour launches so far
 *     "basis_pair"    : -1 auto (default: periodic fcc lattices, single handle or connected z-slab, whose State exceeds L2), 0 never (one launch per
 *                       colour), 1 whenever possible -- the fcc step as TWO launches, colours (0, 1) and (2, 3), each CTA updating
 *                       the first colour on its rows plus one (recomputed) row and then the second colour, from one set of arrays
 *                       to a second one (swapped after every step; twice the State in HBM), no inter-CTA synchronisation
 *                       (heis_basis_pair_kernel); "basis_pair_rows" = rows of a plane per CTA (0 = auto: 48), "basis_pair_chunk" =
 *                       rows after which the CTA switches between its two colours (0 = auto: the fewest that keep every thread busy).  9 % faster than four launches
 *                       on fcc 384^3 (three fat CTAs per SM keep the window between the two colours inside L2); connected fcc
 *                       z-slabs hold both array sets in their one allocation and swap in lock step
 *     "basis_wave"    : 1 whenever the lattice has enough planes; default -1 / 0: one launch per colour -- all 2 / 4 colour
 *                       passes of a periodic bcc / fcc Heisenberg step as ONE persistent cooperative launch whose work items
 *                       (those of the colour launches) are drawn in wave order, colour b a few planes behind colour b-1, so
 *                       that a pass finds its partner sublattices in L2 (basis_wave.cu; also on connected z-slabs with
 *                       neighbours on other devices); "basis_wave_lag" (time slots between consecutive colours, 0 = auto: 2),
 *                       "basis_wave_ipt" (16-byte work items per thread and tile, 0 = auto), "basis_wave_grid" (cap on the
 *                       number of CTAs, 0 = as many as are co-resident).  Opt-in: a third less DRAM traffic, but slower than
 *                       the colour launches so far (per-item synchronisation, no L1 reuse between rows)
 *     "msc_full"      : 1 (default) the Ising colour pass without per-row bounds and predicates whenever the launch grid covers
 *                       the lattice exactly (uniform J > 0, no field, not a slab boundary plane), 0 always the generic variant
 *     "host_pack_min" : fewest spins of an ising_msc lattice for which upload / download / step_host move a sign BITMAP over
 *                       PCIe (host threads convert the int8 State chunk by chunk, copies overlap; VEGAS_HOST_THREADS, default
 *                       min(16, cores)) instead of one byte per spin; default 2^22, 0 = always, -1 = never;
 *                       "host_pack_chunk" = spins per pipelined chunk (default 2^26)
 *     "basis_vec"     : 1 (default) 16-byte accesses in the bcc / fcc colour pass when nx % 4 == 0 (fp64: % 2), 0 scalar
 *     "resident_max"  : largest site count of a general-family lattice that runs batches of steps in ONE launch with
 *                       the State in shared memory (default 8192; 0 = always one launch per colour)
 * Results do not depend on these knobs (same Philox keys, same arithmetic). */
int vegas_gpu_set_tuning(vegas_gpu_t, const char* key, long value);
/* host-only: the unit order of a persistent wave launch of `steps` steps over `n_chunks` chunks with lag `lag`
 * (units[i] = phase << 24 | chunk, phase = 2 * step + colour; 2 * steps * n_chunks entries).  Exposed so that the
 * schedule's invariant -- every unit comes after the three units it waits for -- is tested without a GPU. */
int vegas_gpu_wave_schedule(uint32_t n_chunks, uint32_t lag, uint32_t steps, uint32_t* units, uint64_t capacity, uint64_t* count);
/* host-only: 1 when the unit-cell bond table has the structure the pair launches (heis_basis_pair_kernel) rely on -- every bond
 * between colours 2k and 2k + 1 has dz = 0 and, seen from 2k + 1, dy in {0, +1} -- else 0 (fcc: 1, bcc: 0) */
int vegas_gpu_basis_pair_structure(int unitcell);
/* host-only: the unit order of one wave-ordered bcc / fcc step (basis_wave.cu) over `nz` cell planes with `lag` time slots
 * between consecutive colours: units[i] = colour << 24 | plane, n_basis * nz entries (count = 0: too few planes for the
 * scheme); need[b] bit 2a + r = colour b on plane z waits for colour a on plane (z + r) % nz.  unitcell: VEGAS_BCC / VEGAS_FCC. */
int vegas_gpu_basis_wave_schedule(int unitcell, uint32_t nz, uint32_t lag, uint32_t* units, uint64_t capacity, uint64_t* count, uint32_t need[4]);
/* bytes one upload (or download) of this handle's State moves over PCIe: 24 n (Heisenberg, f64 AoS), n (Ising, one byte per
 * spin) or n / 8 (Ising on the host-packed path, see "host_pack_min") */
uint64_t vegas_gpu_state_transfer_bytes(vegas_gpu_t);
/* host-only: the State <-> sign-bitmap conversion of the host-packed transfer (words[i] bit b = s[32 i + b] > 0), run on
 * `threads` workers in chunks of `chunk_words`, exposed so that it is tested without a GPU */
int vegas_gpu_host_pack(const int8_t* s, uint32_t* words, uint64_t n_words, int threads, uint64_t chunk_words);
int vegas_gpu_host_unpack(const uint32_t* words, int8_t* s, uint64_t n_words, int threads, uint64_t chunk_words);
/* name of the kernel the NEXT step will launch: "heis_fused", "heis_stencil", "ising_msc", ... */
const char* vegas_gpu_step_kernel(vegas_gpu_t);

/* ---- timing hooks for bench.py (CUDA events on the handle's own stream) --------------- */
int vegas_gpu_timer_start(vegas_gpu_t);
int vegas_gpu_timer_stop(vegas_gpu_t, float* elapsed_ms);
uint64_t vegas_gpu_launch_count(vegas_gpu_t);       /* kernels launched by this handle so far */
void* vegas_gpu_stream(vegas_gpu_t);                /* cudaStream_t */

#ifdef __cplusplus
}
#endif
#endif /* VEGAS_GPU_H */
