/*
 * vegas_host.h -- C ABI of the host-side driver that sits above vegas_gpu.h: the reference's Machine,
 * Instrument hooks and Programs (src/machine.rs, src/instrument.rs, src/program.rs) restated in C++
 * (vegas_rs_b200/csrc/vegas_host.{hpp,cpp}) over a device-resident state.
 *
 * Differences from the reference that the GPU design forces:
 *   - the State lives on the device; instruments receive the per-step (E, |M|) pair that the sweep's fused
 *     reduction produced instead of recomputing Hamiltonian::total_energy(&state) on the host
 *     (src/instrument.rs:133-141, :254-262), and a host State only when a StateSensor dump is due;
 *   - sensors are callbacks, so any front end (Rust shim, Python TOML/parquet glue) can own the files.
 * Hook order, stage/step counters, accumulator formulas, validation errors and the StatSensor line are
 * the reference's.
 */
#ifndef VEGAS_HOST_H
#define VEGAS_HOST_H

#include "vegas_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vegas_machine* vegas_machine_t;

/* StatSensor::on_measure_end line, seven "{:.16}" fields (src/instrument.rs:110-131) */
typedef void (*vegas_stat_cb)(void* user, const char* line, double temperature, double field, double mean_e, double cv,
                              double mean_m, double chi, double binder);
/* ObservableSensor batch = one RecordBatch of src/output.rs:59-98: one row per step of the finished stage */
typedef void (*vegas_observable_cb)(void* user, int relax, uint64_t stage, uint64_t n, double temperature, double field,
                                    const double* energy, const double* magnetization, uint64_t len);
/* StateSensor dump = the rows of src/output.rs:149-186: host State in the reference layout
 * (Ising int8[n], Heisenberg double[3n]) */
typedef void (*vegas_state_cb)(void* user, int relax, uint64_t stage, uint64_t step, double temperature, double field,
                               const void* state, uint64_t n);

/* Slab group (multi-GPU, no reference counterpart): sums `len` doubles in place over the ranks of the group; 0 = ok */
typedef int (*vegas_reduce_cb)(void* user, double* values, uint64_t len);

/* Machine::new(Thermostat::new(2.8, Field::zero()), ..) as src/input.rs:273-279 builds it.  The machine
 * borrows the handle; the caller keeps ownership of it. */
int vegas_machine_create(vegas_gpu_t gpu, vegas_machine_t* out);
void vegas_machine_destroy(vegas_machine_t);
const char* vegas_machine_last_error(vegas_machine_t);
/* instruments are called in the order they were added (src/machine.rs:96-98) */
int vegas_machine_add_stat_sensor(vegas_machine_t, vegas_stat_cb, void* user);
int vegas_machine_add_observable_sensor(vegas_machine_t, vegas_observable_cb, void* user);
int vegas_machine_add_state_sensor(vegas_machine_t, uint64_t frequency, vegas_state_cb, void* user);
int vegas_machine_set_thermostat(vegas_machine_t, double temperature, const double field_dir[3], double field_mag);
int vegas_machine_thermostat(vegas_machine_t, double* temperature, double* field_mag);
int vegas_machine_relax_for(vegas_machine_t, uint64_t steps);    /* src/machine.rs:104-113 */
int vegas_machine_measure_for(vegas_machine_t, uint64_t steps);  /* src/machine.rs:116-125 */
uint64_t vegas_machine_steps_done(vegas_machine_t);
/* One Machine per rank, each over its own connected z-slab of ONE lattice (vegas_gpu_slab_connect): the per-step
 * energy and magnetisation projections of a batch -- sums over the rank's sites -- are added up over the group with
 * `reduce` (4 * len doubles at once) before any instrument sees them, and the instruments are told
 * State::len = n_sites_global.  Every rank runs the same program and replays the same hooks, so Relax / CoolDown /
 * HysteresisLoop and StatSensor / ObservableSensor behave as on one GPU (src/machine.rs:91-125); a front end normally
 * lets rank 0 own the output.  A StateSensor receives the rank's OWN slab (local sites).  reduce == NULL: back to one GPU. */
int vegas_machine_set_group(vegas_machine_t, vegas_reduce_cb reduce, void* user, uint64_t n_sites_global);

/* Programs (src/program.rs:97-115, :182-214, :281-336).  Return VEGAS_ERR_NO_STEPS ... as ProgramError. */
int vegas_program_relax(vegas_machine_t, uint64_t steps, double temperature);
int vegas_program_cooldown(vegas_machine_t, double max_temperature, double min_temperature, double cool_rate,
                           uint64_t relax, uint64_t steps);
int vegas_program_hysteresis(vegas_machine_t, uint64_t steps, uint64_t relax, double temperature, double max_field,
                             double field_step);

#ifdef __cplusplus
}
#endif
#endif
