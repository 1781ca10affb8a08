// vegas_host.cpp -- Machine / Programs over the device-resident sweep, and the C ABI of include/vegas_host.h.
#include "vegas_host.hpp"

#include <algorithm>
#include <cstring>

namespace vegas_host {

namespace {
constexpr uint64_t CHUNK = 4096;  // observables kept on the device per batch (vegas_gpu_step_async limit)
}

int Machine::set_thermostat(const Thermostat& th) {
    th_ = th;
    const int rc = vegas_gpu_set_thermostat(gpu_, th_.temperature, th_.field_dir, th_.field_mag);
    if (rc) return fail(rc, vegas_gpu_last_error(gpu_));
    return 0;
}

// The hooks of one finished batch, in the reference's order (src/machine.rs:96-98): per step every instrument's
// after_step with the recorded (E, |M|); the StateSensor dump of the batch's last step just before its after_step.
int Machine::replay_hooks(uint64_t chunk, const std::vector<double>& e, const std::vector<double>& m, bool heis, bool dump,
                          const std::vector<char>& state, uint64_t n_local) {
    for (uint64_t s = 0; s < chunk; ++s) {
        StepView v;
        v.energy = e[s];
        const double mx = m[3 * s], my = m[3 * s + 1], mz = m[3 * s + 2];
        if (heis) {  // HeisenbergSpin::from_projections src/state.rs:150-160
            const double mag = std::sqrt(mx * mx + my * my + mz * mz);
            v.magnetization = std::fabs(mag) < DBL_EPSILON ? 0.0 : std::fabs(mag);
        } else {     // IsingSpin::from_projections src/state.rs:86-92
            v.magnetization = std::fabs(mz);
        }
        for (auto& i : instruments_) {
            if (s + 1 == chunk && dump && i->next_state_dump(1) == 0) {
                const int rc = i->state_dump(state.data(), n_local);
                if (rc) return fail(rc, "state sensor failed");
            }
            const int rc = i->after_step(v);
            if (rc) return fail(rc, "instrument failed");
        }
    }
    return 0;
}

// Machine::run, src/machine.rs:91-101: steps x { state = integrator.step(..); every instrument.after_step(&state) }.
// The steps of a batch run back to back on the device with the observers fused into the sweep; the hooks
// are then replayed in the reference's order with the recorded per-step (E, |M|).  Without a StateSensor (whose step
// counter decides where a batch must end) the hooks of batch k are replayed on the host WHILE batch k + 1 sweeps.
// In a slab group the per-step partial sums are first summed over the ranks (every rank replays the same hooks).
int Machine::run(uint64_t steps) {
    const uint64_t n_local = vegas_gpu_n_sites(gpu_);
    // every Heisenberg family (heis_stencil, heis_general, heis_basis) reports |M| from three projections and dumps
    // 24-byte spins
    const bool heis = std::strncmp(vegas_gpu_kernel_family(gpu_), "heis", 4) == 0;
    bool overlap = true;
    for (auto& i : instruments_) overlap = overlap && !i->dumps_states();
    std::vector<double> e, m, pe, pm, buf;     // current batch; previous batch (hooks pending)
    std::vector<char> state;
    uint64_t remaining = steps, pending = 0;
    while (remaining > 0) {
        uint64_t chunk = std::min(remaining, CHUNK);
        bool record = false;
        for (auto& i : instruments_) {
            record = record || i->wants_observables();
            const int64_t k = i->next_state_dump(chunk);
            if (k >= 0) chunk = std::min<uint64_t>(chunk, (uint64_t)k + 1);
        }
        int rc = vegas_gpu_step_async(gpu_, chunk, record ? 1 : 0);
        if (rc) return fail(rc, vegas_gpu_last_error(gpu_));
        if (pending) {   // the device is busy with this batch: replay the previous one's hooks now
            rc = replay_hooks(pending, pe, pm, heis, false, state, n_local);
            if (rc) return rc;
            pending = 0;
        }
        e.assign(chunk, 0.0); m.assign(3 * chunk, 0.0);
        if (record) {
            rc = vegas_gpu_read_observables(gpu_, chunk, e.data(), m.data());
            if (rc) return fail(rc, vegas_gpu_last_error(gpu_));
            if (reduce_) {   // E and the three projections are sums over sites: add the slabs' partials
                buf.resize(4 * chunk);
                std::copy(e.begin(), e.end(), buf.begin());
                std::copy(m.begin(), m.end(), buf.begin() + chunk);
                if (reduce_(reduce_user_, buf.data(), 4 * chunk)) return fail(VEGAS_ERR_STATE, "slab group reduction failed");
                std::copy(buf.begin(), buf.begin() + chunk, e.begin());
                std::copy(buf.begin() + chunk, buf.end(), m.begin());
            }
        }
        bool dump = false;
        for (auto& i : instruments_) dump = dump || i->next_state_dump(chunk) == (int64_t)chunk - 1;
        if (dump) {  // the only place a host State exists (src/instrument.rs:340-350 needs it)
            state.resize(heis ? n_local * 24 : n_local);
            rc = heis ? vegas_gpu_download_heisenberg(gpu_, (double*)state.data(), n_local)
                      : vegas_gpu_download_ising(gpu_, (int8_t*)state.data(), n_local);
            if (rc) return fail(rc, vegas_gpu_last_error(gpu_));
        }
        remaining -= chunk;
        steps_done_ += chunk;
        if (overlap && remaining > 0) {
            pe.swap(e); pm.swap(m); pending = chunk;
        } else {
            rc = replay_hooks(chunk, e, m, heis, dump, state, n_local);
            if (rc) return rc;
        }
    }
    if (pending) { const int rc = replay_hooks(pending, pe, pm, heis, false, state, n_local); if (rc) return rc; }
    const int rc = vegas_gpu_synchronize(gpu_);
    return rc ? fail(rc, vegas_gpu_last_error(gpu_)) : 0;
}

int Machine::relax_for(uint64_t steps) {  // src/machine.rs:104-113
    for (auto& i : instruments_) { const int rc = i->on_relax_start(th_, n_sites()); if (rc) return rc; }
    const int rc = run(steps);
    if (rc) return rc;
    for (auto& i : instruments_) { const int r2 = i->on_relax_end(); if (r2) return r2; }
    return 0;
}

int Machine::measure_for(uint64_t steps) {  // src/machine.rs:116-125
    for (auto& i : instruments_) { const int rc = i->on_measure_start(th_, n_sites()); if (rc) return rc; }
    const int rc = run(steps);
    if (rc) return rc;
    for (auto& i : instruments_) { const int r2 = i->on_measure_end(); if (r2) return r2; }
    return 0;
}

int Relax::run(Machine& m) const {  // src/program.rs:97-115
    if (steps == 0) return m.fail(VEGAS_ERR_NO_STEPS, "no steps");
    if (temperature < DBL_EPSILON) return m.fail(VEGAS_ERR_ZERO_TEMPERATURE, "zero temperature");
    int rc = m.set_thermostat(m.thermostat().with_temperature(temperature));
    if (rc) return rc;
    return m.relax_for(steps);
}

int CoolDown::run(Machine& m) const {  // src/program.rs:182-214
    if (max_temperature < min_temperature) return m.fail(VEGAS_ERR_TMAX_LT_TMIN, "max temperature less than min temperature");
    if (steps == 0) return m.fail(VEGAS_ERR_NO_STEPS, "no steps");
    if (min_temperature < DBL_EPSILON) return m.fail(VEGAS_ERR_ZERO_TEMPERATURE, "zero temperature");
    if (cool_rate < DBL_EPSILON) return m.fail(VEGAS_ERR_ZERO_COOL_RATE, "zero cool rate");
    double temperature = max_temperature;
    for (;;) {
        int rc = m.set_thermostat(m.thermostat().with_temperature(temperature));
        if (rc) return rc;
        if ((rc = m.relax_for(relax))) return rc;
        if ((rc = m.measure_for(steps))) return rc;
        temperature -= cool_rate;
        if (temperature < min_temperature) break;
    }
    return 0;
}

int HysteresisLoop::run(Machine& m) const {  // src/program.rs:281-336
    if (steps == 0) return m.fail(VEGAS_ERR_NO_STEPS, "no steps");
    if (temperature < DBL_EPSILON) return m.fail(VEGAS_ERR_ZERO_TEMPERATURE, "zero temperature");
    if (max_field < DBL_EPSILON) return m.fail(VEGAS_ERR_ZERO_FIELD, "zero field");
    if (field_step < DBL_EPSILON) return m.fail(VEGAS_ERR_ZERO_FIELD_STEP, "zero field step");
    int rc = m.set_thermostat(m.thermostat().with_temperature(temperature));
    if (rc) return rc;
    const double up[3] = {0.0, 0.0, 1.0};  // Field::new(S::up(), magnitude)
    auto point = [&](double magnitude) {
        int r = m.set_thermostat(m.thermostat().with_field(up, magnitude));
        if (r) return r;
        if ((r = m.relax_for(relax))) return r;
        return m.measure_for(steps);
    };
    double magnitude = 0.0;
    for (;;) { if ((rc = point(magnitude))) return rc; magnitude += field_step; if (magnitude > max_field) break; }
    for (;;) { if ((rc = point(magnitude))) return rc; magnitude -= field_step; if (magnitude < -max_field) break; }
    for (;;) { if ((rc = point(magnitude))) return rc; magnitude += field_step; if (magnitude > max_field) break; }
    return 0;
}

}  // namespace vegas_host

// ========================================================================================= C ABI
struct vegas_machine {
    vegas_host::Machine m;
    explicit vegas_machine(vegas_gpu_t g) : m(g) {}
};

extern "C" {

int vegas_machine_create(vegas_gpu_t gpu, vegas_machine_t* out) {
    if (!gpu || !out) return VEGAS_ERR_INVALID;
    vegas_machine* mm = new vegas_machine(gpu);
    const int rc = mm->m.set_thermostat(vegas_host::Thermostat());  // Thermostat::new(2.8, Field::zero()), src/input.rs:274
    if (rc) { delete mm; return rc; }
    *out = mm;
    return VEGAS_OK;
}

void vegas_machine_destroy(vegas_machine_t m) { delete m; }
const char* vegas_machine_last_error(vegas_machine_t m) { return m ? m->m.error().c_str() : ""; }

int vegas_machine_add_stat_sensor(vegas_machine_t m, vegas_stat_cb cb, void* user) {
    if (!m) return VEGAS_ERR_INVALID;
    m->m.add(std::make_unique<vegas_host::StatSensor>(cb, user));
    return VEGAS_OK;
}
int vegas_machine_add_observable_sensor(vegas_machine_t m, vegas_observable_cb cb, void* user) {
    if (!m) return VEGAS_ERR_INVALID;
    m->m.add(std::make_unique<vegas_host::ObservableSensor>(cb, user));
    return VEGAS_OK;
}
int vegas_machine_add_state_sensor(vegas_machine_t m, uint64_t frequency, vegas_state_cb cb, void* user) {
    if (!m) return VEGAS_ERR_INVALID;
    m->m.add(std::make_unique<vegas_host::StateSensor>(frequency, cb, user));
    return VEGAS_OK;
}
int vegas_machine_set_thermostat(vegas_machine_t m, double temperature, const double field_dir[3], double field_mag) {
    if (!m) return VEGAS_ERR_INVALID;
    vegas_host::Thermostat th = m->m.thermostat().with_temperature(temperature);
    const double up[3] = {0.0, 0.0, 1.0};
    th = th.with_field(field_dir ? field_dir : up, field_mag);
    return m->m.set_thermostat(th);
}
int vegas_machine_thermostat(vegas_machine_t m, double* temperature, double* field_mag) {
    if (!m) return VEGAS_ERR_INVALID;
    if (temperature) *temperature = m->m.thermostat().temperature;
    if (field_mag) *field_mag = m->m.thermostat().field_mag;
    return VEGAS_OK;
}
int vegas_machine_relax_for(vegas_machine_t m, uint64_t steps) { return m ? m->m.relax_for(steps) : VEGAS_ERR_INVALID; }
int vegas_machine_measure_for(vegas_machine_t m, uint64_t steps) { return m ? m->m.measure_for(steps) : VEGAS_ERR_INVALID; }
uint64_t vegas_machine_steps_done(vegas_machine_t m) { return m ? m->m.steps_done() : 0; }
int vegas_machine_set_group(vegas_machine_t m, vegas_reduce_cb reduce, void* user, uint64_t n_sites_global) {
    if (!m || (reduce && n_sites_global == 0)) return VEGAS_ERR_INVALID;
    m->m.set_group(reduce, user, n_sites_global);
    return VEGAS_OK;
}

int vegas_program_relax(vegas_machine_t m, uint64_t steps, double temperature) {
    if (!m) return VEGAS_ERR_INVALID;
    vegas_host::Relax p; p.steps = steps; p.temperature = temperature;
    return p.run(m->m);
}
int vegas_program_cooldown(vegas_machine_t m, double tmax, double tmin, double rate, uint64_t relax, uint64_t steps) {
    if (!m) return VEGAS_ERR_INVALID;
    vegas_host::CoolDown p; p.max_temperature = tmax; p.min_temperature = tmin; p.cool_rate = rate; p.relax = relax; p.steps = steps;
    return p.run(m->m);
}
int vegas_program_hysteresis(vegas_machine_t m, uint64_t steps, uint64_t relax, double temperature, double max_field,
                             double field_step) {
    if (!m) return VEGAS_ERR_INVALID;
    vegas_host::HysteresisLoop p; p.steps = steps; p.relax = relax; p.temperature = temperature; p.max_field = max_field; p.field_step = field_step;
    return p.run(m->m);
}

}  // extern "C"
