// vegas_host.hpp -- host-side mirror of the reference's Machine / Instrument / Program layer, in C++17,
// written above the C ABI of include/vegas_gpu.h (never below it: it only calls vegas_gpu_* functions).
//
// Reference interfaces restated (all under /root/reference):
//   Thermostat        src/thermostat.rs:19-79      Accumulator   src/accumulator.rs:23-64
//   Instrument hooks  src/instrument.rs:19-58      StatSensor    src/instrument.rs:61-142
//   ObservableSensor  src/instrument.rs:145-263    StateSensor   src/instrument.rs:265-351
//   Machine           src/machine.rs:44-125        Programs      src/program.rs:66-336
// The integrator + hamiltonian pair is the GPU handle (GpuMetropolis): like WolffIntegrator
// (src/integrator.rs:146-186) it carries its own model description.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vegas_host.h"

namespace vegas_host {

// src/accumulator.rs:23-64
struct Accumulator {
    double sum = 0.0, sum_sq = 0.0, sum_fourth = 0.0;
    uint64_t count = 0;
    void collect(double v) { sum += v; sum_sq += v * v; sum_fourth += v * v * v * v; count += 1; }
    double mean() const { return sum / (double)count; }
    double variance() const { const double m = mean(); return sum_sq / (double)count - m * m; }
    double binder_cumulant() const {
        const double m2 = sum_sq / (double)count;
        return 1.0 - (sum_fourth / (double)count) / (3.0 * (m2 * m2));
    }
};

// src/thermostat.rs:19-79 with Field (src/state.rs:195-242): magnitude() is |magnitude|
struct Thermostat {
    double temperature = 2.8;
    double field_dir[3] = {0.0, 0.0, 1.0};
    double field_mag = 0.0;
    static double clamp(double t) { return t < DBL_EPSILON ? DBL_EPSILON : t; }
    Thermostat with_temperature(double t) const { Thermostat r = *this; r.temperature = clamp(t); return r; }
    Thermostat with_field(const double dir[3], double mag) const {
        Thermostat r = *this;
        for (int i = 0; i < 3; ++i) r.field_dir[i] = dir[i];
        r.field_mag = mag;
        return r;
    }
    double field_magnitude() const { return std::fabs(field_mag); }
};

// What an instrument sees after a step: the device-reduced observables of that step.
struct StepView {
    double energy;         // Hamiltonian::total_energy in the handle's convention
    double magnetization;  // State::magnetization().magnitude()
};

// src/instrument.rs:19-58
class Instrument {
  public:
    virtual ~Instrument() = default;
    virtual int on_relax_start(const Thermostat&, uint64_t /*n*/) { return 0; }
    virtual int on_relax_end() { return 0; }
    virtual int on_measure_start(const Thermostat&, uint64_t /*n*/) { return 0; }
    virtual int on_measure_end() { return 0; }
    virtual int after_step(const StepView&) { return 0; }
    virtual bool wants_observables() const { return false; }  // needs (E, M) of the coming steps
    virtual int64_t next_state_dump(uint64_t /*steps_ahead*/) const { return -1; }  // steps until a host State is needed
    virtual bool dumps_states() const { return false; }   // true: batch sizes depend on this instrument's step counter
    virtual int state_dump(const void* /*state*/, uint64_t /*n*/) { return 0; }
};

// src/instrument.rs:61-142
class StatSensor : public Instrument {
    vegas_stat_cb cb_; void* user_;
    Accumulator e_, m_;
    bool active_ = false; Thermostat th_; uint64_t n_ = 0;
  public:
    StatSensor(vegas_stat_cb cb, void* user) : cb_(cb), user_(user) {}
    int on_measure_start(const Thermostat& th, uint64_t n) override { th_ = th; n_ = n; active_ = true; return 0; }
    int on_measure_end() override {
        if (active_) {
            const double T = th_.temperature, nn = (double)n_;
            const double row[7] = {T, th_.field_magnitude(), e_.mean(), e_.variance() / (nn * (T * T)), m_.mean(),
                                   m_.variance() / (nn * T), m_.binder_cumulant()};
            char line[512];
            std::snprintf(line, sizeof line, "%.16f %.16f %.16f %.16f %.16f %.16f %.16f", row[0], row[1], row[2], row[3],
                          row[4], row[5], row[6]);
            if (cb_) cb_(user_, line, row[0], row[1], row[2], row[3], row[4], row[5], row[6]);
        }
        active_ = false; e_ = Accumulator(); m_ = Accumulator();
        return 0;
    }
    int after_step(const StepView& v) override {
        if (active_) { e_.collect(v.energy); m_.collect(v.magnetization); }
        return 0;
    }
    bool wants_observables() const override { return active_; }
};

// src/instrument.rs:145-263
class ObservableSensor : public Instrument {
    vegas_observable_cb cb_; void* user_;
    uint64_t stage_ = 0, n_ = 0; bool active_ = false, relax_ = false; Thermostat th_;
    std::vector<double> e_, m_;
    void start(const Thermostat& th, uint64_t n, bool relax) { th_ = th; n_ = n; relax_ = relax; active_ = true; e_.clear(); m_.clear(); }
    int end() {
        if (active_ && cb_) cb_(user_, relax_ ? 1 : 0, stage_, n_, th_.temperature, th_.field_magnitude(), e_.data(), m_.data(), e_.size());
        stage_ += 1; active_ = false; e_.clear(); m_.clear();
        return 0;
    }
  public:
    ObservableSensor(vegas_observable_cb cb, void* user) : cb_(cb), user_(user) {}
    int on_relax_start(const Thermostat& th, uint64_t n) override { start(th, n, true); return 0; }
    int on_relax_end() override { return end(); }
    int on_measure_start(const Thermostat& th, uint64_t n) override { start(th, n, false); return 0; }
    int on_measure_end() override { return end(); }
    int after_step(const StepView& v) override {
        if (active_) { e_.push_back(v.energy); m_.push_back(v.magnetization); }
        return 0;
    }
    bool wants_observables() const override { return active_; }
};

// src/instrument.rs:265-351: dump when step.is_multiple_of(frequency) (frequency 0: only step 0)
class StateSensor : public Instrument {
    vegas_state_cb cb_; void* user_; uint64_t frequency_;
    uint64_t step_ = 0, stage_ = 0; int relax_ = -1; Thermostat th_;
    bool due(uint64_t step) const { return frequency_ == 0 ? step == 0 : step % frequency_ == 0; }
  public:
    StateSensor(uint64_t frequency, vegas_state_cb cb, void* user) : cb_(cb), user_(user), frequency_(frequency) {}
    int on_relax_start(const Thermostat& th, uint64_t) override { relax_ = 1; th_ = th; return 0; }
    int on_relax_end() override { relax_ = -1; step_ = 0; stage_ += 1; return 0; }
    int on_measure_start(const Thermostat& th, uint64_t) override { relax_ = 0; th_ = th; return 0; }
    int on_measure_end() override { relax_ = -1; step_ = 0; stage_ += 1; return 0; }
    int after_step(const StepView&) override { step_ += 1; return 0; }
    // the dump of step index `step_` happens inside after_step in the reference, i.e. with the state AFTER that
    // step: the machine asks how many steps it may run before the state of the last of them is needed.
    int64_t next_state_dump(uint64_t steps_ahead) const override {
        if (relax_ < 0) return -1;
        for (uint64_t k = 0; k < steps_ahead; ++k)
            if (due(step_ + k)) return (int64_t)k;
        return -1;
    }
    int state_dump(const void* state, uint64_t n) override {
        if (cb_ && relax_ >= 0) cb_(user_, relax_, stage_, step_, th_.temperature, th_.field_magnitude(), state, n);
        return 0;
    }
    bool dumps_states() const override { return true; }
};

// src/machine.rs:44-125
class Machine {
    vegas_gpu_t gpu_;
    Thermostat th_;
    std::vector<std::unique_ptr<Instrument>> instruments_;
    uint64_t steps_done_ = 0;
    std::string err_;
    // slab group (one Machine per rank, each over its own z-slab): sums the per-step partials over the ranks
    vegas_reduce_cb reduce_ = nullptr; void* reduce_user_ = nullptr; uint64_t n_global_ = 0;
    int run(uint64_t steps);
    int replay_hooks(uint64_t chunk, const std::vector<double>& e, const std::vector<double>& m, bool heis, bool dump,
                     const std::vector<char>& state, uint64_t n_local);
  public:
    explicit Machine(vegas_gpu_t gpu) : gpu_(gpu) {}
    void set_group(vegas_reduce_cb reduce, void* user, uint64_t n_sites_global) { reduce_ = reduce; reduce_user_ = user; n_global_ = n_sites_global; }
    const Thermostat& thermostat() const { return th_; }
    int set_thermostat(const Thermostat& th);
    void add(std::unique_ptr<Instrument> i) { instruments_.push_back(std::move(i)); }
    int relax_for(uint64_t steps);
    int measure_for(uint64_t steps);
    uint64_t steps_done() const { return steps_done_; }
    uint64_t n_sites() const { return reduce_ ? n_global_ : vegas_gpu_n_sites(gpu_); }   // State::len as the instruments see it
    const std::string& error() const { return err_; }
    int fail(int code, const std::string& msg) { err_ = msg; return code; }
};

// src/program.rs:66-336
struct Relax { uint64_t steps = 1000; double temperature = 3.0; int run(Machine&) const; };
struct CoolDown {
    double max_temperature = 3.0, min_temperature = 0.05, cool_rate = 0.1; uint64_t relax = 1000, steps = 20000;
    int run(Machine&) const;
};
struct HysteresisLoop {
    uint64_t steps = 1000, relax = 1000; double temperature = 3.0, max_field = 1.0, field_step = 0.1;
    int run(Machine&) const;
};

}  // namespace vegas_host
