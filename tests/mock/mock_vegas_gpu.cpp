// mock_vegas_gpu.cpp -- TEST DOUBLE of the device handle, for the CPU-only tests of the C++ host layer
// (vegas_rs_b200/csrc/vegas_host.cpp: Machine, sensors, programs).  It implements the nine vegas_gpu_* entry points the
// host layer calls with a SCRIPTED device: no spins, no Monte Carlo -- step k reports fixed formulas for (E, M) and a
// state pattern that encodes k, and every call is logged, so that tests/test_host_mock.py can check chunking, hook order,
// stage counters, dump scheduling, thermostat sequences and error propagation exactly.
// Test infrastructure only: built under tests/mock/, never linked into or imported by vegas_rs_b200.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vegas_gpu.h"

struct vegas_gpu {
    uint64_t n = 0;
    int heisenberg = 0;
    uint64_t step = 0;            // device steps done so far
    uint64_t pending_first = 0;   // first step of the last recorded batch
    uint64_t fail_at = ~0ull;     // the step_async call covering this step fails
    double T = 0.0, mag = 0.0;
    std::string err;
    std::vector<double> log;      // (kind, a, b, c) per call: 1 set_thermostat(T, mag, dir_z)  2 step_async(n, record, first)
                                  //                           3 download(step)              4 read_observables(n)
};

namespace {
void put(vegas_gpu* h, double k, double a, double b, double c) { h->log.insert(h->log.end(), {k, a, b, c}); }
}

extern "C" {

// ---- scripted observables (mirrored in tests/test_host_mock.py) ----
double mock_energy(uint64_t k, double T) { return -0.5 * (double)k + T; }
void mock_magnetization(uint64_t k, int heisenberg, double out[3]) {
    if (heisenberg) { const double a = (double)(k % 5); out[0] = 3.0 * a; out[1] = 4.0 * a; out[2] = 0.0; }
    else { out[0] = 0.0; out[1] = 0.0; out[2] = (double)((k * 37) % 101) - 50.0; }
}

vegas_gpu_t mock_gpu_create(uint64_t n, int heisenberg, uint64_t fail_at) {
    vegas_gpu* h = new vegas_gpu();
    h->n = n; h->heisenberg = heisenberg; h->fail_at = fail_at;
    return h;
}
void mock_gpu_destroy(vegas_gpu_t h) { delete h; }
uint64_t mock_gpu_log(vegas_gpu_t h, double* out, uint64_t capacity) {
    if (out) std::memcpy(out, h->log.data(), sizeof(double) * (size_t)std::min<uint64_t>(capacity, h->log.size()));
    return h->log.size();
}
uint64_t mock_gpu_steps(vegas_gpu_t h) { return h->step; }

// ---- the entry points vegas_host.cpp uses ----
const char* vegas_gpu_last_error(vegas_gpu_t h) { return h ? h->err.c_str() : ""; }
uint64_t vegas_gpu_n_sites(vegas_gpu_t h) { return h->n; }
const char* vegas_gpu_kernel_family(vegas_gpu_t h) { return h->heisenberg ? "heis_general" : "ising_general"; }

int vegas_gpu_set_thermostat(vegas_gpu_t h, double temperature, const double field_dir[3], double field_mag) {
    h->T = temperature; h->mag = field_mag;
    put(h, 1, temperature, field_mag, field_dir ? field_dir[2] : 0.0);
    return VEGAS_OK;
}

int vegas_gpu_step_async(vegas_gpu_t h, uint64_t n_steps, int record) {
    if (record && n_steps > 4096) { h->err = "step_async records at most 4096 steps per call"; return VEGAS_ERR_INVALID; }
    put(h, 2, (double)n_steps, (double)record, (double)h->step);
    if (h->fail_at >= h->step && h->fail_at < h->step + n_steps) { h->err = "scripted device failure"; return VEGAS_ERR_CUDA; }
    h->pending_first = h->step;
    h->step += n_steps;
    return VEGAS_OK;
}

int vegas_gpu_read_observables(vegas_gpu_t h, uint64_t n_steps, double* energy, double* mag_xyz) {
    put(h, 4, (double)n_steps, 0, 0);
    for (uint64_t s = 0; s < n_steps; ++s) {
        const uint64_t k = h->pending_first + s + 1;   // the state AFTER step s of the batch = device step count k
        if (energy) energy[s] = mock_energy(k, h->T);
        if (mag_xyz) mock_magnetization(k, h->heisenberg, mag_xyz + 3 * s);
    }
    return VEGAS_OK;
}

int vegas_gpu_download_ising(vegas_gpu_t h, int8_t* s, uint64_t n) {
    put(h, 3, (double)h->step, 0, 0);
    for (uint64_t i = 0; i < n; ++i) s[i] = ((h->step + i) & 1ull) ? 1 : -1;
    return VEGAS_OK;
}

int vegas_gpu_download_heisenberg(vegas_gpu_t h, double* sxyz, uint64_t n) {
    put(h, 3, (double)h->step, 0, 0);
    for (uint64_t i = 0; i < n; ++i) { sxyz[3 * i] = 0.0; sxyz[3 * i + 1] = 0.0; sxyz[3 * i + 2] = ((h->step + i) & 1ull) ? 1.0 : -1.0; }
    return VEGAS_OK;
}

int vegas_gpu_synchronize(vegas_gpu_t) { return VEGAS_OK; }

}  // extern "C"
