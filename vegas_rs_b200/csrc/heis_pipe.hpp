// heis_pipe.hpp -- host interface of K3p, the phase-pipelined TMA Heisenberg step for sc lattices (heis_pipe.cu).
//
// Replaces MetropolisIntegrator::step / MetropolisFlipIntegrator::step (src/integrator.rs:66-92, :109-138) for
// HeisenbergSpin on the colour-split SoA layout of heis.cuh.  Same Philox keys, same arithmetic and summation order as
// heis_stencil_kernel / heis_wave_kernel: the trajectories are bit-identical.
#pragma once
#include <string>

#include "heis.cuh"

namespace vg {

struct HeisPipeDesc {
    int device = 0;
    bool f64 = false;
    uint32_t Lx = 0, Ly = 0, Lz = 0, z_offset = 0;
    void* arr[2][3] = {};   // [colour][component]: Lz * Ly * (Lx / 2) elements each
    // connected z-slab: halo planes of the neighbours' boundary planes ([colour][lo 0 / hi 1][component], Ly * Lx/2
    // elements each) and where my own boundary planes go (the neighbours' halos, peer-mapped); null when single handle
    void* halo[2][2][3] = {};
    void* peer[2][2][3] = {};            // [colour][to lower 0 / to upper 1][component]
    unsigned long long* flags = nullptr;       // [2]: written by the lower / upper neighbour: boundary planes it has stored so far
    unsigned long long* peer_flags[2] = {};    // the neighbours' flag words I add to (lower neighbour: its [1], upper: its [0])
    bool slab = false;
    // tuning (0 = automatic)
    uint32_t stages_other = 0, stages_own = 0, tiles = 0, vec = 0;   // vec: sites per consumer thread
    uint32_t lead = 0;      // planes the first colour may run ahead of the last (>= 2 pub_every + 2; bounds the L2 working set)
    uint32_t pub_every = 0; // planes per published progress update (one gpu-scope release fence each)
    uint32_t l2_hints = 1;  // 1: L2 eviction-priority hints on the TMA loads / stores (keep what the next phase reads, drop the rest first)
    uint32_t backoff_consumer = 0, backoff_helper = 0;   // ns slept between failed mbarrier polls of the consumer / helper warps
};

struct HeisPipeState;

// nullptr (and a reason) when the lattice does not fit this kernel; the caller then uses the other Heisenberg kernels
HeisPipeState* heis_pipe_create(const HeisPipeDesc& d, std::string& why_not);
void heis_pipe_destroy(HeisPipeState*);
const char* heis_pipe_describe(const HeisPipeState*);
uint32_t heis_pipe_tiles(const HeisPipeState*);   // bands per colour (= CTAs that signal a boundary plane)

// one Monte Carlo step (both colour passes) on stream `st`; obs_row as heis_stencil_kernel (6 doubles, added to).
// slab_steps: pipelined steps this slab and its neighbours have done since they were connected (all ranks step together)
template <typename real>
int heis_pipe_step(HeisPipeState*, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                   double* obs_row, uint64_t slab_steps, cudaStream_t st, std::string& err);
// after the stream has been synchronised: != 0 (and a message) when a wait inside the kernel timed out
int heis_pipe_check(HeisPipeState*, std::string& err);

}  // namespace vg
