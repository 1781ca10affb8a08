// basis_wave.hpp -- host interface of K4w, the wave-ordered persistent step for periodic bcc / fcc Heisenberg lattices
// (basis_wave.cu).
//
// Replaces MetropolisIntegrator::step / MetropolisFlipIntegrator::step (src/integrator.rs:66-92, :109-138) on the
// basis-split SoA layout of heis_basis.cuh: all 2 / 4 colour (= basis) passes of a step in ONE cooperative launch whose
// work items run in wave order (colour b a few planes behind colour b-1), so that a pass finds its partner sublattices
// in L2.  The work item is the one of heis_basis_vec_kernel (basis_vec_item): same Philox keys, same arithmetic and
// summation order, bit-identical trajectories.
#pragma once
#include <string>
#include <vector>

#include "heis_basis.cuh"

namespace vg {

struct BasisWaveDesc {
    int device = 0;
    bool f64 = false;
    int unitcell = 2;                 // 1 = bcc, 2 = fcc
    uint32_t nx = 0, ny = 0, nz = 0;  // local cells (nz: planes of this slab)
    uint32_t z_offset = 0, nz_global = 0;
    void* arr[4][3] = {};             // [basis][component][cell]; a slab's arrays carry a halo plane below and above
    uint32_t lag = 0, ipt = 0, grid = 0;   // tuning, 0 = automatic
    // z-slab with neighbours on OTHER devices (one process per GPU): boundary planes go straight into the neighbours' halo
    // planes and their completion is signalled through flag words in peer memory
    bool slab = false;
    void* peer_lo = nullptr;          // base of the lower / upper neighbour's allocation (as BasisPeers)
    void* peer_hi = nullptr;
    unsigned long long* flags = nullptr;          // my words: [0..3] colour a's plane 0 of the UPPER neighbour is in my upper halo,
                                                  //           [4..7] colour c's top plane of the LOWER neighbour is in my lower halo (steps so far)
    unsigned long long* peer_flags[2] = {};       // the same words of the lower / upper neighbour
};

struct BasisWaveState;

BasisWaveState* basis_wave_create(const BasisWaveDesc& d, std::string& why_not);   // nullptr: no wave step for this lattice
void basis_wave_destroy(BasisWaveState*);
const char* basis_wave_describe(const BasisWaveState*);
template <typename real>
int basis_wave_step(BasisWaveState*, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                    double* obs_row, cudaStream_t st, std::string& err);
int basis_wave_check(BasisWaveState*, std::string& err);   // after a synchronize: != 0 when a wait inside the kernel timed out (the state
                                                            // then refuses further steps: the caller falls back to colour launches)
bool basis_wave_usable(const BasisWaveState*);
unsigned long long basis_wave_steps_done(const BasisWaveState*);   // slab: value the neighbours' flag words reach after the last step

// host-only: the unit order (colour << 24 | plane) of a step, for tests of the schedule's invariants
std::vector<uint32_t> basis_wave_units(int unitcell, uint32_t nz, uint32_t lag, uint32_t (&need)[4]);

}  // namespace vg
