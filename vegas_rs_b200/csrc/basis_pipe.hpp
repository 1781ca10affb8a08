// basis_pipe.hpp -- host interface of K4p, the phase-pipelined step for periodic bcc / fcc Heisenberg lattices (basis_pipe.cu).
//
// Replaces MetropolisIntegrator::step / MetropolisFlipIntegrator::step (src/integrator.rs:66-92, :109-138) on the
// basis-split SoA layout of heis_basis.cuh: all 2 / 4 colour (= basis) passes of a step in ONE cooperative launch.
// Same Philox keys, same arithmetic and summation order as heis_basis_vec_kernel: bit-identical trajectories.
#pragma once
#include <string>

#include "heis_basis.cuh"

namespace vg {

struct BasisPipeDesc {
    int device = 0;
    bool f64 = false;
    int unitcell = 2;                 // 1 = bcc, 2 = fcc
    uint32_t nx = 0, ny = 0, nz = 0;
    void* arr[4][3] = {};             // [basis][component][cell]
    uint32_t tiles = 0, lead = 0, pub_every = 0;   // tuning, 0 = automatic
};

struct BasisPipeState;

BasisPipeState* basis_pipe_create(const BasisPipeDesc& d, std::string& why_not);   // nullptr: lattice does not fit
void basis_pipe_destroy(BasisPipeState*);
const char* basis_pipe_describe(const BasisPipeState*);
template <typename real>
int basis_pipe_step(BasisPipeState*, const HeisParams<real>& p, bool flip, bool record, uint64_t sweep, const PhiloxKey& pk,
                    double* obs_row, cudaStream_t st, std::string& err);
int basis_pipe_check(BasisPipeState*, std::string& err);   // after a synchronize: != 0 when a wait inside the kernel timed out

}  // namespace vg
