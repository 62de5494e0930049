// Monte-Carlo simulation of a tape program on one long ring tape (see montecarlo.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "rules.h"

namespace tapes {

struct MonteCarlo;

// tape0: n_sites symbols (host).  events_per_substep < 2^20.
MonteCarlo* mc_create(const ProgramTree& tree, int alphabet, uint64_t n_sites, const uint8_t* tape0,
                      uint32_t events_per_substep, uint64_t seed);
void mc_destroy(MonteCarlo* mc);
void mc_run(MonteCarlo* mc, uint64_t n_substeps);
uint64_t mc_substeps_done(const MonteCarlo* mc);
// counts[A^k]: occurrences of every length-k window on the ring (sum = n_sites).
void mc_window_counts(MonteCarlo* mc, int cl_k, int64_t* h_counts);
void mc_fetch(MonteCarlo* mc, uint8_t* h_tape);

// Samples a ring whose length-k window statistics follow `table` (A^k doubles, host): the first
// window from the table, then symbol by symbol from the conditional given the k-1 symbols before.
void mc_sample_ring(int alphabet, int cl_k, const double* table, uint64_t n_sites, uint64_t seed, uint8_t* h_tape);

// The reference's own Monte Carlo of the ferromagnetic chain (examples/ex2_ferromagnet_mc.py:46-122,
// 134-163, 169-191), trial by trial: at every time step `trials_per_step` sites (given, with
// repetitions) are looked at in the state of the previous step and flipped in the new one when their
// uniform number (given) is below the acceptance factor of (equal neighbours, own spin) - two flips
// of one site in a step cancel, as in the reference - and the islands of up-spins of length 1..5
// are counted (ring).  All inputs on the HOST: chain0 [n_trials][chain_length] (0 / 1),
// sites and uniforms [n_trials][n_steps - 1][trials_per_step], accept [3][2] (equal neighbours 0..2,
// spin 0 / 1); counts [n_trials][n_steps][6] (entry 0 unused) receives the island counts, step 0 being
// the initial chain.  One thread block per trial with the chain in shared memory.
void mc_ferromagnet_chains(int64_t n_trials, int64_t chain_length, int64_t n_steps, int64_t trials_per_step,
                           const uint8_t* chain0, const int32_t* sites, const double* uniforms, const double* accept,
                           double* counts);

}  // namespace tapes
