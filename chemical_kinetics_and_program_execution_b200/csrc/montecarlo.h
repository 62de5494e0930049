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

}  // namespace tapes
