// Validation of a subsequence-probability table without the dense eigen-decomposition of
// framework/markov_tapes.py:133-175 (see validate.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace tapes {

struct TableCheck {
  double total = 0;                  // sum of the table
  double marginal_distance = 0;      // |sum over last symbol - sum over first symbol|_2  (mt.py:160-164)
  double stationarity_residual = 0;  // |T pi - pi|_2 for pi = the context marginal
  double power_distance = 0;         // |v - pi|_2 for v = limit of v <- T v from the uniform vector
  double last_change = 0;            // |v_n - v_{n-1}|_2 at the last convergence test
  int iterations = 0;
};

// d_p: A^cl_k doubles on the device.  eps_mpp: the clip of mpp_from_spd (mt.py:101-104).
TableCheck check_table(int alphabet, int cl_k, const double* d_p, double eps_mpp, int max_iterations,
                       double tolerance, cudaStream_t st);

}  // namespace tapes
