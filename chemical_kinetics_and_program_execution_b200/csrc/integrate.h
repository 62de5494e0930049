// Device-resident DOP853 time stepping of the master equation dp/dt = S * w(p).
//
// The reference hands dy/dt to scipy.integrate.solve_ivp(method='DOP853')
// (framework/markov_tapes.py:349-354); every stage vector then crosses the host boundary.  This
// stepper keeps the table, the 16 stage derivatives and the dense-output polynomials in HBM and
// follows SciPy's controller (scipy/integrate/_ivp/rk.py: rk_step, RungeKutta._step_impl,
// DOP853._estimate_error_norm, DOP853._dense_output_impl; common.py: select_initial_step) so that
// trajectories agree with the reference path.  The Butcher tableau is supplied by the caller
// (Python passes scipy's dop853_coefficients).
#pragma once

#include <cstdint>
#include <memory>

#include "engine.h"

namespace tapes {

struct Dop853Tableau {
  double A[16][16];
  double B[12];
  double C[16];
  double E3[13];
  double E5[13];
  double D[4][16];
};

struct Dop853;

// y0: n_states doubles on the HOST.  first_step <= 0 selects the initial step like SciPy.
// peer (may be null): the ranks that evaluate the problem together.  Every rank then runs this
// stepper on the full table: the right-hand side is peer_rhs (identical bits on all ranks), so
// error norms and step sizes agree without further communication.
// The solver shares ownership of the structure: releasing the model or registering its tag again
// while the solver lives leaves the solver working on the structure it was created with.
Dop853* dop853_create(std::shared_ptr<Model> model, PeerGroup* peer, const Dop853Tableau& tab, const double* h_y0,
                      double t0, double t_bound, double rtol, double atol, double max_step, double first_step);
void dop853_destroy(Dop853* s);

// One solver.step(): returns 0 = running, 1 = finished, -1 = failed (step size too small).
int dop853_step(Dop853* s);

// Dense output of the last step evaluated at time t into a DEVICE buffer of n_states doubles.
void dop853_dense_eval(Dop853* s, double t, double* d_out);

// Device pointers to the current state and to the solver's dense-output buffer; the model.
const double* dop853_state(const Dop853* s);
double* dop853_dense_buffer(Dop853* s);
Model& dop853_model(Dop853* s);

// t, t_old, h_abs, nfev, accepted steps, rejected steps.
void dop853_info(const Dop853* s, double out[6]);

// Observables on device: sums[o] = sum_{j < count[o]} y[offset[o] + j * stride[o]], each summed in a
// fixed order by a grid of blocks.  offset/stride/count are HOST arrays; result to a HOST array.
void observe_strided(Model& m, const double* d_y, const int64_t* offset, const int64_t* stride,
                     const int64_t* count, int64_t n_obs, double* h_out);

// seq_prob of framework/markov_tapes.py:190-233 for symbol sequences of any length (HOST arrays:
// sequence o is symbols[seq_ptr[o] .. seq_ptr[o + 1])): up to cl_k symbols a strided sum over the
// leading axes, longer ones extended with the Markov process parameters clipped at eps.
void observe_sequences(Model& m, const double* d_y, int64_t n_seq, const int64_t* seq_ptr, const int32_t* symbols,
                       double eps, double* h_out);

// markov_entropy of framework/markov_tapes.py:178-187 of a DEVICE table.
double markov_entropy(Model& m, const double* d_y);

}  // namespace tapes
