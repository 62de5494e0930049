/* C ABI of the B200 tape-multiverse library (built as tapes_py_interface.so).
 *
 * The first four symbols are exactly what the reference's Python layer binds with ctypes
 * (framework/markov_tapes.py:40-56); each replaces the reference implementation cited beside
 * it.  The tapes_* symbols are additions for device-resident use, introspection and tests.
 *
 * All pointers are plain host or device addresses; there are no torch types in this interface.
 * Unless stated otherwise a function returns 0 on success and non-zero on failure, in which case
 * tapes_last_error() describes the problem.  The library is single-threaded like the reference
 * (global registry, framework/tapes_py_interface.scm:24).
 */
#ifndef TAPES_B200_H_
#define TAPES_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- drop-in symbols ------------------------------------------------------------------- */

/* Replaces setup_gambit (framework/tapes_py_interface_c_glue.c:31-40): selects the CUDA device
 * (LOCAL_RANK modulo the device count when set, else the current device), creates the runtime
 * and returns an opaque handle; NULL when no sm_100 device is usable. */
void* setup_gambit(void);

/* Replaces cleanup_gambit (framework/tapes_py_interface_c_glue.c:42-45): frees every cached
 * model and the handle. */
void cleanup_gambit(void* handle);

/* Replaces c_register_problems (framework/tapes_py_interface.scm:101-112): registers the canary
 * and the problems of framework/problems.scm under the same tags, prints the registry listing
 * (problems.scm:631-638) and returns n + 1. */
int64_t c_register_problems(int64_t n);

/* Replaces c_compute_dy_dt (framework/tapes_py_interface.scm:115-122 -> 80-94): probs_in and
 * probs_out are HOST buffers of A^cl_k doubles (A = alphabet of `tag`).  The structure for
 * (tag, cl_k) is built on first use and cached.  `debug` is accepted and ignored, as in the
 * reference's shipped configuration (framework/tape_multiverse.scm:1448-1449). */
void c_compute_dy_dt(const char* tag, int64_t cl_k, int64_t debug, const double* probs_in,
                     double* probs_out);

/* ---- additions -------------------------------------------------------------------------- */

/* Message of the most recent failure ("" when none); cleared by tapes_clear_error. */
const char* tapes_last_error(void);
void tapes_clear_error(void);

/* Page-locked host memory for tables and results: c_compute_dy_dt moves pinned buffers by plain DMA
 * (about 55 GB/s over PCIe 5 x16) and pageable ones - what NumPy hands the reference's binding -
 * through bounce buffers filled by several host threads (csrc/hostcopy.h).  markov_tapes.get_dy_dt
 * takes its result arrays for large tables from a small pool of such buffers.  NULL on failure. */
void* tapes_host_alloc(int64_t bytes);
int tapes_host_free(void* p);

/* Alphabet size registered for `tag`, or -1. */
int64_t tapes_alphabet_size(const char* tag);

/* Registers a rewrite-rule set as a problem (see csrc/rules.h RewriteRule).  pattern and
 * replacement hold 4 int32 per rule. */
int tapes_register_rules(const char* tag, int64_t alphabet, int64_t n_rules, const int32_t* tape,
                         const int32_t* span, const int32_t* catalyst, const int32_t* pattern,
                         const int32_t* replacement, const double* rate,
                         const double* select_weight);

/* Registers a problem given as the decision tree of its body - the data form of what
 * framework/problems.scm states with tape-get / tape-set! / choose
 * (framework/gambit_macros.scm:99-125); the reference needs an edit of problems.scm and a rebuild of
 * the shared object for a new problem (MAKE.sh:43-47).  Per node: kind (0 end, 1 read, 2 write,
 * 3 choose), a / b / c (read: tape 0 = program 1 = data, cell; write: tape, cell, symbol; choose:
 * number of options), first_child (offset into child: `alphabet` children by symbol read, one after a
 * write, one per option), first_weight (choose: offset into weight, unnormalised weights as in
 * gambit_macros.scm:75-86).  Node 0 is the entry and children come after their parent.
 * markov_tapes.register_program traces a Python function into this form. */
int tapes_register_program(const char* tag, int64_t alphabet, int64_t n_nodes, const int32_t* kind,
                           const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* first_child,
                           const int32_t* first_weight, int64_t n_children, const int32_t* child,
                           int64_t n_weights, const double* weight);

/* Builds (or fetches from the cache) the device structure for (tag, cl_k). NULL on failure.
 * Limits: A^cl_k < 2^32, cl_k <= 32; the forest itself may be of any size that fits the device
 * memory (it is split into structures of at most 2^31 nodes each, see tapes_model_info). */
void* tapes_model(const char* tag, int64_t cl_k);
/* Frees the structure of (tag, cl_k) and of every part of it; 1 when there was none. */
int tapes_release_model(const char* tag, int64_t cl_k);

/* Part `part` of `n_parts` of (tag, cl_k), for ranks that evaluate one problem together: the flux
 * rules (leaf worlds of the program that modify a tape, framework/tape_multiverse.scm:1416-1443)
 * are dealt to the parts by estimated term count, and the part's right-hand side is the sum over
 * its rules only.  The sum of all parts' dy/dt is the dy/dt of the whole problem, which is what
 * tapes_peer_rhs forms.  Works for every registered problem.  NULL on failure. */
void* tapes_model_part(const char* tag, int64_t cl_k, int64_t part, int64_t n_parts);

/* Host-only (no GPU needed): the dealing tapes_model_part uses.  owner[r] = part of flux rule r
 * (order of tapes_rule_table), cost[r] = its estimated number of flux terms; either may be NULL.
 * Returns the number of flux rules, -1 on failure. */
int64_t tapes_rule_parts(const char* tag, int64_t cl_k, int64_t n_parts, int32_t* owner, double* cost);

/* dy/dt for DEVICE buffers; asynchronous on `cuda_stream` (a cudaStream_t, NULL = the model's
 * own stream). */
int tapes_rhs_device(void* model, const double* d_probs_in, double* d_probs_out, void* cuda_stream);

/* The two halves of tapes_rhs_device, for callers that overlap the product with communication:
 * tapes_weights_device evaluates everything that depends on p (marginal tables, leaf-world
 * probabilities, all forest levels); tapes_flux_rows_device then writes dy/dt for the states
 * row_lo <= i < row_hi into d_probs_out[row_lo .. row_hi) from those weights. */
int tapes_weights_device(void* model, const double* d_probs_in, void* cuda_stream);
int tapes_flux_rows_device(void* model, double* d_probs_out, int64_t row_lo, int64_t row_hi,
                           void* cuda_stream);

/* ---- multi-GPU flux exchange over NVLink peer memory (one process per GPU, one node) -------
 * The rule set is dealt to the ranks; every rank evaluates its rules over the full table.  States
 * are owned in contiguous blocks of `block` states (a multiple of 32 * rounds), state i by rank
 * i / block.  tapes_peer_rhs evaluates dy/dt of the whole problem on all ranks together:
 * the product kernel stores this rank's partial dy/dt of a state directly into the owner's staging
 * buffer (slot `rank` of world slots of `block` doubles) by peer stores while it computes; in
 * `rounds` rounds the owners add their slots in rank order and store the sums into every rank's
 * result vector, one round behind the product.  Cross-GPU ordering uses epoch flags in peer
 * memory; a wait that exceeds 20 s sets the group's error flag instead of hanging.
 * Buffers come from tapes_peer_alloc (cudaMalloc, zero-filled, plus a 64-byte CUDA IPC handle to
 * send to the other ranks) and tapes_peer_open (maps a received handle).  Per rank: staging
 * (world * block doubles), result (world * block doubles), flags (2 * world 64-bit words). */
void* tapes_peer_alloc(int64_t n_doubles, void* ipc_handle64);
void* tapes_peer_open(const void* ipc_handle64);
int tapes_peer_close(void* d_ptr);
int tapes_peer_free(void* d_ptr);
/* staging / result / flags: HOST arrays of `world` device pointers, entry q = rank q's buffer as
 * mapped in this process (the own buffers for q = rank).  NULL on failure. */
void* tapes_peer_group_create(int world, int rank, int64_t block, int rounds, void* const* staging,
                              void* const* result, void* const* flags);
void tapes_peer_group_destroy(void* group);
/* Collective over the group: every rank calls it with the same table.  Asynchronous on
 * `cuda_stream`; the sum is in this rank's result buffer when the stream reaches this point. */
int tapes_peer_rhs(void* group, void* model, const double* d_probs_in, void* cuda_stream);
/* Non-zero once a wait timed out (synchronise first). */
int tapes_peer_group_error(void* group);

/* One right-hand side with CUDA events between its phases, recorded on the launching stream;
 * synchronises and writes the phase durations in ms: [0] marginal tables + leaf-world
 * probabilities, [1] forest levels, [2] S * w. */
int tapes_rhs_profile(void* model, const double* d_probs_in, double* d_probs_out, void* cuda_stream,
                      double* phase_ms, int capacity);

/* Blocks until the model's stream is idle. */
int tapes_sync(void* model);

/* Integer facts about a model, in this order: n_states, n_nodes, nnz, n_flux_rules, n_levels,
 * kernel launches per right-hand side, n_terms, n_sum_nodes, worlds_walked, leaf_worlds, seeds,
 * hash_inserts, hash_unique, alphabet, cl_k, CSR-kernel lanes per row, flux format (1 = slices of
 * 32 states, 0 = plain CSR), slices, 32-bit words of the sliced form, runs, entries held by runs,
 * entries held by columns, column slots incl. padding, minimum lanes of a run, loads in flight per
 * thread of the level kernel, forest levels whose parent lists are not arithmetic progressions,
 * left-parent records of all levels, gathers in flight per lane of the product kernel, right
 * children evaluated by the group they feed, groups whose children are evaluated by the next level,
 * structures the model consists of (more than one when the forest exceeds the 31-bit node ids: the
 * flux rules are then split over several structures evaluated one after the other; sizes above are
 * sums over them; TAPES_MAX_PART_TERMS overrides the ~1.7 * 10^9 flux terms a structure may hold), forest
 * levels whose blocks of prefix groups are evaluated in prefix order across seeds, prefix groups that
 * sit in regular blocks of 256 and are evaluated by the plane kernel (csrc/engine.h Level::PlaneBlock),
 * per-step ratio tables (0, 1: right extensions, 2: also left extensions to a full window), entries
 * the flux structure stores (nnz minus the entries of right children, whose flux is evaluated per
 * prefix group from the group sums), 1 when the weights of right children are written per step because
 * a later level reads them (else they exist only after tapes_export_node_weights), levels of the
 * build whose table of prefixes was sized too small at first and redone with the safe size, table /
 * ratio entries one step reads at most once per level (sum over the levels of min(nodes, states):
 * what bench.py counts as the compulsory table reads of the level phase).
 * Returns how many were written. */
int tapes_model_info(void* model, int64_t* out, int capacity);

/* Tuning knobs of a built model: "spmv_lanes" (1, 2, 4, 8 or 16 lanes per row of the plain-CSR
 * kernel), "level_unroll" (1..8 loads in flight per thread of the level kernel), "flux_unroll"
 * (2, 3, 4, 6 or 8 gathers in flight per lane of the sliced product kernel), "interleave_seeds" (1: the
 * level kernel evaluates the blocks of prefix groups of different seeds in prefix order so that they
 * share their reads of the table through L2, 0: in storage order), "ratio_table" (1: the ratios of right extensions are evaluated once per
 * step into a table the level kernel reads, 0: per node), "plane_kernel" (1: regular blocks of 256
 * prefix groups are evaluated by the record-free plane kernel, 0: by the general level kernel), "graphs" (1: for tables of up to
 * 2^22 states the kernels that evaluate the weights are replayed from a CUDA graph captured per input
 * pointer, 0: launched one by one).  Results do not depend on any of them. */
int tapes_model_set(void* model, const char* key, int64_t value);

/* Build timings in ms: host rule enumeration, device expansion, device CSR assembly, slicing, and
 * the part of the expansion spent inside cudaMalloc / cudaFree. */
int tapes_model_timing(void* model, double* out, int capacity);

/* Copies the complete flux structure to host in canonical CSR form: row_ptr has n_states + 1
 * entries, entries has nnz (node id | outflow << 31), ascending inside each row.  It is rebuilt from
 * what the device holds: the sliced form of the stored entries plus the entries of the right
 * children, whose flux the device evaluates per prefix group. */
int tapes_export_csr(void* model, int64_t* row_ptr, uint32_t* entries);

/* Copies the node weights of the most recent right-hand side to host (n_nodes doubles); the weights
 * of right children, which a step does not write, are filled in from the group sums first. */
int tapes_export_node_weights(void* model, double* weights);

/* ---- device-resident time stepping (replaces the SciPy stepper call sites
 * framework/markov_tapes.py:318 and 349-354 for runs that keep the table in HBM) ------------ */

/* Creates a DOP853 solver that follows scipy.integrate.solve_ivp(method='DOP853') step for step.
 * tableau: 374 doubles = A[16][16], B[12], C[16], E3[13], E5[13], D[4][16] (row-major; SciPy's
 * dop853_coefficients).  y0: HOST buffer of n_states doubles.  max_step <= 0 means unbounded,
 * first_step <= 0 means "select like SciPy".  NULL on failure. */
void* tapes_dop853_create(void* model, const double* tableau, const double* y0, double t0,
                          double t_bound, double rtol, double atol, double max_step,
                          double first_step);
/* The same stepper for a group of ranks (tapes_peer_group_create): every rank holds the full table
 * and calls every tapes_dop853_* function in lockstep; right-hand sides are tapes_peer_rhs. */
void* tapes_dop853_create_peer(void* model, void* group, const double* tableau, const double* y0,
                               double t0, double t_bound, double rtol, double atol, double max_step,
                               double first_step);
void tapes_dop853_destroy(void* solver);

/* One solver.step(): 0 running, 1 finished, -1 step size too small, -2 error. */
int tapes_dop853_step(void* solver);

/* Evaluates the dense output of the last step at time t into the solver's device buffer. */
int tapes_dop853_dense(void* solver, double t);

/* Copies the current state (which = 0) or the dense-output buffer (which = 1) to a HOST buffer. */
int tapes_dop853_fetch(void* solver, int which, double* out);

/* Strided sums of the current state (which = 0) or the dense-output buffer (which = 1), computed
 * on the device: out[o] = sum_{j < count[o]} y[offset[o] + j * stride[o]].  A length-L sequence
 * probability (framework/markov_tapes.py:190-222) is offset = index(seq), stride = A^L,
 * count = A^(k-L). */
int tapes_dop853_observe(void* solver, int which, const int64_t* offset, const int64_t* stride,
                         const int64_t* count, int64_t n_obs, double* out);

/* t, t_old, h_abs, nfev, accepted steps, rejected steps. */
int tapes_dop853_info(void* solver, double* out6);

/* Same strided sums for any DEVICE vector of n_states doubles. */
int tapes_observe(void* model, const double* d_y, const int64_t* offset, const int64_t* stride,
                  const int64_t* count, int64_t n_obs, double* out);

/* seq_prob of framework/markov_tapes.py:190-233 evaluated on the device for a DEVICE table of n_states
 * doubles: sequence o is symbols[seq_ptr[o] .. seq_ptr[o + 1]) (HOST arrays).  Sequences of up to cl_k
 * symbols are read off the table (last axes fixed, leading axes summed, markov_tapes.py:215-222); longer
 * ones are extended with the Markov process parameters of markov_tapes.py:81-104 clipped at eps
 * (markov_tapes.py:223-233; the reference's default eps is 1e-100).  out: n_seq doubles (HOST). */
int tapes_observe_sequences(void* model, const double* d_y, int64_t n_seq, const int64_t* seq_ptr,
                            const int32_t* symbols, double eps, double* out);
/* The same for the current state (which = 0) or the dense-output buffer (which = 1) of a solver. */
int tapes_dop853_observe_sequences(void* solver, int which, int64_t n_seq, const int64_t* seq_ptr,
                                   const int32_t* symbols, double eps, double* out);

/* markov_entropy of framework/markov_tapes.py:178-187 (entropy rate of the chain the table describes)
 * of a DEVICE table, evaluated on the device; *out receives the value. */
int tapes_markov_entropy(void* model, const double* d_y, double* out);
int tapes_dop853_entropy(void* solver, int which, double* out);

/* Validation of a table of A^cl_k doubles (HOST buffer, or DEVICE buffer when on_device != 0)
 * without the dense eigen-decomposition of framework/markov_tapes.py:133-175: the context transfer
 * matrix of markov_tapes.py:107-130 has A non-zeros per row and is applied by a streaming kernel.
 * out6: sum of the table; |last-axis marginal - first-axis marginal|_2 (markov_tapes.py:160-164);
 * |T pi - pi|_2 for the context marginal pi; |v - pi|_2 for the limit v of the power iteration
 * v <- T v from the uniform vector (0 iterations: skipped); |v_n - v_{n-1}|_2 at the last
 * convergence test; iterations done.  eps_mpp is the clip of mpp_from_spd (markov_tapes.py:101). */
int tapes_check_table(int64_t alphabet, int64_t cl_k, const double* probs, int on_device, double eps_mpp,
                      int64_t max_iterations, double tolerance, double* out6);

/* ---- Monte-Carlo simulation of a registered problem on one long ring tape: an independent check
 * of the closure behind the master equation (the reference has one for the ferromagnet only,
 * examples/ex2_ferromagnet_mc.py:46-122).  An event puts the program head and the data head on two
 * random sites and runs the program there; events_per_substep events happen at once and a sub-step
 * advances time by events_per_substep / n_sites (every site is visited at rate 1, the normalisation
 * of compute-dy/dt).  Deterministic given the seed; rules in csrc/montecarlo.cu. ---------------- */

/* tape0: HOST array of n_sites symbols (one byte each, alphabet <= 256).  NULL on failure. */
void* tapes_mc_create(const char* tag, int64_t n_sites, const uint8_t* tape0, int64_t events_per_substep,
                      uint64_t seed);
void tapes_mc_destroy(void* mc);
int tapes_mc_run(void* mc, int64_t n_substeps);
/* counts: HOST array of A^cl_k entries: occurrences of every length-cl_k window on the ring. */
int tapes_mc_window_counts(void* mc, int64_t cl_k, int64_t* counts);
/* Copies the ring to a HOST array of n_sites bytes. */
int tapes_mc_fetch(void* mc, uint8_t* tape);
/* Host-only: a ring whose length-cl_k window statistics follow `table` (A^cl_k doubles). */
int tapes_mc_sample_ring(int64_t alphabet, int64_t cl_k, const double* table, int64_t n_sites, uint64_t seed,
                         uint8_t* tape);

/* The reference's own Monte Carlo of the ferromagnetic chain (examples/ex2_ferromagnet_mc.py:46-122
 * `simulate`, 134-163 `island_length_stats`, driver loop 169-191), all trials at once, one thread block
 * per trial with the chain in shared memory.  At every time step trials_per_step sites are looked at in
 * the state of the previous step and flipped in the new one when their uniform number is below
 * accept[equal neighbours][own spin]; two flips of a site within a step cancel, as in the reference.
 * HOST arrays: chain0 [n_trials][chain_length] of 0 / 1; sites (int32) and uniforms (double)
 * [n_trials][n_steps - 1][trials_per_step] - the numbers the reference draws from
 * numpy.random.RandomState(seed) at ex2_ferromagnet_mc.py:93-94; accept [3][2]; counts
 * [n_trials][n_steps][6] receives the numbers of up-spin islands of length 1..5 (entry 0 stays 0), step 0
 * being the initial chain - the layout of the reference's ferromagnet_mc_chain_counts.npz. */
int tapes_mc_ferromagnet_chains(int64_t n_trials, int64_t chain_length, int64_t n_steps, int64_t trials_per_step,
                                const uint8_t* chain0, const int32_t* sites, const double* uniforms, const double* accept,
                                double* counts);

/* Host-only: the decision tree of the body registered under `tag` in the array form of
 * tapes_register_program (any problem: compiled, rewrite rules, or registered as a tree).  Call with
 * kind == NULL to get the lengths in sizes3 = {nodes, children, weights}.  Returns the node count. */
int64_t tapes_program_tree(const char* tag, int64_t* sizes3, int32_t* kind, int32_t* a, int32_t* b, int32_t* c,
                           int32_t* first_child, int32_t* first_weight, int32_t* child, double* weight);

/* Host-only (no GPU needed): the flux-rule table of (tag, cl_k).  Call with all pointers NULL to
 * get sizes: returns the number of rules and stores the total step count in *n_steps. Arrays:
 * rule_ptr[n_rules + 1]; per step kind, length, long_index, short_index, prob; per rule and tape
 * (2 per rule) seed_len, seed_orig, seed_adj.  Returns -1 on failure. */
int64_t tapes_rule_table(const char* tag, int64_t cl_k, int64_t* n_steps, int64_t* rule_ptr,
                         int32_t* step_kind, int32_t* step_len, int64_t* step_long,
                         int64_t* step_short, double* step_prob, int32_t* seed_len,
                         uint64_t* seed_orig, uint64_t* seed_adj, int64_t* walk_stats);

#ifdef __cplusplus
}
#endif

#endif /* TAPES_B200_H_ */
