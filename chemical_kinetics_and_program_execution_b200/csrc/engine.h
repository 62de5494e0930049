// Device engine: the window-extension forest of a rule table, its CSR flux structure, and the
// master-equation right-hand side evaluated on it.  See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "rules.h"

namespace tapes {

// Node kinds of the extension forest (stored in the top two bits of the per-node meta byte; the
// low six bits hold the node's window length).
enum NodeKind : uint8_t {
  NODE_ROOT = 0,   // weight = probability of the leaf world that owns the seed
  NODE_LEFT = 1,   // left extension / left shift (tm.scm:1340-1379): one parent
  NODE_RIGHT = 2,  // right extension (tm.scm:1303-1322): parent is a SUM node
  NODE_SUM = 3,    // sum over the dropped left-context digit of a right-chain prefix
};

struct Level {
  uint64_t base = 0;       // global id of the level's first node
  uint32_t n_plain = 0;    // ROOT / LEFT / RIGHT nodes, ids base .. base + n_plain - 1
  uint32_t n_sum = 0;      // SUM nodes, ids base + n_plain ..
  uint32_t* io = nullptr;       // [n_plain] index of the node's sequence in its marginal table
  uint32_t* parent = nullptr;   // [n_plain] global id of the parent (ROOT: rule number)
  uint8_t* meta = nullptr;      // [n_plain] kind << 6 | length
  uint64_t* sum_ptr = nullptr;      // [n_sum + 1] offsets into sum_parents
  uint32_t* sum_parents = nullptr;  // global ids, ascending inside each SUM node
  uint64_t n_sum_parents = 0;
};

struct BuildStats {
  int64_t worlds_walked = 0, leaf_worlds = 0, flux_rules = 0, seeds = 0;
  int64_t nodes = 0, sum_nodes = 0, terms = 0, levels = 0;
  int64_t hash_inserts = 0, hash_unique = 0;
  double host_enumerate_ms = 0, device_expand_ms = 0, device_csr_ms = 0;
};

struct Model {
  int A = 0, k = 0;
  uint64_t n_states = 0;  // A^k
  uint64_t pow_a[40];
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  // flux rules (probability steps of the leaf worlds)
  uint32_t n_rules = 0;
  uint32_t* rule_ptr = nullptr;    // [n_rules + 1]
  uint8_t* step_kind = nullptr;    // Step::Kind
  uint8_t* step_len = nullptr;
  uint32_t* step_long = nullptr;
  uint32_t* step_short = nullptr;
  double* step_prob = nullptr;
  double* rule_w = nullptr;        // [n_rules] evaluated per right-hand side

  // extension forest
  std::vector<Level> levels;
  uint64_t n_nodes = 0;
  double* node_w = nullptr;        // [n_nodes] evaluated per right-hand side

  // flux structure S (n_states x n_nodes, entries +-1) in CSR by state
  uint64_t nnz = 0;
  uint64_t* row_ptr = nullptr;     // [n_states + 1]
  uint32_t* entries = nullptr;     // [nnz] node id | sign << 31 (1 = outflow)
  int spmv_group = 8;              // lanes per row

  // marginal tables marg_L, L < k, concatenated; marg_off[L] = offset in doubles
  double* marg = nullptr;
  uint64_t marg_off[40];
  uint64_t* d_marg_off = nullptr;
  uint64_t marg_total = 0;

  // staging for the host-buffer entry point
  double* d_in = nullptr;
  double* d_out = nullptr;

  BuildStats stats;
  int64_t launches_per_rhs = 0;

  ~Model();
};

// Builds the device structures for a rule table.  Throws std::runtime_error on failure.
std::unique_ptr<Model> build_model(const RuleTable& table, cudaStream_t stream);

// dy/dt for device-resident p (n_states doubles) into device-resident out.  Asynchronous on
// model.stream (or `stream` when non-null).
void rhs_device(Model& m, const double* d_p, double* d_out, cudaStream_t stream);

// Same, with CUDA events recorded on the launching stream between the phases; synchronises and
// returns ms for (marginals + leaf-world probabilities, forest levels, S * w).
void rhs_device_profiled(Model& m, const double* d_p, double* d_out, cudaStream_t stream, float ms[3]);

// Host-buffer variant (the c_compute_dy_dt path): H2D, rhs_device, D2H, synchronise.
void rhs_host(Model& m, const double* h_p, double* h_out);

// Number of device kernels launched by one rhs_device call.
int64_t rhs_launch_count(const Model& m);

}  // namespace tapes
