// Device engine: the window-extension forest of a rule table, its CSR flux structure, and the
// master-equation right-hand side evaluated on it.  See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rules.h"

namespace tapes {

// Node kinds of the extension forest (build-time only; the top two bits of a frontier node's
// meta byte, the low six bits hold the node's window length).
enum NodeKind : uint8_t {
  NODE_ROOT = 0,   // weight = probability of the leaf world that owns the seed
  NODE_LEFT = 1,   // left extension / left shift (tm.scm:1340-1379)
  NODE_RIGHT = 2,  // right extension (tm.scm:1303-1322)
};

// One level of the forest as it is evaluated every step.  Child nodes are implicit: only their
// parents are stored.
//   level 0:  n_roots nodes, node i has the weight of flux rule root_rule[i].
//   level >0: the A children of left parent r are nodes  base + x * n_left + r  (x = new, most
//             significant digit); the A children of prefix group g are nodes
//             base + A * n_left + g * A + x  (x = new, least significant digit).
// A prefix group collects the right-chain nodes of the previous level that share the window
// prefix the next right extension starts from (the literal recursion visits that prefix once per
// dropped left-context digit, tm.scm:1319-1322); its children get  sum(parents) * ratio.
struct Level {
  uint64_t base = 0;
  uint32_t n_roots = 0, n_left = 0, n_groups = 0;
  uint32_t* root_rule = nullptr;  // [n_roots]
  uint32_t* lp_gid = nullptr;     // [n_left] global id of the parent node
  uint32_t* lp_io = nullptr;      // [n_left] table index shared by the children (before adding x)
  uint8_t* lp_len = nullptr;      // [n_left] window length of the children
  uint32_t* g_prefix = nullptr;   // [n_groups] (k-1)-digit window prefix
  uint32_t* g_adjusted = nullptr; // [n_groups] the same prefix in the adjusted window: the children's inflow goes to rows g_adjusted * A + x
  uint64_t* g_ptr = nullptr;      // [n_groups + 1] offsets into g_parents
  uint32_t* g_parents = nullptr;  // parent node ids, ascending inside each group
  uint64_t n_group_parents = 0;
  // when every parent list of the level is an arithmetic progression (the usual case) the lists
  // are replaced by first + j * stride, j < count, and g_ptr / g_parents are null
  uint32_t* g_first = nullptr;
  uint32_t* g_stride = nullptr;
  uint32_t* g_count = nullptr;    // low 29 bits: parents; flag bits kOwnsParents, kChildrenDeferred, kAllDigits
  // Fused right chain.  A right child is written once and gathered again by the group sum of the
  // next level; when the parents of a group are right children of the previous level, the group
  // computes and stores them itself (flag kOwnsParents: parent j is child  (first - prev_right_base
  // + j * stride) % A  of previous-level group  (first - prev_right_base + j * stride) / A), from
  // that group's sum in prev_total, and the previous level skips the children of such groups
  // (flag kChildrenDeferred) and leaves their sum in g_total instead.  Same values, same order of
  // additions; one 8-byte read per right child less.
  // kAllDigits (with kOwnsParents): the A parents are the A values of the dropped digit in order,
  // so parent j reads table index j * A^(k-1) + prefix and no previous-level prefix is needed.
  static constexpr uint32_t kOwnsParents = 0x80000000u, kChildrenDeferred = 0x40000000u,
                            kAllDigits = 0x20000000u, kCountMask = 0x1fffffffu;
  double* g_total = nullptr;            // [n_groups] sums of the groups whose children are deferred
  uint64_t prev_right_base = 0;         // node id of the previous level's first right child
  const uint32_t* prev_prefix = nullptr;  // previous level's g_prefix
  const double* prev_total = nullptr;     // previous level's g_total
  // Groups are stored seed-major (canonical order), but the chains of different seeds read the same
  // entries of p at the same prefix.  block_order[b] = which block of 256 groups thread block b of
  // the group part evaluates: blocks sorted by the prefix they start at, so that the seeds walk
  // through p together and the second reader of an entry finds it in L2.  Null when the level has
  // one seed only (identity) or the option is off.
  const uint32_t* block_order = nullptr;
  // Plane blocks.  Most prefix groups of a large problem come in long regular stretches: one seed's
  // groups at one level.  Their first parents are consecutive node ids, they share one stride, one
  // parent count (A parents that are the A values of the dropped digit, or the single parent of a
  // group whose dropped digit is a fixed cell of the view) and one set of flags, and their prefixes
  // are consecutive except where the fixed cells of the seed's view sit inside the prefix: there the
  // prefix jumps by a constant every `period` groups.  A block of 256 consecutive groups that is
  // regular in this sense needs no per-group records at all: one 48-byte record describes it
  // (PlaneBlock), and plane_kernel evaluates it with every load of a thread independent of the
  // others.  plane_blocks lists those blocks, general_blocks the others (evaluated by level_kernel
  // as before), both in prefix order like block_order.
  struct PlaneBlock {
    uint32_t group_block;  // groups 256 * group_block ... + 255
    uint32_t prefix0;      // prefix of the first group; group t: prefix0 + t + jump * (jumps before t)
    uint32_t first0;       // first parent of the first group; group t has first0 + t
    uint32_t long_off;     // table index of parent j of a group = its prefix + long_off + j * A^(k-1)
    uint32_t stride;       // node-id distance between the parents of a group
    uint32_t meta;         // parents per group (1 or A) | kPlaneDeferred | kPlaneGather
    uint32_t jump_at;      // last group before the first jump of the prefix (0xffffffff: no jump in the block)
    uint32_t period;       // groups between jumps
    uint32_t jump;         // what a jump adds on top of the usual + 1
    uint32_t pad[3];
  };
  static constexpr uint32_t kPlaneDeferred = 0x80000000u;
  static constexpr uint32_t kPlaneGather = 0x40000000u;  // the single parent is a stored node (not owned): its weight is read
  const PlaneBlock* plane_blocks = nullptr;
  uint32_t n_plane_blocks = 0;
  const uint32_t* general_blocks = nullptr;
  uint32_t n_general_blocks = 0;
};

// The flux structure cut into slices of 32 consecutive states, one warp lane per state.  Consecutive
// states gather consecutive forest nodes, so most of a slice is "runs":
// (first node id | sign << 31, lane mask) meaning lane l, if its mask bit is set, holds the entry
// first + popcount(mask below l).  Entries outside long enough runs are stored as padded columns of
// 32 (0xffffffff = none).  Words of slice s start at words[slice_ptr[s]]: slice_runs[s] run pairs
// (padded to an even count), then the columns.
struct FluxSlices {
  uint64_t n_slices = 0;
  uint64_t* slice_ptr = nullptr;   // [n_slices + 1]
  uint32_t* slice_runs = nullptr;  // [n_slices]
  uint32_t* words = nullptr;
  uint64_t n_words = 0;
  // facts about the encoding
  uint64_t runs = 0, run_entries = 0, column_entries = 0, column_slots = 0;
  int min_run_lanes = 0;
};

// Device memory that lives as long as a model: a bump allocator over a few large cudaMalloc
// chunks.  A model has a few hundred arrays; asking the driver for each of them cost 45-780 ms per
// build (more the more memory the process already holds), most of the build time of the
// reference's small problems.
struct DeviceArena {
  std::vector<void*> chunks;
  char* cursor = nullptr;
  size_t left = 0, next_chunk = 32u << 20, bytes = 0;
  DeviceArena() {}
  DeviceArena(const DeviceArena&) = delete;
  DeviceArena& operator=(const DeviceArena&) = delete;
  ~DeviceArena() { release(); }
  void* take(size_t n_bytes);  // 256-byte aligned; throws std::runtime_error when out of memory
  template <typename T>
  T* array(size_t n) { return (T*)take((n ? n : 1) * sizeof(T)); }
  void release();
};

struct BuildStats {
  int64_t worlds_walked = 0, leaf_worlds = 0, flux_rules = 0, seeds = 0;
  int64_t nodes = 0, sum_nodes = 0, terms = 0, levels = 0;  // sum_nodes = prefix groups
  int64_t hash_inserts = 0, hash_unique = 0;
  int64_t hash_retries = 0;  // levels whose table of prefixes was sized too small at first (engine.cu, pass 1)
  int64_t irregular_levels = 0;  // levels whose parent lists are kept explicitly
  int64_t left_parents = 0;      // parent records of left extensions / left shifts, all levels
  int64_t owned_parents = 0;     // right children computed by the group they feed (fused right chain)
  int64_t deferred_groups = 0;   // groups whose children are computed by the next level
  int64_t plane_groups = 0;      // groups evaluated by plane_kernel (in blocks of 256)
  double host_enumerate_ms = 0, device_expand_ms = 0, device_csr_ms = 0, device_slices_ms = 0;
  double expand_alloc_ms = 0;  // part of device_expand_ms spent inside cudaMalloc / cudaFree
};

struct Model {
  DeviceArena arena;  // owns every array below except `entries`
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
  uint64_t nnz = 0;                // 2 per flux term
  uint64_t nnz_stored = 0;         // entries held by row_ptr / entries / slices: the outflow of right children is not (out_ptr)
  uint64_t* row_ptr = nullptr;     // [n_states + 1]
  uint32_t* entries = nullptr;     // [nnz] node id | sign << 31 (1 = outflow); freed once the sliced
                                   // form below exists (tapes_export_csr rebuilds it on demand)
  int spmv_group = 8;              // lanes per row of the CSR kernel

  // the same structure in the form the product kernel streams (flux.cu): slices of 32 states
  FluxSlices slices;
  int flux_format = 1;             // 1 = slices, 0 = plain CSR
  int level_unroll = 4;            // loads in flight per thread in level_kernel
  int flux_unroll = 3;             // gathers in flight per lane in flux_slices_kernel (3: 32 registers, 8 blocks per SM)
  int interleave_seeds = 1;        // level kernel: use Level::block_order
  int ratio_table = 1;             // extension ratios of full windows evaluated once per step into tables
  double* ratio_right = nullptr;   // [n_states] p[i] / max(p[i], marg_{k-1}[i / A]), 0 where p[i] == 0
  // Flux of the right children, evaluated per prefix group instead of per stored term (engine.cu
  // build_model).  A right child x of group g (prefix q, adjusted prefix q') has weight
  // sum(g) * ratio_right[q * A + x]; it leaves row q * A + x and enters row q' * A + x.
  //   outflow: all right children that leave a row share the factor ratio_right[row], so per prefix q
  //            the groups of all levels and seeds with prefix q are listed (out_ptr / out_ids, numbers
  //            into g_total_all, ascending) and out_sum[q] = the sum of their sums is formed per step;
  //            the product subtracts out_sum[row / A] * ratio_right[row];
  //   inflow:  per adjusted prefix q' the groups with that adjusted prefix are listed with their own
  //            prefix (in_ptr / in_pairs); the product adds, for row q' * A + x,
  //            sum(g) * ratio_right[own prefix * A + x] over the list.
  // Neither kind of term is stored in row_ptr / entries / slices, and the weights of right children are
  // not written at all in regular blocks (Level::plane_blocks) unless a later level reads them
  // (materialize_right) - tapes_export_node_weights fills them in on demand.
  uint64_t n_groups_all = 0;
  double* g_total_all = nullptr;   // [n_groups_all] the levels' g_total, one after the other
  uint64_t* out_ptr = nullptr;     // [A^(k-1) + 1]
  uint32_t* out_ids = nullptr;     // [n_groups_all]
  double* out_sum = nullptr;       // [A^(k-1)], rewritten each step
  uint64_t* in_ptr = nullptr;      // [A^(k-1) + 1]
  uint2* in_pairs = nullptr;       // [n_groups_all] (group number, the group's own prefix), group numbers ascending inside each list
  int materialize_right = 0;       // 1: some level reads stored weights of right children, so all are written
  double* ratio_left = nullptr;    // [n_states] p[i] / max(p[i], marg_{k-1}[i % A^(k-1)]): left extensions / shifts to a full window
  int plane_kernel = 1;            // regular blocks of prefix groups go to plane_kernel (Level::plane_blocks)
  int fuse_marginal_ratio = 1;     // marg_{k-1} and ratio_right from one pass over the table (marginal_ratio_kernel)
  // Small problems: the whole right-hand side in one launch by one thread block or one cluster of
  // thread blocks (engine.cu fused_rhs_kernel).  fused_cluster: 0 = not eligible, else blocks used.
  int fused_small = 1;
  int fused_cluster = 0;
  void* d_fused_levels = nullptr;

  // marginal tables marg_L, L < k, concatenated; marg_off[L] = offset in doubles
  double* marg = nullptr;
  uint64_t marg_off[40];
  uint64_t* d_marg_off = nullptr;
  uint64_t marg_total = 0;

  // staging for the host-buffer entry point
  static constexpr int kCopyBlocks = 8;
  double* d_in = nullptr;
  double* d_out = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_events[kCopyBlocks] = {};
  // small tables: the caller's (pageable) buffers are copied through pinned memory of the model, so
  // that both transfers are plain DMA instead of the driver's staged path
  static constexpr uint64_t kPinnedStagingStates = 1ull << 20;
  double* h_pinned = nullptr;  // [2 * n_states]

  // device copy of the last set of strided sums asked for (sequence observables are evaluated at
  // thousands of output times with the same set)
  std::vector<int64_t> obs_spec;
  int64_t* d_obs_spec = nullptr;
  double* d_obs_out = nullptr;
  size_t obs_capacity = 0;
  double* d_obs_partial = nullptr;  // per-block partial sums of the observables
  size_t obs_partial_capacity = 0;

  BuildStats stats;
  int64_t launches_per_rhs = 0;

  // Every right-hand side writes scratch of the model (rule_w, node_w, marg, the ratio tables,
  // g_total) and the caller picks the stream.  `busy` is recorded on the stream of the most recent
  // call; a call that arrives on another stream waits for it first (begin_use / end_use below), so
  // work of two streams never overlaps on the scratch.
  cudaEvent_t busy = nullptr;
  cudaStream_t last_stream = nullptr;
  bool in_use = false;

  // A problem whose forest would not fit the 31-bit node ids of one structure is built as several
  // structures over disjoint shares of its flux rules (rule_table_part): this one plus `more`, all
  // on this model's stream.  Every right-hand-side entry point below evaluates them one after the
  // other, the later ones adding their flux to the result of the earlier ones (fixed order, so
  // results stay reproducible).  Each part keeps its own marginal tables and node weights.
  std::vector<std::unique_ptr<Model>> more;

  // Small problems are launch-bound (about 20 kernels of a few microseconds per right-hand side):
  // everything that depends on p only through the input pointer - marginal tables, leaf-world
  // probabilities, all forest levels, of all parts - is captured once per input pointer into a
  // CUDA graph and replayed.  The product stays a plain launch (its arguments change per call).
  struct WeightsGraph {
    const double* d_p = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t last_use = 0;
  };
  std::vector<WeightsGraph> weight_graphs;  // at most kMaxWeightGraphs, least recently used replaced
  static constexpr size_t kMaxWeightGraphs = 16;
  static constexpr uint64_t kGraphMaxStates = 1ull << 22;  // larger tables are bandwidth-bound
  uint64_t graph_clock = 0;
  int use_graphs = 1;
  void drop_weight_graphs();  // after a change of launch options

  ~Model();
};

// Frees the scratch memory small builds leave behind for the next build and returns the unused
// part of the library's memory pool to the driver.
void release_build_scratch();

// Stream-ordered allocations (build temporaries, solver vectors) come from a pool of the library's
// own, one per device, so that the process-wide default pool keeps its settings and what this
// library hoards stays bounded: freed memory above the release threshold (2 GiB, TAPES_POOL_KEEP_MB)
// goes back to the driver at the next synchronisation.
cudaMemPool_t library_pool();
void* pool_alloc(size_t bytes, cudaStream_t st);  // throws std::runtime_error

// Ordering between streams that use one model (see Model::busy).  Every entry point that touches
// the model's scratch calls begin_use before its first launch and end_use after its last.
void begin_use(Model& m, cudaStream_t st);
void end_use(Model& m, cudaStream_t st);

// Thrown by build_model when a forest does not fit the 32-bit indices of one structure.
struct TooLarge : std::runtime_error {
  explicit TooLarge(const std::string& what) : std::runtime_error(what) {}
};

// Builds the device structures for a rule table.  Throws std::runtime_error on failure.
std::unique_ptr<Model> build_model(const RuleTable& table, cudaStream_t stream);

// dy/dt for device-resident p (n_states doubles) into device-resident out.  Asynchronous on
// model.stream (or `stream` when non-null).
void rhs_device(Model& m, const double* d_p, double* d_out, cudaStream_t stream);

// Runge-Kutta stage update fused into the product: after dy/dt[i] has been summed, the same thread
// forms  stage[i] = y[i] + h * (sum_j coef[j] * vec[j][i] + coef_self * dy/dt[i])  (terms in this
// order), which is the argument of the next stage.  Saves one pass over the newest stage vector
// and one launch per stage.
struct StageUpdate {
  int n = 0;
  double coef[12];
  const double* vec[12];
  double coef_self = 0.0;
  const double* y = nullptr;
  double* stage = nullptr;
  double h = 0.0;
};
void rhs_device_fused(Model& m, const double* d_p, double* d_out, const StageUpdate& up, cudaStream_t stream);

// The two halves of rhs_device: all p-dependent weights, then the product for a range of states.
void weights_device(Model& m, const double* d_p, cudaStream_t stream);
void flux_rows_device(Model& m, double* d_out, uint64_t row_lo, uint64_t row_hi, cudaStream_t stream);

// Same, with CUDA events recorded on the launching stream between the phases; synchronises and
// returns ms for (marginals + leaf-world probabilities, forest levels, S * w).
void rhs_device_profiled(Model& m, const double* d_p, double* d_out, cudaStream_t stream, float ms[3]);

// Host-buffer variant (the c_compute_dy_dt path): H2D, weights, then the product in row blocks
// with the D2H copy of each block overlapping the product of the next; synchronises.
void rhs_host(Model& m, const double* h_p, double* h_out);

// Number of device kernels launched by one rhs_device call.
int64_t rhs_launch_count(const Model& m);

// flux.cu ---------------------------------------------------------------------------------------
// Encodes m.row_ptr / m.entries as slices (runs need at least min_run_lanes lanes).
void build_flux_slices(Model& m, int min_run_lanes, cudaStream_t st);
// dy/dt for the states row_lo <= i < row_hi from the node weights of the last weights pass; `up`
// (may be null) fuses a Runge-Kutta stage update into the same pass.
// accumulate: dy/dt[i] = d_out[i] + sum instead of the sum alone (later parts of a composite model).
void launch_flux_slices(Model& m, double* d_out, uint64_t row_lo, uint64_t row_hi, cudaStream_t st,
                        const StageUpdate* up, bool accumulate);
// Multi-GPU exchange fused into the product (flux.cu).  Device pointers into the memory of the
// ranks of one NVLink domain, opened through CUDA IPC by the caller.
struct PeerPointers {
  static constexpr int kMax = 16;
  double* ptr[kMax] = {};
};
struct PeerFlags {
  unsigned long long* ptr[PeerPointers::kMax] = {};
};

// The ranks that evaluate one problem together (rule set dealt to the ranks, states owned in
// contiguous blocks of `block` states).  One right-hand side:
//   main stream: weights; then per round c: product of sub-block c of every owner, stored straight
//                into the owner's staging slot `rank` (peer stores), signal "round c landed";
//   side stream: per round c: wait until every rank signalled round c, add the world slots of the
//                own sub-block in rank order, store the sums into every rank's result vector;
//                finally signal / wait "sums landed", and the main stream joins.
// The owner's half of round c runs while the main stream computes round c + 1.  Signals are
// 64-bit epochs in peer-visible flag arrays (release / acquire at system scope); waits time out
// and raise an error flag instead of hanging.
struct PeerGroup {
  static constexpr int kMaxRounds = 16;
  int world = 1, rank = 0, rounds = 4;
  uint64_t block = 0;          // states per owner; a multiple of 32 * rounds
  PeerPointers staging;        // [o]: world * block doubles in rank o's memory
  PeerPointers result;         // [q]: world * block doubles in rank q's memory (full dy/dt)
  PeerFlags flags;             // [q]: 2 * world epochs in rank q's memory: landed[r], summed[r]
  unsigned long long epoch = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int* d_error = nullptr;      // set by a wait that timed out
  ~PeerGroup();
};
PeerGroup* peer_group_create(int world, int rank, uint64_t block, int rounds, void* const* staging,
                             void* const* result, void* const* flags);
// dy/dt of the whole problem into result.ptr[rank] on every rank; asynchronous on `st` (the side
// stream is joined before returning control of `st`).
void peer_rhs(PeerGroup& g, Model& m, const double* d_p, cudaStream_t st);
// Non-zero once a wait timed out (a rank died or fell more than the timeout behind).
int peer_group_error(PeerGroup& g);

// Rebuilds the stored CSR entries (ascending inside each row) from the slices into a device
// buffer of nnz_stored words (single-structure models only).
void expand_flux_slices(Model& m, uint32_t* d_entries, cudaStream_t st);

// Fills in the weights of the right children from the group sums of the last evaluation (they are
// not written per step unless Model::materialize_right); synchronises.
void materialize_node_weights(Model& m);

// The complete flux structure on the HOST in canonical form: row_ptr [n_states + 1], entries [nnz]
// (node id | outflow << 31, ascending inside each row) - the stored entries plus the entries of the
// right children, whose flux the device evaluates per prefix group (Model::out_ptr, in_ptr).
void export_full_csr(Model& m, int64_t* h_row_ptr, uint32_t* h_entries);

}  // namespace tapes
