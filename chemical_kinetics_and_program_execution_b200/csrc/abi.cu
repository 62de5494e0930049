// C ABI (include/tapes_b200.h).  The four drop-in symbols keep the names and signatures the
// reference's ctypes layer binds (framework/markov_tapes.py:40-56).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>

#include "../../include/tapes_b200.h"
#include "engine.h"
#include "hostcopy.h"
#include "integrate.h"
#include "montecarlo.h"
#include "rules.h"
#include "validate.h"

namespace {

struct Runtime {
  int device = -1;
  bool cuda_ok = false;
};

std::string g_error;
Runtime* g_runtime = nullptr;

// What the C ABI hands out as "model": a small record that outlives the structure it names.  The
// structure itself is shared between the cache and whoever still works on it (a live stepper), so
// re-registering a tag or releasing a model invalidates the handle instead of leaving a dangling
// pointer: later calls through it fail with a message, and a stepper created before keeps its
// structure until it is destroyed.  Handles are never freed (a few bytes per build).
struct ModelHandle {
  static constexpr uint64_t kMagic = 0x7461706573423230ull;  // "tapesB20"
  uint64_t magic = kMagic;
  std::shared_ptr<tapes::Model> model;
  std::string what;  // tag and cl_k, for messages
};

// (tag, cl_k, part, n_parts); the whole problem is part 0 of 1
typedef std::tuple<std::string, int, int, int> ModelKey;
std::map<ModelKey, ModelHandle*> g_models;

void fail(const std::string& msg) {
  g_error = msg;
  std::fprintf(stderr, "tapes_b200: %s\n", msg.c_str());
  std::fflush(stderr);
}

bool ensure_cuda() {
  if (g_runtime && g_runtime->cuda_ok) return true;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    fail(std::string("no CUDA device available (") + cudaGetErrorString(err) +
         "); this library has no CPU fallback");
    return false;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (const char* lr = std::getenv("LOCAL_RANK")) dev = std::atoi(lr) % count;
  if (const char* td = std::getenv("TAPES_DEVICE")) dev = std::atoi(td) % count;
  if (cudaSetDevice(dev) != cudaSuccess) { fail("cudaSetDevice failed"); return false; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major < 10) {
    fail(std::string("device '") + prop.name + "' is not sm_100; kernels are built for sm_100a only");
    return false;
  }
  if (!g_runtime) g_runtime = new Runtime();
  g_runtime->device = dev;
  g_runtime->cuda_ok = true;
  return true;
}

void invalidate(ModelHandle* h) {
  if (h->model && h->model->stream) cudaStreamSynchronize(h->model->stream);
  h->model.reset();  // the structure goes once the last stepper working on it is destroyed
}

// The structure behind a handle, or null (with the error set) when the handle is not one of ours
// or its structure was released / its tag registered again.
std::shared_ptr<tapes::Model> resolve(void* handle) {
  ModelHandle* h = (ModelHandle*)handle;
  if (!h) { fail("null model"); return nullptr; }
  if (h->magic != ModelHandle::kMagic) { fail("not a model handle"); return nullptr; }
  if (!h->model) {
    fail("model " + h->what + " was invalidated (its tag was registered again, or it was released); fetch it again with tapes_model");
    return nullptr;
  }
  return h->model;
}

ModelHandle* get_model(const char* tag, int64_t cl_k, int64_t part = 0, int64_t n_parts = 1) {
  tapes::register_builtin_problems();
  if (!tag) { fail("null tag"); return nullptr; }
  if (n_parts < 1 || part < 0 || part >= n_parts) { fail("part must be in 0..n_parts-1"); return nullptr; }
  const ModelKey key(std::string(tag), (int)cl_k, (int)part, (int)n_parts);
  auto it = g_models.find(key);
  if (it != g_models.end()) return it->second;
  const tapes::Problem* prob = tapes::find_problem(tag);
  if (!prob) { fail(std::string("unknown problem tag: ") + tag); return nullptr; }
  if (!ensure_cuda()) return nullptr;
  try {
    auto t0 = std::chrono::steady_clock::now();
    tapes::RuleTable table = tapes::enumerate_rules(*prob, (int)cl_k);
    if (n_parts > 1) table = tapes::rule_table_part(table, (int)part, (int)n_parts);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // node ids have 31 bits: a larger forest is built as several structures over disjoint shares of
    // the rules, evaluated one after the other (engine.h Model::more).  Nodes are estimated from the
    // flux terms (interior nodes add about 1 / (A - 1)); if a structure still comes out too large
    // the split is refined.
    const double a = (double)std::max(table.alphabet, 2);
    double limit = 1.9e9 * (a - 1.0) / a, total = 0.0;
    if (const char* e = std::getenv("TAPES_MAX_PART_TERMS")) limit = std::max(1.0, std::atof(e));
    for (double c : tapes::flux_rule_costs(table)) total += c;
    const int max_pieces = (int)std::max<size_t>(table.rules.size(), 1);
    int pieces = (int)std::min<double>(std::max(1.0, std::ceil(total / limit)), (double)max_pieces);
    std::unique_ptr<tapes::Model> m;
    for (;;) {
      try {
        if (pieces <= 1) {
          m = tapes::build_model(table, nullptr);
        } else {
          m = tapes::build_model(tapes::rule_table_part(table, 0, pieces), nullptr);
          for (int i = 1; i < pieces; ++i)
            m->more.push_back(tapes::build_model(tapes::rule_table_part(table, i, pieces), m->stream));
        }
        break;
      } catch (const tapes::TooLarge&) {
        m.reset();
        if (pieces >= max_pieces) throw;
        pieces = std::min(max_pieces, pieces + std::max(1, pieces / 2));
      }
    }
    m->stats.host_enumerate_ms = ms;
    ModelHandle* h = new ModelHandle();
    h->model = std::shared_ptr<tapes::Model>(m.release());
    h->what = std::string(tag) + " (cl_k = " + std::to_string(cl_k) + ")";
    g_models[key] = h;
    return h;
  } catch (const std::exception& ex) {
    fail(std::string("building ") + tag + " k=" + std::to_string(cl_k) + " failed: " + ex.what());
    return nullptr;
  }
}

// Invalidates the handles of every cached structure of `tag` (cl_k < 0: all window lengths).
int drop_models_of(const std::string& tag, int cl_k) {
  int dropped = 0;
  for (auto it = g_models.begin(); it != g_models.end();) {
    if (std::get<0>(it->first) == tag && (cl_k < 0 || std::get<1>(it->first) == cl_k)) {
      invalidate(it->second);
      it = g_models.erase(it);
      ++dropped;
    } else {
      ++it;
    }
  }
  return dropped;
}

}  // namespace

extern "C" {

void* setup_gambit(void) {
  if (!ensure_cuda()) return nullptr;
  return (void*)g_runtime;
}

void cleanup_gambit(void* handle) {
  (void)handle;
  for (auto& kv : g_models) invalidate(kv.second);
  g_models.clear();
  tapes::staged_copy_shutdown();
  if (g_runtime && g_runtime->cuda_ok) tapes::release_build_scratch();
  if (g_runtime) { delete g_runtime; g_runtime = nullptr; }
}

int64_t c_register_problems(int64_t n) {
  tapes::register_builtin_problems();
  if (!std::getenv("MARKOV_TAPES_QUIET")) {  // the listing of problems.scm:631-638
    std::printf("=== Registered Problems ===\n");
    for (const std::string& tag : tapes::registered_tags()) std::printf("%s\n", tag.c_str());
    std::printf("======\n");
    std::fflush(stdout);
  }
  return n + 1;
}

// The reference has no error channel here: a failing compute-dy/dt never returns (the Scheme error
// handler drops into a REPL, framework/tapes_py_interface.scm:42-44, 81).  The binding of the
// unmodified framework/markov_tapes.py:279-288 hands in a zero-filled result buffer and checks
// nothing, so returning with the buffer untouched would let odeint / solve_ivp integrate dy/dt = 0
// and report a plausible trajectory.  A failed call therefore poisons the result: every entry NaN
// when the table size is known, the first entry when it is not (the buffer has at least one), and
// with TAPES_ABORT_ON_ERROR=1 the process stops like the reference does.
static void poison_result(const char* tag, int64_t cl_k, double* probs_out) {
  if (const char* e = std::getenv("TAPES_ABORT_ON_ERROR"))
    if (std::atoi(e) != 0) std::abort();
  if (!probs_out) return;
  const double nan = std::nan("");
  uint64_t n = 1;
  const tapes::Problem* prob = tag ? tapes::find_problem(tag) : nullptr;
  if (prob && prob->alphabet >= 1 && cl_k >= 1 && cl_k <= 32) {
    for (int64_t i = 0; i < cl_k && n < (1ull << 32); ++i) n *= (uint64_t)prob->alphabet;
    if (n >= (1ull << 32)) n = 1;  // no such table can have been allocated for this library
  }
  for (uint64_t i = 0; i < n; ++i) probs_out[i] = nan;
}

void c_compute_dy_dt(const char* tag, int64_t cl_k, int64_t debug, const double* probs_in,
                     double* probs_out) {
  (void)debug;
  g_error.clear();  // the message describes this call, not an earlier unchecked one
  ModelHandle* h = get_model(tag, cl_k);
  std::shared_ptr<tapes::Model> m = h ? resolve(h) : nullptr;
  if (!m) { poison_result(tag, cl_k, probs_out); return; }
  if (!probs_in || !probs_out) { fail("c_compute_dy_dt: null buffer"); poison_result(tag, cl_k, probs_out); return; }
  try {
    tapes::rhs_host(*m, probs_in, probs_out);
  } catch (const std::exception& ex) {
    fail(ex.what());
    cudaGetLastError();
    poison_result(tag, cl_k, probs_out);
  }
}

void* tapes_host_alloc(int64_t bytes) {
  if (!ensure_cuda()) return nullptr;
  void* p = nullptr;
  if (bytes < 0 || cudaHostAlloc(&p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fail("host_alloc: cudaHostAlloc failed");
    return nullptr;
  }
  return p;
}

int tapes_host_free(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); fail("host_free failed"); return 1; }
  return 0;
}

const char* tapes_last_error(void) { return g_error.c_str(); }
void tapes_clear_error(void) { g_error.clear(); }

int64_t tapes_alphabet_size(const char* tag) {
  tapes::register_builtin_problems();
  const tapes::Problem* p = tapes::find_problem(tag);
  return p ? p->alphabet : -1;
}

int tapes_register_rules(const char* tag, int64_t alphabet, int64_t n_rules, const int32_t* tape,
                         const int32_t* span, const int32_t* catalyst, const int32_t* pattern,
                         const int32_t* replacement, const double* rate, const double* select_weight) {
  try {
    tapes::register_builtin_problems();
    std::vector<tapes::RewriteRule> rules((size_t)n_rules);
    for (int64_t i = 0; i < n_rules; ++i) {
      tapes::RewriteRule& r = rules[(size_t)i];
      r.tape = tape[i]; r.span = span[i]; r.catalyst = catalyst[i];
      for (int j = 0; j < 4; ++j) { r.pattern[j] = pattern[4 * i + j]; r.replacement[j] = replacement[4 * i + j]; }
      r.rate = rate[i]; r.select_weight = select_weight[i];
    }
    tapes::register_problem(tag, (int)alphabet, tapes::body_from_rewrite_rules(std::move(rules)));
    drop_models_of(tag, -1);  // a re-registered tag invalidates cached structures
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_register_program(const char* tag, int64_t alphabet, int64_t n_nodes, const int32_t* kind,
                           const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* first_child,
                           const int32_t* first_weight, int64_t n_children, const int32_t* child,
                           int64_t n_weights, const double* weight) {
  try {
    tapes::register_builtin_problems();
    if (alphabet < 1 || alphabet > 65535) throw std::runtime_error("alphabet size must be in 1..65535");
    if (n_nodes < 1 || n_children < 0 || n_weights < 0) throw std::runtime_error("negative array length");
    tapes::ProgramTree t;
    t.kind.assign(kind, kind + n_nodes); t.a.assign(a, a + n_nodes); t.b.assign(b, b + n_nodes);
    t.c.assign(c, c + n_nodes); t.first_child.assign(first_child, first_child + n_nodes);
    t.first_weight.assign(first_weight, first_weight + n_nodes);
    t.child.assign(child, child + n_children);
    t.weight.assign(weight, weight + n_weights);
    tapes::register_problem(tag, (int)alphabet, tapes::body_from_program(std::move(t), (int)alphabet));
    drop_models_of(tag, -1);  // a re-registered tag invalidates cached structures
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("register_program: ") + ex.what());
    return 1;
  }
}

void* tapes_model(const char* tag, int64_t cl_k) { return (void*)get_model(tag, cl_k); }

void* tapes_model_part(const char* tag, int64_t cl_k, int64_t part, int64_t n_parts) {
  return (void*)get_model(tag, cl_k, part, n_parts);
}

int tapes_release_model(const char* tag, int64_t cl_k) {  // the whole problem and every part of it
  if (!tag) return 1;
  return drop_models_of(tag, (int)cl_k) ? 0 : 1;
}

int64_t tapes_rule_parts(const char* tag, int64_t cl_k, int64_t n_parts, int32_t* owner, double* cost) {
  try {
    tapes::register_builtin_problems();
    const tapes::Problem* prob = tapes::find_problem(tag);
    if (!prob) { fail(std::string("unknown problem tag: ") + tag); return -1; }
    const tapes::RuleTable t = tapes::enumerate_rules(*prob, (int)cl_k);
    const std::vector<double> costs = tapes::flux_rule_costs(t);
    const std::vector<int> owners = tapes::deal_flux_rules(costs, (int)n_parts);
    for (size_t r = 0; r < costs.size(); ++r) {
      if (owner) owner[r] = owners[r];
      if (cost) cost[r] = costs[r];
    }
    return (int64_t)costs.size();
  } catch (const std::exception& ex) {
    fail(ex.what());
    return -1;
  }
}

int tapes_rhs_device(void* model, const double* d_probs_in, double* d_probs_out, void* cuda_stream) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  try {
    tapes::rhs_device(*mp, d_probs_in, d_probs_out, (cudaStream_t)cuda_stream);
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_weights_device(void* model, const double* d_probs_in, void* cuda_stream) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  try {
    tapes::weights_device(*mp, d_probs_in, (cudaStream_t)cuda_stream);
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_flux_rows_device(void* model, double* d_probs_out, int64_t row_lo, int64_t row_hi, void* cuda_stream) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  try {
    tapes::flux_rows_device(*mp, d_probs_out, (uint64_t)row_lo, (uint64_t)row_hi,
                            (cudaStream_t)cuda_stream);
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

void* tapes_peer_alloc(int64_t n_doubles, void* ipc_handle64) {
  if (!ensure_cuda()) return nullptr;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
  void* p = nullptr;
  if (n_doubles < 0 || cudaMalloc(&p, (size_t)(n_doubles > 0 ? n_doubles : 1) * 8) != cudaSuccess) {
    fail("peer_alloc: cudaMalloc failed");
    return nullptr;
  }
  cudaMemset(p, 0, (size_t)(n_doubles > 0 ? n_doubles : 1) * 8);
  cudaIpcMemHandle_t h;
  cudaError_t err = cudaIpcGetMemHandle(&h, p);
  if (err != cudaSuccess) {
    fail(std::string("peer_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(err));
    cudaFree(p);
    return nullptr;
  }
  if (ipc_handle64) std::memcpy(ipc_handle64, &h, 64);
  return p;
}

void* tapes_peer_open(const void* ipc_handle64) {
  if (!ensure_cuda() || !ipc_handle64) return nullptr;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle64, 64);
  void* p = nullptr;
  cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (err != cudaSuccess) {
    fail(std::string("peer_open: cudaIpcOpenMemHandle: ") + cudaGetErrorString(err));
    return nullptr;
  }
  return p;
}

int tapes_peer_close(void* d_ptr) {
  if (d_ptr && cudaIpcCloseMemHandle(d_ptr) != cudaSuccess) { fail("peer_close failed"); return 1; }
  return 0;
}

int tapes_peer_free(void* d_ptr) {
  if (d_ptr && cudaFree(d_ptr) != cudaSuccess) { fail("peer_free failed"); return 1; }
  return 0;
}

void* tapes_peer_group_create(int world, int rank, int64_t block, int rounds, void* const* staging,
                              void* const* result, void* const* flags) {
  if (!ensure_cuda()) return nullptr;
  if (!staging || !result || !flags) { fail("peer_group_create: null pointer table"); return nullptr; }
  try {
    return (void*)tapes::peer_group_create(world, rank, (uint64_t)block, rounds, staging, result, flags);
  } catch (const std::exception& ex) {
    fail(std::string("peer_group_create: ") + ex.what());
    return nullptr;
  }
}

void tapes_peer_group_destroy(void* group) { delete (tapes::PeerGroup*)group; }

int tapes_peer_rhs(void* group, void* model, const double* d_probs_in, void* cuda_stream) {
  if (!group) { fail("null group"); return 1; }
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  try {
    tapes::Model& m = *mp;
    tapes::peer_rhs(*(tapes::PeerGroup*)group, m, d_probs_in, cuda_stream ? (cudaStream_t)cuda_stream : m.stream);
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_peer_group_error(void* group) {
  if (!group) { fail("null group"); return 1; }
  try {
    return tapes::peer_group_error(*(tapes::PeerGroup*)group);
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_rhs_profile(void* model, const double* d_probs_in, double* d_probs_out, void* cuda_stream,
                      double* phase_ms, int capacity) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  try {
    float ms[3] = {0, 0, 0};
    tapes::rhs_device_profiled(*mp, d_probs_in, d_probs_out, (cudaStream_t)cuda_stream, ms);
    for (int i = 0; i < 3 && i < capacity; ++i) phase_ms[i] = ms[i];
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_sync(void* model) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  cudaError_t err = cudaStreamSynchronize(mp->stream);
  if (err != cudaSuccess) { fail(cudaGetErrorString(err)); return 1; }
  return 0;
}

int tapes_model_info(void* model, int64_t* out, int capacity) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 0;
  const tapes::Model& head = *mp;
  const int kFields = 38;
  // sizes add up over the parts of a composite model; facts shared by all parts come from the first
  static const bool adds[kFields] = {false, true, true, true, false, false, true, true, false, false, true, true,
                                     true, false, false, false, false, false, true, true, true, true,
                                     true, false, false, true, true, false, true, true, true, true,
                                     true, false, true, false, true, true};
  int64_t total[kFields] = {};
  for (size_t part = 0; part <= head.more.size(); ++part) {
    const tapes::Model& m = part == 0 ? head : *head.more[part - 1];
    int64_t interleaved = 0;
    int64_t distinct_reads = 0;  // table / ratio entries a level can read at most once each: min(nodes of the level, states)
    for (const tapes::Level& lv : m.levels) {
      interleaved += lv.block_order ? 1 : 0;
      const uint64_t level_nodes = (uint64_t)lv.n_roots + ((uint64_t)lv.n_left + lv.n_groups) * (uint64_t)m.A;
      distinct_reads += (int64_t)std::min<uint64_t>(level_nodes, m.n_states);
    }
    const int64_t v[kFields] = {(int64_t)m.n_states, (int64_t)m.n_nodes, (int64_t)m.nnz, (int64_t)m.n_rules,
                                (int64_t)m.levels.size(), 0, m.stats.terms, m.stats.sum_nodes,
                                m.stats.worlds_walked, m.stats.leaf_worlds, m.stats.seeds, m.stats.hash_inserts,
                                m.stats.hash_unique, (int64_t)m.A, (int64_t)m.k, (int64_t)m.spmv_group,
                                (int64_t)m.flux_format, (int64_t)m.slices.n_slices, (int64_t)m.slices.n_words,
                                (int64_t)m.slices.runs, (int64_t)m.slices.run_entries, (int64_t)m.slices.column_entries,
                                (int64_t)m.slices.column_slots, (int64_t)m.slices.min_run_lanes, (int64_t)m.level_unroll,
                                m.stats.irregular_levels, m.stats.left_parents, (int64_t)m.flux_unroll,
                                m.stats.owned_parents, m.stats.deferred_groups, 1, interleaved,
                                m.stats.plane_groups, (int64_t)((m.ratio_right ? 1 : 0) + (m.ratio_left ? 1 : 0)),
                                (int64_t)m.nnz_stored, (int64_t)m.materialize_right, m.stats.hash_retries,
                                distinct_reads};
    for (int i = 0; i < kFields; ++i) {
      if (part == 0) total[i] = v[i];
      else if (adds[i]) total[i] += v[i];
    }
    if (v[4] > total[4]) total[4] = v[4];  // forest levels: the deepest part
  }
  total[5] = tapes::rhs_launch_count(head);  // kernel launches of one right-hand side, all parts
  int n = kFields;
  if (n > capacity) n = capacity;
  for (int i = 0; i < n; ++i) out[i] = total[i];
  return n;
}

int tapes_model_set(void* model, const char* key, int64_t value) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  tapes::Model& head = *mp;
  int tapes::Model::*field = nullptr;
  if (std::strcmp(key, "spmv_lanes") == 0 && (value == 1 || value == 2 || value == 4 || value == 8 || value == 16))
    field = &tapes::Model::spmv_group;
  if (std::strcmp(key, "flux_unroll") == 0 && (value == 2 || value == 3 || value == 4 || value == 6 || value == 8))
    field = &tapes::Model::flux_unroll;
  if (std::strcmp(key, "level_unroll") == 0 && value >= 1 && value <= 8) field = &tapes::Model::level_unroll;
  if (std::strcmp(key, "interleave_seeds") == 0 && (value == 0 || value == 1)) field = &tapes::Model::interleave_seeds;
  if (std::strcmp(key, "ratio_table") == 0 && (value == 0 || value == 1)) field = &tapes::Model::ratio_table;
  if (std::strcmp(key, "plane_kernel") == 0 && (value == 0 || value == 1)) field = &tapes::Model::plane_kernel;
  if (std::strcmp(key, "fuse_marginal_ratio") == 0 && (value == 0 || value == 1)) field = &tapes::Model::fuse_marginal_ratio;
  if (std::strcmp(key, "fused_small") == 0 && (value == 0 || value == 1)) field = &tapes::Model::fused_small;
  if (std::strcmp(key, "fused_cluster") == 0 && (value == 0 || value == 1 || value == 2 || value == 4 || value == 8 || value == 16)) {
    if (head.flux_format != 1 || head.n_rules == 0) { fail("fused_cluster: this model has no single-launch form"); return 1; }
    field = &tapes::Model::fused_cluster;
  }
  if (std::strcmp(key, "graphs") == 0 && (value == 0 || value == 1)) {
    cudaStreamSynchronize(head.stream);
    head.drop_weight_graphs();
    head.use_graphs = (int)value;
    return 0;
  }
  if (field) {
    cudaStreamSynchronize(head.stream);
    head.drop_weight_graphs();  // captured launches carry the old setting
    head.*field = (int)value;
    for (auto& part : head.more) (*part).*field = (int)value;
    head.launches_per_rhs = tapes::rhs_launch_count(head);
    return 0;
  }
  fail(std::string("unknown option or value: ") + key);
  return 1;
}

int tapes_model_timing(void* model, double* out, int capacity) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 0;
  const tapes::Model& head = *mp;
  double v[5] = {0, 0, 0, 0, 0};
  for (size_t part = 0; part <= head.more.size(); ++part) {  // the parts of a composite model are built in turn
    const tapes::Model& m = part == 0 ? head : *head.more[part - 1];
    v[0] += m.stats.host_enumerate_ms; v[1] += m.stats.device_expand_ms; v[2] += m.stats.device_csr_ms;
    v[3] += m.stats.device_slices_ms; v[4] += m.stats.expand_alloc_ms;
  }
  int n = 5 > capacity ? capacity : 5;
  for (int i = 0; i < n; ++i) out[i] = v[i];
  return n;
}

int tapes_export_csr(void* model, int64_t* row_ptr, uint32_t* entries) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  tapes::Model& m = *mp;
  if (!m.more.empty()) { fail("export_csr: composite model (forest above the 31-bit node ids); export shares made with tapes_model_part instead"); return 1; }
  try {
    tapes::export_full_csr(m, row_ptr, entries);
  } catch (const std::exception& ex) {
    cudaGetLastError();
    fail(std::string("export_csr: ") + ex.what());
    return 1;
  }
  return 0;
}

int tapes_export_node_weights(void* model, double* weights) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  tapes::Model& m = *mp;
  if (!m.more.empty()) { fail("export_node_weights: composite model; export shares made with tapes_model_part instead"); return 1; }
  try {
    tapes::materialize_node_weights(m);  // right children are not written per step
  } catch (const std::exception& ex) {
    fail(std::string("export_node_weights: ") + ex.what());
    return 1;
  }
  if (m.n_nodes && cudaMemcpy(weights, m.node_w, m.n_nodes * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
    fail("export_node_weights: copy failed");
    return 1;
  }
  return 0;
}

void* tapes_dop853_create_peer(void* model, void* group, const double* tableau, const double* y0, double t0,
                               double t_bound, double rtol, double atol, double max_step, double first_step);

void* tapes_dop853_create(void* model, const double* tableau, const double* y0, double t0, double t_bound,
                          double rtol, double atol, double max_step, double first_step) {
  return tapes_dop853_create_peer(model, nullptr, tableau, y0, t0, t_bound, rtol, atol, max_step, first_step);
}

void* tapes_dop853_create_peer(void* model, void* group, const double* tableau, const double* y0, double t0,
                               double t_bound, double rtol, double atol, double max_step, double first_step) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return nullptr;
  try {
    tapes::Dop853Tableau tab;
    const double* p = tableau;
    std::memcpy(tab.A, p, sizeof(tab.A)); p += 256;
    std::memcpy(tab.B, p, sizeof(tab.B)); p += 12;
    std::memcpy(tab.C, p, sizeof(tab.C)); p += 16;
    std::memcpy(tab.E3, p, sizeof(tab.E3)); p += 13;
    std::memcpy(tab.E5, p, sizeof(tab.E5)); p += 13;
    std::memcpy(tab.D, p, sizeof(tab.D));
    return (void*)tapes::dop853_create(mp, (tapes::PeerGroup*)group, tab, y0, t0, t_bound, rtol,
                                       atol, max_step, first_step);
  } catch (const std::exception& ex) {
    fail(ex.what());
    return nullptr;
  }
}

void tapes_dop853_destroy(void* solver) { tapes::dop853_destroy((tapes::Dop853*)solver); }

int tapes_dop853_step(void* solver) {
  if (!solver) { fail("null solver"); return -2; }
  try {
    return tapes::dop853_step((tapes::Dop853*)solver);
  } catch (const std::exception& ex) {
    fail(ex.what());
    return -2;
  }
}

int tapes_dop853_dense(void* solver, double t) {
  if (!solver) { fail("null solver"); return 1; }
  try {
    tapes::Dop853* s = (tapes::Dop853*)solver;
    tapes::dop853_dense_eval(s, t, tapes::dop853_dense_buffer(s));
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_dop853_fetch(void* solver, int which, double* out) {
  if (!solver) { fail("null solver"); return 1; }
  tapes::Dop853* s = (tapes::Dop853*)solver;
  const double* src = which == 0 ? tapes::dop853_state(s) : tapes::dop853_dense_buffer(s);
  tapes::Model& m = tapes::dop853_model(s);
  if (cudaMemcpyAsync(out, src, m.n_states * 8, cudaMemcpyDeviceToHost, m.stream) != cudaSuccess ||
      cudaStreamSynchronize(m.stream) != cudaSuccess) {
    fail("dop853_fetch: copy failed");
    return 1;
  }
  return 0;
}

int tapes_dop853_observe(void* solver, int which, const int64_t* offset, const int64_t* stride,
                         const int64_t* count, int64_t n_obs, double* out) {
  if (!solver) { fail("null solver"); return 1; }
  try {
    tapes::Dop853* s = (tapes::Dop853*)solver;
    const double* src = which == 0 ? tapes::dop853_state(s) : tapes::dop853_dense_buffer(s);
    tapes::observe_strided(tapes::dop853_model(s), src, offset, stride, count, n_obs, out);
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_dop853_info(void* solver, double* out6) {
  if (!solver) { fail("null solver"); return 1; }
  tapes::dop853_info((tapes::Dop853*)solver, out6);
  return 0;
}

int tapes_observe(void* model, const double* d_y, const int64_t* offset, const int64_t* stride,
                  const int64_t* count, int64_t n_obs, double* out) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  try {
    tapes::observe_strided(*mp, d_y, offset, stride, count, n_obs, out);
    return 0;
  } catch (const std::exception& ex) {
    fail(ex.what());
    return 1;
  }
}

int tapes_observe_sequences(void* model, const double* d_y, int64_t n_seq, const int64_t* seq_ptr,
                            const int32_t* symbols, double eps, double* out) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  if (!d_y || !seq_ptr || !symbols || !out) { fail("observe_sequences: null argument"); return 1; }
  try {
    tapes::observe_sequences(*mp, d_y, n_seq, seq_ptr, symbols, eps, out);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("observe_sequences: ") + ex.what());
    return 1;
  }
}

int tapes_dop853_observe_sequences(void* solver, int which, int64_t n_seq, const int64_t* seq_ptr,
                                   const int32_t* symbols, double eps, double* out) {
  if (!solver) { fail("null solver"); return 1; }
  if (!seq_ptr || !symbols || !out) { fail("observe_sequences: null argument"); return 1; }
  try {
    tapes::Dop853* s = (tapes::Dop853*)solver;
    const double* src = which == 0 ? tapes::dop853_state(s) : tapes::dop853_dense_buffer(s);
    tapes::observe_sequences(tapes::dop853_model(s), src, n_seq, seq_ptr, symbols, eps, out);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("observe_sequences: ") + ex.what());
    return 1;
  }
}

int tapes_markov_entropy(void* model, const double* d_y, double* out) {
  std::shared_ptr<tapes::Model> mp = resolve(model);
  if (!mp) return 1;
  if (!d_y || !out) { fail("markov_entropy: null argument"); return 1; }
  try {
    *out = tapes::markov_entropy(*mp, d_y);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("markov_entropy: ") + ex.what());
    return 1;
  }
}

int tapes_dop853_entropy(void* solver, int which, double* out) {
  if (!solver || !out) { fail("null solver"); return 1; }
  try {
    tapes::Dop853* s = (tapes::Dop853*)solver;
    *out = tapes::markov_entropy(tapes::dop853_model(s), which == 0 ? tapes::dop853_state(s) : tapes::dop853_dense_buffer(s));
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("markov_entropy: ") + ex.what());
    return 1;
  }
}

int tapes_check_table(int64_t alphabet, int64_t cl_k, const double* probs, int on_device, double eps_mpp,
                      int64_t max_iterations, double tolerance, double* out6) {
  if (!probs || !out6) { fail("check_table: null argument"); return 1; }
  if (!ensure_cuda()) return 1;
  double* staged = nullptr;
  try {
    if (alphabet < 1 || cl_k < 2 || cl_k > 32) throw std::runtime_error("needs alphabet >= 1 and 2 <= cl_k <= 32");
    double states = 1;
    for (int64_t i = 0; i < cl_k; ++i) states *= (double)alphabet;
    if (states >= 4294967296.0) throw std::runtime_error("A^cl_k must be below 2^32");
    const double* d_p = probs;
    if (!on_device) {
      const size_t bytes = (size_t)states * sizeof(double);
      if (cudaMalloc((void**)&staged, bytes) != cudaSuccess) throw std::runtime_error("out of device memory");
      if (cudaMemcpy(staged, probs, bytes, cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("copy failed");
      d_p = staged;
    }
    const tapes::TableCheck c = tapes::check_table((int)alphabet, (int)cl_k, d_p, eps_mpp, (int)max_iterations,
                                                   tolerance, nullptr);
    out6[0] = c.total; out6[1] = c.marginal_distance; out6[2] = c.stationarity_residual;
    out6[3] = c.power_distance; out6[4] = c.last_change; out6[5] = (double)c.iterations;
    if (staged) cudaFree(staged);
    return 0;
  } catch (const std::exception& ex) {
    if (staged) cudaFree(staged);
    fail(std::string("check_table: ") + ex.what());
    return 1;
  }
}

void* tapes_mc_create(const char* tag, int64_t n_sites, const uint8_t* tape0, int64_t events_per_substep,
                      uint64_t seed) {
  try {
    tapes::register_builtin_problems();
    const tapes::Problem* prob = tapes::find_problem(tag);
    if (!prob) throw std::runtime_error(std::string("unknown problem tag: ") + tag);
    if (!tape0 || n_sites < 1 || events_per_substep < 1) throw std::runtime_error("bad ring or event count");
    if (!ensure_cuda()) return nullptr;
    const tapes::ProgramTree tree = tapes::trace_body(prob->body, prob->alphabet);
    return (void*)tapes::mc_create(tree, prob->alphabet, (uint64_t)n_sites, tape0, (uint32_t)std::min<int64_t>(events_per_substep, 1 << 20),
                                   seed);
  } catch (const std::exception& ex) {
    fail(std::string("mc_create: ") + ex.what());
    return nullptr;
  }
}

void tapes_mc_destroy(void* mc) { tapes::mc_destroy((tapes::MonteCarlo*)mc); }

int tapes_mc_run(void* mc, int64_t n_substeps) {
  if (!mc || n_substeps < 0) { fail("mc_run: bad argument"); return 1; }
  try {
    tapes::mc_run((tapes::MonteCarlo*)mc, (uint64_t)n_substeps);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("mc_run: ") + ex.what());
    return 1;
  }
}

int tapes_mc_window_counts(void* mc, int64_t cl_k, int64_t* counts) {
  if (!mc || !counts) { fail("mc_window_counts: null argument"); return 1; }
  try {
    tapes::mc_window_counts((tapes::MonteCarlo*)mc, (int)cl_k, counts);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("mc_window_counts: ") + ex.what());
    return 1;
  }
}

int tapes_mc_fetch(void* mc, uint8_t* tape) {
  if (!mc || !tape) { fail("mc_fetch: null argument"); return 1; }
  try {
    tapes::mc_fetch((tapes::MonteCarlo*)mc, tape);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("mc_fetch: ") + ex.what());
    return 1;
  }
}

int tapes_mc_sample_ring(int64_t alphabet, int64_t cl_k, const double* table, int64_t n_sites, uint64_t seed,
                         uint8_t* tape) {
  if (!table || !tape || n_sites < 1) { fail("mc_sample_ring: bad argument"); return 1; }
  try {
    tapes::mc_sample_ring((int)alphabet, (int)cl_k, table, (uint64_t)n_sites, seed, tape);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("mc_sample_ring: ") + ex.what());
    return 1;
  }
}

int tapes_mc_ferromagnet_chains(int64_t n_trials, int64_t chain_length, int64_t n_steps, int64_t trials_per_step,
                                const uint8_t* chain0, const int32_t* sites, const double* uniforms, const double* accept,
                                double* counts) {
  if (!chain0 || !accept || !counts || (n_steps > 1 && (!sites || !uniforms))) { fail("mc_ferromagnet_chains: null argument"); return 1; }
  if (!ensure_cuda()) return 1;
  try {
    tapes::mc_ferromagnet_chains(n_trials, chain_length, n_steps, trials_per_step, chain0, sites, uniforms, accept, counts);
    return 0;
  } catch (const std::exception& ex) {
    fail(std::string("mc_ferromagnet_chains: ") + ex.what());
    return 1;
  }
}

int64_t tapes_program_tree(const char* tag, int64_t* sizes3, int32_t* kind, int32_t* a, int32_t* b, int32_t* c,
                           int32_t* first_child, int32_t* first_weight, int32_t* child, double* weight) {
  try {
    tapes::register_builtin_problems();
    const tapes::Problem* prob = tapes::find_problem(tag);
    if (!prob) throw std::runtime_error(std::string("unknown problem tag: ") + tag);
    const tapes::ProgramTree t = tapes::trace_body(prob->body, prob->alphabet);
    if (sizes3) { sizes3[0] = (int64_t)t.kind.size(); sizes3[1] = (int64_t)t.child.size(); sizes3[2] = (int64_t)t.weight.size(); }
    if (kind) {
      std::copy(t.kind.begin(), t.kind.end(), kind); std::copy(t.a.begin(), t.a.end(), a);
      std::copy(t.b.begin(), t.b.end(), b); std::copy(t.c.begin(), t.c.end(), c);
      std::copy(t.first_child.begin(), t.first_child.end(), first_child);
      std::copy(t.first_weight.begin(), t.first_weight.end(), first_weight);
      std::copy(t.child.begin(), t.child.end(), child); std::copy(t.weight.begin(), t.weight.end(), weight);
    }
    return (int64_t)t.kind.size();
  } catch (const std::exception& ex) {
    fail(std::string("program_tree: ") + ex.what());
    return -1;
  }
}

int64_t tapes_rule_table(const char* tag, int64_t cl_k, int64_t* n_steps, int64_t* rule_ptr,
                         int32_t* step_kind, int32_t* step_len, int64_t* step_long, int64_t* step_short,
                         double* step_prob, int32_t* seed_len, uint64_t* seed_orig, uint64_t* seed_adj,
                         int64_t* walk_stats) {
  try {
    tapes::register_builtin_problems();
    const tapes::Problem* prob = tapes::find_problem(tag);
    if (!prob) { fail(std::string("unknown problem tag: ") + tag); return -1; }
    tapes::RuleTable t = tapes::enumerate_rules(*prob, (int)cl_k);
    int64_t steps = 0;
    for (auto& r : t.rules) steps += (int64_t)r.steps.size();
    if (n_steps) *n_steps = steps;
    if (walk_stats) { walk_stats[0] = t.worlds_walked; walk_stats[1] = t.leaf_worlds; }
    if (rule_ptr) {
      int64_t s = 0;
      for (size_t i = 0; i < t.rules.size(); ++i) {
        rule_ptr[i] = s;
        for (const tapes::Step& st : t.rules[i].steps) {
          step_kind[s] = st.kind; step_len[s] = st.length; step_long[s] = st.long_index;
          step_short[s] = st.short_index; step_prob[s] = st.prob;
          ++s;
        }
        for (int tp = 0; tp < 2; ++tp) {
          seed_len[2 * i + tp] = t.rules[i].tape[tp].length;
          seed_orig[2 * i + tp] = t.rules[i].tape[tp].orig;
          seed_adj[2 * i + tp] = t.rules[i].tape[tp].adjusted;
        }
      }
      rule_ptr[t.rules.size()] = s;
    }
    return (int64_t)t.rules.size();
  } catch (const std::exception& ex) {
    fail(ex.what());
    return -1;
  }
}

}  // extern "C"
