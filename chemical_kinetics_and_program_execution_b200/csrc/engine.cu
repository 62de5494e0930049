// Device engine (sm_100a): frontier expansion of the window-extension forest, CSR assembly of the
// flux structure, and the per-step evaluation  dy/dt = S * w(p).
//
// Reference behaviour being reproduced (framework/tape_multiverse.scm):
//   lr-rec-extend-1 1249-1401   window extension, one node per (seed, window, context)
//   get-prob-relative 1263-1269 per-node ratio  p_long / max(p_long, p_short), 0 when p_long == 0
//   accumulate-dp/dt 1271-1301  -w at the original window index, +w at the adjusted one
//   sp-table-marginal 362-385   last-axis marginals, sequential sums
//   mv-state-unfold-for-tape-get 482-588, -choose 594-626   leaf-world probabilities
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "flux_device.cuh"
#include "hostcopy.h"
#include "primitives.cuh"

namespace tapes {

namespace {

constexpr uint8_t FL_TERM = 1;   // contributes a flux term (window is full)
constexpr uint8_t FL_LEFT = 2;   // expands leftwards
constexpr uint8_t FL_RIGHT = 4;  // hands its weight to the right chain
constexpr uint32_t kNoRank = 0xffffffffu;
const double* const kLinkLater = reinterpret_cast<const double*>(uintptr_t(8));  // Level::prev_total until the sums exist
constexpr uint32_t kOutflowBit = 0x80000000u;
constexpr int kThreads = 256;

template <typename T>
T* dalloc(size_t n, cudaStream_t st) {
  return (T*)pool_alloc(std::max<size_t>(n, 1) * sizeof(T), st);
}
template <typename T>
void dfree(T*& p, cudaStream_t st) {
  if (p) cudaFreeAsync((void*)p, st);
  p = nullptr;
}
double g_alloc_ms = 0;  // host time spent in cudaMalloc / cudaFree during the current build
struct AllocTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~AllocTimer() { g_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

template <typename T>
T* dkeep(Model& m, size_t n) {  // arrays that live as long as the model
  AllocTimer timer;
  return m.arena.array<T>(n);
}

template <typename T>
T* dtemp(size_t n) {  // large temporaries handed back with cudaFree
  AllocTimer timer;
  void* p = nullptr;
  TAPES_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return (T*)p;
}

// Scratch for one phase of a level: sizes are declared first, then one cudaMalloc backs them all.
// The allocation is kept and reused by later levels as long as it is large enough (the deep
// levels of a forest have similar sizes), so the driver is asked for memory a handful of times.
struct Slab {
  char* base = nullptr;
  size_t capacity = 0, bytes = 0;
  std::vector<size_t> offsets;
  Slab() {}
  Slab(const Slab&) = delete;
  Slab& operator=(const Slab&) = delete;
  ~Slab() { release(); }
  void swap(Slab& o) {
    std::swap(base, o.base); std::swap(capacity, o.capacity); std::swap(bytes, o.bytes);
    offsets.swap(o.offsets);
  }
  void reset() { bytes = 0; offsets.clear(); }  // forget the layout, keep the memory
  size_t want(size_t b) {
    bytes = (bytes + 255) & ~(size_t)255;
    offsets.push_back(bytes);
    bytes += std::max<size_t>(b, 1);
    return offsets.size() - 1;
  }
  void commit() {
    if (bytes <= capacity) return;
    AllocTimer timer;
    if (base) cudaFree(base);
    base = nullptr; capacity = 0;
    void* p = nullptr;
    TAPES_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
    base = (char*)p;
    capacity = std::max<size_t>(bytes, 256);
  }
  template <typename T>
  T* at(size_t i) const { return (T*)(base + offsets[i]); }
  void release() {
    AllocTimer timer;
    if (base) cudaFree(base);
    base = nullptr; capacity = 0; reset();
  }
};

struct Consts {
  uint32_t A;
  int k;
  uint32_t M;       // A^(k-1)
  uint32_t pw[33];  // pw[i] = A^i for i < k (fits u32 since A^k < 2^32)
};

// Build-time frontier: the nodes of one level with everything needed to expand them.
struct Frontier {
  uint64_t n = 0;
  uint32_t* io = nullptr;    // original window index
  uint32_t* ia = nullptr;    // adjusted window index
  uint32_t* seed = nullptr;  // which (leaf world, tape) the node descends from
  uint8_t* meta = nullptr;   // kind << 6 | window length
  uint8_t* flags = nullptr;  // FL_*
  // A level that consists of the right children of prefix groups only (every level after the last left
  // extension: right children extend to the right only) is not materialised: node i is child i % A of group
  // i / A, with the window indices prefix * A + x and adjusted prefix * A + x, full length, FL_TERM | FL_RIGHT.
  bool right_only = false;
  const uint32_t* g_prefix = nullptr;    // [n / A] of this level's groups
  const uint32_t* g_adjusted = nullptr;
  const uint32_t* g_seed = nullptr;
};

// ---------------------------------------------------------------------------------------------
// Expansion pass 1: classify every frontier node and hash-insert right-chain prefixes.
// ---------------------------------------------------------------------------------------------
template <bool VOTE, bool RIGHT_ONLY>
__global__ void __launch_bounds__(kThreads) classify_kernel(Frontier f, Consts c, HashSet hs,
                                                            uint32_t* __restrict__ ltflag,
                                                            uint32_t* __restrict__ keyslot, uint64_t* __restrict__ unique,
                                                            unsigned long long* __restrict__ n_unique) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < f.n;
  bool haskey = false;
  uint64_t key = 0;
  uint32_t pa = 0;
  if (valid && RIGHT_ONLY) {  // no left children, no stored terms: nothing to rank (ltflag is not written)
    const uint32_t g = (uint32_t)i / c.A, x = (uint32_t)i - g * c.A;
    const uint32_t po = (f.g_prefix[g] * c.A + x) % c.M;
    pa = (f.g_adjusted[g] * c.A + x) % c.M;
    haskey = po != pa;
    key = ((uint64_t)f.g_seed[g] << 32) | po;
  } else if (valid) {
    const uint8_t meta = f.meta[i], fl = f.flags[i];
    const int len = meta & 63;
    const uint32_t io = f.io[i], ia = f.ia[i];
    uint32_t left = 0;
    if (fl & FL_LEFT) {
      if (len < c.k) left = 1;                        // tm.scm:1340-1357
      else if (io / c.A != ia / c.A) left = 1;        // tm.scm:1358-1379, entry test 1331
    }
    if (fl & FL_RIGHT) {                              // tm.scm:1393-1397 / 1319-1322
      const uint32_t po = io % c.M;
      pa = ia % c.M;
      haskey = po != pa;                              // tm.scm:1308-1309
      key = ((uint64_t)f.seed[i] << 32) | po;
    }
    // bit 0: the node has left children; bit 1: its flux term is stored as a pair of edges (outflow,
    // inflow).  The terms of right children are not stored per node at all - their flux is evaluated
    // per prefix group (Model::out_ptr, in_ptr).  One scan ranks both (exclusive_scan_u32<true>).
    ltflag[i] = left | (((fl & FL_TERM) && (meta >> 6) != NODE_RIGHT) ? 2u : 0u);
  }
  // keys seen for the first time are compacted into the new frontier's group list as they are
  // inserted (block scan, one atomic per block); the list is put in canonical order afterwards
  uint32_t slot = 0;
  const bool fresh = VOTE ? hash_insert_warp(hs, haskey, key, pa, &slot) : hash_insert_lane(hs, haskey, key, pa, &slot);
  if (valid) keyslot[i] = haskey ? slot : kNoRank;  // emit_kernel finds the key's group without probing
  block_append_u64(fresh, key, unique, n_unique);
}

// ---------------------------------------------------------------------------------------------
// Expansion pass 2: write the left children, the parent records the per-step evaluation reads,
// the flux edges, and the prefix-group rank of every node that feeds the right chain.  A left
// extension adds the MOST significant digit, so left children are stored digit-major
// (x * n_left_parents + parent rank): consecutive nodes then have consecutive table indices.
// ---------------------------------------------------------------------------------------------
__global__ void emit_kernel(Frontier f, Consts c, uint64_t cur_base, const uint32_t* __restrict__ ltflag,
                            const uint32_t* __restrict__ keyslot, const uint64_t* __restrict__ ltrank,
                            uint64_t n_left_parents, HashSet hs, Frontier next, uint32_t* __restrict__ lp_gid,
                            uint32_t* __restrict__ lp_io, uint8_t* __restrict__ lp_len,
                            uint32_t* __restrict__ edge_row, uint32_t* __restrict__ edge_val,
                            uint32_t* __restrict__ keyrank, uint4* __restrict__ gstat,
                            unsigned long long* __restrict__ chain_starts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f.n) return;
  const uint8_t meta = f.meta[i];
  const int len = meta & 63;
  const uint32_t io = f.io[i], ia = f.ia[i], seed = f.seed[i];
  const uint32_t gid = (uint32_t)(cur_base + i);
  const uint32_t lt = ltflag[i];
  {  // left children of window length k - 1 start right chains: their number bounds the prefixes they add
    // to the next level, which sizes that level's table (build, pass 1)
    const unsigned active = __activemask();
    const unsigned starters = __ballot_sync(active, (lt & 1u) && len == c.k - 2);
    if (starters && (threadIdx.x & 31) == (unsigned)(__ffs(active) - 1))
      atomicAdd(chain_starts, (unsigned long long)__popc(starters) * c.A);
  }
  const uint64_t ranks = lt ? ltrank[i] : 0ull;  // low half: rank among the left parents, high half: among the stored terms
  if (lt & 1u) {
    const uint64_t r = (uint32_t)ranks;
    uint32_t bo, ba, step;
    int nl;
    uint8_t nfl;
    if (len < c.k) {                   // left extension: x * A^len + index
      bo = io; ba = ia; step = c.pw[len];
      nl = len + 1;
      nfl = FL_LEFT | (nl == c.k ? FL_TERM : 0) | (nl == c.k - 1 ? FL_RIGHT : 0);
    } else {                           // left shift of a full window: x * A^(k-1) + index / A
      bo = io / c.A; ba = ia / c.A; step = c.M;
      nl = c.k;
      nfl = FL_TERM | FL_LEFT;
    }
    lp_gid[r] = gid;
    lp_io[r] = bo;
    lp_len[r] = (uint8_t)nl;
    const uint8_t nmeta = (uint8_t)((NODE_LEFT << 6) | nl);
    for (uint32_t x = 0; x < c.A; ++x) {
      const uint64_t ci = (uint64_t)x * n_left_parents + r;
      next.io[ci] = bo + x * step;
      next.ia[ci] = ba + x * step;
      next.seed[ci] = seed;
      next.meta[ci] = nmeta;
      next.flags[ci] = nfl;
    }
  }
  if (lt & 2u) {
    const uint64_t e = 2 * (ranks >> 32);
    edge_row[e] = io;     edge_val[e] = gid | kOutflowBit;   // -w at the original window
    edge_row[e + 1] = ia; edge_val[e + 1] = gid;             // +w at the adjusted window
  }
  const uint32_t slot = keyslot[i];
  if (slot != kNoRank) {
    // the node is a parent of prefix group g: parent lists are almost always arithmetic progressions
    // of node ids, which smallest id, largest id and count describe completely
    const uint32_t g = hs.ranks[slot];
    keyrank[i] = g;
    // one 16-byte record per group, cleared to all ones: smallest id, complement of the largest id,
    // complement of the count - the three updates of a parent land in one sector
    atomicMin(&gstat[g].x, gid);
    atomicMin(&gstat[g].y, ~gid);
    atomicSub(&gstat[g].z, 1u);
  } else {
    keyrank[i] = kNoRank;
  }
}

// emit_kernel for a level of right children only: nothing but the group each node is a parent of.
__global__ void emit_right_only_kernel(uint64_t n, uint64_t cur_base, const uint32_t* __restrict__ keyslot, HashSet hs,
                                       uint32_t* __restrict__ keyrank, uint4* __restrict__ gstat) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = keyslot[i];
  if (slot != kNoRank) {
    const uint32_t g = hs.ranks[slot], gid = (uint32_t)(cur_base + i);
    keyrank[i] = g;
    atomicMin(&gstat[g].x, gid);
    atomicMin(&gstat[g].y, ~gid);
    atomicSub(&gstat[g].z, 1u);
  } else {
    keyrank[i] = kNoRank;
  }
}

// (first, stride, count) of every group from the extremes of its parent ids; a group whose extremes
// cannot belong to a progression of `count` ids is counted as irregular.  facts: [1] irregular
// groups, [2] parents of all groups.
__global__ void derive_progressions_kernel(const uint4* __restrict__ gstat, uint64_t n_groups,
                                           uint32_t* __restrict__ first, uint32_t* __restrict__ stride,
                                           uint32_t* __restrict__ count, unsigned long long* __restrict__ facts) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = 0;
  bool bad = false;
  if (g < n_groups) {
    const uint4 st = gstat[g];
    n = ~st.z;
    const uint32_t lo = n ? st.x : 0u, span = n ? ~st.y - lo : 0u;
    const uint32_t d = n > 1 ? span / (n - 1) : 0u;
    bad = n > 1 && (d == 0 || span % (n - 1) != 0);
    first[g] = lo; stride[g] = d; count[g] = n;
  }
  block_add(n, &facts[2]);
  block_add(bad ? 1ull : 0ull, &facts[1]);
}

// Every parent must sit on its group's progression; together with distinct ids, matching extremes
// and the count this makes the parent set exactly {first + j * stride : j < count}.
__global__ void verify_progressions_kernel(const uint32_t* __restrict__ keyrank, uint64_t n, uint64_t id_base,
                                           const uint32_t* __restrict__ first, const uint32_t* __restrict__ stride,
                                           unsigned long long* __restrict__ facts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (i < n && keyrank[i] != kNoRank) {
    const uint32_t g = keyrank[i], id = (uint32_t)(id_base + i);
    const uint32_t d = stride[g], off = id - first[g];
    bad = d ? (off % d != 0) : (off != 0);
  }
  const unsigned bad_lanes = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad_lanes) atomicAdd(&facts[1], (unsigned long long)__popc(bad_lanes));
}

// The A right children of every prefix group (tm.scm:1310-1322), group-major so that consecutive
// nodes read consecutive entries of p.  A block takes kThreads groups: every thread finds the slot
// of one group's key (the only probe of the table after the insertions: the slot's rank is recorded
// here for emit_kernel, and the adjusted prefix comes out of it), then the block writes the
// kThreads * A children with consecutive threads on consecutive children.
__global__ void __launch_bounds__(kThreads) emit_groups_kernel(const uint64_t* __restrict__ sorted_keys, uint32_t n_keys,
                                                               HashSet hs, Consts c, Frontier next, uint64_t first,
                                                               uint32_t* __restrict__ g_prefix,
                                                               uint32_t* __restrict__ g_adjusted,
                                                               uint32_t* __restrict__ g_seed) {
  __shared__ uint32_t s_po[kThreads], s_pa[kThreads], s_seed[kThreads];
  const uint32_t g0 = blockIdx.x * kThreads, g = g0 + threadIdx.x;
  if (g < n_keys) {
    const uint64_t key = sorted_keys[g];
    const uint64_t slot = hash_slot(hs, key);
    const uint32_t po = (uint32_t)key, pa = hs.vals[slot];
    hs.ranks[slot] = g;
    g_prefix[g] = po;
    g_adjusted[g] = pa;
    if (g_seed) g_seed[g] = (uint32_t)(key >> 32);
    s_po[threadIdx.x] = po; s_pa[threadIdx.x] = pa; s_seed[threadIdx.x] = (uint32_t)(key >> 32);
  }
  if (g_seed) return;  // the next level is not materialised (Frontier::right_only)
  __syncthreads();
  const uint32_t groups = min((uint32_t)kThreads, n_keys - g0), children = groups * c.A;
  const uint8_t nmeta = (uint8_t)((NODE_RIGHT << 6) | c.k);
  const uint64_t base = first + (uint64_t)g0 * c.A;
  for (uint32_t j = threadIdx.x; j < children; j += kThreads) {
    const uint32_t q = j / c.A, x = j - q * c.A;
    const uint64_t ci = base + j;
    next.io[ci] = s_po[q] * c.A + x;
    next.ia[ci] = s_pa[q] * c.A + x;
    next.seed[ci] = s_seed[q];
    next.meta[ci] = nmeta;
    next.flags[ci] = FL_TERM | FL_RIGHT;
  }
}

// ---------------------------------------------------------------------------------------------
// Grouping: (group, value) pairs -> CSR lists with ascending values inside each group.
// ---------------------------------------------------------------------------------------------
__global__ void group_count_kernel(const uint32_t* __restrict__ group, uint64_t n,
                                   uint32_t* __restrict__ counts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && group[i] != kNoRank) atomicAdd(&counts[group[i]], 1u);
}

__global__ void group_fill_ids_kernel(const uint32_t* __restrict__ group, uint64_t n, uint64_t id_base,
                                      const uint64_t* __restrict__ ptr, uint32_t* __restrict__ cursor,
                                      uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && group[i] != kNoRank) {
    const uint32_t g = group[i];
    out[ptr[g] + atomicAdd(&cursor[g], 1u)] = (uint32_t)(id_base + i);
  }
}

__global__ void group_fill_vals_kernel(const uint32_t* __restrict__ group, const uint32_t* __restrict__ val,
                                       uint64_t n, const uint64_t* __restrict__ ptr,
                                       uint32_t* __restrict__ cursor, uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t g = group[i];
    out[ptr[g] + atomicAdd(&cursor[g], 1u)] = val[i];
  }
}

// (group number, the group's own prefix) for every entry of the inflow lists
__global__ void pair_with_prefix_kernel(const uint32_t* __restrict__ prefix_all, const uint32_t* __restrict__ ids,
                                        uint2* __restrict__ out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_uint2(ids[i], prefix_all[ids[i]]);
}

// Groups that do not own their parents gather stored weights; counts those with a parent among the
// right children of the previous level (node ids from prev_right_base on).
__global__ void count_right_readers_kernel(const uint32_t* __restrict__ first, const uint32_t* __restrict__ stride,
                                           const uint32_t* __restrict__ count, uint64_t n_groups, uint64_t prev_right_base,
                                           unsigned long long* __restrict__ readers) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool reads = false;
  if (g < n_groups && !(count[g] & Level::kOwnsParents)) {
    const uint32_t n = count[g] & Level::kCountMask;
    reads = n > 0 && (uint64_t)first[g] + (uint64_t)(n - 1) * stride[g] >= prev_right_base;
  }
  const unsigned lanes = __ballot_sync(0xffffffffu, reads);
  if ((threadIdx.x & 31) == 0 && lanes) atomicAdd(readers, (unsigned long long)__popc(lanes));
}

constexpr uint32_t kShortGroup = 64;  // groups up to this length are sorted by one thread

__global__ void group_sort_kernel(const uint64_t* __restrict__ ptr, uint64_t n_groups,
                                  uint32_t* __restrict__ vals, uint32_t* __restrict__ long_groups,
                                  unsigned long long* __restrict__ n_long) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const uint64_t lo = ptr[g], hi = ptr[g + 1];
  if (hi - lo > kShortGroup) {  // left to group_sort_long_kernel (a state many rules flow through)
    if (long_groups) long_groups[atomicAdd(n_long, 1ull)] = (uint32_t)g;
    return;
  }
  for (uint64_t a = lo + 1; a < hi; ++a) {  // insertion sort
    const uint32_t v = vals[a];
    uint64_t b = a;
    while (b > lo && vals[b - 1] > v) { vals[b] = vals[b - 1]; --b; }
    vals[b] = v;
  }
}

// The same sort with the rows of a block staged in shared memory: 256 consecutive groups own one
// contiguous stretch of `vals`, which is loaded and stored coalesced, and every thread sorts its row
// in shared memory.  (One thread sorting a 100-byte row straight in global memory touches a different
// sector on every step: group_sort_kernel took 19 % of the build's kernel time, profiles/r01_n_*.)
// A block whose stretch does not fit the tile falls back to the global-memory sort of its rows.
constexpr uint32_t kSortTileWords = 11264;  // 44 KB
__global__ void __launch_bounds__(kThreads) group_sort_tile_kernel(const uint64_t* __restrict__ ptr, uint64_t n_groups,
                                                                   uint32_t* __restrict__ vals, uint32_t* __restrict__ long_groups,
                                                                   unsigned long long* __restrict__ n_long) {
  __shared__ uint32_t tile[kSortTileWords];
  const uint64_t g0 = (uint64_t)blockIdx.x * kThreads;
  const uint64_t g = g0 + threadIdx.x;
  const uint64_t g_end = min(g0 + (uint64_t)kThreads, n_groups);
  const uint64_t base = ptr[g0], stop = ptr[g_end];
  const bool staged = stop - base <= kSortTileWords;
  uint64_t lo = 0, hi = 0;
  if (g < n_groups) { lo = ptr[g]; hi = ptr[g + 1]; }
  const bool is_long = hi - lo > kShortGroup;
  if (is_long && long_groups) long_groups[atomicAdd(n_long, 1ull)] = (uint32_t)g;  // left to group_sort_long_kernel
  if (staged) {
    for (uint64_t i = base + threadIdx.x; i < stop; i += kThreads) tile[i - base] = vals[i];
    __syncthreads();
    if (!is_long) {
      uint32_t* v = tile + (lo - base);
      const uint32_t len = (uint32_t)(hi - lo);
      for (uint32_t a = 1; a < len; ++a) {  // insertion sort
        const uint32_t x = v[a];
        uint32_t b = a;
        while (b > 0 && v[b - 1] > x) { v[b] = v[b - 1]; --b; }
        v[b] = x;
      }
    }
    __syncthreads();
    for (uint64_t i = base + threadIdx.x; i < stop; i += kThreads) vals[i] = tile[i - base];
  } else if (g < n_groups && !is_long) {
    for (uint64_t a = lo + 1; a < hi; ++a) {
      const uint32_t x = vals[a];
      uint64_t b = a;
      while (b > lo && vals[b - 1] > x) { vals[b] = vals[b - 1]; --b; }
      vals[b] = x;
    }
  }
}

// One block per long group: bitonic network in its all-ascending form (the first step of every
// merge compares mirrored positions), so positions past the end act as +infinity without ever
// being touched.
__global__ void __launch_bounds__(kThreads) group_sort_long_kernel(const uint64_t* __restrict__ ptr,
                                                                   const uint32_t* __restrict__ long_groups,
                                                                   uint32_t* __restrict__ vals) {
  const uint32_t g = long_groups[blockIdx.x];
  uint32_t* v = vals + ptr[g];
  const uint64_t len = ptr[g + 1] - ptr[g];
  uint64_t padded = 1;
  while (padded < len) padded <<= 1;
  auto exchange = [&](uint64_t i, uint64_t partner) {
    if (partner > i && partner < len) {
      const uint32_t a = v[i], b = v[partner];
      if (a > b) { v[i] = b; v[partner] = a; }
    }
  };
  for (uint64_t k = 2; k <= padded; k <<= 1) {
    for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) exchange(i, i ^ (k - 1));
    __syncthreads();
    for (uint64_t j = k >> 2; j > 0; j >>= 1) {
      for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) exchange(i, i ^ j);
      __syncthreads();
    }
  }
}

// Fused right chain, build side.  A group may own its parents when they are right children of the
// previous level, A apart in group index (stride % A == 0, so all are the same digit x of their
// groups).  consumed[g'] counts how many children of previous-level group g' found such an owner.
__global__ void mark_owned_parents_kernel(const uint32_t* __restrict__ first, const uint32_t* __restrict__ stride,
                                          uint32_t* __restrict__ count, const uint32_t* __restrict__ prefix,
                                          const uint32_t* __restrict__ prev_prefix, uint64_t n_groups,
                                          uint64_t prev_right_base, uint32_t A, uint32_t M,
                                          uint32_t* __restrict__ consumed) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const uint32_t n = count[g] & Level::kCountMask, f = first[g], d = stride[g];
  if (n == 0 || f < prev_right_base || (n > 1 && d % A != 0)) return;
  const uint32_t rel = f - (uint32_t)prev_right_base;
  bool all_digits = n == A;  // parent j is the group with prefix  j * M / A + prefix / A ?
  for (uint32_t j = 0; j < n; ++j) {
    const uint32_t gp = (rel + j * d) / A;
    atomicAdd(&consumed[gp], 1u);
    all_digits = all_digits && prev_prefix[gp] == j * (M / A) + prefix[g] / A;
  }
  count[g] = n | Level::kOwnsParents | (all_digits ? Level::kAllDigits : 0u);
}

// Every previous-level group must have all A children owned, or none.
__global__ void check_consumed_kernel(const uint32_t* __restrict__ consumed, uint64_t n_prev_groups, uint32_t A,
                                      unsigned long long* __restrict__ partial) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  block_add((g < n_prev_groups && consumed[g] != 0 && consumed[g] != A) ? 1ull : 0ull, partial);
}

__global__ void clear_owned_parents_kernel(uint32_t* __restrict__ count, uint64_t n_groups) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) count[g] &= Level::kCountMask;
}

__global__ void mark_deferred_kernel(uint32_t* __restrict__ prev_count, const uint32_t* __restrict__ consumed,
                                     uint64_t n_prev_groups, uint32_t A, unsigned long long* __restrict__ n_deferred) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool deferred = g < n_prev_groups && consumed[g] == A;
  if (deferred) prev_count[g] |= Level::kChildrenDeferred;
  block_add(deferred ? 1ull : 0ull, n_deferred);
}

// ---------------------------------------------------------------------------------------------
// Per-step kernels.
// ---------------------------------------------------------------------------------------------

// marg_{L}[i] = sum_j marg_{L+1}[i * A + j], j ascending from an exact 0 (tm.scm:378-384).
__global__ void marginal_kernel(const double* __restrict__ src, double* __restrict__ dst, uint64_t n_out,
                                uint32_t A) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const double* s = src + i * A;
  double total = 0.0;
  for (uint32_t j = 0; j < A; ++j) total = total + s[j];
  dst[i] = total;
}

// The short tables (marg_top .. marg_0) in one block.
__global__ void marginal_tail_kernel(const double* __restrict__ p, double* __restrict__ marg,
                                     const uint64_t* __restrict__ off_in, int k, int top, uint32_t A) {
  __shared__ uint64_t off[40];
  if (threadIdx.x < 40) off[threadIdx.x] = off_in[threadIdx.x];
  __syncthreads();
  uint64_t n_out = 1;
  for (int i = 0; i < top; ++i) n_out *= A;
  for (int L = top; L >= 0; --L) {
    const double* src = (L + 1 == k) ? p : marg + off[L + 1];
    double* dst = marg + off[L];
    for (uint64_t i = threadIdx.x; i < n_out; i += blockDim.x) {
      double total = 0.0;
      for (uint32_t j = 0; j < A; ++j) total = total + src[i * A + j];
      dst[i] = total;
    }
    n_out /= A;
    __syncthreads();
  }
}

struct Tables {
  const double* p;       // marg_k
  const double* marg;    // marg_L, L < k, at marg + off[L]
  uint64_t off[34];
  int k;
};
__device__ __forceinline__ const double* table(const Tables& t, int L) {
  return L == t.k ? t.p : t.marg + t.off[L];
}

// Leaf-world probabilities: product of unfold ratios (tm.scm:556-565) and choice
// probabilities (tm.scm:617-618) in program order.
__device__ __forceinline__ double rule_weight(const Tables& t, uint32_t r, const uint32_t* rule_ptr, const uint8_t* kind,
                                              const uint8_t* len, const uint32_t* ilong, const uint32_t* ishort,
                                              const double* prob) {
  double w = 1.0;
  for (uint32_t s = rule_ptr[r]; s < rule_ptr[r + 1]; ++s) {
    if (kind[s] == Step::UNFOLD) {
      const int L = len[s];
      const double p_here = fmax(0.0, table(t, L)[ilong[s]]);
      const double p_marg = table(t, L - 1)[ishort[s]];
      const double rel = p_here == 0.0 ? 0.0 : p_here / fmax(p_here, p_marg);
      w = w * rel;
      if (!(w > 0.0)) { w = 0.0; break; }  // pruned branch (tm.scm:565)
    } else {
      w = fmax(0.0, prob[s]) * w;
    }
  }
  return w;
}

__global__ void rule_weight_kernel(Tables t, uint32_t n_rules, const uint32_t* __restrict__ rule_ptr,
                                   const uint8_t* __restrict__ kind, const uint8_t* __restrict__ len,
                                   const uint32_t* __restrict__ ilong, const uint32_t* __restrict__ ishort,
                                   const double* __restrict__ prob, double* __restrict__ rule_w) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rules) rule_w[r] = rule_weight(t, r, rule_ptr, kind, len, ilong, ishort, prob);
}

__global__ void root_kernel(const uint32_t* __restrict__ root_rule, uint32_t n_roots,
                            const double* __restrict__ rule_w, double* __restrict__ w) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_roots) w[i] = rule_w[root_rule[i]];
}

// w_child = w_parent * ratio, pruned when the ratio is not > 0 (tm.scm:1316, 1350, 1373).
__device__ __forceinline__ double child_weight(double w_parent, double p_long, double p_short) {
  if (p_long == 0.0) return 0.0;                       // tm.scm:1266
  const double r = p_long / fmax(p_long, p_short);     // tm.scm:1267-1269
  return r > 0.0 ? w_parent * r : 0.0;
}

// The ratio of a right extension depends on the window alone: p[i] against marg_{k-1}[i / A] (the
// window without its last cell).  It is evaluated once per step for the whole table; the chains of all
// seeds and levels then read one number and multiply instead of reading two and dividing.
__device__ __forceinline__ double extension_ratio(double p_long, double p_short) {
  return p_long == 0.0 ? 0.0 : p_long / fmax(p_long, p_short);
}

constexpr int kRatioBatch = 4;  // entries per thread: loads and divisions of a batch overlap
// Both per-step ratio tables from one read of the table: right extensions divide by the marginal of
// the window without its last cell, left extensions / left shifts to a full window by the marginal
// of the window without its first cell (both last-axis marginals marg_{k-1}, tm.scm:1263-1269 with
// the short indices of 1305-1311 and 1341-1366).  ratio_left may be null.
__global__ void __launch_bounds__(kThreads) ratio_tables_kernel(const double* __restrict__ p, const double* __restrict__ short_table,
                                                                double* __restrict__ ratio_right, double* __restrict__ ratio_left,
                                                                uint64_t n, uint32_t A, uint32_t M) {
  const uint64_t base = (uint64_t)blockIdx.x * (kThreads * kRatioBatch) + threadIdx.x;
  double p_long[kRatioBatch], right_short[kRatioBatch], left_short[kRatioBatch];
#pragma unroll
  for (int u = 0; u < kRatioBatch; ++u) {
    const uint64_t i = base + (uint64_t)u * kThreads;
    p_long[u] = i < n ? p[i] : 0.0;
    right_short[u] = i < n ? short_table[i / A] : 0.0;
    left_short[u] = (ratio_left && i < n) ? short_table[i % M] : 0.0;
  }
#pragma unroll
  for (int u = 0; u < kRatioBatch; ++u) {
    const uint64_t i = base + (uint64_t)u * kThreads;
    if (i < n) {
      ratio_right[i] = extension_ratio(p_long[u], right_short[u]);
      if (ratio_left) ratio_left[i] = extension_ratio(p_long[u], left_short[u]);
    }
  }
}

// marg_{k-1} and the right-extension ratios in one pass over the table: a block stages 256 rows of A
// entries in shared memory, thread r adds up row r in the reference's order (j ascending from an
// exact 0, tm.scm:378-384), then every entry is divided by its row's sum - the operands
// ratio_tables_kernel reads from global memory - and stored coalesced.  Saves the second read of the
// table (0.8 GB of 2.5 GB at 10^8 states).
// BULK: the block's stretch of the table is one contiguous piece of global memory, which one thread
// hands to the bulk-copy engine (cp.async.bulk, completion on an mbarrier) instead of 2560 per-thread
// loads; it needs 16-byte alignment and a multiple of 16 bytes, which the host checks.  Otherwise the
// threads load it, into rows padded to an odd pitch so that the row sums run without bank conflicts.
constexpr int kRowsPerBlock = kThreads;
template <bool BULK>
__global__ void __launch_bounds__(kThreads) marginal_ratio_kernel(const double* __restrict__ p, double* __restrict__ marg,
                                                                  double* __restrict__ ratio, uint64_t n_rows, uint32_t A) {
  extern __shared__ __align__(16) double tile[];  // [kRowsPerBlock * pitch] entries, then [kRowsPerBlock] row sums
  __shared__ __align__(8) unsigned long long arrived;
  const uint32_t pitch = BULK ? A : (A | 1u);
  double* sums = tile + (size_t)kRowsPerBlock * (A | 1u);
  const uint64_t row0 = (uint64_t)blockIdx.x * kRowsPerBlock;
  const uint32_t rows = (uint32_t)min((uint64_t)kRowsPerBlock, n_rows - row0);
  const uint32_t entries = rows * A;
  const double* src = p + row0 * A;
  if (BULK) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&arrived);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
    const uint32_t bytes = entries * (uint32_t)sizeof(double);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(bytes), "r"(bar)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred ok;\n\tmbarrier.try_wait.parity.shared::cta.b64 ok, [%1], 0;\n\tselp.u32 %0, 1, 0, ok;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  } else {
    for (uint32_t e = threadIdx.x; e < entries; e += kThreads) {
      const uint32_t r = e / A;
      tile[r * pitch + (e - r * A)] = src[e];
    }
    __syncthreads();
  }
  if (threadIdx.x < rows) {
    const double* mine = tile + threadIdx.x * pitch;
    double total = 0.0;
    for (uint32_t j = 0; j < A; ++j) total = total + mine[j];
    sums[threadIdx.x] = total;
    marg[row0 + threadIdx.x] = total;
  }
  __syncthreads();
  double* dst = ratio + row0 * A;
  for (uint32_t e = threadIdx.x; e < entries; e += kThreads) {
    const uint32_t r = e / A;
    dst[e] = extension_ratio(tile[r * pitch + (e - r * A)], sums[r]);
  }
}

// A group that owns its parents (fused right chain, engine.h Level): evaluates parent j = child
// x_prev of previous-level group g_prev + j * g_step from that group's sum, stores it at its node
// and returns the sum over j in ascending order.  DIRECT: the parents are the A values of the
// dropped digit in order, so their table indices follow from the group's own prefix.
template <int UO, bool DIRECT, bool RATIO>
__device__ __forceinline__ double own_parents(const Level& lv, const Consts& c, const double* p,
                                              const double* short_table, double* ww,
                                              uint64_t g, uint32_t first, uint32_t stride, uint32_t n,
                                              uint32_t g_prev, uint32_t g_step, uint32_t x_prev) {
  double total = 0.0;
  const uint32_t mine = DIRECT ? lv.g_prefix[g] : 0u;
  const uint32_t mine_short = mine / c.A, short_step = c.M / c.A;
  for (uint32_t e = 0; e < n; e += UO) {
    uint32_t i_short[UO], i_long[UO];  // table indices stay below A^k < 2^32
    double sum_prev[UO], p_long[UO], p_marg[UO];
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      const uint32_t gp = g_prev + (e + u) * g_step;
      if (DIRECT) {
        i_long[u] = (e + u) * c.M + mine;
        i_short[u] = (e + u) * short_step + mine_short;
      } else {
        const uint32_t pre = e + u < n ? lv.prev_prefix[gp] : 0u;
        i_long[u] = pre * c.A + x_prev;
        i_short[u] = pre;
      }
      sum_prev[u] = e + u < n ? lv.prev_total[gp] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      p_long[u] = e + u < n ? p[i_long[u]] : 0.0;  // RATIO: p is the ratio table, i_short = i_long / A is built in
      p_marg[u] = !RATIO && e + u < n ? short_table[i_short[u]] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      const double v = e + u >= n ? 0.0 : (RATIO ? weight_from_ratio(sum_prev[u], p_long[u])
                                                 : child_weight(sum_prev[u], p_long[u], p_marg[u]));
      if (e + u < n) ww[first + (e + u) * stride] = v;
      total += v;
    }
  }
  return total;
}

// One level of the forest.  Blocks [0, left_blocks) evaluate the children of the left parents,
// one parent per thread (the A children of consecutive parents are consecutive in memory for every
// digit x).  The remaining blocks evaluate the right children, 32 prefix groups per warp: each
// lane first adds up one group's parents in ascending id order, then the warp writes the 32 * A
// children cooperatively so that the reads of p and the writes of w are contiguous.
// Loads are issued U at a time before the divisions and stores that depend on them: the kernel is
// bound by HBM latency x bandwidth, and one load in flight per thread reaches ~60 % of peak only
// (profiles/r01_c_*).  wr and ww are the same vector: reads touch earlier levels only.
// The body takes the block index and the thread index inside the block (256 threads) as arguments so
// that the single-launch kernel of small problems (fused_rhs_kernel) can run the same code over
// virtual blocks.  Its pointers carry no __restrict__: the kernels that inline it say what may alias.
template <int U, int UO, bool PROGRESSIONS, bool RATIO>
__device__ __forceinline__ void level_body(const Tables& t, const Consts& c, const Level& lv, uint32_t left_blocks,
                                           uint32_t warp_step_q, uint32_t warp_step_r, const double* wr, double* ww,
                                           const double* ratio, const double* ratio_left, uint32_t block, uint32_t tid) {
  if (block < left_blocks) {
    const uint32_t r = block * (uint32_t)kThreads + tid;
    if (r >= lv.n_left) return;
    const int len = lv.lp_len[r];
    const uint32_t bo = lv.lp_io[r];
    const double wp = wr[lv.lp_gid[r]];
    const uint32_t step = c.pw[len - 1];
    double* out = ww + lv.base + r;
    if (RATIO && ratio_left && len == c.k) {
      // children with a full window: child x has table index bo + x * A^(k-1) and divides by
      // marg_{k-1}[bo] - exactly the pair of operands ratio_left holds the quotient of
      for (uint32_t x0 = 0; x0 < c.A; x0 += U) {
        double rl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) rl[u] = x0 + u < c.A ? ratio_left[bo + (x0 + u) * step] : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (x0 + u < c.A) out[(uint64_t)(x0 + u) * lv.n_left] = weight_from_ratio(wp, rl[u]);
      }
      return;
    }
    const double* tl = table(t, len);
    const double p_short = table(t, len - 1)[bo];
    for (uint32_t x0 = 0; x0 < c.A; x0 += U) {
      double p_long[U];
#pragma unroll
      for (int u = 0; u < U; ++u) p_long[u] = x0 + u < c.A ? tl[bo + (x0 + u) * step] : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (x0 + u < c.A) out[(uint64_t)(x0 + u) * lv.n_left] = child_weight(wp, p_long[u], p_short);
    }
  } else {
    const unsigned lane = tid & 31;
    const uint32_t group_block = lv.block_order ? lv.block_order[block - left_blocks] : block - left_blocks;
    const uint32_t warp = group_block * (kThreads / 32) + (tid >> 5);
    const uint64_t g0 = (uint64_t)warp * 32;
    if (g0 >= lv.n_groups) return;
    const uint64_t g = g0 + lane;
    double total = 0.0, p_short = 0.0;
    uint32_t prefix = 0;
    bool deferred = true;  // lanes without a group have no children to write
    // right extensions: the table itself, or (RATIO) the per-window ratios ratio_right_kernel left behind
    const double* p = RATIO ? ratio : t.p;
    if (g < lv.n_groups) {
      if (PROGRESSIONS) {  // parents first, first + stride, ...
        const uint32_t first = lv.g_first[g], stride = lv.g_stride[g], packed = lv.g_count[g];
        const uint32_t n = packed & Level::kCountMask;
        deferred = (packed & Level::kChildrenDeferred) != 0;
        if (packed & Level::kOwnsParents) {
          // the parents are children x_prev of the previous level's groups g_prev, g_prev + g_step, ...:
          // evaluate and store them here (tm.scm:1310-1318 for the previous shift), then add them up
          const uint32_t rel = first - (uint32_t)lv.prev_right_base;
          const uint32_t g_prev = rel / c.A, x_prev = rel - g_prev * c.A, g_step = stride / c.A;
          const double* short_table = table(t, c.k - 1);
          if (packed & Level::kAllDigits)
            total = own_parents<UO, true, RATIO>(lv, c, p, short_table, ww, g, first, stride, n, g_prev, g_step, x_prev);
          else
            total = own_parents<UO, false, RATIO>(lv, c, p, short_table, ww, g, first, stride, n, g_prev, g_step, x_prev);
        } else {
          for (uint32_t e = 0; e < n; e += U) {
            double v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = e + u < n ? wr[first + (e + u) * stride] : 0.0;  // node ids stay below 2^31
#pragma unroll
            for (int u = 0; u < U; ++u) total += v[u];  // ascending id order; + 0.0 leaves the sum as it is
          }
        }
      } else {
        deferred = false;
        const uint64_t lo = lv.g_ptr[g], hi = lv.g_ptr[g + 1];
        for (uint64_t e = lo; e < hi; e += U) {
          uint32_t id[U];
          double v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) id[u] = e + u < hi ? lv.g_parents[e + u] : kNoRank;
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = id[u] != kNoRank ? wr[id[u]] : 0.0;
#pragma unroll
          for (int u = 0; u < U; ++u) total += v[u];
        }
      }
      lv.g_total[g] = total;  // read by the next level when it owns this group's children, and by prefix_sums_kernel
      if (!deferred) {
        prefix = lv.g_prefix[g];
        if (!RATIO) p_short = table(t, c.k - 1)[prefix];
      }
    }
    const uint32_t skip = __ballot_sync(0xffffffffu, deferred);  // bit gl: children of group gl are not ours
    if (skip == 0xffffffffu) return;
    const uint32_t groups_here = (uint32_t)min((uint64_t)32, (uint64_t)lv.n_groups - g0);
    const uint32_t children = groups_here * c.A;
    double* out = ww + lv.base + (uint64_t)c.A * lv.n_left + g0 * c.A;
    // child j = group gl, digit x with j = gl * A + x; a step of 32 children advances (gl, x) by
    // (warp_step_q, warp_step_r) = divmod(32, A) with a carry
    uint32_t gl = lane / c.A, x = lane - gl * c.A;
    for (uint32_t j = lane; j < ((children + 31) & ~31u); j += 32 * U) {
      uint32_t gu[U], xu[U];
      double p_long[U];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        gu[u] = gl; xu[u] = x;
        gl += warp_step_q; x += warp_step_r;
        if (x >= c.A) { x -= c.A; ++gl; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        live[u] = j + 32 * u < children;
        if (!live[u]) gu[u] = 0;
        live[u] = live[u] && !((skip >> gu[u]) & 1u);
        const uint32_t pre = __shfl_sync(0xffffffffu, prefix, gu[u]);
        p_long[u] = live[u] ? p[pre * c.A + xu[u]] : 0.0;  // index below A^k < 2^32
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double wp = __shfl_sync(0xffffffffu, total, gu[u]);
        if (RATIO) {
          if (live[u]) out[j + 32 * u] = weight_from_ratio(wp, p_long[u]);
        } else {
          const double ps = __shfl_sync(0xffffffffu, p_short, gu[u]);
          if (live[u]) out[j + 32 * u] = child_weight(wp, p_long[u], ps);
        }
      }
    }
  }
}

template <int U, int UO, bool PROGRESSIONS, int MIN_BLOCKS, bool RATIO>
__global__ void __launch_bounds__(kThreads, MIN_BLOCKS) level_kernel(Tables t, Consts c, Level lv, uint32_t left_blocks,
                                                         uint32_t warp_step_q, uint32_t warp_step_r,
                                                         const double* __restrict__ wr, double* __restrict__ ww,
                                                         const double* __restrict__ ratio, const double* __restrict__ ratio_left) {
  level_body<U, UO, PROGRESSIONS, RATIO>(t, c, lv, left_blocks, warp_step_q, warp_step_r, wr, ww, ratio, ratio_left, blockIdx.x,
                                         threadIdx.x);
}

// out_sum[q] = sum of the sums of all prefix groups with prefix q, in ascending group number (fixed
// order): the common factor of the outflow of every right child that leaves a row q * A + x
// (tm.scm:1310-1318: child weight = sum * ratio; accumulate-dp/dt 1288: -w at the original window).
__global__ void __launch_bounds__(kThreads) prefix_sums_kernel(const uint64_t* __restrict__ out_ptr, const uint32_t* __restrict__ out_ids,
                                                               const double* __restrict__ totals, double* __restrict__ out_sum,
                                                               uint64_t n_prefixes) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_prefixes) return;
  const uint64_t lo = out_ptr[q], hi = out_ptr[q + 1];
  double total = 0.0;
  for (uint64_t e = lo; e < hi; e += 4) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = e + u < hi ? totals[out_ids[e + u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) total += v[u];
  }
  out_sum[q] = total;
}

// One regular block of 256 prefix groups (engine.h Level::PlaneBlock): thread t evaluates group t of
// the block.  Same operands, same operations and the same order of additions as level_kernel's
// own_parents path (tm.scm:1310-1322), so weights keep their bits; what is gone are the four
// per-group records, the bounds tests, and the dependence of the parent loads on one another:
// with the alphabet size known at compile time (A_ > 0) all 2 * A loads of a thread are in flight
// at once.  A_ = 0: any alphabet, loads in batches of four.
// STORE: the weights of the right children (the groups' parents, the children of groups nobody
// takes over) are written; they are needed only when a later level reads stored weights
// (Model::materialize_right) - the flux of right children is evaluated from the group sums.
template <int A_, bool STORE>
__global__ void __launch_bounds__(kThreads, A_ > 8 || A_ == 0 ? 3 : 4) plane_kernel(Consts c, Level lv, const double* __restrict__ ratio,
                                                                       const double* __restrict__ wr, double* __restrict__ ww) {
  const uint4* rec4 = (const uint4*)(lv.plane_blocks + blockIdx.x);
  const uint4 r0 = rec4[0], r1 = rec4[1], r2 = rec4[2];
  const uint32_t A = A_ ? (uint32_t)A_ : c.A;
  const uint32_t tid = threadIdx.x;
  const uint32_t g = r0.x * (uint32_t)kThreads + tid, first = r0.z + tid;
  uint32_t prefix = r0.y + tid;
  if (tid > r1.z) prefix += r2.x * (1u + (tid - r1.z - 1u) / r1.w);  // r1.z = jump_at, r1.w = period, r2.x = jump
  const uint32_t i_long = prefix + r0.w;
  const uint32_t stride = r1.x, n_par = r1.y & 0xffffu;
  const bool deferred = (r1.y & Level::kPlaneDeferred) != 0;
  double total = 0.0;
  if (r1.y & Level::kPlaneGather) {
    total = wr[first];  // the parent is a stored node of the previous level
  } else {
    const uint32_t rel = first - (uint32_t)lv.prev_right_base;
    const uint32_t gp = rel / A;
    const double* __restrict__ prev_total = lv.prev_total;
    if (n_par == 1) {
      total = weight_from_ratio(prev_total[gp], ratio[i_long]);
      if (STORE) ww[first] = total;
    } else {
      const uint32_t g_step = stride / A;
      if (A_ > 0) {
        double r[A_ > 0 ? A_ : 1], t[A_ > 0 ? A_ : 1];
#pragma unroll
        for (int j = 0; j < A_; ++j) {
          r[j] = ratio[i_long + (uint32_t)j * c.M];
          t[j] = prev_total[gp + (uint32_t)j * g_step];
        }
#pragma unroll
        for (int j = 0; j < A_; ++j) {
          const double v = weight_from_ratio(t[j], r[j]);
          if (STORE) ww[first + (uint32_t)j * stride] = v;
          total += v;
        }
      } else {
        for (uint32_t j0 = 0; j0 < A; j0 += 4) {
          double r[4], t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            r[u] = j0 + u < A ? ratio[i_long + (j0 + u) * c.M] : 0.0;
            t[u] = j0 + u < A ? prev_total[gp + (j0 + u) * g_step] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (j0 + u < A) {
              const double v = weight_from_ratio(t[u], r[u]);
              if (STORE) ww[first + (j0 + u) * stride] = v;
              total += v;
            }
          }
        }
      }
    }
  }
  lv.g_total[g] = total;  // read by the next level when it owns this group's children, and by the product
  if (deferred || !STORE) return;
  // the 32 * A children of the warp's 32 groups: contiguous in the weight vector, and in the table
  // wherever the prefixes are
  const uint32_t lane = tid & 31, warp_first = tid - lane;
  double* out = ww + lv.base + (uint64_t)A * lv.n_left + (uint64_t)(r0.x * (uint32_t)kThreads + warp_first) * A;
#pragma unroll 5
  for (uint32_t j = lane; j < 32 * A; j += 32) {
    const uint32_t gl = j / A, x = j - gl * A;
    const double wp = __shfl_sync(0xffffffffu, total, gl);
    const uint32_t pre = __shfl_sync(0xffffffffu, prefix, gl);
    out[j] = weight_from_ratio(wp, ratio[(uint64_t)pre * A + x]);
  }
}

// Build side of the plane blocks: is the block of 256 groups starting at 256 * b regular (engine.h
// Level::PlaneBlock)?  rec[b] is written for every block (prefix0 also orders the other blocks).
__global__ void __launch_bounds__(kThreads) classify_plane_blocks_kernel(Level lv, Consts c, uint32_t n_blocks,
                                                                         Level::PlaneBlock* __restrict__ rec,
                                                                         uint32_t* __restrict__ is_plane) {
  __shared__ uint32_t base[5];  // first, stride, packed, prefix, long - prefix of the block's first group
  __shared__ uint32_t pre[kThreads];
  __shared__ uint32_t jump_at, jump_next;
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  const uint64_t g = (uint64_t)b * kThreads + tid;
  const bool full = (uint64_t)(b + 1) * kThreads <= lv.n_groups;
  uint32_t first = 0, stride = 0, packed = 0, prefix = 0, long_off = 0;
  bool ok = full && lv.g_first != nullptr;
  if (g < lv.n_groups) prefix = lv.g_prefix[g];
  pre[tid] = prefix;
  if (tid == 0) { jump_at = 0xffffffffu; jump_next = 0xffffffffu; }
  if (ok) {
    first = lv.g_first[g]; stride = lv.g_stride[g]; packed = lv.g_count[g];
    const uint32_t n = packed & Level::kCountMask;
    if (packed & Level::kOwnsParents) {
      ok = lv.prev_total != nullptr;
      if (ok && n == c.A && (packed & Level::kAllDigits) && c.A > 1) {
        long_off = 0;  // parent j reads j * A^(k-1) + prefix
        ok = stride % c.A == 0;
      } else if (ok && n == 1) {
        const uint32_t rel = first - (uint32_t)lv.prev_right_base;
        const uint64_t at = (uint64_t)lv.prev_prefix[rel / c.A] * c.A + rel % c.A;
        long_off = (uint32_t)at - prefix;
      } else {
        ok = false;
      }
    } else {
      ok = n == 1;  // one stored parent: its weight is gathered
    }
  }
  if (tid == 0) { base[0] = first; base[1] = stride; base[2] = packed; base[3] = prefix; base[4] = long_off; }
  __syncthreads();
  // where the prefix does not simply count up: first such place, then the next one
  const bool jumps_here = tid + 1 < (uint32_t)kThreads && pre[tid + 1] != pre[tid] + 1u;
  if (jumps_here) atomicMin(&jump_at, tid);
  __syncthreads();
  const uint32_t a = jump_at;
  if (jumps_here && tid > a) atomicMin(&jump_next, tid);
  __syncthreads();
  const uint32_t period = jump_next != 0xffffffffu ? jump_next - a : 0x40000000u;
  const uint32_t jump = a != 0xffffffffu ? pre[a + 1] - pre[a] - 1u : 0u;
  uint32_t expect = base[3] + tid;
  if (tid > a) expect += jump * (1u + (tid - a - 1u) / period);
  ok = ok && first == base[0] + tid && stride == base[1] && packed == base[2] && prefix == expect && long_off == base[4];
  const int bad = __syncthreads_or(ok ? 0 : 1);
  if (tid == 0) {
    Level::PlaneBlock out;
    out.group_block = b; out.prefix0 = base[3]; out.first0 = base[0]; out.long_off = base[4]; out.stride = base[1];
    out.meta = (base[2] & Level::kCountMask) | ((base[2] & Level::kChildrenDeferred) ? Level::kPlaneDeferred : 0u) |
               ((base[2] & Level::kOwnsParents) ? 0u : Level::kPlaneGather);
    out.jump_at = a; out.period = period; out.jump = jump;
    out.pad[0] = out.pad[1] = out.pad[2] = 0;
    rec[b] = out;
    is_plane[b] = bad ? 0u : 1u;
  }
}

// Order of the blocks of 256 groups of a level and their split into plane blocks and the others, on
// the device (the host used to fetch the records of every level, sort them and send three arrays back:
// 14 ms of copies and synchronisation for 0.7 ms of kernels at the bench size).
__global__ void plane_keys_kernel(const Level::PlaneBlock* __restrict__ rec, uint32_t n_blocks, uint64_t* __restrict__ keys) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_blocks) keys[b] = ((uint64_t)rec[b].prefix0 << 32) | b;  // by prefix, ties in block order
}

// order[i] = block at position i; taken[i] = whether that block is a plane block; facts[1] counts the
// positions that are not their own block (0: the order is the identity).
__global__ void plane_order_kernel(const uint64_t* __restrict__ sorted, uint32_t n_blocks, const uint32_t* __restrict__ is_plane,
                                   uint32_t* __restrict__ order, uint32_t* __restrict__ taken,
                                   unsigned long long* __restrict__ facts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool moved = false;
  if (i < n_blocks) {
    const uint32_t b = (uint32_t)sorted[i];
    order[i] = b;
    taken[i] = is_plane[b] ? 1u : 0u;
    moved = b != i;
  }
  block_add(moved ? 1ull : 0ull, &facts[1]);
}

// Plane records and the numbers of the other blocks, both in the order of `order`; facts[0] = plane blocks.
__global__ void plane_partition_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ taken,
                                       const uint64_t* __restrict__ position, const Level::PlaneBlock* __restrict__ rec,
                                       uint32_t n_blocks, Level::PlaneBlock* __restrict__ planes,
                                       uint32_t* __restrict__ general, unsigned long long* __restrict__ facts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_blocks) return;
  const uint64_t before = position[i];  // plane blocks before position i
  if (taken[i]) planes[before] = rec[order[i]];
  else general[i - before] = order[i];
  if (i == 0) facts[0] = position[n_blocks];
}

// ---------------------------------------------------------------------------------------------
// Small problems: the whole right-hand side in ONE launch.
//
// The reference's shipped problems have 8 ... 10^5 states; about 20 dependent kernels of a few
// microseconds each were launch latency and nothing else (ex2 at cl_k = 7: 57.6 us on the GPU
// against 40 us for the CPU port, profiles/r01_l_*).  Here one thread block, or one cluster of up to
// 16 thread blocks (distributed over as many SMs, synchronised by the hardware cluster barrier), walks
// through the phases - marginal tables, leaf-world probabilities and ratio table, every forest
// level, the product - with a barrier between them.  The phases run the very code of the separate
// kernels (level_body, slice_sum, rule_weight), over virtual blocks, so every result keeps its bits.
// ---------------------------------------------------------------------------------------------
struct FusedLevel {
  Level lv;
  uint32_t left_blocks, group_blocks, warp_step_q, warp_step_r;
};

struct FusedArgs {
  const FusedLevel* levels;
  int n_levels;
  uint32_t n_rules;
  const uint32_t* rule_ptr;
  const uint8_t* step_kind;
  const uint8_t* step_len;
  const uint32_t* step_long;
  const uint32_t* step_short;
  const double* step_prob;
  double* rule_w;
  double* marg;
  double* ratio_right;
  double* node_w;
  const uint64_t* slice_ptr;
  const uint32_t* slice_runs;
  const uint32_t* words;
  uint64_t n_states, n_slices;
  double* out;
  int fused_update;
  const uint64_t* out_ptr;   // flux of the right children (engine.h Model::out_ptr, in_ptr); out_sum null: none
  const uint32_t* out_ids;
  const double* g_total_all;
  double* out_sum;
  uint64_t n_prefixes;
  const uint64_t* in_ptr;
  const uint2* in_pairs;
};

constexpr int kFusedThreads = 1024;

template <bool CLUSTER>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_rhs_kernel(Tables t, Consts c, FusedArgs a, StageUpdate up) {
  uint32_t cta = 0, n_cta = 1;
  if (CLUSTER) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n_cta));
  }
  auto phase_barrier = [&]() {
    if (CLUSTER) {
      // release / acquire at cluster scope: the global writes of the phase are visible to all blocks
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      __syncthreads();
    }
  };
  const uint32_t gtid = cta * kFusedThreads + threadIdx.x, gthreads = n_cta * kFusedThreads;
  const double* p = t.p;
  // marginal tables, longest first (tm.scm:378-384: j ascending from an exact 0)
  // (32-bit index arithmetic throughout: a 64-bit division per phase and thread is a third of a phase
  // of a tiny problem, profiles/r02_aj_small_kernel_phase_cycles.txt)
  for (int L = c.k - 1; L >= 0; --L) {
    const double* src = (L + 1 == c.k) ? p : a.marg + t.off[L + 1];
    double* dst = a.marg + t.off[L];
    const uint32_t n_out = c.pw[L];  // A^L
    for (uint32_t i = gtid; i < n_out; i += gthreads) {
      const double* s = src + i * c.A;
      double total = 0.0;
      for (uint32_t j = 0; j < c.A; ++j) total = total + s[j];
      dst[i] = total;
    }
    phase_barrier();
  }
  // leaf-world probabilities and the right-extension ratios
  for (uint32_t r = gtid; r < a.n_rules; r += gthreads)
    a.rule_w[r] = rule_weight(t, r, a.rule_ptr, a.step_kind, a.step_len, a.step_long, a.step_short, a.step_prob);
  if (a.ratio_right) {
    const double* short_table = a.marg + t.off[c.k - 1];
    const uint32_t n32 = (uint32_t)a.n_states;
    for (uint32_t i = gtid; i < n32; i += gthreads) a.ratio_right[i] = extension_ratio(p[i], short_table[i / c.A]);
  }
  phase_barrier();
  // forest levels over virtual blocks of 256 threads
  const uint32_t sub = threadIdx.x / kThreads, tid = threadIdx.x % kThreads;
  constexpr uint32_t kSubs = kFusedThreads / kThreads;
  for (int l = 0; l < a.n_levels; ++l) {
    const FusedLevel& fl = a.levels[l];
    if (fl.lv.n_roots) {
      for (uint32_t i = gtid; i < fl.lv.n_roots; i += gthreads) a.node_w[i] = a.rule_w[fl.lv.root_rule[i]];
    } else {
      const uint32_t blocks = fl.left_blocks + fl.group_blocks;
      if (fl.lv.g_first != nullptr || fl.lv.n_groups == 0) {
        for (uint32_t vb = cta * kSubs + sub; vb < blocks; vb += n_cta * kSubs)
          level_body<4, 2, true, true>(t, c, fl.lv, fl.left_blocks, fl.warp_step_q, fl.warp_step_r, a.node_w, a.node_w,
                                       a.ratio_right, nullptr, vb, tid);
      } else {
        for (uint32_t vb = cta * kSubs + sub; vb < blocks; vb += n_cta * kSubs)
          level_body<4, 1, false, true>(t, c, fl.lv, fl.left_blocks, fl.warp_step_q, fl.warp_step_r, a.node_w, a.node_w,
                                        a.ratio_right, nullptr, vb, tid);
      }
    }
    phase_barrier();
  }
  // what leaves the rows through right children: one sum per prefix (prefix_sums_kernel)
  if (a.out_sum) {
    const uint32_t n_prefixes = (uint32_t)a.n_prefixes;
    for (uint32_t q = gtid; q < n_prefixes; q += gthreads) {
      double total = 0.0;
      for (uint64_t e = a.out_ptr[q]; e < a.out_ptr[q + 1]; ++e) total += a.g_total_all[a.out_ids[e]];
      a.out_sum[q] = total;
    }
    phase_barrier();
  }
  // the product, one slice of 32 states per warp
  const unsigned lane = threadIdx.x & 31;
  RightFlux right;
  right.out_sum = a.out_sum; right.ratio = a.ratio_right; right.totals = a.g_total_all;
  right.in_ptr = a.in_ptr; right.in_pairs = a.in_pairs; right.A = c.A;
  for (uint64_t s = gtid >> 5; s < a.n_slices; s += gthreads >> 5) {
    double acc = slice_sum<4>(a.slice_ptr, a.slice_runs, a.words, a.node_w, s, lane);
    const uint64_t row = s * 32 + lane;
    if (row < a.n_states) {
      acc = acc + right_flux<4>(right, row);
      a.out[row] = acc;
      if (a.fused_update) {  // Runge-Kutta stage update, terms in tableau order (as in flux_slices_kernel)
        double sum = 0.0;
        for (int j = 0; j < up.n; ++j) sum += up.vec[j][row] * up.coef[j];
        sum += acc * up.coef_self;
        up.stage[row] = up.y[row] + sum * up.h;
      }
    }
  }
}

// The weights of the right children of one level from the group sums of the last evaluation - what
// the level kernels write when Model::materialize_right is set (same operands, same product).
__global__ void materialize_right_kernel(Level lv, Consts c, const double* __restrict__ ratio, double* __restrict__ w) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (uint64_t)lv.n_groups * c.A) return;
  const uint64_t g = j / c.A, x = j - g * c.A;
  w[lv.base + (uint64_t)c.A * lv.n_left + j] = weight_from_ratio(lv.g_total[g], ratio[(uint64_t)lv.g_prefix[g] * c.A + x]);
}

// dy/dt[row] = sum over the row's entries of +-w[node].  G lanes share a row; every lane keeps
// kSpmvUnroll independent entry loads and weight gathers in flight, then the lanes of a row are
// combined by a shuffle reduction.  The summation order is fixed by (G, unroll), so results are
// reproducible run to run.
constexpr int kSpmvUnroll = 4;

template <int G, bool FUSED>
__global__ void __launch_bounds__(256) spmv_kernel(const uint64_t* __restrict__ row_ptr,
                                                   const uint32_t* __restrict__ entries,
                                                   const double* __restrict__ w, double* __restrict__ out,
                                                   uint64_t n_rows, StageUpdate up, int accumulate, RightFlux right,
                                                   uint64_t first_row) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t row = t / G;
  const int sub = (int)(t % G);
  double acc = 0.0;
  if (row < n_rows) {
    const uint64_t lo = row_ptr[row], hi = row_ptr[row + 1];
    for (uint64_t e = lo + sub; e < hi; e += (uint64_t)G * kSpmvUnroll) {
      uint32_t v[kSpmvUnroll];
#pragma unroll
      for (int u = 0; u < kSpmvUnroll; ++u) {
        const uint64_t eu = e + (uint64_t)u * G;
        v[u] = eu < hi ? entries[eu] : 0xffffffffu;
      }
      double x[kSpmvUnroll];
#pragma unroll
      for (int u = 0; u < kSpmvUnroll; ++u) x[u] = v[u] != 0xffffffffu ? w[v[u] & ~kOutflowBit] : 0.0;
#pragma unroll
      for (int u = 0; u < kSpmvUnroll; ++u) acc += (v[u] & kOutflowBit) ? -x[u] : x[u];
    }
  }
  if (G > 1) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d, G);
  }
  if (sub == 0 && row < n_rows) {
    acc = acc + right_flux<kSpmvUnroll>(right, first_row + row);
    if (accumulate) acc = out[row] + acc;  // a later part of a composite model
    out[row] = acc;
    if (FUSED) {  // Runge-Kutta stage update for this state (same term order as the unfused kernel)
      double a = 0.0;
      for (int j = 0; j < up.n; ++j) a += up.vec[j][row] * up.coef[j];
      a += acc * up.coef_self;
      up.stage[row] = up.y[row] + a * up.h;
    }
  }
}

double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

struct HostRoot {
  uint32_t io, ia, seed, rule;
  uint8_t meta, flags;
};

Consts make_consts(const Model& m) {
  Consts c;
  c.A = (uint32_t)m.A;
  c.k = m.k;
  c.M = (uint32_t)m.pow_a[m.k - 1];
  for (int i = 0; i < 33; ++i) c.pw[i] = i < m.k ? (uint32_t)m.pow_a[i] : 0u;
  return c;
}

}  // namespace

cudaMemPool_t library_pool() {
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  TAPES_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) throw std::runtime_error("device index out of range");
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    TAPES_CUDA_CHECK(cudaMemPoolCreate(&pools[dev], &props));
    uint64_t keep = 2048ull << 20;
    if (const char* e = std::getenv("TAPES_POOL_KEEP_MB")) keep = std::strtoull(e, nullptr, 10) << 20;
    TAPES_CUDA_CHECK(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  return pools[dev];
}

void* pool_alloc(size_t bytes, cudaStream_t st) {
  void* p = nullptr;
  TAPES_CUDA_CHECK(cudaMallocFromPoolAsync(&p, std::max<size_t>(bytes, 1), library_pool(), st));
  return p;
}

namespace {
bool capturing(cudaStream_t st) {
  if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) return false;
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &status) != cudaSuccess) { cudaGetLastError(); return false; }
  return status != cudaStreamCaptureStatusNone;
}
}  // namespace

void begin_use(Model& m, cudaStream_t st) {
  if (!m.in_use || m.last_stream == st || capturing(st)) return;  // a capturing caller orders its graph itself
  TAPES_CUDA_CHECK(cudaStreamWaitEvent(st, m.busy, 0));
}

void end_use(Model& m, cudaStream_t st) {
  if (capturing(st)) return;
  if (!m.busy) TAPES_CUDA_CHECK(cudaEventCreateWithFlags(&m.busy, cudaEventDisableTiming));
  TAPES_CUDA_CHECK(cudaEventRecord(m.busy, st));
  m.last_stream = st;
  m.in_use = true;
}

void* DeviceArena::take(size_t n_bytes) {
  n_bytes = (std::max<size_t>(n_bytes, 1) + 255) & ~(size_t)255;
  bytes += n_bytes;
  if (n_bytes <= left) {
    void* p = cursor;
    cursor += n_bytes; left -= n_bytes;
    return p;
  }
  void* p = nullptr;
  if (n_bytes >= next_chunk / 2) {  // large arrays get a chunk of their own; the open chunk stays open
    TAPES_CUDA_CHECK(cudaMalloc(&p, n_bytes));
    chunks.push_back(p);
    return p;
  }
  TAPES_CUDA_CHECK(cudaMalloc(&p, next_chunk));
  chunks.push_back(p);
  cursor = (char*)p + n_bytes;
  left = next_chunk - n_bytes;
  next_chunk = std::min<size_t>(next_chunk * 2, (size_t)1 << 30);
  return p;
}

void DeviceArena::release() {
  for (void* p : chunks) cudaFree(p);
  chunks.clear();
  cursor = nullptr; left = 0; bytes = 0; next_chunk = 32u << 20;
}

void Model::drop_weight_graphs() {
  for (WeightsGraph& g : weight_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  weight_graphs.clear();
}

Model::~Model() {
  if (stream) cudaStreamSynchronize(stream);
  drop_weight_graphs();
  more.clear();  // the other parts of a composite model use this model's stream: they go first
  cudaStream_t st = stream;
  if (st) cudaStreamSynchronize(st);
  if (entries) cudaFree(entries);
  entries = nullptr;
  if (h_pinned) cudaFreeHost(h_pinned);
  if (d_obs_spec) cudaFree(d_obs_spec);
  if (d_obs_out) cudaFree(d_obs_out);
  if (d_obs_partial) cudaFree(d_obs_partial);
  if (busy) cudaEventDestroy(busy);
  arena.release();
  if (copy_stream) {
    cudaStreamDestroy(copy_stream);
    for (cudaEvent_t e : copy_events) if (e) cudaEventDestroy(e);
  }
  if (own_stream && stream) cudaStreamDestroy(stream);
}

namespace {
// Never destroyed (no cudaFree during static destruction); release_build_scratch() empties them.
Slab* build_scratch() {
  static Slab* slabs = new Slab[3];
  return slabs;
}

// Ascending order inside every group of a CSR-like list: short groups by one thread each, long
// ones (found on the way) by one block each.
void sort_groups(const uint64_t* ptr, uint64_t n_groups, uint32_t* vals, cudaStream_t st) {
  if (n_groups == 0) return;
  uint32_t* long_groups = dalloc<uint32_t>(n_groups, st);
  unsigned long long* n_long = dalloc<unsigned long long>(1, st);
  TAPES_CUDA_CHECK(cudaMemsetAsync(n_long, 0, 8, st));
  if (std::getenv("TAPES_SORT_IN_GLOBAL"))
    group_sort_kernel<<<grid_for(n_groups, kThreads), kThreads, 0, st>>>(ptr, n_groups, vals, long_groups, n_long);
  else
    group_sort_tile_kernel<<<grid_for(n_groups, kThreads), kThreads, 0, st>>>(ptr, n_groups, vals, long_groups, n_long);
  unsigned long long h_long = 0;
  TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_long, n_long, 8, cudaMemcpyDeviceToHost, st));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_long) group_sort_long_kernel<<<(unsigned)h_long, kThreads, 0, st>>>(ptr, long_groups, vals);
  TAPES_CUDA_CHECK(cudaGetLastError());
  dfree(long_groups, st); dfree(n_long, st);
}
}  // namespace

void release_build_scratch() {
  Slab* slabs = build_scratch();
  for (int i = 0; i < 3; ++i) slabs[i].release();
  if (cudaMemPoolTrimTo(library_pool(), 0) != cudaSuccess) cudaGetLastError();
}

std::unique_ptr<Model> build_model(const RuleTable& table, cudaStream_t stream) {
  std::unique_ptr<Model> mp(new Model());
  Model& m = *mp;
  m.A = table.alphabet;
  m.k = table.cl_k;
  if (m.k < 1 || m.k > 32) throw std::runtime_error("cl_k must be in 1..32");
  if (m.A < 1 || m.A > 65535) throw std::runtime_error("alphabet size must be in 1..65535");
  m.pow_a[0] = 1;
  for (int i = 1; i <= m.k; ++i) m.pow_a[i] = m.pow_a[i - 1] * (uint64_t)m.A;
  m.n_states = m.pow_a[m.k];
  if (m.n_states >= (1ull << 32)) throw std::runtime_error("A^cl_k must be below 2^32");
  if (stream) {
    m.stream = stream;
  } else {
    TAPES_CUDA_CHECK(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    m.own_stream = true;
  }
  cudaStream_t st = m.stream;

  const Consts c = make_consts(m);
  const uint64_t W = m.n_states, M = m.pow_a[m.k - 1];

  m.stats.worlds_walked = table.worlds_walked;
  m.stats.leaf_worlds = table.leaf_worlds;
  m.stats.flux_rules = (int64_t)table.rules.size();

  // ---- rule steps to the device ----
  {
    std::vector<uint32_t> ptr(1, 0), ilong, ishort;
    std::vector<uint8_t> kind, len;
    std::vector<double> prob;
    for (const FluxRule& r : table.rules) {
      for (const Step& s : r.steps) {
        kind.push_back(s.kind); len.push_back(s.length);
        ilong.push_back(s.long_index); ishort.push_back(s.short_index); prob.push_back(s.prob);
      }
      ptr.push_back((uint32_t)kind.size());
    }
    m.n_rules = (uint32_t)table.rules.size();
    auto up = [&](auto*& dptr, const auto& h) {
      typedef typename std::remove_const<typename std::remove_reference<decltype(h[0])>::type>::type T;
      dptr = dkeep<T>(m, h.size());
      if (!h.empty())
        TAPES_CUDA_CHECK(cudaMemcpyAsync((void*)dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    };
    up(m.rule_ptr, ptr); up(m.step_kind, kind); up(m.step_len, len);
    up(m.step_long, ilong); up(m.step_short, ishort); up(m.step_prob, prob);
    m.rule_w = dkeep<double>(m, m.n_rules);
    TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  }

  // ---- roots (finish-fn-eval-fast-fixed, tm.scm:1416-1443; top call 1398-1401) ----
  std::vector<HostRoot> roots;
  uint32_t n_seeds = 0;
  for (size_t r = 0; r < table.rules.size(); ++r) {
    for (int t = 0; t < 2; ++t) {
      const Seed& sd = table.rules[r].tape[t];
      if (!sd.changed()) continue;
      const uint32_t seed = n_seeds++;
      uint64_t io = sd.orig, ia = sd.adjusted;
      int len = sd.length;
      if (len >= m.k - 1) {  // right chain starts from the right-most k-1 digits
        const uint64_t po = io % M, pa = ia % M;
        if (po != pa)
          roots.push_back({(uint32_t)po, (uint32_t)pa, seed, (uint32_t)r, (uint8_t)((NODE_ROOT << 6) | (m.k - 1)), FL_RIGHT});
      }
      bool alive = true;
      while (len > m.k) {  // tm.scm:1380-1390: accumulate, drop the right-most digit
        if (io == ia) { alive = false; break; }
        const uint64_t s = io % W, d = ia % W;
        if (s != d)
          roots.push_back({(uint32_t)s, (uint32_t)d, seed, (uint32_t)r, (uint8_t)((NODE_ROOT << 6) | m.k), FL_TERM});
        io /= (uint64_t)m.A; ia /= (uint64_t)m.A; --len;
      }
      if (alive && io != ia) {
        uint8_t fl = FL_LEFT;
        if (len == m.k) fl |= FL_TERM;
        roots.push_back({(uint32_t)io, (uint32_t)ia, seed, (uint32_t)r, (uint8_t)((NODE_ROOT << 6) | len), fl});
      }
    }
  }
  m.stats.seeds = n_seeds;

  g_alloc_ms = 0;
  // ---- level-synchronous expansion ----
  // Scratch memory comes in two slabs per level (one sized before the level's counts are known,
  // one after): growing the stream-ordered pool by hundreds of small requests cost 2.9 s of a
  // 3.2 s first build (profiles/r01_e_*).
  auto t_expand = std::chrono::steady_clock::now();
  struct EdgeChunk { uint32_t* row; uint32_t* val; uint64_t n; };
  struct EdgeChunks : std::vector<EdgeChunk> {  // freed on every way out, a failed build included
    ~EdgeChunks() { for (EdgeChunk& ec : *this) if (ec.row) cudaFree(ec.row); }
  } edge_chunks;

  Frontier cur;
  cur.n = roots.size();
  Level cur_level;
  cur_level.base = 0;
  cur_level.n_roots = (uint32_t)roots.size();
  // kept between builds while small, so that the reference's small problems do not pay for
  // driver allocations at all
  Slab* slabs = build_scratch();
  Slab& cur_slab = slabs[0];  // owns the arrays of `cur`
  Slab& s1 = slabs[1];        // scratch sized before the counts of a level are known
  Slab& s2 = slabs[2];        // ... and after; owns `next`
  cur_slab.reset(); s1.reset(); s2.reset();
  auto plan_frontier = [](Slab& slab, uint64_t n, size_t idx[5]) {
    idx[0] = slab.want(n * 4); idx[1] = slab.want(n * 4); idx[2] = slab.want(n * 4);
    idx[3] = slab.want(n); idx[4] = slab.want(n);
  };
  auto bind_frontier = [](const Slab& slab, const size_t idx[5], Frontier& f) {
    f.io = slab.at<uint32_t>(idx[0]); f.ia = slab.at<uint32_t>(idx[1]); f.seed = slab.at<uint32_t>(idx[2]);
    f.meta = slab.at<uint8_t>(idx[3]); f.flags = slab.at<uint8_t>(idx[4]);
  };
  if (cur.n) {
    std::vector<uint32_t> h_io(cur.n), h_ia(cur.n), h_seed(cur.n), h_rule(cur.n);
    std::vector<uint8_t> h_meta(cur.n), h_fl(cur.n);
    for (size_t i = 0; i < cur.n; ++i) {
      h_io[i] = roots[i].io; h_ia[i] = roots[i].ia; h_seed[i] = roots[i].seed; h_rule[i] = roots[i].rule;
      h_meta[i] = roots[i].meta; h_fl[i] = roots[i].flags;
    }
    cur_level.root_rule = dkeep<uint32_t>(m, cur.n);
    size_t fi[5];
    plan_frontier(cur_slab, cur.n, fi);
    cur_slab.commit();
    bind_frontier(cur_slab, fi, cur);
    TAPES_CUDA_CHECK(cudaMemcpyAsync(cur.io, h_io.data(), cur.n * 4, cudaMemcpyHostToDevice, st));
    TAPES_CUDA_CHECK(cudaMemcpyAsync(cur.ia, h_ia.data(), cur.n * 4, cudaMemcpyHostToDevice, st));
    TAPES_CUDA_CHECK(cudaMemcpyAsync(cur.seed, h_seed.data(), cur.n * 4, cudaMemcpyHostToDevice, st));
    TAPES_CUDA_CHECK(cudaMemcpyAsync(cur_level.root_rule, h_rule.data(), cur.n * 4, cudaMemcpyHostToDevice, st));
    TAPES_CUDA_CHECK(cudaMemcpyAsync(cur.meta, h_meta.data(), cur.n, cudaMemcpyHostToDevice, st));
    TAPES_CUDA_CHECK(cudaMemcpyAsync(cur.flags, h_fl.data(), cur.n, cudaMemcpyHostToDevice, st));
    TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  }

  uint64_t seed_bits = 0;
  while ((1ull << seed_bits) < std::max<uint64_t>(n_seeds, 1)) ++seed_bits;
  uint64_t prefix_bits = 0;
  while ((1ull << prefix_bits) < M) ++prefix_bits;
  const uint64_t significant = ((prefix_bits ? ((1ull << prefix_bits) - 1) : 0ull)) |
                               ((seed_bits ? ((1ull << seed_bits) - 1) : 0ull) << 32);

  uint64_t total_terms = 0, total_edges = 0;
  uint64_t node_limit = 0x7fffffffull;  // node id + sign bit in 32 bits; lowered by tests of the splitting
  int hash_vote = -1;  // -1: by table shape
  bool virtual_right_levels = true;  // TAPES_VIRTUAL_RIGHT_LEVELS=0: write every level out (A/B and tests)
  if (const char* e = std::getenv("TAPES_VIRTUAL_RIGHT_LEVELS")) virtual_right_levels = std::atoi(e) != 0;
  int hash_run_bits = 2;
  if (const char* e = std::getenv("TAPES_HASH_RUN_BITS")) hash_run_bits = std::max(0, std::min(8, std::atoi(e)));
  int hash_guess = 1;  // 0: always 2 n slots; 1: sized on a guess; 2: start every level with the smallest table (tests of the retry)
  if (const char* e = std::getenv("TAPES_HASH_GUESS")) hash_guess = std::atoi(e);
  uint64_t prev_groups = 0;
  unsigned long long chain_starts = 0;
  bool chain_starts_known = false;
  const bool build_trace = std::getenv("TAPES_BUILD_TRACE") != nullptr;
  if (const char* e = std::getenv("TAPES_HASH_VOTE")) hash_vote = std::atoi(e);
  if (const char* e = std::getenv("TAPES_MAX_NODES")) node_limit = std::min<uint64_t>(node_limit, std::strtoull(e, nullptr, 10));
  while (cur.n > 0) {
    if (cur_level.base + cur.n >= node_limit) throw TooLarge("extension forest exceeds 2^31 nodes");
    m.stats.levels++;
    const uint64_t n = cur.n;

    // pass 1: classify + hash-dedup of the right-chain prefixes.  A table of 2 n slots is always enough
    // but several times larger than needed (most prefixes are reached from A parents), which costs in
    // cleared and randomly touched memory: the table is sized on an estimate of the number of distinct
    // prefixes instead, and the pass starts over with the safe size when the estimate was too small.
    uint64_t cap_safe = 1024;
    while (cap_safe < 2 * n) cap_safe <<= 1;
    uint64_t cap = cap_safe;
    if (hash_guess == 2) cap = 1024;
    else if (hash_guess && chain_starts_known) {
      // nodes that start right chains bring one prefix each at most; the right children of the G groups of
      // the level before bring A G prefixes that usually merge A to one (same prefix but for the first symbol)
      const double guess = (double)chain_starts + 1.5 * (double)prev_groups + 4096.0;
      cap = 1024;
      while (cap < cap_safe && (double)cap < 2.0 * guess) cap <<= 1;
    }
    HashSet hs;
    uint32_t *ltflag = nullptr, *keyslot = nullptr;
    uint64_t *ltrank = nullptr, *scan_tmp = nullptr, *keys_a = nullptr;
    // [0] unique keys, [1] irregular groups, [2] parents, [3] table too small, [4] nodes of the next level that start right chains
    unsigned long long* counters = nullptr;
    uint64_t NL = 0, NT = 0, NG = 0;
    for (;;) {
      s1.reset();
      const size_t i_keys = s1.want(cap * 8), i_vals = s1.want(cap * 4), i_ranks = s1.want(cap * 4);
      const uint64_t n_ranked = cur.right_only ? 0 : n;  // right-only levels rank nothing
      const size_t i_ltflag = s1.want(n_ranked * 4), i_kflag = s1.want(n * 4);
      const size_t i_ltrank = s1.want((n_ranked + 1) * 8);
      const size_t i_scan = s1.want(scan_tmp_elems(std::max<uint64_t>(n, 256ull * 1184)) * 8);
      const size_t i_unique = s1.want(n * 8), i_counters = s1.want(64);
      s1.commit();
      hs.keys = s1.at<uint64_t>(i_keys); hs.vals = s1.at<uint32_t>(i_vals); hs.ranks = s1.at<uint32_t>(i_ranks);
      hs.mask = cap - 1;
      ltflag = s1.at<uint32_t>(i_ltflag);
      keyslot = s1.at<uint32_t>(i_kflag);
      ltrank = s1.at<uint64_t>(i_ltrank);
      scan_tmp = s1.at<uint64_t>(i_scan);
      keys_a = s1.at<uint64_t>(i_unique);
      counters = s1.at<unsigned long long>(i_counters);
      hs.full = counters + 3;
      hs.run_bits = hash_run_bits;
      TAPES_CUDA_CHECK(cudaMemsetAsync(hs.keys, 0xff, cap * 8, st));
      TAPES_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64, st));
      // lanes of a warp hold different prefixes unless the level is tiny or the table short (M < 32):
      // the warp vote that merges equal keys before the insertion only pays there
      const bool vote = hash_vote < 0 ? c.M < 32 : hash_vote != 0;
      const unsigned grid = grid_for(n, kThreads);
      if (cur.right_only) {  // nothing to rank: no node has left children or a stored term
        if (vote) classify_kernel<true, true><<<grid, kThreads, 0, st>>>(cur, c, hs, ltflag, keyslot, keys_a, counters);
        else classify_kernel<false, true><<<grid, kThreads, 0, st>>>(cur, c, hs, ltflag, keyslot, keys_a, counters);
      } else {
        if (vote) classify_kernel<true, false><<<grid, kThreads, 0, st>>>(cur, c, hs, ltflag, keyslot, keys_a, counters);
        else classify_kernel<false, false><<<grid, kThreads, 0, st>>>(cur, c, hs, ltflag, keyslot, keys_a, counters);
        exclusive_scan_u32<true>(ltflag, n, ltrank, scan_tmp, st);
      }
      uint64_t h_tot[3] = {0, 0, 0};
      if (!cur.right_only) TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_tot[0], ltrank + n, 8, cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_tot[1], counters + 3, 8, cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_tot[2], counters, 8, cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
      NL = (uint32_t)h_tot[0]; NT = 2 * (h_tot[0] >> 32); NG = h_tot[2];  // NT counts edges, two per stored term
      if (build_trace)
        std::fprintf(stderr, "[tapes] level %d: %llu nodes, %llu left parents, %llu prefixes, table %llu slots%s\n",
                     (int)m.stats.levels, (unsigned long long)n, (unsigned long long)NL, (unsigned long long)NG,
                     (unsigned long long)cap, (cap != cap_safe && (h_tot[1] != 0 || NG * 10 > cap * 7)) ? " (too small)" : "");
      if (cap == cap_safe && h_tot[1] != 0)  // cannot happen with 2 n slots for at most n keys unless probing degenerates
        throw std::runtime_error("the prefix table of the expansion overflowed at its safe size");
      if (cap == cap_safe || (h_tot[1] == 0 && NG * 10 <= cap * 7)) break;
      cap = cap_safe;  // the guess was too small: the slots handed out are not to be trusted
      m.stats.hash_retries++;
    }
    prev_groups = NG;
    chain_starts_known = false;  // read below when this level has prefix groups; otherwise the next table gets the safe size
    if ((NL + NG) * (uint64_t)m.A >= 0xffffffffull) throw TooLarge("a level of the extension forest exceeds 2^32 nodes");

    // pass 2: children, parent records, flux edges
    Frontier next;
    next.n = (NL + NG) * (uint64_t)m.A;
    Level next_level;
    next_level.base = cur_level.base + n;
    next_level.n_left = (uint32_t)NL;
    next_level.n_groups = (uint32_t)NG;
    const RadixPlan plan = radix_plan(std::max<uint64_t>(NG, 1));
    s2.reset();  // owns `next` until the end of the next level; the memory of the level before is reused
    // a level without left parents is followed by right children only: that level is not written out
    // (Frontier::right_only), only the seed of every group is kept beside the model's group arrays
    next.right_only = NL == 0 && virtual_right_levels;
    size_t fi[5];
    plan_frontier(s2, next.right_only ? 0 : next.n, fi);
    const size_t i_gseed = s2.want(next.right_only ? NG * 4 : 0);
    const size_t i_keys_b = s2.want(NG * 8), i_rh = s2.want(256ull * plan.blocks * 4);
    const size_t i_ro = s2.want((256ull * plan.blocks + 1) * 8), i_keyrank = s2.want(n * 4), i_gstat = s2.want(NG * 16);
    const size_t i_consumed = s2.want((size_t)cur_level.n_groups * 4);
    const size_t i_gptr = s2.want((NG + 1) * 8), i_gparents = s2.want(n * 4);  // parent lists (at most n parents)
    s2.commit();
    bind_frontier(s2, fi, next);
    uint32_t* keyrank = s2.at<uint32_t>(i_keyrank);

    // unique right-chain prefixes in canonical (seed, prefix) order, and their rank by slot
    uint64_t* sorted = keys_a;
    if (NG) {
      sorted = radix_sort_u64(keys_a, s2.at<uint64_t>(i_keys_b), NG, significant, s2.at<uint32_t>(i_rh),
                              s2.at<uint64_t>(i_ro), scan_tmp, st);
      // the children of the groups; records every group's rank in the table for emit_kernel
      next_level.g_prefix = dkeep<uint32_t>(m, NG);
      next_level.g_adjusted = dkeep<uint32_t>(m, NG);
      uint32_t* g_seed = next.right_only ? s2.at<uint32_t>(i_gseed) : nullptr;
      emit_groups_kernel<<<grid_for(NG, kThreads), kThreads, 0, st>>>(sorted, (uint32_t)NG, hs, c, next,
                                                                     NL * (uint64_t)m.A, next_level.g_prefix,
                                                                     next_level.g_adjusted, g_seed);
      next.g_prefix = next_level.g_prefix; next.g_adjusted = next_level.g_adjusted; next.g_seed = g_seed;
    }
    if (NL) {
      next_level.lp_gid = dkeep<uint32_t>(m, NL);
      next_level.lp_io = dkeep<uint32_t>(m, NL);
      next_level.lp_len = dkeep<uint8_t>(m, NL);
    }
    uint4* gstat = s2.at<uint4>(i_gstat);  // per group: smallest parent id, ~largest, ~count (emit_kernel)
    if (NG) TAPES_CUDA_CHECK(cudaMemsetAsync(gstat, 0xff, NG * 16, st));
    EdgeChunk ec{nullptr, nullptr, NT};  // NT = stored flux edges of this level (two per term that is not a right child)
    if (NT) {  // one allocation for both arrays, owned by the list from here on
      ec.row = dtemp<uint32_t>(2 * NT); ec.val = ec.row + NT;
      edge_chunks.push_back(ec);
    }
    if (cur.right_only)
      emit_right_only_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(n, cur_level.base, keyslot, hs, keyrank, gstat);
    else
      emit_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(cur, c, cur_level.base, ltflag, keyslot, ltrank,
                                                            NL, hs, next, next_level.lp_gid, next_level.lp_io,
                                                            next_level.lp_len, ec.row, ec.val, keyrank, gstat, counters + 4);
    if (NG) {
      // parent lists of the prefix groups: as progressions (first, stride, count) when every list is
      // one - found from the extremes and counts emit_kernel collected, then checked parent by parent
      next_level.g_first = dkeep<uint32_t>(m, NG);
      next_level.g_stride = dkeep<uint32_t>(m, NG);
      next_level.g_count = dkeep<uint32_t>(m, NG);
      derive_progressions_kernel<<<grid_for(NG, kThreads), kThreads, 0, st>>>(
          gstat, NG, next_level.g_first, next_level.g_stride, next_level.g_count, counters);
      verify_progressions_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(
          keyrank, n, cur_level.base, next_level.g_first, next_level.g_stride, counters);
      unsigned long long h_counts[2] = {0, 0};  // irregular groups or parents off their progression, parents
      TAPES_CUDA_CHECK(cudaMemcpyAsync(h_counts, counters + 1, 16, cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaMemcpyAsync(&chain_starts, counters + 4, 8, cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
      chain_starts_known = true;
      const uint64_t n_par = h_counts[1];
      next_level.n_group_parents = n_par;
      if (h_counts[0] != 0 || std::getenv("TAPES_KEEP_PARENT_LISTS")) {
        // some list is not a progression: this level keeps explicit lists, ascending inside each group
        uint32_t* cnt = (uint32_t*)gstat;  // the records have been read: their memory serves as the fill cursors
        uint64_t* g_ptr = s2.at<uint64_t>(i_gptr);
        uint32_t* g_parents = s2.at<uint32_t>(i_gparents);
        exclusive_scan_u32(next_level.g_count, NG, g_ptr, scan_tmp, st);
        TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, NG * 4, st));
        group_fill_ids_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(keyrank, n, cur_level.base, g_ptr, cnt, g_parents);
        sort_groups(g_ptr, NG, g_parents, st);
        next_level.g_first = next_level.g_stride = next_level.g_count = nullptr;  // their memory stays in the arena
        next_level.g_ptr = dkeep<uint64_t>(m, NG + 1);
        next_level.g_parents = dkeep<uint32_t>(m, n_par);
        TAPES_CUDA_CHECK(cudaMemcpyAsync(next_level.g_ptr, g_ptr, (NG + 1) * 8, cudaMemcpyDeviceToDevice, st));
        TAPES_CUDA_CHECK(cudaMemcpyAsync(next_level.g_parents, g_parents, n_par * 4, cudaMemcpyDeviceToDevice, st));
        m.stats.irregular_levels++;
      }
      m.stats.hash_inserts += (int64_t)n_par;
      m.stats.hash_unique += (int64_t)NG;
      m.stats.sum_nodes += (int64_t)NG;
      // fused right chain: groups whose parents are right children of this level take them over
      const bool fuse = !(std::getenv("TAPES_LEVEL_FUSE") && std::atoi(std::getenv("TAPES_LEVEL_FUSE")) == 0);
      if (fuse && next_level.g_first && cur_level.g_first && cur_level.n_groups) {
        const uint64_t PG = cur_level.n_groups;
        const uint64_t right_base = cur_level.base + (uint64_t)m.A * cur_level.n_left;
        uint32_t* consumed = s2.at<uint32_t>(i_consumed);
        TAPES_CUDA_CHECK(cudaMemsetAsync(consumed, 0, PG * 4, st));
        TAPES_CUDA_CHECK(cudaMemsetAsync(counters, 0, 16, st));
        mark_owned_parents_kernel<<<grid_for(NG, kThreads), kThreads, 0, st>>>(
            next_level.g_first, next_level.g_stride, next_level.g_count, next_level.g_prefix, cur_level.g_prefix,
            NG, right_base, (uint32_t)m.A, c.M, consumed);
        check_consumed_kernel<<<grid_for(PG, kThreads), kThreads, 0, st>>>(consumed, PG, (uint32_t)m.A, counters);
        unsigned long long h_partial = 0;
        TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_partial, counters, 8, cudaMemcpyDeviceToHost, st));
        TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_partial != 0) {  // some group would be half taken over: leave this pair of levels as it is
          clear_owned_parents_kernel<<<grid_for(NG, kThreads), kThreads, 0, st>>>(next_level.g_count, NG);
        } else {
          mark_deferred_kernel<<<grid_for(PG, kThreads), kThreads, 0, st>>>(cur_level.g_count, consumed, PG,
                                                                           (uint32_t)m.A, counters + 1);
          unsigned long long h_deferred = 0;
          TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_deferred, counters + 1, 8, cudaMemcpyDeviceToHost, st));
          TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
          if (h_deferred) {
            next_level.prev_right_base = right_base;
            next_level.prev_prefix = cur_level.g_prefix;
            next_level.prev_total = kLinkLater;  // the previous level's sums: all levels' sums share one array, made below
            m.stats.deferred_groups += (int64_t)h_deferred;
            m.stats.owned_parents += (int64_t)h_deferred * m.A;
          }
        }
      }
    }
    TAPES_CUDA_CHECK(cudaGetLastError());
    TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
    total_edges += NT;
    total_terms += NT / 2 + (uint64_t)cur_level.n_groups * (uint64_t)m.A;  // every right child is a term
    m.stats.left_parents += (int64_t)NL;

    m.levels.push_back(cur_level);
    cur_slab.swap(s2);  // ownership of `next` moves with the frontier; the old frontier's memory becomes scratch
    cur = next;
    cur_level = next_level;
  }
  if (cur_slab.capacity + s1.capacity + s2.capacity > ((size_t)1 << 30)) {
    cur_slab.release(); s1.release(); s2.release();
  }
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  m.stats.device_expand_ms = ms_since(t_expand);
  m.stats.expand_alloc_ms = g_alloc_ms;
  auto t_lists = std::chrono::steady_clock::now();
  auto trace_phase = [&](const char* what, std::chrono::steady_clock::time_point& since) {
    if (!build_trace) return;
    TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
    std::fprintf(stderr, "[tapes] %s: %.2f ms\n", what, ms_since(since));
    since = std::chrono::steady_clock::now();
  };
  auto t_phase = t_lists;
  // ---- the sums of all prefix groups in one array; per-prefix lists of the groups (right-chain outflow) ----
  // A right child of group g with prefix q leaves row q * A + x with weight sum(g) * ratio[q * A + x]: all
  // right children of all groups with prefix q share the factor ratio[row], so the rows' outflow
  // through right children is ratio[row] * (sum over those groups' sums) - one number per prefix,
  // out_sum[q], added up per step by prefix_sums_kernel in a fixed order.  These 40 % of all flux
  // edges are therefore not stored in the flux structure at all.
  {
    uint64_t n_all = 0;
    for (const Level& lv : m.levels) n_all += lv.n_groups;
    m.n_groups_all = n_all;
    if (n_all >= 0xffffffffull) throw TooLarge("more than 2^32 prefix groups");
    if (n_all) {
      m.g_total_all = dkeep<double>(m, n_all);
      uint64_t at = 0;
      for (size_t l = 0; l < m.levels.size(); ++l) {
        Level& lv = m.levels[l];
        if (lv.n_groups) lv.g_total = m.g_total_all + at;
        if (lv.prev_total == kLinkLater) lv.prev_total = m.levels[l - 1].g_total;
        at += lv.n_groups;
      }
      const uint64_t B = M;  // prefixes
      uint32_t* cnt = dalloc<uint32_t>(B, st);
      TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, B * 4, st));
      for (const Level& lv : m.levels)
        if (lv.n_groups) group_count_kernel<<<grid_for(lv.n_groups, kThreads), kThreads, 0, st>>>(lv.g_prefix, lv.n_groups, cnt);
      m.out_ptr = dkeep<uint64_t>(m, B + 1);
      uint64_t* scan_tmp = dalloc<uint64_t>(scan_tmp_elems(B), st);
      exclusive_scan_u32(cnt, B, m.out_ptr, scan_tmp, st);
      m.out_ids = dkeep<uint32_t>(m, n_all);
      TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, B * 4, st));
      at = 0;
      for (const Level& lv : m.levels) {
        if (lv.n_groups)
          group_fill_ids_kernel<<<grid_for(lv.n_groups, kThreads), kThreads, 0, st>>>(lv.g_prefix, lv.n_groups, at, m.out_ptr, cnt,
                                                                                    m.out_ids);
        at += lv.n_groups;
      }
      sort_groups(m.out_ptr, B, m.out_ids, st);  // ascending group number: the order of the additions
      m.out_sum = dkeep<double>(m, B);
      // inflow: the same groups listed by their ADJUSTED prefix, each with its own prefix
      TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, B * 4, st));
      for (const Level& lv : m.levels)
        if (lv.n_groups) group_count_kernel<<<grid_for(lv.n_groups, kThreads), kThreads, 0, st>>>(lv.g_adjusted, lv.n_groups, cnt);
      m.in_ptr = dkeep<uint64_t>(m, B + 1);
      exclusive_scan_u32(cnt, B, m.in_ptr, scan_tmp, st);
      uint32_t* in_ids = dalloc<uint32_t>(n_all, st);
      m.in_pairs = dkeep<uint2>(m, n_all);
      TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, B * 4, st));
      uint32_t* prefix_all = dalloc<uint32_t>(n_all, st);  // the levels' g_prefix, one after the other
      at = 0;
      for (const Level& lv : m.levels) {
        if (lv.n_groups) {
          group_fill_ids_kernel<<<grid_for(lv.n_groups, kThreads), kThreads, 0, st>>>(lv.g_adjusted, lv.n_groups, at, m.in_ptr, cnt,
                                                                                    in_ids);
          TAPES_CUDA_CHECK(cudaMemcpyAsync(prefix_all + at, lv.g_prefix, (size_t)lv.n_groups * 4, cudaMemcpyDeviceToDevice, st));
        }
        at += lv.n_groups;
      }
      sort_groups(m.in_ptr, B, in_ids, st);
      pair_with_prefix_kernel<<<grid_for(n_all, kThreads), kThreads, 0, st>>>(prefix_all, in_ids, m.in_pairs, n_all);
      // does any level read stored weights of right children (parents of a group that does not own them)?
      unsigned long long* readers = dalloc<unsigned long long>(1, st);
      TAPES_CUDA_CHECK(cudaMemsetAsync(readers, 0, 8, st));
      bool explicit_lists = false;
      for (size_t l = 1; l < m.levels.size(); ++l) {
        const Level& lv = m.levels[l];
        if (!lv.n_groups) continue;
        if (!lv.g_first) { explicit_lists = true; continue; }
        const uint64_t prev_right_base = m.levels[l - 1].base + (uint64_t)m.A * m.levels[l - 1].n_left;
        count_right_readers_kernel<<<grid_for(lv.n_groups, kThreads), kThreads, 0, st>>>(lv.g_first, lv.g_stride, lv.g_count, lv.n_groups,
                                                                                       prev_right_base, readers);
      }
      unsigned long long h_readers = 0;
      TAPES_CUDA_CHECK(cudaMemcpyAsync(&h_readers, readers, 8, cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
      m.materialize_right = (h_readers != 0 || explicit_lists) ? 1 : 0;
      if (const char* e = std::getenv("TAPES_MATERIALIZE_RIGHT")) m.materialize_right = std::atoi(e) != 0 || m.materialize_right;
      dfree(prefix_all, st); dfree(readers, st); dfree(in_ids, st);
      dfree(cnt, st); dfree(scan_tmp, st);
      TAPES_CUDA_CHECK(cudaGetLastError());
      TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
    }
  }

  trace_phase("per-prefix lists of the groups", t_phase);
  // seeds walk through p together (Level::block_order): sort the blocks of 256 groups of every
  // level by the prefix they start at; a level whose order comes out as the identity keeps none.
  // The same pass finds the regular blocks (Level::plane_blocks) and lists the others.  Everything
  // stays on the device; the host reads two numbers per level at the end.
  {
    std::vector<size_t> which;  // levels with more than one block
    for (size_t l = 0; l < m.levels.size(); ++l)
      if (m.levels[l].g_prefix && m.levels[l].n_groups > (uint32_t)kThreads) which.push_back(l);
    unsigned long long* facts = which.empty() ? nullptr : dalloc<unsigned long long>(2 * which.size(), st);
    if (facts) TAPES_CUDA_CHECK(cudaMemsetAsync(facts, 0, 16 * which.size(), st));
    std::vector<uint32_t*> orders(which.size()), generals(which.size());
    std::vector<Level::PlaneBlock*> planes(which.size());
    for (size_t w = 0; w < which.size(); ++w) {
      Level& lv = m.levels[which[w]];
      const size_t n_blocks = ((size_t)lv.n_groups + kThreads - 1) / kThreads;
      const unsigned grid = grid_for(n_blocks, kThreads);
      Level::PlaneBlock* d_rec = dalloc<Level::PlaneBlock>(n_blocks, st);
      uint32_t* d_flag = dalloc<uint32_t>(n_blocks, st);
      classify_plane_blocks_kernel<<<(unsigned)n_blocks, kThreads, 0, st>>>(lv, c, (uint32_t)n_blocks, d_rec, d_flag);
      const RadixPlan plan = radix_plan(n_blocks);
      uint64_t* keys = dalloc<uint64_t>(n_blocks, st);
      uint64_t* keys_alt = dalloc<uint64_t>(n_blocks, st);
      uint32_t* hist = dalloc<uint32_t>(256ull * plan.blocks, st);
      uint64_t* offsets = dalloc<uint64_t>(256ull * plan.blocks + 1, st);
      uint64_t* scan_tmp = dalloc<uint64_t>(scan_tmp_elems(std::max<uint64_t>(n_blocks, 256ull * plan.blocks)), st);
      plane_keys_kernel<<<grid, kThreads, 0, st>>>(d_rec, (uint32_t)n_blocks, keys);
      uint64_t block_bits = 1;
      while ((1ull << block_bits) < n_blocks) ++block_bits;
      const uint64_t key_bits = ((1ull << block_bits) - 1) | (((prefix_bits ? ((1ull << prefix_bits) - 1) : 0ull)) << 32);
      uint64_t* sorted = radix_sort_u64(keys, keys_alt, n_blocks, key_bits, hist, offsets, scan_tmp, st);
      orders[w] = dkeep<uint32_t>(m, n_blocks);
      planes[w] = dkeep<Level::PlaneBlock>(m, n_blocks);  // room for the case that every block is one
      generals[w] = dkeep<uint32_t>(m, n_blocks);
      uint32_t* taken = dalloc<uint32_t>(n_blocks, st);
      uint64_t* position = dalloc<uint64_t>(n_blocks + 1, st);
      plane_order_kernel<<<grid, kThreads, 0, st>>>(sorted, (uint32_t)n_blocks, d_flag, orders[w], taken, facts + 2 * w);
      exclusive_scan_u32(taken, n_blocks, position, scan_tmp, st);
      plane_partition_kernel<<<grid, kThreads, 0, st>>>(orders[w], taken, position, d_rec, (uint32_t)n_blocks, planes[w],
                                                       generals[w], facts + 2 * w);
      dfree(d_rec, st); dfree(d_flag, st); dfree(keys, st); dfree(keys_alt, st); dfree(hist, st); dfree(offsets, st);
      dfree(scan_tmp, st); dfree(taken, st); dfree(position, st);
    }
    if (facts) {
      std::vector<unsigned long long> h_facts(2 * which.size());
      TAPES_CUDA_CHECK(cudaMemcpyAsync(h_facts.data(), facts, 16 * which.size(), cudaMemcpyDeviceToHost, st));
      TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
      dfree(facts, st);
      for (size_t w = 0; w < which.size(); ++w) {
        Level& lv = m.levels[which[w]];
        const size_t n_blocks = ((size_t)lv.n_groups + kThreads - 1) / kThreads;
        const uint64_t n_planes = h_facts[2 * w], moved = h_facts[2 * w + 1];
        if (moved) lv.block_order = orders[w];
        if (n_planes) {
          lv.plane_blocks = planes[w]; lv.n_plane_blocks = (uint32_t)n_planes;
          lv.general_blocks = generals[w]; lv.n_general_blocks = (uint32_t)(n_blocks - n_planes);
          m.stats.plane_groups += (int64_t)n_planes * kThreads;
        }
      }
    }
    TAPES_CUDA_CHECK(cudaGetLastError());
  }
  m.n_nodes = cur_level.base;  // base of the (empty) level after the last
  m.stats.nodes = (int64_t)m.n_nodes;
  m.stats.terms = (int64_t)total_terms;
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  const double lists_and_blocks_ms = ms_since(t_lists);  // per-prefix lists, block order, plane blocks

  trace_phase("block order and plane blocks", t_phase);
  // ---- CSR assembly: count per state, scan, fill, sort inside each row ----
  auto t_csr = std::chrono::steady_clock::now();
  const uint64_t n = m.n_states;
  m.nnz = 2 * total_terms;
  m.nnz_stored = total_edges;
  {
    uint32_t* cnt = dalloc<uint32_t>(n, st);
    TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, n * 4, st));
    for (const EdgeChunk& ec : edge_chunks)
      group_count_kernel<<<grid_for(ec.n, kThreads), kThreads, 0, st>>>(ec.row, ec.n, cnt);
    m.row_ptr = dkeep<uint64_t>(m, n + 1);
    uint64_t* scan_tmp = dalloc<uint64_t>(scan_tmp_elems(n), st);
    exclusive_scan_u32(cnt, n, m.row_ptr, scan_tmp, st);
    trace_phase("CSR: count + scan", t_phase);
    m.entries = dtemp<uint32_t>(m.nnz_stored);
    trace_phase("CSR: allocation of the entries", t_phase);
    TAPES_CUDA_CHECK(cudaMemsetAsync(cnt, 0, n * 4, st));
    for (EdgeChunk& ec : edge_chunks) {
      group_fill_vals_kernel<<<grid_for(ec.n, kThreads), kThreads, 0, st>>>(ec.row, ec.val, ec.n, m.row_ptr, cnt, m.entries);
      TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
      cudaFree(ec.row);  // row and val share one allocation
      ec.row = ec.val = nullptr;
    }
    trace_phase("CSR: fill, edge chunks freed", t_phase);
    sort_groups(m.row_ptr, n, m.entries, st);
    trace_phase("CSR: sort inside the rows", t_phase);
    dfree(cnt, st); dfree(scan_tmp, st);
    TAPES_CUDA_CHECK(cudaGetLastError());
  }
  const double per_row = n ? (double)m.nnz_stored / (double)n : 0.0;
  m.spmv_group = per_row > 128 ? 4 : (per_row > 64 ? 2 : 1);
  if (const char* g = std::getenv("TAPES_SPMV_LANES")) {
    const int v = std::atoi(g);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) m.spmv_group = v;
  }
  // batch size that wastes the fewest slots of the last batch; measured on B200 (A = 10, n = 1e8,
  // 24 rules): 1: 7.79, 2: 6.29, 4: 6.05, 5: 5.48, 8: 6.40 ms
  m.level_unroll = m.A <= 2 ? 2 : ((m.A + 4) / 5 * 5 - m.A <= (m.A + 3) / 4 * 4 - m.A ? 5 : 4);
  if (const char* g = std::getenv("TAPES_LEVEL_UNROLL")) m.level_unroll = std::max(1, std::atoi(g));
  if (const char* g = std::getenv("TAPES_FLUX_UNROLL")) m.flux_unroll = std::atoi(g);
  if (const char* g = std::getenv("TAPES_INTERLEAVE_SEEDS")) m.interleave_seeds = std::atoi(g) != 0;
  if (const char* g = std::getenv("TAPES_GRAPHS")) m.use_graphs = std::atoi(g) != 0;
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  m.stats.device_csr_ms = ms_since(t_csr) + lists_and_blocks_ms;

  // ---- the form the product kernel streams: slices of 32 states (flux.cu) ----
  m.flux_format = 1;
  if (const char* f = std::getenv("TAPES_FLUX_FORMAT")) m.flux_format = std::strcmp(f, "csr") == 0 ? 0 : 1;
  if (m.flux_format == 1) {
    int min_lanes = 32;  // measured (n = 1e8, 24 rules): 32: 4.22, 16: 4.23, 8: 4.54, 4: 4.76, 2: 6.97 ms; CSR 6.10 ms
    if (const char* g = std::getenv("TAPES_RUN_MIN_LANES")) min_lanes = std::atoi(g);
    build_flux_slices(m, min_lanes, st);
    cudaFree(m.entries);  // tapes_export_csr rebuilds the plain entries from the slices
    m.entries = nullptr;
  }

  // ---- per-step buffers ----
  m.marg_total = 0;
  for (int L = 0; L < m.k; ++L) { m.marg_off[L] = m.marg_total; m.marg_total += m.pow_a[L]; }
  m.marg_off[m.k] = m.marg_total;
  for (int L = m.k + 1; L < 40; ++L) m.marg_off[L] = 0;
  m.marg = dkeep<double>(m, m.marg_total);
  m.d_marg_off = dkeep<uint64_t>(m, 40);
  TAPES_CUDA_CHECK(cudaMemcpyAsync(m.d_marg_off, m.marg_off, 40 * 8, cudaMemcpyHostToDevice, st));
  m.node_w = dkeep<double>(m, m.n_nodes);
  if (const char* g = std::getenv("TAPES_RATIO_TABLE")) m.ratio_table = std::atoi(g) != 0;
  bool any_groups = false;
  for (const Level& lv : m.levels) any_groups = any_groups || lv.n_groups > 0;
  if (any_groups && m.k >= 2) m.ratio_right = dkeep<double>(m, m.n_states);  // right-extension ratios, per step
  bool any_full_left = false;  // left parents whose children have a full window
  for (const Level& lv : m.levels) any_full_left = any_full_left || lv.n_left > 0;
  // Off unless asked for: at the bench size the second table costs 0.21 ms per step (0.8 GB written) and
  // the division-free left part of the level kernel gains 0.11 ms (profiles/r02_c_sweep*.log): the
  // level kernel is not bound by its divisions.
  if (m.ratio_right && any_full_left && std::getenv("TAPES_RATIO_LEFT") && std::atoi(std::getenv("TAPES_RATIO_LEFT")) != 0)
    m.ratio_left = dkeep<double>(m, m.n_states);
  if (const char* g = std::getenv("TAPES_PLANE_KERNEL")) m.plane_kernel = std::atoi(g) != 0;
  if (const char* g = std::getenv("TAPES_FUSE_MARGINAL_RATIO")) m.fuse_marginal_ratio = std::atoi(g) != 0;
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  // Tiny problems: one launch per right-hand side by ONE thread block.  Work = what the phases touch.
  // Measured on B200 (profiles/r02_e_time_small.log, device-resident, us per right-hand side, multi-launch
  // graph replay -> single launch): ex2 k=3 (work 136) 16.4 -> 11.2, ex2 k=7 (3.7e3) 34.8 -> 22.5; ex3 k=6
  // (4.1e4) 38.3 -> 36.1 with a cluster of 4; ex5 (1.1e6) 79.4 -> 112 and ex4 (3.0e6) 71.7 -> 100 with a
  // cluster of 16: beyond a few thousand nodes 16 SMs lose against the whole GPU even with ~4 us of launch
  // latency per kernel, so only the tiny ones take the single launch by default (clusters stay
  // available through "fused_cluster" / TAPES_FUSED_CLUSTER).
  {
    const uint64_t work = m.n_nodes + m.nnz + m.n_states;
    m.fused_cluster = work <= (1ull << 13) ? 1 : 0;
    if (m.flux_format != 1 || m.n_rules == 0 || m.k < 1) m.fused_cluster = 0;
    if (const char* e = std::getenv("TAPES_FUSED_CLUSTER")) {
      const int v = std::atoi(e);
      if (v == 0 || v == 1 || v == 2 || v == 4 || v == 8 || v == 16) m.fused_cluster = m.flux_format == 1 && m.n_rules ? v : 0;
    }
    if (const char* e = std::getenv("TAPES_FUSED_SMALL")) m.fused_small = std::atoi(e) != 0;
  }
  m.launches_per_rhs = rhs_launch_count(m);
  return mp;
}

namespace {
// Tables with at most this many entries are produced by the single-block tail kernel.
constexpr uint64_t kTailEntries = 4096;
int marginal_tail_top(const Model& m) {
  int top = -1;
  for (int L = 0; L < m.k; ++L)
    if (m.pow_a[L] <= kTailEntries) top = L;
  return top;
}

void launch_weights(Model& m, const double* d_p, cudaStream_t st, cudaEvent_t* ev) {
  const Consts c = make_consts(m);
  Tables t;
  t.p = d_p; t.marg = m.marg; t.k = m.k;
  for (int i = 0; i < 34; ++i) t.off[i] = i <= m.k ? m.marg_off[i] : 0;

  // marginal tables, longest first
  const int top = marginal_tail_top(m);
  const bool need_ratio = m.ratio_right && m.k >= 2;     // the product reads the ratios (right-chain outflow)
  const bool use_ratio = m.ratio_table && need_ratio;    // ... and the levels, unless switched off
  // the longest marginal table and the right-extension ratios come from one pass over p when both
  // are wanted and a tile of 256 rows fits shared memory
  const size_t tile_bytes = ((size_t)kRowsPerBlock * (c.A | 1u) + kRowsPerBlock) * sizeof(double);
  const bool one_pass = need_ratio && !m.ratio_left && m.fuse_marginal_ratio && m.k - 1 > top && tile_bytes <= 48 * 1024;
  for (int L = m.k - 1; L > top; --L) {
    const double* src = (L + 1 == m.k) ? d_p : m.marg + m.marg_off[L + 1];
    if (one_pass && L == m.k - 1) {
      // bulk copies need 16-byte aligned pieces of a multiple of 16 bytes: every block's stretch starts
      // at a multiple of 256 * A entries; the last block's length is rows * A entries
      const uint64_t last_rows = m.pow_a[L] % kRowsPerBlock;
      const bool bulk = ((uintptr_t)d_p % 16 == 0) && ((last_rows * c.A) % 2 == 0) && !std::getenv("TAPES_NO_BULK_COPY");
      if (bulk)
        marginal_ratio_kernel<true><<<grid_for(m.pow_a[L], kRowsPerBlock), kThreads, tile_bytes, st>>>(
            d_p, m.marg + m.marg_off[L], m.ratio_right, m.pow_a[L], c.A);
      else
        marginal_ratio_kernel<false><<<grid_for(m.pow_a[L], kRowsPerBlock), kThreads, tile_bytes, st>>>(
            d_p, m.marg + m.marg_off[L], m.ratio_right, m.pow_a[L], c.A);
    } else
      marginal_kernel<<<grid_for(m.pow_a[L], kThreads), kThreads, 0, st>>>(src, m.marg + m.marg_off[L], m.pow_a[L], c.A);
  }
  if (top >= 0)
    marginal_tail_kernel<<<1, 1024, 0, st>>>(d_p, m.marg, m.d_marg_off, m.k, top, c.A);
  if (m.n_rules)
    rule_weight_kernel<<<grid_for(m.n_rules, 128), 128, 0, st>>>(t, m.n_rules, m.rule_ptr, m.step_kind, m.step_len,
                                                                m.step_long, m.step_short, m.step_prob, m.rule_w);
  if (need_ratio && !one_pass)
    ratio_tables_kernel<<<grid_for(m.n_states, kThreads * kRatioBatch), kThreads, 0, st>>>(
        d_p, m.marg + m.marg_off[m.k - 1], m.ratio_right, m.ratio_left, m.n_states, c.A, c.M);
  if (ev) TAPES_CUDA_CHECK(cudaEventRecord(ev[1], st));
  for (const Level& stored : m.levels) {
    Level lv = stored;
    if (!m.interleave_seeds) lv.block_order = nullptr;
    if (lv.n_roots) {
      root_kernel<<<grid_for(lv.n_roots, kThreads), kThreads, 0, st>>>(lv.root_rule, lv.n_roots, m.rule_w, m.node_w);
    } else if (lv.n_left + lv.n_groups) {
      const unsigned left_blocks = lv.n_left ? grid_for(lv.n_left, kThreads) : 0;
      unsigned group_blocks = lv.n_groups ? grid_for(((uint64_t)lv.n_groups + 31) / 32 * 32, kThreads) : 0;
      const bool planes = use_ratio && m.plane_kernel && lv.n_plane_blocks > 0;
      if (planes) {
        // the regular blocks of this level go to plane_kernel, level_kernel keeps the others
        lv.block_order = lv.general_blocks;
        group_blocks = lv.n_general_blocks;
#define TAPES_PLANE(A_)                                                                                               \
  (m.materialize_right ? plane_kernel<A_, true><<<lv.n_plane_blocks, kThreads, 0, st>>>(c, lv, m.ratio_right, m.node_w, m.node_w) \
                       : plane_kernel<A_, false><<<lv.n_plane_blocks, kThreads, 0, st>>>(c, lv, m.ratio_right, m.node_w, m.node_w))
        switch (c.A) {
          case 10: TAPES_PLANE(10); break;
          case 4: TAPES_PLANE(4); break;
          case 2: TAPES_PLANE(2); break;
          default: TAPES_PLANE(0); break;
        }
#undef TAPES_PLANE
      }
      const uint32_t q = 32u / c.A, r = 32u % c.A;
      const unsigned grid = left_blocks + group_blocks;
      if (grid == 0) continue;
      const bool prog = lv.g_first != nullptr;
      // U loads in flight per thread; a group that owns its parents evaluates UO of them at a time
      // (each needs three loads).  Measured at n = 1e8, A = 10 (profiles/r01_g_sweep_fused_right_chain.log):
      // (U, UO) = (5, 3) 5.42 ms, (5, 2) 5.47, (4, 2) 5.66, (2, 2) 5.83, (5, 5) 6.20 (spills), 64 registers
      // at 4 blocks per SM 5.63-5.94; 40 registers at 6 blocks per SM (120 B of spills) 5.37-5.55: no gain.
      // Separate lean kernels for the pure right-chain levels (one thread per group, 32-62 registers, 4-8
      // blocks per SM, then one thread per child) measured 5.59-6.35 ms against 5.51 for this kernel
      // (profiles/r01_m_sweep_lean_chain_kernels.log): the deep levels are not bound by registers.
#define TAPES_LEVEL_R(U_, UO_, B_, R_)                                                                       \
  (prog ? level_kernel<U_, UO_, true, B_, R_><<<grid, kThreads, 0, st>>>(t, c, lv, left_blocks, q, r, m.node_w, m.node_w, \
                                                                          m.ratio_right, m.ratio_left)             \
        : level_kernel<U_, 1, false, 5, R_><<<grid, kThreads, 0, st>>>(t, c, lv, left_blocks, q, r, m.node_w, m.node_w, \
                                                                        m.ratio_right, m.ratio_left))
#define TAPES_LEVEL(U_, UO_, B_) (use_ratio ? TAPES_LEVEL_R(U_, UO_, B_, true) : TAPES_LEVEL_R(U_, UO_, B_, false))
      if (m.level_unroll >= 8) TAPES_LEVEL(8, 4, 4);
      else if (m.level_unroll >= 5) TAPES_LEVEL(5, 3, 5);
      else if (m.level_unroll >= 4) TAPES_LEVEL(4, 2, 5);
      else if (m.level_unroll >= 2) TAPES_LEVEL(2, 2, 5);
      else TAPES_LEVEL(1, 1, 5);
#undef TAPES_LEVEL
#undef TAPES_LEVEL_R
    }
  }
  if (m.out_sum)  // what leaves the rows through right children: one sum per prefix
    prefix_sums_kernel<<<grid_for(m.pow_a[m.k - 1], kThreads), kThreads, 0, st>>>(m.out_ptr, m.out_ids, m.g_total_all, m.out_sum,
                                                                                 m.pow_a[m.k - 1]);
  TAPES_CUDA_CHECK(cudaGetLastError());
}

RightFlux right_flux_of(const Model& m) {
  RightFlux f;
  f.out_sum = m.out_sum; f.ratio = m.ratio_right; f.totals = m.g_total_all;
  f.in_ptr = m.in_ptr; f.in_pairs = m.in_pairs; f.A = (uint32_t)m.A;
  return f;
}

template <bool FUSED>
void launch_flux_impl(Model& m, double* d_out, uint64_t row_lo, uint64_t row_hi, cudaStream_t st,
                      const StageUpdate& up0, bool accumulate) {
  if (row_hi <= row_lo) return;
  if (m.flux_format == 1) {
    launch_flux_slices(m, d_out, row_lo, row_hi, st, FUSED ? &up0 : nullptr, accumulate);
    return;
  }
  const uint64_t rows = row_hi - row_lo;
  const uint64_t threads = rows * (uint64_t)m.spmv_group;
  const uint64_t* rp = m.row_ptr + row_lo;
  double* out = d_out + row_lo;
  StageUpdate up = up0;
  if (FUSED) {  // the kernel indexes every vector by the row inside the range
    for (int j = 0; j < up.n; ++j) up.vec[j] += row_lo;
    up.y += row_lo;
    up.stage += row_lo;
  }
  const unsigned grid = grid_for(threads, kThreads);
  const int acc = accumulate ? 1 : 0;
  const RightFlux of = right_flux_of(m);
  switch (m.spmv_group) {
    case 1: spmv_kernel<1, FUSED><<<grid, kThreads, 0, st>>>(rp, m.entries, m.node_w, out, rows, up, acc, of, row_lo); break;
    case 2: spmv_kernel<2, FUSED><<<grid, kThreads, 0, st>>>(rp, m.entries, m.node_w, out, rows, up, acc, of, row_lo); break;
    case 4: spmv_kernel<4, FUSED><<<grid, kThreads, 0, st>>>(rp, m.entries, m.node_w, out, rows, up, acc, of, row_lo); break;
    case 8: spmv_kernel<8, FUSED><<<grid, kThreads, 0, st>>>(rp, m.entries, m.node_w, out, rows, up, acc, of, row_lo); break;
    default: spmv_kernel<16, FUSED><<<grid, kThreads, 0, st>>>(rp, m.entries, m.node_w, out, rows, up, acc, of, row_lo); break;
  }
  TAPES_CUDA_CHECK(cudaGetLastError());
}

// The product of every part of a (possibly composite) model for a range of states: the first part
// writes, the others add to it in order.  `up` (may be null) rides on the last part's kernel, which
// is the one that sees the complete dy/dt.
void launch_flux(Model& m, double* d_out, uint64_t row_lo, uint64_t row_hi, cudaStream_t st,
                 const StageUpdate* up = nullptr) {
  const size_t parts = 1 + m.more.size();
  for (size_t i = 0; i < parts; ++i) {
    Model& part = i == 0 ? m : *m.more[i - 1];
    if (up && i + 1 == parts) launch_flux_impl<true>(part, d_out, row_lo, row_hi, st, *up, i > 0);
    else launch_flux_impl<false>(part, d_out, row_lo, row_hi, st, StageUpdate(), i > 0);
  }
}

// The whole right-hand side of a small single-structure model in one launch; false when the model is
// not eligible (the caller then takes the multi-launch path).
bool launch_fused(Model& m, const double* d_p, double* d_out, cudaStream_t st, const StageUpdate* up) {
  if (!m.fused_small || m.fused_cluster <= 0 || !m.more.empty() || m.flux_format != 1) return false;
  if (m.ratio_right == nullptr) {  // the fused kernel reads right-extension ratios from the table
    for (const Level& lv : m.levels)
      if (lv.n_groups) return false;
  }
  const Consts c = make_consts(m);
  if (!m.d_fused_levels) {
    std::vector<FusedLevel> host(m.levels.size());
    for (size_t i = 0; i < m.levels.size(); ++i) {
      FusedLevel& fl = host[i];
      fl.lv = m.levels[i];
      if (!m.interleave_seeds) fl.lv.block_order = nullptr;
      fl.left_blocks = fl.lv.n_left ? grid_for(fl.lv.n_left, kThreads) : 0;
      fl.group_blocks = fl.lv.n_groups ? grid_for(((uint64_t)fl.lv.n_groups + 31) / 32 * 32, kThreads) : 0;
      fl.warp_step_q = 32u / c.A; fl.warp_step_r = 32u % c.A;
    }
    FusedLevel* d = dkeep<FusedLevel>(m, host.size());
    TAPES_CUDA_CHECK(cudaMemcpy(d, host.data(), host.size() * sizeof(FusedLevel), cudaMemcpyHostToDevice));
    m.d_fused_levels = d;
    static bool allowed = false;
    if (!allowed) {
      TAPES_CUDA_CHECK(cudaFuncSetAttribute(fused_rhs_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      allowed = true;
    }
  }
  Tables t;
  t.p = d_p; t.marg = m.marg; t.k = m.k;
  for (int i = 0; i < 34; ++i) t.off[i] = i <= m.k ? m.marg_off[i] : 0;
  FusedArgs a;
  a.levels = (const FusedLevel*)m.d_fused_levels; a.n_levels = (int)m.levels.size();
  a.n_rules = m.n_rules; a.rule_ptr = m.rule_ptr; a.step_kind = m.step_kind; a.step_len = m.step_len;
  a.step_long = m.step_long; a.step_short = m.step_short; a.step_prob = m.step_prob; a.rule_w = m.rule_w;
  a.marg = m.marg; a.ratio_right = m.k >= 2 ? m.ratio_right : nullptr; a.node_w = m.node_w;
  a.slice_ptr = m.slices.slice_ptr; a.slice_runs = m.slices.slice_runs; a.words = m.slices.words;
  a.n_states = m.n_states; a.n_slices = m.slices.n_slices; a.out = d_out; a.fused_update = up ? 1 : 0;
  a.out_ptr = m.out_ptr; a.out_ids = m.out_ids; a.g_total_all = m.g_total_all; a.out_sum = m.out_sum;
  a.n_prefixes = m.pow_a[m.k - 1];
  a.in_ptr = m.in_ptr; a.in_pairs = m.in_pairs;
  StageUpdate upd = up ? *up : StageUpdate();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)m.fused_cluster); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (m.fused_cluster > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)m.fused_cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TAPES_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fused_rhs_kernel<true>, t, c, a, upd));
  } else {
    TAPES_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fused_rhs_kernel<false>, t, c, a, upd));
  }
  return true;
}

void launch_all_weights_plain(Model& m, const double* d_p, cudaStream_t st) {
  launch_weights(m, d_p, st, nullptr);
  for (auto& part : m.more) launch_weights(*part, d_p, st, nullptr);
}

// Replays (after capturing it on first use) the graph of launch_all_weights_plain for this input
// pointer; false when graphs do not apply and the caller has to launch directly.
bool launch_weights_graph(Model& m, const double* d_p, cudaStream_t st) {
  if (!m.use_graphs || m.n_states > Model::kGraphMaxStates) return false;
  if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) return false;  // cannot be captured
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &status) != cudaSuccess || status != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;  // the caller is capturing this stream itself
  }
  ++m.graph_clock;
  for (Model::WeightsGraph& g : m.weight_graphs) {
    if (g.d_p == d_p) {
      g.last_use = m.graph_clock;
      TAPES_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
      return true;
    }
  }
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = true;
  try {
    launch_all_weights_plain(m, d_p, st);
  } catch (const std::exception&) {
    ok = false;
  }
  if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) ok = false;
  if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError();
    if (exec) cudaGraphExecDestroy(exec);
    m.use_graphs = 0;  // do not try again on every call
    return false;
  }
  Model::WeightsGraph entry;
  entry.d_p = d_p; entry.exec = exec; entry.last_use = m.graph_clock;
  if (m.weight_graphs.size() < Model::kMaxWeightGraphs) {
    m.weight_graphs.push_back(entry);
  } else {
    size_t oldest = 0;
    for (size_t i = 1; i < m.weight_graphs.size(); ++i)
      if (m.weight_graphs[i].last_use < m.weight_graphs[oldest].last_use) oldest = i;
    // the replaced graph may still be running on some stream: launches are ordered on the streams
    // they went to, and destroying an executable graph waits for nothing, so wait here
    TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaGraphExecDestroy(m.weight_graphs[oldest].exec);
    m.weight_graphs[oldest] = entry;
  }
  TAPES_CUDA_CHECK(cudaGraphLaunch(exec, st));
  return true;
}

void launch_all_weights(Model& m, const double* d_p, cudaStream_t st) {
  if (!launch_weights_graph(m, d_p, st)) launch_all_weights_plain(m, d_p, st);
}
}  // namespace

int64_t rhs_launch_count(const Model& m) {
  if (m.fused_small && m.fused_cluster > 0 && m.more.empty() && m.flux_format == 1) {
    bool table_ok = m.ratio_right != nullptr;
    for (const Level& lv : m.levels) table_ok = table_ok || lv.n_groups == 0;
    if (table_ok || m.levels.empty()) return 1;  // fused_rhs_kernel
  }
  int64_t launches = 0;
  const int top = marginal_tail_top(m);
  launches += (m.k - 1 - top);          // one kernel per long marginal table
  if (top >= 0) launches += 1;          // tail tables
  if (m.n_rules) launches += 1;         // leaf-world probabilities
  {
    const size_t tile_bytes = ((size_t)kRowsPerBlock * ((size_t)m.A | 1u) + kRowsPerBlock) * sizeof(double);
    const bool one_pass = m.ratio_right && m.k >= 2 && !m.ratio_left && m.fuse_marginal_ratio &&
                          m.k - 1 > top && tile_bytes <= 48 * 1024;
    if (m.ratio_right && m.k >= 2 && !one_pass) launches += 1;  // extension ratios in a pass of their own
  }
  const bool use_ratio = m.ratio_table && m.ratio_right && m.k >= 2;
  for (const Level& lv : m.levels) {
    if (lv.n_roots) { launches += 1; continue; }
    const bool planes = use_ratio && m.plane_kernel && lv.n_plane_blocks > 0;
    if (planes) launches += 1;
    if (lv.n_left + (planes ? lv.n_general_blocks : lv.n_groups)) launches += 1;
  }
  if (m.out_sum) launches += 1;         // per-prefix sums of the group sums (right-chain outflow)
  launches += 1;                        // S * w
  for (const auto& part : m.more) launches += rhs_launch_count(*part);
  return launches;
}

void rhs_device(Model& m, const double* d_p, double* d_out, cudaStream_t stream) {
  cudaStream_t st = stream ? stream : m.stream;
  begin_use(m, st);
  if (!launch_fused(m, d_p, d_out, st, nullptr)) {
    launch_all_weights(m, d_p, st);
    launch_flux(m, d_out, 0, m.n_states, st);
  }
  end_use(m, st);
}

void rhs_device_fused(Model& m, const double* d_p, double* d_out, const StageUpdate& up, cudaStream_t stream) {
  cudaStream_t st = stream ? stream : m.stream;
  begin_use(m, st);
  if (!launch_fused(m, d_p, d_out, st, &up)) {
    launch_all_weights(m, d_p, st);
    launch_flux(m, d_out, 0, m.n_states, st, &up);
  }
  end_use(m, st);
}

void weights_device(Model& m, const double* d_p, cudaStream_t stream) {
  cudaStream_t st = stream ? stream : m.stream;
  begin_use(m, st);
  launch_all_weights(m, d_p, st);
  end_use(m, st);
}

void flux_rows_device(Model& m, double* d_out, uint64_t row_lo, uint64_t row_hi, cudaStream_t stream) {
  if (row_hi > m.n_states || row_lo > row_hi) throw std::runtime_error("row range outside the state table");
  cudaStream_t st = stream ? stream : m.stream;
  begin_use(m, st);
  launch_flux(m, d_out, row_lo, row_hi, st);
  end_use(m, st);
}

void rhs_device_profiled(Model& m, const double* d_p, double* d_out, cudaStream_t stream, float ms[3]) {
  cudaStream_t st = stream ? stream : m.stream;
  begin_use(m, st);
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) TAPES_CUDA_CHECK(cudaEventCreate(&ev[i]));
  ms[0] = ms[1] = ms[2] = 0.0f;
  const size_t parts = 1 + m.more.size();
  for (size_t i = 0; i < parts; ++i) {  // part by part, so that the phases can be told apart
    Model& part = i == 0 ? m : *m.more[i - 1];
    TAPES_CUDA_CHECK(cudaEventRecord(ev[0], st));
    launch_weights(part, d_p, st, ev);
    TAPES_CUDA_CHECK(cudaEventRecord(ev[2], st));
    launch_flux_impl<false>(part, d_out, 0, part.n_states, st, StageUpdate(), i > 0);
    TAPES_CUDA_CHECK(cudaEventRecord(ev[3], st));
    TAPES_CUDA_CHECK(cudaEventSynchronize(ev[3]));
    for (int j = 0; j < 3; ++j) {
      float t = 0.0f;
      TAPES_CUDA_CHECK(cudaEventElapsedTime(&t, ev[j], ev[j + 1]));
      ms[j] += t;
    }
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  end_use(m, st);
}

namespace {
void rhs_host_impl(Model& m, const double* h_p, double* h_out);
}
void rhs_host(Model& m, const double* h_p, double* h_out) {
  begin_use(m, m.stream);
  rhs_host_impl(m, h_p, h_out);
  m.in_use = false;  // both streams were synchronised: nothing of this model is in flight
}
namespace {
void rhs_host_impl(Model& m, const double* h_p, double* h_out) {
  const uint64_t n = m.n_states;
  const size_t bytes = (size_t)n * 8;
  if (!m.d_in) {
    m.d_in = dkeep<double>(m, n);
    m.d_out = dkeep<double>(m, n);
    TAPES_CUDA_CHECK(cudaStreamCreateWithFlags(&m.copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : m.copy_events) TAPES_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const bool small = n <= Model::kPinnedStagingStates;
  if (small) {
    // small tables: through pinned memory of the model, so that both transfers are plain DMA
    if (!m.h_pinned) TAPES_CUDA_CHECK(cudaHostAlloc((void**)&m.h_pinned, 2 * bytes, cudaHostAllocDefault));
    std::memcpy(m.h_pinned, h_p, bytes);
    TAPES_CUDA_CHECK(cudaMemcpyAsync(m.d_in, m.h_pinned, bytes, cudaMemcpyHostToDevice, m.stream));
    rhs_device(m, m.d_in, m.d_out, m.stream);
    TAPES_CUDA_CHECK(cudaMemcpyAsync(m.h_pinned + n, m.d_out, bytes, cudaMemcpyDeviceToHost, m.stream));
    TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
    std::memcpy(h_out, m.h_pinned + n, bytes);
    return;
  }
  // large tables.  Pinned caller buffers go by DMA on the model's streams; pageable ones (what the
  // reference's binding hands over, framework/markov_tapes.py:278-279) through the threaded staging
  // of hostcopy.h.  The result goes back in row blocks while the product of the next block runs.
  if (is_pinned_host(h_p)) {
    TAPES_CUDA_CHECK(cudaMemcpyAsync(m.d_in, h_p, bytes, cudaMemcpyHostToDevice, m.stream));
  } else {
    TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));  // d_in is idle: the copy runs on the stagers' streams
    staged_h2d(m.d_in, h_p, bytes);
  }
  const bool out_pinned = is_pinned_host(h_out);
  const int blocks = n >= (1ull << 22) ? Model::kCopyBlocks : 1;
  launch_all_weights(m, m.d_in, m.stream);
  const uint64_t per = ((n + blocks - 1) / blocks + 31) & ~31ull;
  int launched = 0;
  for (int b = 0; b < blocks; ++b) {
    const uint64_t lo = std::min<uint64_t>(n, per * b), hi = std::min<uint64_t>(n, lo + per);
    if (hi <= lo) break;
    launch_flux(m, m.d_out, lo, hi, m.stream);
    TAPES_CUDA_CHECK(cudaEventRecord(m.copy_events[b], m.stream));
    ++launched;
    if (out_pinned) {
      TAPES_CUDA_CHECK(cudaStreamWaitEvent(m.copy_stream, m.copy_events[b], 0));
      TAPES_CUDA_CHECK(cudaMemcpyAsync(h_out + lo, m.d_out + lo, (hi - lo) * 8, cudaMemcpyDeviceToHost, m.copy_stream));
    }
  }
  if (!out_pinned) {
    for (int b = 0; b < launched; ++b) {  // block b leaves while the products of the later blocks run
      const uint64_t lo = std::min<uint64_t>(n, per * b), hi = std::min<uint64_t>(n, lo + per);
      TAPES_CUDA_CHECK(cudaEventSynchronize(m.copy_events[b]));
      staged_d2h(h_out + lo, m.d_out + lo, (hi - lo) * 8);
    }
  }
  TAPES_CUDA_CHECK(cudaStreamSynchronize(m.copy_stream));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
}
}  // namespace

void materialize_node_weights(Model& m) {
  if (!m.ratio_right) return;
  const Consts c = make_consts(m);
  for (const Level& lv : m.levels)
    if (lv.n_groups)
      materialize_right_kernel<<<grid_for((uint64_t)lv.n_groups * c.A, kThreads), kThreads, 0, m.stream>>>(lv, c, m.ratio_right, m.node_w);
  TAPES_CUDA_CHECK(cudaGetLastError());
  TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
}

void export_full_csr(Model& m, int64_t* h_row_ptr, uint32_t* h_entries) {
  if (!m.more.empty()) throw std::runtime_error("composite model: export its parts one by one");
  const uint64_t n = m.n_states, A = (uint64_t)m.A;
  TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
  std::vector<uint64_t> rp(n + 1);
  TAPES_CUDA_CHECK(cudaMemcpy(rp.data(), m.row_ptr, (n + 1) * 8, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> stored(m.nnz_stored);
  if (m.nnz_stored) {
    uint32_t* d_entries = m.entries;
    uint32_t* rebuilt = nullptr;
    if (!d_entries) {  // only the sliced form is resident: expand it back
      if (cudaMalloc((void**)&rebuilt, m.nnz_stored * 4) != cudaSuccess) throw std::runtime_error("out of device memory");
      try {
        expand_flux_slices(m, rebuilt, m.stream);
        TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
      } catch (...) {
        cudaFree(rebuilt);
        throw;
      }
      d_entries = rebuilt;
    }
    const cudaError_t err = cudaMemcpy(stored.data(), d_entries, m.nnz_stored * 4, cudaMemcpyDeviceToHost);
    if (rebuilt) cudaFree(rebuilt);
    TAPES_CUDA_CHECK(err);
  }
  // the entries of the right children, whose flux the device evaluates per prefix group
  std::vector<std::vector<uint32_t>> prefixes(m.levels.size()), adjusted(m.levels.size());
  std::vector<int64_t> count(n + 1, 0);
  for (uint64_t r = 0; r < n; ++r) count[r] = (int64_t)(rp[r + 1] - rp[r]);
  for (size_t l = 0; l < m.levels.size(); ++l) {
    const Level& lv = m.levels[l];
    if (!lv.n_groups) continue;
    prefixes[l].resize(lv.n_groups);
    TAPES_CUDA_CHECK(cudaMemcpy(prefixes[l].data(), lv.g_prefix, (size_t)lv.n_groups * 4, cudaMemcpyDeviceToHost));
    adjusted[l].resize(lv.n_groups);
    TAPES_CUDA_CHECK(cudaMemcpy(adjusted[l].data(), lv.g_adjusted, (size_t)lv.n_groups * 4, cudaMemcpyDeviceToHost));
    for (uint32_t g = 0; g < lv.n_groups; ++g)
      for (uint64_t x = 0; x < A; ++x) {
        count[(uint64_t)prefixes[l][g] * A + x] += 1;
        count[(uint64_t)adjusted[l][g] * A + x] += 1;
      }
  }
  h_row_ptr[0] = 0;
  for (uint64_t r = 0; r < n; ++r) h_row_ptr[r + 1] = h_row_ptr[r] + count[r];
  if ((uint64_t)h_row_ptr[n] != m.nnz) throw std::runtime_error("export: entry count does not match the flux terms");
  std::vector<int64_t> cursor(h_row_ptr, h_row_ptr + n);
  for (uint64_t r = 0; r < n; ++r)
    for (uint64_t e = rp[r]; e < rp[r + 1]; ++e) h_entries[cursor[r]++] = stored[e];
  for (size_t l = 0; l < m.levels.size(); ++l) {
    const Level& lv = m.levels[l];
    const uint64_t right_base = lv.base + A * lv.n_left;
    for (uint32_t g = 0; g < lv.n_groups; ++g)
      for (uint64_t x = 0; x < A; ++x) {
        const uint32_t node = (uint32_t)(right_base + (uint64_t)g * A + x);
        h_entries[cursor[(uint64_t)prefixes[l][g] * A + x]++] = node | kOutflowBit;
        h_entries[cursor[(uint64_t)adjusted[l][g] * A + x]++] = node;
      }
  }
  for (uint64_t r = 0; r < n; ++r) std::sort(h_entries + h_row_ptr[r], h_entries + h_row_ptr[r + 1]);
}

}  // namespace tapes
