// Sliced flux structure (sm_100a): the product  dy/dt = S * w  of the master-equation step.
//
// Reference behaviour being reproduced: accumulate-dp/dt (framework/tape_multiverse.scm:1271-1301)
// adds -w at the original and +w at the adjusted window index of every flux term.  Here every
// state gathers its terms instead (no atomics, fixed summation order).
//
// Why slices: with one lane per state and plain CSR, a warp's entry loads are 32 separate sectors
// (ncu: 17 sectors per request, L1 throughput 88 %, profiles/r01_c_*).  Consecutive states gather
// consecutive forest nodes, so 32 states are cut into runs "lane l holds first + rank(l)" that cost
// 8 bytes per run instead of 4 bytes per entry, and the gathers of a run are one contiguous
// segment of the weight vector.  What does not fall into runs is stored column-major (coalesced).
#include <cuda/atomic>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <utility>
#include <vector>

#include "engine.h"
#include "flux_device.cuh"
#include "primitives.cuh"

namespace tapes {

namespace {

RightFlux right_flux_of(const Model& m) {
  RightFlux f;
  f.out_sum = m.out_sum; f.ratio = m.ratio_right; f.totals = m.g_total_all;
  f.in_ptr = m.in_ptr; f.in_pairs = m.in_pairs; f.A = (uint32_t)m.A;
  return f;
}

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

template <typename T>
T* dalloc(size_t n, cudaStream_t st) {
  return (T*)pool_alloc(std::max<size_t>(n, 1) * sizeof(T), st);
}

// One warp merges the (ascending) rows of its 32 states into one ascending stream and cuts it into
// runs: a run continues while the next entry is the previous one + 1 and sits in a higher lane.
// Runs with at least min_lanes lanes are kept as (first, mask); the entries of shorter ones go to
// the lane's column list.  FILL = false counts, FILL = true writes.
template <bool FILL>
__global__ void __launch_bounds__(kThreads) encode_slices_kernel(
    const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ entries, uint64_t n_rows,
    uint64_t n_slices, int min_lanes, uint32_t* __restrict__ slice_runs, uint32_t* __restrict__ slice_cols,
    const uint64_t* __restrict__ slice_ptr, uint32_t* __restrict__ words) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t s = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= n_slices) return;
  const uint64_t row = s * 32 + lane;
  uint64_t cur = 0, end = 0;
  if (row < n_rows) { cur = row_ptr[row]; end = row_ptr[row + 1]; }
  uint32_t head = cur < end ? entries[cur] : kNone;
  uint32_t next = cur + 1 < end ? entries[cur + 1] : kNone;  // one entry ahead hides the load latency

  uint64_t run_at = 0, col_at = 0;
  if (FILL) {
    run_at = slice_ptr[s];
    col_at = run_at + 2ull * ((slice_runs[s] + 1u) & ~1u);
  }
  const uint32_t below = (1u << lane) - 1u;
  uint32_t n_runs = 0, my_cols = 0;
  uint32_t first = 0, mask = 0, prev = 0;
  int prev_lane = -1;

  auto close = [&]() {
    if (mask == 0) return;
    if (__popc(mask) >= min_lanes) {
      if (FILL && lane == 0) { words[run_at + 2ull * n_runs] = first; words[run_at + 2ull * n_runs + 1] = mask; }
      ++n_runs;
    } else if ((mask >> lane) & 1u) {
      if (FILL) words[col_at + 32ull * my_cols + lane] = first + __popc(mask & below);
      ++my_cols;
    }
  };

  for (;;) {
    const uint32_t m = __reduce_min_sync(0xffffffffu, head);
    if (m == kNone) break;
    const int win = __ffs(__ballot_sync(0xffffffffu, head == m)) - 1;  // entries are unique
    const bool extend = mask != 0 && m == prev + 1 && ((m ^ prev) & kSignBit) == 0 && win > prev_lane;
    if (!extend) { close(); first = m; mask = 0; }
    mask |= 1u << win;
    prev = m;
    prev_lane = win;
    if ((int)lane == win) {
      ++cur;
      head = next;
      next = cur + 1 < end ? entries[cur + 1] : kNone;
    }
  }
  close();

  if (!FILL) {
    const uint32_t cols = __reduce_max_sync(0xffffffffu, my_cols);
    if (lane == 0) { slice_runs[s] = n_runs; slice_cols[s] = cols; }
  } else {
    if (lane == 0 && (n_runs & 1u)) {  // pad to an even number of run pairs (16-byte alignment)
      words[run_at + 2ull * n_runs] = 0; words[run_at + 2ull * n_runs + 1] = 0;
    }
  }
}

// The same encoding for min_lanes = 32, where a run is exactly "value v in lane 0 and v + l in every
// lane l": only lane 0's entries have to be visited (tens of steps instead of one per entry of the
// slice).  Every lane walks its own ascending row once; the entries it matched are remembered in a
// bit mask (rows longer than 128 entries keep their tail in columns).
template <bool FILL>
__global__ void __launch_bounds__(kThreads) encode_full_runs_kernel(
    const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ entries, uint64_t n_rows,
    uint64_t n_slices, uint32_t* __restrict__ slice_runs, uint32_t* __restrict__ slice_cols,
    const uint64_t* __restrict__ slice_ptr, uint32_t* __restrict__ words) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t s = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= n_slices) return;
  const uint64_t row = s * 32 + lane;
  uint64_t lo = 0, hi = 0;
  if (row < n_rows) { lo = row_ptr[row]; hi = row_ptr[row + 1]; }
  const uint32_t len = (uint32_t)(hi - lo);
  const uint32_t len0 = __shfl_sync(0xffffffffu, len, 0);

  uint64_t run_at = 0, col_at = 0;
  if (FILL) {
    run_at = slice_ptr[s];
    col_at = run_at + 2ull * ((slice_runs[s] + 1u) & ~1u);
  }
  uint64_t used_lo = 0, used_hi = 0;  // entries of this lane that went into runs
  uint32_t n_runs = 0, at = 0;        // at: first entry of the row not yet passed
  uint32_t head = at < len ? entries[lo + at] : kNone;
  for (uint32_t j = 0; j < len0; ++j) {
    uint32_t v0 = lane == 0 ? entries[lo + j] : 0u;
    v0 = __shfl_sync(0xffffffffu, v0, 0);
    const uint32_t want = v0 + lane;
    while (head < want) {  // kNone is larger than any entry
      ++at;
      head = at < len ? entries[lo + at] : kNone;
    }
    const bool match = head == want && ((want ^ v0) & kSignBit) == 0 && at < 128;
    if (__all_sync(0xffffffffu, match)) {
      if (at < 64) used_lo |= 1ull << at; else used_hi |= 1ull << (at - 64);
      if (FILL && lane == 0) { words[run_at + 2ull * n_runs] = v0; words[run_at + 2ull * n_runs + 1] = 0xffffffffu; }
      ++n_runs;
      ++at;
      head = at < len ? entries[lo + at] : kNone;
    }
  }
  const uint32_t my_cols = len - (uint32_t)__popcll(used_lo) - (uint32_t)__popcll(used_hi);
  if (!FILL) {
    const uint32_t cols = __reduce_max_sync(0xffffffffu, my_cols);
    if (lane == 0) { slice_runs[s] = n_runs; slice_cols[s] = cols; }
  } else {
    uint32_t c = 0;
    for (uint32_t e = 0; e < len; ++e) {
      const bool used = e < 64 ? (used_lo >> e) & 1ull : (e < 128 ? (used_hi >> (e - 64)) & 1ull : 0ull);
      if (!used) { words[col_at + 32ull * c + lane] = entries[lo + e]; ++c; }
    }
    if (lane == 0 && (n_runs & 1u)) {
      words[run_at + 2ull * n_runs] = 0; words[run_at + 2ull * n_runs + 1] = 0;
    }
  }
}

__global__ void slice_sizes_kernel(const uint32_t* __restrict__ slice_runs, const uint32_t* __restrict__ slice_cols,
                                   uint64_t n_slices, uint32_t* __restrict__ sizes) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_slices) sizes[s] = 2u * ((slice_runs[s] + 1u) & ~1u) + 32u * slice_cols[s];
}

// Totals for the encoding facts: [0] runs, [1] entries held by runs, [2] run pairs incl. padding.
__global__ void run_facts_kernel(const uint64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ slice_runs,
                                 const uint32_t* __restrict__ words, uint64_t n_slices,
                                 unsigned long long* __restrict__ facts) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long runs = 0, held = 0, pairs = 0;
  if (s < n_slices) {
    const uint64_t at = slice_ptr[s];
    runs = slice_runs[s];
    pairs = (runs + 1ull) & ~1ull;
    for (uint32_t j = 0; j < runs; ++j) held += __popc(words[at + 2ull * j + 1]);
  }
  for (int d = 16; d > 0; d >>= 1) {
    runs += __shfl_down_sync(0xffffffffu, runs, d);
    held += __shfl_down_sync(0xffffffffu, held, d);
    pairs += __shfl_down_sync(0xffffffffu, pairs, d);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&facts[0], runs); atomicAdd(&facts[1], held); atomicAdd(&facts[2], pairs); }
}

template <int U, bool FUSED, int MIN_BLOCKS>
__global__ void __launch_bounds__(kThreads, MIN_BLOCKS) flux_slices_kernel(
    const uint64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ slice_runs,
    const uint32_t* __restrict__ words, const double* __restrict__ w, double* __restrict__ out,
    uint64_t slice_lo, uint64_t slice_hi, uint64_t row_lo, uint64_t row_hi, StageUpdate up, int accumulate,
    RightFlux right) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t s = slice_lo + (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= slice_hi) return;
  double acc = slice_sum<U>(slice_ptr, slice_runs, words, w, s, lane);
  const uint64_t row = s * 32 + lane;
  if (row >= row_lo && row < row_hi) {
    acc = acc + right_flux<U>(right, row);  // the terms of right children, per prefix group (engine.h Model::out_ptr)
    if (accumulate) acc = out[row] + acc;  // a later part of a composite model (engine.h Model::more)
    out[row] = acc;
    if (FUSED) {  // Runge-Kutta stage update for this state (terms in tableau order)
      double a = 0.0;
      for (int j = 0; j < up.n; ++j) a += up.vec[j][row] * up.coef[j];
      a += acc * up.coef_self;
      up.stage[row] = up.y[row] + a * up.h;
    }
  }
}

// The product fused with the flux exchange of a multi-GPU run: states are owned in contiguous
// blocks of `block` states (a multiple of 32, so a slice has one owner), and this rank's partial
// dy/dt of a state goes straight into the owner's staging memory, slot `rank`, with peer stores
// over NVLink while the rest of the kernel computes.  staging.ptr[o] is owner o's buffer of
// world * block doubles.  One launch handles sub-block `round` (sub_slices slices) of every owner.
// Every rank starts with a different owner (rank r with owner r + 1, then r + 2, ...): started in
// the same order, all ranks would store into one owner's memory at a time and queue up at its
// NVLink ingress (measured at 8 GPUs: +1.5 ms on a 3.6 ms kernel).
template <int U, int MIN_BLOCKS>
__global__ void __launch_bounds__(kThreads, MIN_BLOCKS) flux_slices_scatter_kernel(
    const uint64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ slice_runs,
    const uint32_t* __restrict__ words, const double* __restrict__ w,
    const __grid_constant__ PeerPointers staging, uint32_t world, uint32_t rank, uint64_t block,
    uint64_t sub_slices, uint64_t round, uint64_t n_slices, uint64_t n_rows, RightFlux right) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t idx = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (idx >= sub_slices * world) return;
  const uint64_t hop = idx / sub_slices, within = idx - hop * sub_slices;
  uint64_t owner = rank + 1 + hop;
  if (owner >= world) owner -= world;
  const uint64_t s = owner * (block / 32) + round * sub_slices + within;
  if (s >= n_slices) return;  // ragged end of the table
  const double acc = slice_sum<U>(slice_ptr, slice_runs, words, w, s, lane);
  const uint64_t row = s * 32 + lane;
  if (row < n_rows) staging.ptr[owner][(uint64_t)rank * block + (row - owner * block)] = acc + right_flux<U>(right, row);
}

// The owner's half of the exchange: adds the world slots of rows [j_lo, j_hi) of its block in rank
// order (so every rank of a run, and every run, gets the same bits) and stores the sums into the
// full vector of every rank (peer stores).  result.ptr[q] is rank q's full dy/dt vector.
__global__ void __launch_bounds__(kThreads) sum_slots_broadcast_kernel(const double* __restrict__ slots,
                                                                       const __grid_constant__ PeerPointers result,
                                                                       uint32_t world, uint32_t rank, uint64_t block,
                                                                       uint64_t j_lo, uint64_t j_hi, uint64_t n_rows) {
  // grid-stride: the launch may use fewer blocks than rows / 256 (peer_rhs: a round's sums are pushed at a
  // moderate rate while the product of the next round runs, so that the product's own peer stores do not
  // queue behind a burst on the same NVLink egress)
  for (uint64_t j = j_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j_hi; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = (uint64_t)rank * block + j;
    if (row >= n_rows) return;
    double total = 0.0;
    for (uint32_t r = 0; r < world; ++r) total += slots[(uint64_t)r * block + j];
    for (uint32_t q = 0; q < world; ++q) result.ptr[q][row] = total;
  }
}

// Cross-GPU signalling through peer-visible flag arrays.  flags.ptr[q] holds 2 * world epochs in
// rank q's memory: [r] "rank r's partial flux of the current round has landed here", [world + r]
// "rank r's sums of the current right-hand side have landed here".  A kernel boundary orders the
// data stores of the preceding kernel before the signal.
__global__ void peer_signal_kernel(const __grid_constant__ PeerFlags flags, uint32_t world, uint32_t rank,
                                   uint32_t set, unsigned long long epoch) {
  const uint32_t q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();
  cuda::atomic_ref<unsigned long long, cuda::thread_scope_system> slot(flags.ptr[q][set * world + rank]);
  slot.store(epoch, cuda::std::memory_order_release);
}

__global__ void peer_wait_kernel(unsigned long long* __restrict__ mine, uint32_t world, uint32_t set,
                                 unsigned long long epoch, unsigned long long timeout_ns, int* __restrict__ error) {
  const uint32_t r = threadIdx.x;
  if (r >= world) return;
  cuda::atomic_ref<unsigned long long, cuda::thread_scope_system> slot(mine[set * world + r]);
  unsigned long long start;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
  while (slot.load(cuda::std::memory_order_acquire) < epoch) {
    if (*(volatile int*)error) break;  // an earlier wait already gave up: do not stack timeouts
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - start > timeout_ns) { atomicExch(error, 1); break; }
    __nanosleep(200);
  }
  __threadfence_system();
}

// Writes every lane's entries back to CSR positions (runs, then columns) and sorts the row.
__global__ void __launch_bounds__(kThreads) expand_slices_kernel(
    const uint64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ slice_runs,
    const uint32_t* __restrict__ words, const uint64_t* __restrict__ row_ptr, uint64_t n_rows,
    uint64_t n_slices, uint32_t* __restrict__ entries) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t s = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= n_slices) return;
  const uint64_t row = s * 32 + lane;
  if (row >= n_rows) return;
  const uint64_t at = slice_ptr[s], stop = slice_ptr[s + 1];
  const uint32_t n_runs = slice_runs[s];
  const uint32_t n_pairs = (n_runs + 1u) & ~1u;
  const uint32_t n_cols = (uint32_t)((stop - at - 2ull * n_pairs) >> 5);
  const uint32_t below = (1u << lane) - 1u;
  const uint64_t lo = row_ptr[row];
  uint64_t pos = lo;
  for (uint32_t j = 0; j < n_runs; ++j) {
    const uint32_t first = words[at + 2ull * j], mask = words[at + 2ull * j + 1];
    if ((mask >> lane) & 1u) entries[pos++] = first + __popc(mask & below);
  }
  for (uint32_t c = 0; c < n_cols; ++c) {
    const uint32_t v = words[at + 2ull * n_pairs + 32ull * c + lane];
    if (v != kNone) entries[pos++] = v;
  }
  for (uint64_t a = lo + 1; a < pos; ++a) {  // insertion sort: canonical ascending order
    const uint32_t v = entries[a];
    uint64_t b = a;
    while (b > lo && entries[b - 1] > v) { entries[b] = entries[b - 1]; --b; }
    entries[b] = v;
  }
}

}  // namespace

void build_flux_slices(Model& m, int min_run_lanes, cudaStream_t st) {
  auto t0 = std::chrono::steady_clock::now();
  FluxSlices& fs = m.slices;
  const uint64_t n = m.n_states;
  fs.n_slices = (n + 31) / 32;
  fs.min_run_lanes = std::max(1, std::min(32, min_run_lanes));
  const uint64_t S = fs.n_slices;
  const unsigned grid = grid_for(S * 32, kThreads);
  fs.slice_runs = m.arena.array<uint32_t>(S);
  uint32_t* cols = dalloc<uint32_t>(S, st);
  uint32_t* sizes = dalloc<uint32_t>(S, st);
  const bool full_only = fs.min_run_lanes == 32;
  if (full_only)
    encode_full_runs_kernel<false><<<grid, kThreads, 0, st>>>(m.row_ptr, m.entries, n, S, fs.slice_runs, cols,
                                                             nullptr, nullptr);
  else
    encode_slices_kernel<false><<<grid, kThreads, 0, st>>>(m.row_ptr, m.entries, n, S, fs.min_run_lanes,
                                                          fs.slice_runs, cols, nullptr, nullptr);
  slice_sizes_kernel<<<grid_for(S, kThreads), kThreads, 0, st>>>(fs.slice_runs, cols, S, sizes);
  fs.slice_ptr = m.arena.array<uint64_t>(S + 1);
  uint64_t* scan_tmp = dalloc<uint64_t>(scan_tmp_elems(S), st);
  exclusive_scan_u32(sizes, S, fs.slice_ptr, scan_tmp, st);
  TAPES_CUDA_CHECK(cudaMemcpyAsync(&fs.n_words, fs.slice_ptr + S, 8, cudaMemcpyDeviceToHost, st));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  fs.words = m.arena.array<uint32_t>(fs.n_words);
  TAPES_CUDA_CHECK(cudaMemsetAsync(fs.words, 0xff, fs.n_words * 4, st));  // columns default to "none"
  if (full_only)
    encode_full_runs_kernel<true><<<grid, kThreads, 0, st>>>(m.row_ptr, m.entries, n, S, fs.slice_runs, cols,
                                                            fs.slice_ptr, fs.words);
  else
    encode_slices_kernel<true><<<grid, kThreads, 0, st>>>(m.row_ptr, m.entries, n, S, fs.min_run_lanes,
                                                         fs.slice_runs, cols, fs.slice_ptr, fs.words);
  unsigned long long* facts = dalloc<unsigned long long>(3, st);
  TAPES_CUDA_CHECK(cudaMemsetAsync(facts, 0, 24, st));
  run_facts_kernel<<<grid_for(S, kThreads), kThreads, 0, st>>>(fs.slice_ptr, fs.slice_runs, fs.words, S, facts);
  unsigned long long h_facts[3] = {0, 0, 0};
  TAPES_CUDA_CHECK(cudaMemcpyAsync(h_facts, facts, 24, cudaMemcpyDeviceToHost, st));
  TAPES_CUDA_CHECK(cudaGetLastError());
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  fs.runs = h_facts[0];
  fs.run_entries = h_facts[1];
  fs.column_entries = m.nnz_stored - fs.run_entries;
  fs.column_slots = fs.n_words - 2 * h_facts[2];
  cudaFreeAsync(cols, st); cudaFreeAsync(sizes, st); cudaFreeAsync(scan_tmp, st); cudaFreeAsync(facts, st);
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  m.stats.device_slices_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void launch_flux_slices(Model& m, double* d_out, uint64_t row_lo, uint64_t row_hi, cudaStream_t st,
                        const StageUpdate* up, bool accumulate) {
  if (row_hi <= row_lo) return;
  const FluxSlices& fs = m.slices;
  const uint64_t slice_lo = row_lo / 32, slice_hi = (row_hi + 31) / 32;
  const unsigned grid = grid_for((slice_hi - slice_lo) * 32, kThreads);
  const int acc_flag = accumulate ? 1 : 0;
  const RightFlux of = right_flux_of(m);
#define TAPES_FLUX(U_, B_)                                                                                     \
  (up ? flux_slices_kernel<U_, true, B_><<<grid, kThreads, 0, st>>>(fs.slice_ptr, fs.slice_runs, fs.words, m.node_w, \
                                                                   d_out, slice_lo, slice_hi, row_lo, row_hi, *up, acc_flag, of) \
      : flux_slices_kernel<U_, false, B_><<<grid, kThreads, 0, st>>>(fs.slice_ptr, fs.slice_runs, fs.words, m.node_w, \
                                                                    d_out, slice_lo, slice_hi, row_lo, row_hi, StageUpdate(), acc_flag, of))
  // measured on B200 (n = 1e8, 24 rules): occupancy beats depth.  Round 1 (all terms stored): 4 gathers per
  // lane at 40 registers 3.55 ms, 6 at 48 registers 3.93 ms, 8 at 56 registers 4.22 ms.  Round 2 (right
  // children per prefix group, the kernel walks lists and is bound by latency): 3 at 32 registers and 8
  // blocks per SM 2.73 ms, 4 at 40 registers 2.96 ms, 2: 3.07 ms, 6: 3.24 ms (profiles/r02_m_*).  Issuing the
  // first loads of the per-group lists before the slice is walked (two chains of dependent loads side by
  // side) needs registers the kernel does not have: 96 B of spills, 3.32 ms; dropped again.
  if (m.flux_unroll >= 8) TAPES_FLUX(8, 4);
  else if (m.flux_unroll >= 6) TAPES_FLUX(6, 5);
  else if (m.flux_unroll >= 4) TAPES_FLUX(4, 6);
  else if (m.flux_unroll == 3) TAPES_FLUX(3, 8);
  else TAPES_FLUX(2, 8);
#undef TAPES_FLUX
  TAPES_CUDA_CHECK(cudaGetLastError());
}

PeerGroup::~PeerGroup() {
  if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); }
  if (fork) cudaEventDestroy(fork);
  if (join) cudaEventDestroy(join);
  if (d_error) cudaFree(d_error);
}

PeerGroup* peer_group_create(int world, int rank, uint64_t block, int rounds, void* const* staging,
                             void* const* result, void* const* flags) {
  if (world < 1 || world > PeerPointers::kMax || rank < 0 || rank >= world) throw std::runtime_error("bad world / rank");
  if (rounds < 1 || rounds > PeerGroup::kMaxRounds) throw std::runtime_error("rounds out of range");
  if (block == 0 || block % (32ull * rounds) != 0)
    throw std::runtime_error("ownership blocks must be multiples of 32 * rounds states");
  std::unique_ptr<PeerGroup> g(new PeerGroup());
  g->world = world; g->rank = rank; g->block = block; g->rounds = rounds;
  for (int i = 0; i < world; ++i) {
    if (!staging[i] || !result[i] || !flags[i]) throw std::runtime_error("null peer buffer");
    g->staging.ptr[i] = (double*)staging[i];
    g->result.ptr[i] = (double*)result[i];
    g->flags.ptr[i] = (unsigned long long*)flags[i];
  }
  // highest priority: the owner's small kernels must get the block slots the product grid frees,
  // not queue behind the rest of that grid
  int least = 0, greatest = 0;
  TAPES_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  TAPES_CUDA_CHECK(cudaStreamCreateWithPriority(&g->side, cudaStreamNonBlocking, greatest));
  TAPES_CUDA_CHECK(cudaEventCreateWithFlags(&g->fork, cudaEventDisableTiming));
  TAPES_CUDA_CHECK(cudaEventCreateWithFlags(&g->join, cudaEventDisableTiming));
  TAPES_CUDA_CHECK(cudaMalloc((void**)&g->d_error, sizeof(int)));
  TAPES_CUDA_CHECK(cudaMemset(g->d_error, 0, sizeof(int)));
  return g.release();
}

void peer_rhs(PeerGroup& g, Model& m, const double* d_p, cudaStream_t st) {
  if (m.flux_format != 1) throw std::runtime_error("the fused exchange needs the sliced flux structure");
  if (!m.more.empty())
    throw std::runtime_error("this rank's share is a composite model (forest above 2^31 nodes): deal the problem to more ranks");
  if (g.block * (uint64_t)g.world < m.n_states) throw std::runtime_error("ownership blocks do not cover the table");
  const FluxSlices& fs = m.slices;
  const uint32_t world = (uint32_t)g.world, rank = (uint32_t)g.rank;
  const uint64_t sub = g.block / (uint64_t)g.rounds, sub_slices = sub / 32;
  const unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
  unsigned long long* mine = g.flags.ptr[g.rank];
  // TAPES_PEER_TRACE=1: time stamps (CUDA events on both streams) of every kernel of the exchange, printed
  // by rank 0 two calls later - a diagnostic
  static const bool trace = std::getenv("TAPES_PEER_TRACE") != nullptr;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  auto mark = [&](const char* what, cudaStream_t on) {
    if (!trace) return;
    cudaEvent_t e;
    TAPES_CUDA_CHECK(cudaEventCreate(&e));
    TAPES_CUDA_CHECK(cudaEventRecord(e, on));
    marks.push_back({what, e});
  };
  mark("start", st);
  weights_device(m, d_p, st);
  // the side stream must not run ahead of work already queued on the main stream
  TAPES_CUDA_CHECK(cudaEventRecord(g.fork, st));
  TAPES_CUDA_CHECK(cudaStreamWaitEvent(g.side, g.fork, 0));
  const unsigned scatter_grid = grid_for(sub_slices * world * 32, kThreads);
  const RightFlux of = right_flux_of(m);
  const unsigned long long base = g.epoch;
  // blocks of the owner's sum + broadcast kernel in the rounds that run beside a product (0 = one thread per
  // row): measured at 2 GPUs, 4 rounds: uncapped 6.33 ms per step, 296 blocks 6.09, 148 blocks 6.08, 64 blocks
  // 6.76 (too slow to keep up), profiles/r02_aq_sum_blocks_n2.log
  static const int sum_blocks = std::getenv("TAPES_PEER_SUM_BLOCKS") ? std::atoi(std::getenv("TAPES_PEER_SUM_BLOCKS")) : 296;
  mark("weights done", st);
  for (int c = 0; c < g.rounds; ++c) {
    if (m.flux_unroll >= 4)
      flux_slices_scatter_kernel<4, 6><<<scatter_grid, kThreads, 0, st>>>(
          fs.slice_ptr, fs.slice_runs, fs.words, m.node_w, g.staging, world, rank, g.block, sub_slices, (uint64_t)c,
          fs.n_slices, m.n_states, of);
    else  // like the product of one rank: 3 gathers in flight at 8 blocks per SM
      flux_slices_scatter_kernel<3, 8><<<scatter_grid, kThreads, 0, st>>>(
          fs.slice_ptr, fs.slice_runs, fs.words, m.node_w, g.staging, world, rank, g.block, sub_slices, (uint64_t)c,
          fs.n_slices, m.n_states, of);
    mark("scatter done", st);
    peer_signal_kernel<<<1, 32, 0, st>>>(g.flags, world, rank, 0u, base + c + 1);
    peer_wait_kernel<<<1, 32, 0, g.side>>>(mine, world, 0u, base + c + 1, timeout_ns, g.d_error);
    mark("all partials here", g.side);
    // every round but the last runs beside the next round's product
    const unsigned sum_grid = (c + 1 < g.rounds && sum_blocks > 0) ? std::min<unsigned>(grid_for(sub, kThreads), (unsigned)sum_blocks)
                                                                    : grid_for(sub, kThreads);
    sum_slots_broadcast_kernel<<<sum_grid, kThreads, 0, g.side>>>(
        g.staging.ptr[g.rank], g.result, world, rank, g.block, sub * c, sub * (c + 1), m.n_states);
    mark("sums sent", g.side);
  }
  g.epoch = base + g.rounds;
  // every owner's sums have landed everywhere (and its slots may be overwritten) before st goes on
  peer_signal_kernel<<<1, 32, 0, g.side>>>(g.flags, world, rank, 1u, g.epoch);
  peer_wait_kernel<<<1, 32, 0, g.side>>>(mine, world, 1u, g.epoch, timeout_ns, g.d_error);
  mark("all sums here", g.side);
  TAPES_CUDA_CHECK(cudaEventRecord(g.join, g.side));
  TAPES_CUDA_CHECK(cudaStreamWaitEvent(st, g.join, 0));
  TAPES_CUDA_CHECK(cudaGetLastError());
  if (trace) {
    // printed two calls later, when the events have completed anyway: the calls stay free-running
    static std::vector<std::vector<std::pair<const char*, cudaEvent_t>>> pending;
    static cudaEvent_t previous_end = nullptr;
    pending.push_back(marks);
    if (pending.size() > 2) {
      auto done = pending.front();
      pending.erase(pending.begin());
      TAPES_CUDA_CHECK(cudaEventSynchronize(done.back().second));
      if (rank == 0) std::fprintf(stderr, "[peer trace] rounds=%d:", g.rounds);
      float ms = 0;
      if (previous_end && cudaEventElapsedTime(&ms, previous_end, done.front().second) == cudaSuccess && rank == 0)
        std::fprintf(stderr, " since the end of the call before %.3f;", ms);
      for (auto& mk : done) {
        const cudaError_t err = cudaEventElapsedTime(&ms, done.front().second, mk.second);
        if (rank == 0) std::fprintf(stderr, " %s %.3f%s;", mk.first, ms, err == cudaSuccess ? "" : cudaGetErrorName(err));
      }
      if (rank == 0) std::fprintf(stderr, "\n");
      if (previous_end) cudaEventDestroy(previous_end);
      previous_end = done.back().second;
      for (size_t i = 0; i + 1 < done.size(); ++i) cudaEventDestroy(done[i].second);
    }
  }
  end_use(m, st);  // the product kernels read the model's weights after weights_device's own record
}

int peer_group_error(PeerGroup& g) {
  int h = 0;
  TAPES_CUDA_CHECK(cudaMemcpy(&h, g.d_error, sizeof(int), cudaMemcpyDeviceToHost));
  return h;
}

void expand_flux_slices(Model& m, uint32_t* d_entries, cudaStream_t st) {
  if (!m.more.empty()) throw std::runtime_error("composite model: export its parts one by one");
  const FluxSlices& fs = m.slices;
  expand_slices_kernel<<<grid_for(fs.n_slices * 32, kThreads), kThreads, 0, st>>>(
      fs.slice_ptr, fs.slice_runs, fs.words, m.row_ptr, m.n_states, fs.n_slices, d_entries);
  TAPES_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tapes
