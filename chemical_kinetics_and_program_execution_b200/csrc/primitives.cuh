// Device-wide building blocks used by the structure build (sm_100a): exclusive scan, stable LSD
// radix sort of 64-bit keys, and an open-addressing hash set with warp-aggregated inserts.
// All launches go to the caller's stream; nothing here synchronises the host.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "cuda_check.h"

namespace tapes {
namespace {  // kernels here have internal linkage: every translation unit gets its own copy

// ------------------------------------------------------------------------------------------
// Exclusive scan: u32 counts -> u64 offsets.  Three kernels: tile sums, scan of the tile sums by
// one block, tile-local scan plus tile offset.  The grand total lands in out[n].
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t warp_inclusive_sum(uint64_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// Block-wide exclusive prefix of one value per thread; returns the prefix and the block total.
__device__ __forceinline__ uint64_t block_exclusive_sum(uint64_t v, uint64_t* warp_totals /*32*/,
                                                        uint64_t* block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  uint64_t inc = warp_inclusive_sum(v, lane);
  if (lane == 31) warp_totals[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t t = lane < nwarps ? warp_totals[lane] : 0;
    uint64_t ti = warp_inclusive_sum(t, lane);
    warp_totals[lane] = ti - t;  // exclusive
    if (lane == 31) *block_total = ti;
  }
  __syncthreads();
  uint64_t res = inc - v + warp_totals[warp];
  __syncthreads();
  return res;
}

// PAIR: the input words carry two flags (bit 0, bit 1) that are counted side by side, bit 0 in the
// low and bit 1 in the high half of the 64-bit sums (each count below 2^32): one scan for two ranks.
template <bool PAIR>
__device__ __forceinline__ uint64_t scan_value(uint32_t v) {
  return PAIR ? ((uint64_t)(v & 1u) | ((uint64_t)((v >> 1) & 1u) << 32)) : (uint64_t)v;
}

template <bool PAIR>
__global__ void scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                      uint64_t* __restrict__ tile_sums) {
  __shared__ uint64_t wt[32];
  __shared__ uint64_t total;
  const uint64_t tile0 = (uint64_t)blockIdx.x * kScanTile;
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = tile0 + (uint64_t)j * kScanThreads + threadIdx.x;
    if (i < n) s += scan_value<PAIR>(in[i]);
  }
  (void)block_exclusive_sum(s, wt, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_tile_offsets_kernel(uint64_t* __restrict__ tile_sums, uint64_t ntiles,
                                         uint64_t* __restrict__ grand_total) {
  __shared__ uint64_t wt[32];
  __shared__ uint64_t total;
  uint64_t carry = 0;
  for (uint64_t base = 0; base < ntiles; base += blockDim.x) {
    uint64_t i = base + threadIdx.x;
    uint64_t v = i < ntiles ? tile_sums[i] : 0;
    uint64_t ex = block_exclusive_sum(v, wt, &total);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

// Position of item i of a tile in the staging buffer: one pad word per 16 keeps both access
// patterns below free of bank conflicts (striped: consecutive threads, consecutive words; blocked:
// thread t touches word 16 * t + j, which lands 17 words after thread t - 1's).
__device__ __forceinline__ int scan_slot(int i) { return i + (i >> 4); }

template <bool PAIR>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                  const uint64_t* __restrict__ tile_offsets,
                                                                  uint64_t* __restrict__ out) {
  __shared__ uint64_t wt[32];
  __shared__ uint64_t total;
  __shared__ uint64_t stage[kScanTile + kScanTile / 16];
  const uint64_t tile0 = (uint64_t)blockIdx.x * kScanTile;
  // global memory is touched in striped order (coalesced both ways); the per-thread prefix needs
  // kScanItems consecutive items per thread, so the tile is transposed through shared memory
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int local = j * kScanThreads + (int)threadIdx.x;
    const uint64_t i = tile0 + (uint64_t)local;
    stage[scan_slot(local)] = i < n ? in[i] : 0u;
  }
  __syncthreads();
  uint32_t v[kScanItems];
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = (uint32_t)stage[scan_slot((int)threadIdx.x * kScanItems + j)];
    s += scan_value<PAIR>(v[j]);
  }
  uint64_t ex = block_exclusive_sum(s, wt, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    stage[scan_slot((int)threadIdx.x * kScanItems + j)] = ex;
    ex += scan_value<PAIR>(v[j]);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int local = j * kScanThreads + (int)threadIdx.x;
    const uint64_t i = tile0 + (uint64_t)local;
    if (i < n) out[i] = stage[scan_slot(local)];
  }
}

// `tmp` must hold ceil(n / kScanTile) + 1 u64; out must hold n + 1 u64 (out[n] = total).
inline size_t scan_tmp_elems(uint64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile + 1); }

template <bool PAIR = false>
inline void exclusive_scan_u32(const uint32_t* in, uint64_t n, uint64_t* out, uint64_t* tmp,
                               cudaStream_t st) {
  if (n == 0) {
    TAPES_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(uint64_t), st));
    return;
  }
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  scan_tile_sums_kernel<PAIR><<<(unsigned)ntiles, kScanThreads, 0, st>>>(in, n, tmp);
  scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(tmp, ntiles, out + n);
  scan_apply_kernel<PAIR><<<(unsigned)ntiles, kScanThreads, 0, st>>>(in, n, tmp, out);
  TAPES_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
// Open-addressing hash set of 64-bit keys in HBM.  The first inserter of a key also stores a
// 32-bit value; `ranks` is filled later with the position of the key in the sorted key list.
// Lanes of a warp that carry the same key elect one inserter (warp-aggregated insert).
// ------------------------------------------------------------------------------------------
constexpr uint64_t kEmptyKey = ~0ull;

struct HashSet {
  uint64_t* keys = nullptr;
  uint32_t* vals = nullptr;
  uint32_t* ranks = nullptr;
  uint64_t mask = 0;  // capacity - 1 (capacity is a power of two)
  // set by an insertion that gave up (the table was sized on a guess of the number of distinct keys
  // and the guess was too small): the caller starts over with a larger table
  unsigned long long* full = nullptr;
  int run_bits = 2;  // keys that differ in this many low bits only are neighbours in the table (hash_home)
};

constexpr uint32_t kMaxProbes = 1u << 12;

// Advances a probe; false when the insertion has to be abandoned.
__device__ __forceinline__ bool hash_next_probe(const HashSet& hs, uint64_t& h, uint32_t& probes) {
  h = (h + 1) & hs.mask;
  if ((++probes & 63u) == 0 && (probes >= kMaxProbes || *(volatile unsigned long long*)hs.full)) {
    *(volatile unsigned long long*)hs.full = 1ull;
    return false;
  }
  return true;
}

__device__ __forceinline__ uint64_t hash_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// Home slot of a key.  Keys that differ in their lowest bits only - the prefixes of consecutive nodes
// of a level - get neighbouring slots, so that the probes of a warp share DRAM sectors (a sector holds
// 4 keys) instead of touching 32 of them; the rest of the key is mixed.
__device__ __forceinline__ uint64_t hash_home(const HashSet& hs, uint64_t key) {
  return ((hash_mix(key >> hs.run_bits) << hs.run_bits) | (key & ((1ull << hs.run_bits) - 1))) & hs.mask;
}

// All 32 lanes must call this; `active` says whether the lane has a key.  Returns true in the
// lane that claimed an empty slot for a key not seen before; *slot receives the slot of the lane's
// key (every active lane, so that later passes need not probe again).
__device__ __forceinline__ bool hash_insert_warp(const HashSet& hs, bool active, uint64_t key,
                                                 uint32_t val, uint32_t* slot) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(0xffffffffu, active ? key : (kEmptyKey - lane));
  const int leader_lane = __ffs(peers) - 1;
  const bool leader = active && (leader_lane == (int)lane);
  bool fresh = false;
  uint64_t h = 0;
  if (leader) {
    h = hash_home(hs, key);
    uint32_t probes = 0;
    for (;;) {
      unsigned long long prev =
          atomicCAS((unsigned long long*)&hs.keys[h], (unsigned long long)kEmptyKey,
                    (unsigned long long)key);
      if (prev == kEmptyKey) { hs.vals[h] = val; fresh = true; break; }
      if (prev == key) break;
      if (!hash_next_probe(hs, h, probes)) break;
    }
  }
  *slot = __shfl_sync(0xffffffffu, (uint32_t)h, leader_lane);  // capacity is at most 2^31
  return fresh;
}

// The same without the warp vote, for key streams in which the lanes of a warp hold different keys
// anyway (consecutive nodes of a level: their prefixes differ): every lane with a key inserts it.
__device__ __forceinline__ bool hash_insert_lane(const HashSet& hs, bool active, uint64_t key, uint32_t val,
                                                 uint32_t* slot) {
  if (!active) { *slot = 0; return false; }
  uint64_t h = hash_home(hs, key);
  bool fresh = false;
  uint32_t probes = 0;
  for (;;) {
    // most keys are in the table already (a prefix is reached from several parents): look before
    // the atomic.  A slot is written once, so a key read here is final; "empty" is settled by the CAS.
    unsigned long long seen = __ldcg((const unsigned long long*)&hs.keys[h]);
    if (seen == kEmptyKey) {
      seen = atomicCAS((unsigned long long*)&hs.keys[h], (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (seen == kEmptyKey) { hs.vals[h] = val; fresh = true; break; }
    }
    if (seen == key) break;
    if (!hash_next_probe(hs, h, probes)) break;
  }
  *slot = (uint32_t)h;
  return fresh;
}

// Slot of a key that is known to be present.
__device__ __forceinline__ uint64_t hash_slot(const HashSet& hs, uint64_t key) {
  uint64_t h = hash_home(hs, key);
  while (hs.keys[h] != key) h = (h + 1) & hs.mask;
  return h;
}

// Adds the block's sum of v to *counter with one global atomic per block (a counter that every warp
// of a large grid adds to is a serial bottleneck in L2).  All threads of the block must call this.
__device__ __forceinline__ void block_add(unsigned long long v, unsigned long long* __restrict__ counter) {
  __shared__ unsigned long long block_sum;
  if (threadIdx.x == 0) block_sum = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(&block_sum, v);
  __syncthreads();
  if (threadIdx.x == 0 && block_sum) atomicAdd(counter, block_sum);
}

// Block-wide compaction: every thread with `take` appends `item` to `list`; one atomic per block
// on the shared counter.  All threads of the block must call this.
__device__ __forceinline__ void block_append_u64(bool take, uint64_t item, uint64_t* __restrict__ list,
                                                 unsigned long long* __restrict__ counter) {
  __shared__ uint32_t warp_count[32];
  __shared__ unsigned long long block_base;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned takers = __ballot_sync(0xffffffffu, take);
  if (lane == 0) warp_count[warp] = __popc(takers);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (unsigned w = 0; w < (blockDim.x + 31) / 32; ++w) { const uint32_t c = warp_count[w]; warp_count[w] = total; total += c; }
    block_base = total ? atomicAdd(counter, (unsigned long long)total) : 0ull;
  }
  __syncthreads();
  if (take) list[block_base + warp_count[warp] + __popc(takers & ((1u << lane) - 1u))] = item;
}

// ------------------------------------------------------------------------------------------
// Stable LSD radix sort of u64 keys, 8 bits per pass, only over the byte positions the caller
// marks as significant.  Each block owns one contiguous chunk and ranks it tile by tile.
// ------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;

__global__ void radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint64_t chunk,
                                  int shift, uint32_t* __restrict__ hist /*[256][gridDim.x]*/) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t lo = (uint64_t)blockIdx.x * chunk;
  uint64_t hi = lo + chunk;
  if (hi > n) hi = n;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += kSortThreads)
    atomicAdd(&h[(keys[i] >> shift) & 255], 1u);
  __syncthreads();
  hist[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

__global__ void radix_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t* __restrict__ out,
                                     uint64_t n, uint64_t chunk, int shift,
                                     const uint64_t* __restrict__ offsets /*[256][gridDim.x]*/) {
  __shared__ uint64_t off[256];
  off[threadIdx.x] = offsets[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t lo = (uint64_t)blockIdx.x * chunk;
  uint64_t hi = lo + chunk;
  if (hi > n) hi = n;
  for (uint64_t t0 = lo; t0 < hi; t0 += kSortThreads) {
    const uint64_t i = t0 + threadIdx.x;
    const bool active = i < hi;
    const uint64_t key = active ? keys[i] : 0;
    const unsigned d = active ? (unsigned)((key >> shift) & 255) : (256u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    uint64_t base = 0;
    // warps take turns so that earlier items get smaller positions (stability)
    for (int w = 0; w < kSortThreads / 32; ++w) {
      if ((int)warp == w && active && (int)lane == leader) {
        base = off[d];
        off[d] = base + __popc(peers);
      }
      __syncthreads();
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (active) out[base + rank] = key;
  }
}

// tmp_hist: 256 * blocks u32; tmp_off: 256 * blocks + 1 u64; scan_tmp per scan_tmp_elems.
struct RadixPlan {
  unsigned blocks = 0;
  uint64_t chunk = 0;
};
inline RadixPlan radix_plan(uint64_t n) {
  RadixPlan p;
  uint64_t blocks = (n + 4095) / 4096;
  if (blocks > 1184) blocks = 1184;  // 8 x 148 SMs
  if (blocks == 0) blocks = 1;
  p.blocks = (unsigned)blocks;
  p.chunk = (n + blocks - 1) / blocks;
  return p;
}

// Sorts `keys` (n items) using `alt` as ping-pong space; returns the buffer holding the result.
inline uint64_t* radix_sort_u64(uint64_t* keys, uint64_t* alt, uint64_t n, uint64_t significant_bits,
                                uint32_t* tmp_hist, uint64_t* tmp_off, uint64_t* scan_tmp,
                                cudaStream_t st) {
  if (n <= 1) return keys;
  const RadixPlan plan = radix_plan(n);
  uint64_t* src = keys;
  uint64_t* dst = alt;
  for (int byte = 0; byte < 8; ++byte) {
    if (((significant_bits >> (8 * byte)) & 0xff) == 0) continue;
    const int shift = 8 * byte;
    radix_hist_kernel<<<plan.blocks, kSortThreads, 0, st>>>(src, n, plan.chunk, shift, tmp_hist);
    exclusive_scan_u32(tmp_hist, 256ull * plan.blocks, tmp_off, scan_tmp, st);
    radix_scatter_kernel<<<plan.blocks, kSortThreads, 0, st>>>(src, dst, n, plan.chunk, shift,
                                                              tmp_off);
    uint64_t* t = src; src = dst; dst = t;
  }
  TAPES_CUDA_CHECK(cudaGetLastError());
  return src;
}

}  // namespace
}  // namespace tapes
