// Monte-Carlo simulation of a tape program (sm_100a): an independent check of the closure the
// master equation rests on (the reference has one for the
// ferromagnet only, examples/ex2_ferromagnet_mc.py:46-122).
//
// One ring tape of n_sites cells carries both roles: an event puts the program head on a random
// site and the data head on another, independently, and runs the program there - reads see the
// ring, `choose` is sampled, writes go to the ring.  Every site is visited at rate 1, which is the
// normalisation of compute-dy/dt (a leaf world's probability is the probability of finding its
// cells at a random position).  Time advances in sub-steps of `events` simultaneous events
// (dt = events / n_sites): all events of a sub-step read the ring as it was at its start; where
// several write the same cell, the one with the highest event number wins (within an event, its last
// write).  That makes the process deterministic given the seed - tests compare the ring with a NumPy
// implementation of the same rules bit for bit - and the bias of simultaneous events is O(dt).
// Random numbers are counter-based: draw(seed, sub-step, event, index), index 0 / 1 = the heads,
// 2 + i = the i-th choice of the event.
#include "montecarlo.h"

#include <algorithm>
#include <stdexcept>
#include <vector>

#include "cuda_check.h"

namespace tapes {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxWrites = 16;
constexpr int kMaxWalk = 4096;  // nodes an event may visit: trees are acyclic, this guards corrupt ones

__host__ __device__ inline uint64_t mc_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__host__ __device__ inline uint64_t mc_draw(uint64_t seed, uint64_t substep, uint64_t event, uint64_t index) {
  uint64_t x = mc_mix(seed + 0x9E3779B97F4A7C15ull);
  x = mc_mix(x ^ substep);
  x = mc_mix(x + event * 0x9E3779B97F4A7C15ull);
  return mc_mix(x ^ index);
}

struct DeviceTree {
  const int32_t* kind; const int32_t* a; const int32_t* b; const int32_t* c;
  const int32_t* first_child; const int32_t* first_weight; const int32_t* child; const double* weight;
};

__global__ void __launch_bounds__(kThreads) mc_events_kernel(DeviceTree t, const uint8_t* __restrict__ tape,
                                                             uint32_t* __restrict__ stamp, uint64_t n_sites,
                                                             uint32_t events, uint64_t seed, uint64_t substep,
                                                             int* __restrict__ error) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= events) return;
  const uint64_t head[2] = {mc_draw(seed, substep, e, 0) % n_sites, mc_draw(seed, substep, e, 1) % n_sites};
  uint64_t wpos[kMaxWrites];
  uint32_t wsym[kMaxWrites];
  int n_writes = 0;
  uint64_t picks = 0;
  int node = 0;
  for (int walked = 0; walked < kMaxWalk; ++walked) {
    const int kind = t.kind[node];
    if (kind == ProgramTree::END) break;
    if (kind == ProgramTree::PICK) {
      const int n = t.a[node];
      const double* w = t.weight + t.first_weight[node];
      double total = 0.0;
      for (int j = 0; j < n; ++j) total = total + w[j];
      const double u = (double)(mc_draw(seed, substep, e, 2 + picks) >> 11) * 0x1.0p-53;
      ++picks;
      const double x = u * total;
      double cum = 0.0;
      int option = n - 1;
      for (int j = 0; j < n; ++j) {
        cum = cum + w[j];
        if (x < cum) { option = j; break; }
      }
      node = t.child[t.first_child[node] + option];
      continue;
    }
    // cell b relative to the head of tape a, on the ring
    const int64_t off = t.b[node];
    const uint64_t pos = (head[t.a[node]] + (uint64_t)(off % (int64_t)n_sites + (int64_t)n_sites)) % n_sites;
    if (kind == ProgramTree::READ) {
      uint32_t sym = tape[pos];
      for (int w = 0; w < n_writes; ++w)  // the event sees its own writes, the last one first
        if (wpos[w] == pos) sym = wsym[w];
      node = t.child[t.first_child[node] + (int)sym];
    } else {
      if (n_writes == kMaxWrites) { atomicExch(error, 1); return; }
      wpos[n_writes] = pos;
      wsym[n_writes] = (uint32_t)t.c[node];
      ++n_writes;
      node = t.child[t.first_child[node]];
    }
  }
  for (int w = 0; w < n_writes; ++w)
    atomicMax(&stamp[wpos[w]], ((e + 1u) << 12) | ((uint32_t)w << 8) | wsym[w]);
}

__global__ void mc_apply_kernel(uint8_t* __restrict__ tape, uint32_t* __restrict__ stamp, uint64_t n_sites) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sites) return;
  const uint32_t s = stamp[i];
  if (s) { tape[i] = (uint8_t)(s & 255u); stamp[i] = 0; }
}

__global__ void mc_count_kernel(const uint8_t* __restrict__ tape, uint64_t n_sites, uint32_t A, int k,
                                unsigned long long* __restrict__ hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sites) return;
  uint64_t idx = 0;
  for (int o = 0; o < k; ++o) {
    uint64_t pos = i + (uint64_t)o;
    if (pos >= n_sites) pos -= n_sites;
    idx = idx * A + tape[pos];
  }
  atomicAdd(&hist[idx], 1ull);
}

template <typename T>
T* to_device(const std::vector<T>& h) {
  void* p = nullptr;
  TAPES_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) TAPES_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return (T*)p;
}

}  // namespace

struct MonteCarlo {
  int A = 0;
  uint64_t n_sites = 0, seed = 0, substep = 0;
  uint32_t events = 0;
  uint8_t* d_tape = nullptr;
  uint32_t* d_stamp = nullptr;
  int* d_error = nullptr;
  int32_t *kind = nullptr, *a = nullptr, *b = nullptr, *c = nullptr, *first_child = nullptr, *first_weight = nullptr,
          *child = nullptr;
  double* weight = nullptr;
  cudaStream_t st = nullptr;
};

MonteCarlo* mc_create(const ProgramTree& tree, int alphabet, uint64_t n_sites, const uint8_t* tape0,
                      uint32_t events_per_substep, uint64_t seed) {
  if (alphabet < 1 || alphabet > 256) throw std::runtime_error("the simulator stores one byte per cell: alphabet <= 256");
  if (n_sites < 1) throw std::runtime_error("the ring needs at least one site");
  if (events_per_substep < 1 || events_per_substep >= (1u << 20)) throw std::runtime_error("events per sub-step must be in 1..2^20-1");
  for (uint64_t i = 0; i < n_sites; ++i)
    if (tape0[i] >= alphabet) throw std::runtime_error("initial ring holds a symbol outside the alphabet");
  MonteCarlo* mc = new MonteCarlo();
  try {
    mc->A = alphabet; mc->n_sites = n_sites; mc->seed = seed; mc->events = events_per_substep;
    TAPES_CUDA_CHECK(cudaStreamCreateWithFlags(&mc->st, cudaStreamNonBlocking));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&mc->d_tape, n_sites));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&mc->d_stamp, n_sites * 4));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&mc->d_error, sizeof(int)));
    TAPES_CUDA_CHECK(cudaMemcpy(mc->d_tape, tape0, n_sites, cudaMemcpyHostToDevice));
    TAPES_CUDA_CHECK(cudaMemset(mc->d_stamp, 0, n_sites * 4));
    TAPES_CUDA_CHECK(cudaMemset(mc->d_error, 0, sizeof(int)));
    mc->kind = to_device(tree.kind); mc->a = to_device(tree.a); mc->b = to_device(tree.b); mc->c = to_device(tree.c);
    mc->first_child = to_device(tree.first_child); mc->first_weight = to_device(tree.first_weight);
    mc->child = to_device(tree.child); mc->weight = to_device(tree.weight);
  } catch (...) {
    mc_destroy(mc);
    throw;
  }
  return mc;
}

void mc_destroy(MonteCarlo* mc) {
  if (!mc) return;
  if (mc->st) cudaStreamSynchronize(mc->st);
  for (void* p : {(void*)mc->d_tape, (void*)mc->d_stamp, (void*)mc->d_error, (void*)mc->kind, (void*)mc->a, (void*)mc->b,
                  (void*)mc->c, (void*)mc->first_child, (void*)mc->first_weight, (void*)mc->child, (void*)mc->weight})
    if (p) cudaFree(p);
  if (mc->st) cudaStreamDestroy(mc->st);
  delete mc;
}

void mc_run(MonteCarlo* mc, uint64_t n_substeps) {
  const DeviceTree t{mc->kind, mc->a, mc->b, mc->c, mc->first_child, mc->first_weight, mc->child, mc->weight};
  for (uint64_t s = 0; s < n_substeps; ++s) {
    mc_events_kernel<<<grid_for(mc->events, kThreads), kThreads, 0, mc->st>>>(t, mc->d_tape, mc->d_stamp, mc->n_sites,
                                                                             mc->events, mc->seed, mc->substep, mc->d_error);
    mc_apply_kernel<<<grid_for(mc->n_sites, kThreads), kThreads, 0, mc->st>>>(mc->d_tape, mc->d_stamp, mc->n_sites);
    ++mc->substep;
  }
  TAPES_CUDA_CHECK(cudaGetLastError());
  int err = 0;
  TAPES_CUDA_CHECK(cudaMemcpyAsync(&err, mc->d_error, sizeof(int), cudaMemcpyDeviceToHost, mc->st));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(mc->st));
  if (err) throw std::runtime_error("an event wrote more than 16 cells");
}

uint64_t mc_substeps_done(const MonteCarlo* mc) { return mc->substep; }

void mc_window_counts(MonteCarlo* mc, int cl_k, int64_t* h_counts) {
  double states = 1;
  for (int i = 0; i < cl_k; ++i) states *= mc->A;
  if (cl_k < 1 || states > 268435456.0) throw std::runtime_error("window table too large");
  const size_t n = (size_t)states;
  unsigned long long* hist = nullptr;
  TAPES_CUDA_CHECK(cudaMalloc((void**)&hist, n * 8));
  TAPES_CUDA_CHECK(cudaMemsetAsync(hist, 0, n * 8, mc->st));
  mc_count_kernel<<<grid_for(mc->n_sites, kThreads), kThreads, 0, mc->st>>>(mc->d_tape, mc->n_sites, (uint32_t)mc->A, cl_k, hist);
  cudaError_t err = cudaMemcpyAsync(h_counts, hist, n * 8, cudaMemcpyDeviceToHost, mc->st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(mc->st);
  cudaFree(hist);
  TAPES_CUDA_CHECK(err);
}

void mc_fetch(MonteCarlo* mc, uint8_t* h_tape) {
  TAPES_CUDA_CHECK(cudaMemcpyAsync(h_tape, mc->d_tape, mc->n_sites, cudaMemcpyDeviceToHost, mc->st));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(mc->st));
}

void mc_sample_ring(int alphabet, int cl_k, const double* table, uint64_t n_sites, uint64_t seed, uint8_t* h_tape) {
  if (alphabet < 1 || alphabet > 256 || cl_k < 1) throw std::runtime_error("needs 1 <= alphabet <= 256 and cl_k >= 1");
  uint64_t n = 1, m = 1;
  for (int i = 0; i < cl_k; ++i) { n *= (uint64_t)alphabet; if (i + 1 < cl_k) m *= (uint64_t)alphabet; }
  auto uniform = [&](uint64_t i) { return (double)(mc_draw(seed, ~0ull, i, 0) >> 11) * 0x1.0p-53; };
  // first window from the table itself
  double total = 0.0;
  for (uint64_t i = 0; i < n; ++i) total += table[i] > 0 ? table[i] : 0.0;
  if (!(total > 0)) throw std::runtime_error("the table has no positive entry");
  double x = uniform(0) * total, cum = 0.0;
  uint64_t first = n - 1;
  for (uint64_t i = 0; i < n; ++i) {
    cum += table[i] > 0 ? table[i] : 0.0;
    if (x < cum) { first = i; break; }
  }
  uint64_t context = first % m;  // the last k-1 symbols
  for (int i = cl_k - 1; i >= 0; --i) {
    if ((uint64_t)i < n_sites) h_tape[i] = (uint8_t)(first % (uint64_t)alphabet);
    first /= (uint64_t)alphabet;
  }
  for (uint64_t pos = (uint64_t)cl_k; pos < n_sites; ++pos) {
    const double* row = table + context * (uint64_t)alphabet;
    double row_total = 0.0;
    for (int s = 0; s < alphabet; ++s) row_total += row[s] > 0 ? row[s] : 0.0;
    int sym = 0;
    if (row_total > 0) {
      x = uniform(pos) * row_total; cum = 0.0; sym = alphabet - 1;
      for (int s = 0; s < alphabet; ++s) {
        cum += row[s] > 0 ? row[s] : 0.0;
        if (x < cum) { sym = s; break; }
      }
    } else {  // a context the table gives no weight: continue uniformly
      sym = (int)(mc_draw(seed, ~0ull, pos, 1) % (uint64_t)alphabet);
    }
    h_tape[pos] = (uint8_t)sym;
    context = (context * (uint64_t)alphabet + (uint64_t)sym) % m;
  }
}

namespace {

constexpr int kChainThreads = 512;

// One trial per block.  Shared memory: previous and current chain (bytes, padded to words) and six
// counters.  Per time step: the trials of the step read `prev` and xor into `cur`; islands of `cur`
// are counted from their first site; `cur` becomes `prev`.
__global__ void __launch_bounds__(kChainThreads) ferromagnet_chain_kernel(int64_t chain_length, int64_t n_steps,
                                                                          int64_t trials_per_step,
                                                                          const uint8_t* __restrict__ chain0,
                                                                          const int32_t* __restrict__ sites,
                                                                          const double* __restrict__ uniforms,
                                                                          const double* __restrict__ accept,
                                                                          double* __restrict__ counts) {
  extern __shared__ unsigned int smem[];
  const int64_t words = (chain_length + 3) / 4;
  unsigned int* prev_w = smem;
  unsigned int* cur_w = smem + words;
  unsigned int* tally = smem + 2 * words;  // [6]
  uint8_t* prev = (uint8_t*)prev_w;
  uint8_t* cur = (uint8_t*)cur_w;
  const int64_t trial = blockIdx.x;
  const int64_t n = chain_length;
  double acc[6];
  for (int i = 0; i < 6; ++i) acc[i] = accept[i];
  for (int64_t w = threadIdx.x; w < words; w += blockDim.x) { prev_w[w] = 0; cur_w[w] = 0; }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { prev[i] = chain0[trial * n + i]; cur[i] = prev[i]; }
  __syncthreads();
  for (int64_t step = 0; step < n_steps; ++step) {
    if (step > 0) {
      const int32_t* my_sites = sites + (trial * (n_steps - 1) + (step - 1)) * trials_per_step;
      const double* my_u = uniforms + (trial * (n_steps - 1) + (step - 1)) * trials_per_step;
      for (int64_t t = threadIdx.x; t < trials_per_step; t += blockDim.x) {
        const int64_t i = my_sites[t];
        const int64_t left = i == 0 ? n - 1 : i - 1, right = i + 1 == n ? 0 : i + 1;
        const int mid = prev[i];
        const int equal = (prev[left] == mid) + (prev[right] == mid);
        if (my_u[t] < acc[2 * equal + mid]) atomicXor(&cur_w[i >> 2], 1u << (8 * (i & 3)));
      }
      __syncthreads();
    }
    if (threadIdx.x < 6) tally[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      if (cur[i] && !cur[i == 0 ? n - 1 : i - 1]) {  // first site of an island
        int len = 1;
        int64_t at = i + 1 == n ? 0 : i + 1;
        while (len < 6 && cur[at]) { ++len; at = at + 1 == n ? 0 : at + 1; }
        if (len < 6) atomicAdd(&tally[len], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < 6) counts[(trial * n_steps + step) * 6 + threadIdx.x] = threadIdx.x ? (double)tally[threadIdx.x] : 0.0;
    for (int64_t w = threadIdx.x; w < words; w += blockDim.x) prev_w[w] = cur_w[w];
    __syncthreads();
  }
}

}  // namespace

void mc_ferromagnet_chains(int64_t n_trials, int64_t chain_length, int64_t n_steps, int64_t trials_per_step,
                           const uint8_t* chain0, const int32_t* sites, const double* uniforms, const double* accept,
                           double* counts) {
  if (n_trials < 1 || chain_length < 3 || n_steps < 1 || trials_per_step < 1) throw std::runtime_error("bad sizes");
  const size_t smem = (size_t)(2 * ((chain_length + 3) / 4) + 8) * sizeof(unsigned int);
  if (smem > 200 * 1024) throw std::runtime_error("the chain must fit shared memory twice (at most ~100 000 sites)");
  for (int64_t i = 0; i < n_trials * (n_steps - 1) * trials_per_step; ++i)
    if (sites[i] < 0 || sites[i] >= chain_length) throw std::runtime_error("site outside the chain");
  TAPES_CUDA_CHECK(cudaFuncSetAttribute(ferromagnet_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t n_draws = (size_t)(n_trials * (n_steps - 1) * trials_per_step);
  uint8_t* d_chain = nullptr; int32_t* d_sites = nullptr; double* d_u = nullptr; double* d_acc = nullptr; double* d_counts = nullptr;
  auto release = [&]() { cudaFree(d_chain); cudaFree(d_sites); cudaFree(d_u); cudaFree(d_acc); cudaFree(d_counts); };
  try {
    TAPES_CUDA_CHECK(cudaMalloc((void**)&d_chain, (size_t)(n_trials * chain_length)));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&d_sites, std::max<size_t>(n_draws, 1) * 4));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&d_u, std::max<size_t>(n_draws, 1) * 8));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&d_acc, 6 * 8));
    TAPES_CUDA_CHECK(cudaMalloc((void**)&d_counts, (size_t)(n_trials * n_steps * 6) * 8));
    TAPES_CUDA_CHECK(cudaMemcpy(d_chain, chain0, (size_t)(n_trials * chain_length), cudaMemcpyHostToDevice));
    if (n_draws) {
      TAPES_CUDA_CHECK(cudaMemcpy(d_sites, sites, n_draws * 4, cudaMemcpyHostToDevice));
      TAPES_CUDA_CHECK(cudaMemcpy(d_u, uniforms, n_draws * 8, cudaMemcpyHostToDevice));
    }
    TAPES_CUDA_CHECK(cudaMemcpy(d_acc, accept, 6 * 8, cudaMemcpyHostToDevice));
    ferromagnet_chain_kernel<<<(unsigned)n_trials, kChainThreads, smem>>>(chain_length, n_steps, trials_per_step, d_chain, d_sites,
                                                                         d_u, d_acc, d_counts);
    TAPES_CUDA_CHECK(cudaGetLastError());
    TAPES_CUDA_CHECK(cudaMemcpy(counts, d_counts, (size_t)(n_trials * n_steps * 6) * 8, cudaMemcpyDeviceToHost));
  } catch (...) {
    release();
    throw;
  }
  release();
}

}  // namespace tapes
