// Validation of a subsequence-probability table at scale (sm_100a).
//
// The reference checks a starting table with get_ctm_eigenvalue1_eigenspace
// (framework/markov_tapes.py:133-175): left and right (k-1)-marginals must agree, and the context
// marginal must lie in the eigenvalue-1 eigenspace of the context transfer matrix
// (ctm_from_mpp, markov_tapes.py:107-130), found by a dense eigen-decomposition of an
// A^(k-1) x A^(k-1) matrix ("takes some time (and a nontrivial amount of RAM)" at 9^4,
// examples/ex4_chemical_turing.py:88-95; impossible at 10^7 contexts).  The transfer matrix has
// only A non-zeros per row (context (s1..s_{k-1}) -> (s2..s_{k-1}, s) with probability
// mpp[s1..s_{k-1}, s], mpp_from_spd, markov_tapes.py:81-104), so the same questions are answered
// by streaming kernels: the marginal distance, the residual |T pi - pi| of the context marginal,
// and a power iteration v <- T v from the uniform vector whose limit is compared with pi
// (a second stationary vector, i.e. an eigenspace of dimension > 1, shows up as a limit != pi).
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "cuda_check.h"
#include "validate.h"

namespace tapes {

namespace {

constexpr int kThreads = 256;
constexpr int kBlocks = 1184;  // 8 x 148 SMs

// right[j] = sum_s p[j A + s] (last-axis marginal), clipped[j] = sum_s clip(p[j A + s], eps, 1)
// (the normaliser of mpp_from_spd), left[j] = sum_x p[x M + j] (first-axis marginal).
__global__ void context_marginals_kernel(const double* __restrict__ p, double* __restrict__ right,
                                         double* __restrict__ clipped, double* __restrict__ left, uint64_t M,
                                         uint32_t A, double eps) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  double r = 0.0, c = 0.0, l = 0.0;
  for (uint32_t s = 0; s < A; ++s) {
    const double v = p[j * A + s];
    r += v;
    c += fmin(fmax(v, eps), 1.0);
  }
  for (uint32_t x = 0; x < A; ++x) l += p[(uint64_t)x * M + j];
  right[j] = r; clipped[j] = c; left[j] = l;
}

// out[(q, s)] = sum_x v[(x, q)] * mpp[(x, q), s]   with q = the k-2 digits two contexts share
__global__ void transfer_apply_kernel(const double* __restrict__ p, const double* __restrict__ clipped,
                                      const double* __restrict__ v, double* __restrict__ out, uint64_t M,
                                      uint32_t A, double eps) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const uint64_t q = j / A, stride = M / A;
  const uint32_t s = (uint32_t)(j - q * A);
  double acc = 0.0;
  for (uint32_t x = 0; x < A; ++x) {
    const uint64_t i = (uint64_t)x * stride + q;
    acc += v[i] * (fmin(fmax(p[i * A + s], eps), 1.0) / clipped[i]);
  }
  out[j] = acc;
}

__global__ void fill_kernel(double* __restrict__ v, uint64_t n, double value) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = value;
}

// partial[b] = sum (a - b)^2, partial[gridDim.x + b] = sum a  over the block's elements
__global__ void diff_partials_kernel(const double* __restrict__ a, const double* __restrict__ b, uint64_t n,
                                     double* __restrict__ partial) {
  __shared__ double s2[kThreads], s1[kThreads];
  double d2 = 0.0, d1 = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double d = a[i] - b[i];
    d2 += d * d;
    d1 += a[i];
  }
  s2[threadIdx.x] = d2; s1[threadIdx.x] = d1;
  __syncthreads();
  for (int h = kThreads / 2; h > 0; h >>= 1) {
    if ((int)threadIdx.x < h) { s2[threadIdx.x] += s2[threadIdx.x + h]; s1[threadIdx.x] += s1[threadIdx.x + h]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[blockIdx.x] = s2[0]; partial[gridDim.x + blockIdx.x] = s1[0]; }
}

struct Scratch {
  std::vector<void*> owned;
  double* take(uint64_t n) {
    void* p = nullptr;
    TAPES_CUDA_CHECK(cudaMalloc(&p, std::max<uint64_t>(n, 1) * sizeof(double)));
    owned.push_back(p);
    return (double*)p;
  }
  ~Scratch() { for (void* p : owned) cudaFree(p); }
};

// sqrt(sum (a - b)^2) and sum a
void distance_and_mass(const double* a, const double* b, uint64_t n, double* d_partial, cudaStream_t st,
                       double* distance, double* mass) {
  diff_partials_kernel<<<kBlocks, kThreads, 0, st>>>(a, b, n, d_partial);
  std::vector<double> h(2 * kBlocks);
  TAPES_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_partial, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(st));
  double s2 = 0.0, s1 = 0.0;
  for (int i = 0; i < kBlocks; ++i) { s2 += h[i]; s1 += h[kBlocks + i]; }
  *distance = std::sqrt(s2);
  if (mass) *mass = s1;
}

}  // namespace

TableCheck check_table(int alphabet, int cl_k, const double* d_p, double eps_mpp, int max_iterations,
                       double tolerance, cudaStream_t st) {
  if (alphabet < 1 || cl_k < 2) throw std::runtime_error("check_table needs an alphabet and cl_k >= 2");
  uint64_t M = 1;
  for (int i = 0; i + 1 < cl_k; ++i) {
    M *= (uint64_t)alphabet;
    if (M >= (1ull << 32)) throw std::runtime_error("A^cl_k must be below 2^32");
  }
  const uint32_t A = (uint32_t)alphabet;
  Scratch scratch;
  double* right = scratch.take(M);
  double* clipped = scratch.take(M);
  double* left = scratch.take(M);
  double* v0 = scratch.take(M);
  double* v1 = scratch.take(M);
  double* partial = scratch.take(2 * kBlocks);
  const unsigned grid = grid_for(M, kThreads);
  TableCheck out;
  context_marginals_kernel<<<grid, kThreads, 0, st>>>(d_p, right, clipped, left, M, A, eps_mpp);
  distance_and_mass(right, left, M, partial, st, &out.marginal_distance, &out.total);
  // residual of the context marginal under one application of the transfer matrix
  transfer_apply_kernel<<<grid, kThreads, 0, st>>>(d_p, clipped, left, v1, M, A, eps_mpp);
  distance_and_mass(v1, left, M, partial, st, &out.stationarity_residual, nullptr);
  // power iteration from the uniform vector
  out.iterations = 0;
  out.last_change = 0.0;
  out.power_distance = 0.0;
  if (max_iterations > 0) {
    fill_kernel<<<grid, kThreads, 0, st>>>(v0, M, 1.0 / (double)M);
    for (int it = 0; it < max_iterations; ++it) {
      transfer_apply_kernel<<<grid, kThreads, 0, st>>>(d_p, clipped, v0, v1, M, A, eps_mpp);
      std::swap(v0, v1);
      ++out.iterations;
      if ((it & 7) == 7 || it + 1 == max_iterations) {
        distance_and_mass(v0, v1, M, partial, st, &out.last_change, nullptr);
        if (out.last_change <= tolerance) break;
      }
    }
    distance_and_mass(v0, left, M, partial, st, &out.power_distance, nullptr);
  }
  TAPES_CUDA_CHECK(cudaGetLastError());
  return out;
}

}  // namespace tapes
