// Probe of the deep forest levels (GPU box): one thread per prefix group, A parents per group, the
// access pattern of own_parents<., true> on synthetic arrays, with parts of the work removed one at
// a time to see what bounds the kernel.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chain_probe chain_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ double child_weight(double w_parent, double p_long, double p_short) {
  if (p_long == 0.0) return 0.0;
  const double r = p_long / fmax(p_long, p_short);
  return r > 0.0 ? w_parent * r : 0.0;
}

// MODE 0: as in the product; 1: no division (multiply instead); 2: no stores of the parent weights;
// 3: parents' table reads contiguous per warp but one stream only (j-major launch: thread = (j, g))
template <int UO, int MODE, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) probe(const double* __restrict__ p, const double* __restrict__ short_table,
                                                         const uint32_t* __restrict__ g_first, const uint32_t* __restrict__ g_stride,
                                                         const uint32_t* __restrict__ g_prefix, const double* __restrict__ prev_total,
                                                         double* __restrict__ g_total, double* __restrict__ ww, uint32_t n_groups,
                                                         uint32_t A, uint32_t M, uint32_t base) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const uint32_t first = g_first[g], stride = g_stride[g], mine = g_prefix[g];
  const uint32_t rel = first - base;
  const uint32_t g_prev = rel / A, g_step = stride / A;
  const uint32_t mine_short = mine / A, short_step = M / A;
  double total = 0.0;
  for (uint32_t e = 0; e < A; e += UO) {
    double sum_prev[UO], p_long[UO], p_marg[UO];
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      const bool live = e + u < A;
      sum_prev[u] = live ? prev_total[g_prev + (e + u) * g_step] : 0.0;
      p_long[u] = live ? p[(e + u) * M + mine] : 0.0;
      p_marg[u] = live ? short_table[(e + u) * short_step + mine_short] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      const bool live = e + u < A;
      double v;
      if (MODE == 1) v = live ? sum_prev[u] * p_long[u] * p_marg[u] : 0.0;
      else v = live ? child_weight(sum_prev[u], p_long[u], p_marg[u]) : 0.0;
      if (MODE != 2 && live) ww[first + (e + u) * stride] = v;
      total += v;
    }
  }
  g_total[g] = total;
}

// copy-like reference: one thread per parent node, same bytes (read p, write ww), no group structure
__global__ void __launch_bounds__(256) flat(const double* __restrict__ p, const double* __restrict__ short_table,
                                            const double* __restrict__ prev_total, double* __restrict__ ww, uint64_t n,
                                            uint32_t A) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ww[i] = child_weight(prev_total[i / A], p[i], short_table[i / A]);
}

int main() {
  const uint32_t A = 10, M = 10000000;  // A^(k-1), k = 8
  const uint32_t G = 2 * M;             // two seeds' worth of groups, like a deep level of the bench
  const uint64_t n_nodes = (uint64_t)G * A;
  double *p, *short_table, *prev_total, *g_total, *ww;
  uint32_t *g_first, *g_stride, *g_prefix;
  cudaMalloc(&p, (uint64_t)A * M * 8); cudaMalloc(&short_table, (uint64_t)M * 8);
  cudaMalloc(&prev_total, (uint64_t)G * 8); cudaMalloc(&g_total, (uint64_t)G * 8);
  cudaMalloc(&ww, n_nodes * 8);
  cudaMalloc(&g_first, G * 4); cudaMalloc(&g_stride, G * 4); cudaMalloc(&g_prefix, G * 4);
  std::vector<uint32_t> h_first(G), h_stride(G), h_prefix(G);
  for (uint32_t g = 0; g < G; ++g) {
    const uint32_t seed = g / M, prefix = g % M;
    // parent j = child x = prefix % A of previous group (seed, j * M / A + prefix / A): node id
    h_first[g] = (seed * M + prefix / A) * A + prefix % A;
    h_stride[g] = (M / A) * A;
    h_prefix[g] = prefix;
  }
  cudaMemcpy(g_first, h_first.data(), G * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(g_stride, h_stride.data(), G * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(g_prefix, h_prefix.data(), G * 4, cudaMemcpyHostToDevice);
  std::vector<double> h((uint64_t)A * M);
  for (uint64_t i = 0; i < h.size(); ++i) h[i] = 1e-8 * (1 + (i * 2654435761u) % 1000);
  cudaMemcpy(p, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  for (uint64_t i = 0; i < M; ++i) h[i] = 1e-7 * (1 + (i * 40503u) % 1000);
  cudaMemcpy(short_table, h.data(), (uint64_t)M * 8, cudaMemcpyHostToDevice);
  for (uint64_t i = 0; i < G; ++i) h[i] = 1e-3;
  cudaMemcpy(prev_total, h.data(), (uint64_t)G * 8, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = n_nodes * 16.0 + G * (12.0 + 4 + 8 + 8);
  auto time = [&](const char* name, auto launch) {
    for (int i = 0; i < 2; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("%-64s %7.3f ms  %6.0f GB/s of %.2f GB  (%s)\n", name, ms, bytes / ms * 1e-6, bytes * 1e-9, cudaGetErrorString(cudaGetLastError()));
  };
  const unsigned grid = (G + 255) / 256;
#define RUN(UO, MODE, MB, label) time(label, [&] { probe<UO, MODE, MB><<<grid, 256>>>(p, short_table, g_first, g_stride, g_prefix, prev_total, g_total, ww, G, A, M, 0); })
  RUN(3, 0, 5, "as the product: 3 parents per batch, 5 blocks/SM");
  RUN(3, 1, 5, "no division (multiply), 3 per batch, 5 blocks/SM");
  RUN(3, 2, 5, "no parent stores, 3 per batch, 5 blocks/SM");
  RUN(2, 0, 6, "2 per batch, 6 blocks/SM");
  RUN(5, 0, 4, "5 per batch, 4 blocks/SM");
  RUN(10, 0, 2, "10 per batch, 2 blocks/SM");
  RUN(10, 0, 3, "10 per batch, 3 blocks/SM");
  RUN(10, 1, 3, "no division, 10 per batch, 3 blocks/SM");
  RUN(5, 1, 4, "no division, 5 per batch, 4 blocks/SM");
  RUN(1, 0, 8, "1 per batch, 8 blocks/SM");
  time("flat: one thread per node, contiguous reads and writes", [&] { flat<<<(unsigned)((n_nodes + 255) / 256), 256>>>(p, short_table, prev_total, ww, n_nodes / 2, A); });
  cudaDeviceSynchronize();
  return 0;
}
