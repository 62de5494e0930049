// Copies between pageable host memory and the device through pinned bounce buffers, with several
// host threads doing the memcpy halves in parallel.
//
// The reference's binding hands c_compute_dy_dt plain NumPy buffers (framework/markov_tapes.py:275-288):
// pageable memory.  cudaMemcpy stages such copies through one driver-internal buffer on the calling
// thread, about 8 GB/s on the bench hosts - 100 ms each way for a 10^8-state table, ten times the
// right-hand side itself.  Here `workers` threads each own a stream and two pinned buffers and move
// interleaved pieces, so the host-side memcpy runs at several threads' bandwidth and overlaps the DMA.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace tapes {

// True when `p` is host memory the driver can DMA from / to directly (cudaHostAlloc, cudaHostRegister).
bool is_pinned_host(const void* p);

// Blocking: returns when all bytes have arrived.  Throw std::runtime_error on CUDA errors.
void staged_h2d(void* d_dst, const void* h_src, size_t bytes);
void staged_d2h(void* h_dst, const void* d_src, size_t bytes);

// Stops the worker threads and frees their buffers (cleanup_gambit).
void staged_copy_shutdown();

}  // namespace tapes
