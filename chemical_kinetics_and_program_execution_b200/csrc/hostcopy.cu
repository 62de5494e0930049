// See hostcopy.h.
#include "hostcopy.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "cuda_check.h"

namespace tapes {

bool is_pinned_host(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeHost;
}

namespace {

constexpr size_t kBounceBytes = 2u << 20;  // per buffer; two per worker
constexpr size_t kPieceBytes = 8u << 20;   // unit of work handed to a worker

struct Job {
  char* device = nullptr;
  char* host = nullptr;
  size_t bytes = 0;
  bool to_device = true;
  std::atomic<size_t> next_piece{0};
  size_t n_pieces = 0;
  std::atomic<int> pending{0};
  std::string error;
  std::mutex error_mutex;
};

class Pool {
 public:
  explicit Pool(int device, int n_workers) : device_(device) {
    for (int i = 0; i < n_workers; ++i) threads_.emplace_back([this] { work(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      stop_ = true;
    }
    wake_.notify_all();
    for (std::thread& t : threads_) t.join();
  }
  int device() const { return device_; }

  void run(Job& job) {
    job.n_pieces = (job.bytes + kPieceBytes - 1) / kPieceBytes;
    job.pending = (int)threads_.size();
    {
      std::lock_guard<std::mutex> lock(mutex_);
      job_ = &job;
      ++generation_;
    }
    wake_.notify_all();
    std::unique_lock<std::mutex> lock(mutex_);
    done_.wait(lock, [&] { return job.pending.load() == 0; });
    job_ = nullptr;
    if (!job.error.empty()) throw std::runtime_error(job.error);
  }

 private:
  void work() {
    cudaSetDevice(device_);
    cudaStream_t stream = nullptr;
    char* bounce[2] = {nullptr, nullptr};
    cudaEvent_t free_again[2] = {nullptr, nullptr};
    bool ok = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaHostAlloc((void**)&bounce[i], kBounceBytes, cudaHostAllocDefault) == cudaSuccess &&
           cudaEventCreateWithFlags(&free_again[i], cudaEventDisableTiming) == cudaSuccess;
    uint64_t seen = 0;
    for (;;) {
      Job* job = nullptr;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&] { return stop_ || generation_ != seen; });
        if (stop_) break;
        seen = generation_;
        job = job_;
      }
      if (!job) continue;
      std::string problem = ok ? "" : "staged copy: could not allocate a worker's pinned buffers";
      int slot = 0;
      bool used[2] = {false, false};
      while (problem.empty()) {
        const size_t piece = job->next_piece.fetch_add(1);
        if (piece >= job->n_pieces) break;
        const size_t lo = piece * kPieceBytes, hi = std::min(job->bytes, lo + kPieceBytes);
        const size_t n_sub = (hi - lo + kBounceBytes - 1) / kBounceBytes;
        auto sub_len = [&](size_t i) { return std::min(kBounceBytes, hi - (lo + i * kBounceBytes)); };
        cudaError_t err = cudaSuccess;
        if (job->to_device) {
          // memcpy of sub-chunk i + 1 into one buffer while the DMA of sub-chunk i reads the other
          for (size_t i = 0; i < n_sub && err == cudaSuccess; ++i) {
            const size_t at = lo + i * kBounceBytes;
            if (used[slot]) err = cudaEventSynchronize(free_again[slot]);  // its last DMA has read the buffer
            if (err != cudaSuccess) break;
            std::memcpy(bounce[slot], job->host + at, sub_len(i));
            err = cudaMemcpyAsync(job->device + at, bounce[slot], sub_len(i), cudaMemcpyHostToDevice, stream);
            if (err == cudaSuccess) err = cudaEventRecord(free_again[slot], stream);
            used[slot] = true;
            slot ^= 1;
          }
        } else {
          // DMA of sub-chunk i + 1 into one buffer while sub-chunk i is copied out of the other
          auto fetch = [&](size_t i) {
            cudaError_t e = cudaMemcpyAsync(bounce[i & 1], job->device + lo + i * kBounceBytes, sub_len(i),
                                            cudaMemcpyDeviceToHost, stream);
            return e == cudaSuccess ? cudaEventRecord(free_again[i & 1], stream) : e;
          };
          err = fetch(0);
          for (size_t i = 0; i < n_sub && err == cudaSuccess; ++i) {
            if (i + 1 < n_sub) err = fetch(i + 1);
            if (err == cudaSuccess) err = cudaEventSynchronize(free_again[i & 1]);
            if (err == cudaSuccess) std::memcpy(job->host + lo + i * kBounceBytes, bounce[i & 1], sub_len(i));
          }
          used[0] = used[1] = false;  // every DMA of this piece has been waited for
        }
        if (err != cudaSuccess) problem = std::string("staged copy: ") + cudaGetErrorString(err);
      }
      if (ok && cudaStreamSynchronize(stream) != cudaSuccess && problem.empty()) problem = "staged copy: stream failed";
      if (!problem.empty()) {
        cudaGetLastError();
        std::lock_guard<std::mutex> lock(job->error_mutex);
        if (job->error.empty()) job->error = problem;
      }
      {
        std::lock_guard<std::mutex> lock(mutex_);
        job->pending.fetch_sub(1);
      }
      done_.notify_all();
    }
    for (int i = 0; i < 2; ++i) {
      if (bounce[i]) cudaFreeHost(bounce[i]);
      if (free_again[i]) cudaEventDestroy(free_again[i]);
    }
    if (stream) cudaStreamDestroy(stream);
  }

  int device_;
  std::vector<std::thread> threads_;
  std::mutex mutex_;
  std::condition_variable wake_, done_;
  Job* job_ = nullptr;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

Pool* g_pool = nullptr;

Pool& pool() {
  int dev = 0;
  TAPES_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_pool && g_pool->device() != dev) { delete g_pool; g_pool = nullptr; }
  if (!g_pool) {
    int workers = (int)std::min<unsigned>(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
    if (const char* e = std::getenv("TAPES_COPY_THREADS")) workers = std::max(1, std::min(64, std::atoi(e)));
    g_pool = new Pool(dev, workers);
  }
  return *g_pool;
}

void run(void* device, void* host, size_t bytes, bool to_device) {
  if (bytes == 0) return;
  Job job;
  job.device = (char*)device; job.host = (char*)host; job.bytes = bytes; job.to_device = to_device;
  pool().run(job);
}

}  // namespace

void staged_h2d(void* d_dst, const void* h_src, size_t bytes) { run(d_dst, const_cast<void*>(h_src), bytes, true); }
void staged_d2h(void* h_dst, const void* d_src, size_t bytes) { run(const_cast<void*>(d_src), h_dst, bytes, false); }

void staged_copy_shutdown() {
  delete g_pool;
  g_pool = nullptr;
}

}  // namespace tapes
