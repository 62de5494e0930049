// Device-resident DOP853 (see integrate.h).  All vector work is elementwise fp64 streaming, bound
// by HBM bandwidth; reductions use a fixed two-pass order so that step sizes are reproducible.
#include "integrate.h"

#include <cmath>
#include <cstdlib>
#include <limits>
#include <vector>

#include "cuda_check.h"

namespace tapes {

namespace {

constexpr int kThreads = 256;
constexpr int kReduceBlocks = 1184;  // 8 x 148 SMs

struct Terms {  // a linear combination  sum_j coef[j] * vec[j]
  int n = 0;
  double coef[16];
  const double* vec[16];
};

// out = y + (sum_j coef_j * K_j) * h        (rk_step: dy = np.dot(K[:s].T, a[:s]) * h)
__global__ void lincomb_kernel(double* __restrict__ out, const double* __restrict__ y, Terms t, double h,
                               uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
#pragma unroll 4
  for (int j = 0; j < t.n; ++j) acc += t.vec[j][i] * t.coef[j];
  out[i] = y[i] + acc * h;
}

// out = (sum_j coef_j * K_j) * h            (F[3:] = h * np.dot(D, K))
__global__ void scaled_sum_kernel(double* __restrict__ out, Terms t, double h, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
#pragma unroll 4
  for (int j = 0; j < t.n; ++j) acc += t.vec[j][i] * t.coef[j];
  out[i] = h * acc;
}

__device__ __forceinline__ double block_sum(double v, double* smem) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0.0;
    for (int d = 16; d > 0; d >>= 1) r += __shfl_down_sync(0xffffffffu, r, d);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// partial[b] = sum (err5/scale)^2, partial[B + b] = sum (err3/scale)^2 over the block's elements,
// scale = atol + max(|y|, |y_new|) * rtol   (RungeKutta._step_impl, DOP853._estimate_error_norm)
__global__ void error_partials_kernel(Terms e5, Terms e3, const double* __restrict__ y,
                                      const double* __restrict__ y_new, double rtol, double atol, uint64_t n,
                                      double* __restrict__ partial) {
  __shared__ double smem[32];
  double s5 = 0.0, s3 = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    double a5 = 0.0, a3 = 0.0;
    for (int j = 0; j < e5.n; ++j) a5 += e5.vec[j][i] * e5.coef[j];
    for (int j = 0; j < e3.n; ++j) a3 += e3.vec[j][i] * e3.coef[j];
    const double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
    a5 /= scale; a3 /= scale;
    s5 += a5 * a5; s3 += a3 * a3;
  }
  const double b5 = block_sum(s5, smem);
  const double b3 = block_sum(s3, smem);
  if (threadIdx.x == 0) { partial[blockIdx.x] = b5; partial[gridDim.x + blockIdx.x] = b3; }
}

// partial[b] = sum ((a - b_or_0) / (atol + |y| rtol))^2   (select_initial_step norms)
__global__ void scaled_norm_partials_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                            const double* __restrict__ y, double rtol, double atol, uint64_t n,
                                            double* __restrict__ partial) {
  __shared__ double smem[32];
  double s = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double v = (a[i] - (b ? b[i] : 0.0)) / (atol + fabs(y[i]) * rtol);
    s += v * v;
  }
  const double r = block_sum(s, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void final_sum_kernel(const double* __restrict__ partial, int n_partials, int n_sums,
                                 double* __restrict__ out) {
  __shared__ double smem[32];
  for (int k = 0; k < n_sums; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) s += partial[k * n_partials + i];
    const double r = block_sum(s, smem);
    if (threadIdx.x == 0) out[k] = r;
  }
}

// F[0] = y - y_old ; F[1] = h f_old - F[0] ; F[2] = 2 F[0] - h (f + f_old)
__global__ void dense_head_kernel(double* __restrict__ F0, double* __restrict__ F1, double* __restrict__ F2,
                                  const double* __restrict__ y, const double* __restrict__ y_old,
                                  const double* __restrict__ f_old, const double* __restrict__ f, double h,
                                  uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double delta = y[i] - y_old[i];
  F0[i] = delta;
  F1[i] = h * f_old[i] - delta;
  F2[i] = 2 * delta - h * (f[i] + f_old[i]);
}

struct SevenVecs { const double* F[7]; };

// Dop853DenseOutput._call_impl for one time point.
__global__ void dense_eval_kernel(double* __restrict__ out, SevenVecs fv, const double* __restrict__ y_old,
                                  double x, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double y = 0.0;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    y += fv.F[6 - r][i];
    y *= (r % 2 == 0) ? x : (1 - x);
  }
  out[i] = y + y_old[i];
}

// Strided sums over the whole grid: block (b, o) adds the elements  b * chunk <= j < (b + 1) * chunk  of
// observable o (chunk = ceil(count / gridDim.x)), then one block per observable adds the partial sums in
// block order.  Fixed launch shape per (count, n_obs), so the sums are reproducible.  A one-symbol
// observable at 10^8 states touches every sector of the table: one block per observable (the first
// version) read it through one SM.
__global__ void __launch_bounds__(kThreads) observe_partials_kernel(const double* __restrict__ y,
                                                                    const int64_t* __restrict__ offset,
                                                                    const int64_t* __restrict__ stride,
                                                                    const int64_t* __restrict__ count,
                                                                    double* __restrict__ partial) {
  __shared__ double smem[32];
  const int o = blockIdx.y;
  const int64_t off = offset[o], st = stride[o], cnt = count[o];
  const int64_t chunk = (cnt + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = lo + chunk < cnt ? lo + chunk : cnt;
  double s = 0.0;
  for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) s += y[off + j * st];
  const double r = block_sum(s, smem);
  if (threadIdx.x == 0) partial[(int64_t)o * gridDim.x + blockIdx.x] = r;
}

__global__ void observe_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ out) {
  __shared__ double smem[32];
  const int o = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < blocks; i += blockDim.x) s += partial[(int64_t)o * blocks + i];
  const double r = block_sum(s, smem);
  if (threadIdx.x == 0) out[o] = r;
}

// Probability of a sequence longer than the window (framework/markov_tapes.py:223-233): the table entry
// of the first k symbols times, for every later window, the Markov process parameter
// clip(p[window], eps, 1) / sum_s clip(p[prefix of the window, s], eps, 1)  (markov_tapes.py:81-104),
// multiplied in sequence order.  One block per sequence, one thread per window.
__global__ void observe_long_kernel(const double* __restrict__ y, const int64_t* __restrict__ seq_ptr,
                                    const int32_t* __restrict__ symbols, const int32_t* __restrict__ which,
                                    int k, uint32_t A, double eps, double* __restrict__ out) {
  extern __shared__ double factor[];
  const int o = blockIdx.x;
  const int32_t* seq = symbols + seq_ptr[o];
  const int len = (int)(seq_ptr[o + 1] - seq_ptr[o]);
  const int windows = len - k + 1;
  for (int j = threadIdx.x; j < windows; j += blockDim.x) {
    uint64_t idx = 0;
    for (int c = 0; c < k; ++c) idx = idx * A + (uint32_t)seq[j + c];
    if (j == 0) {
      factor[0] = y[idx];
    } else {
      const uint64_t base = idx - idx % A;
      double total = 0.0;
      for (uint32_t sym = 0; sym < A; ++sym) total += fmin(fmax(y[base + sym], eps), 1.0);
      factor[j] = fmin(fmax(y[idx], eps), 1.0) / total;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double current = factor[0];
    for (int j = 1; j < windows; ++j) current = factor[j] * current;
    out[which[o]] = current;
  }
}

// Entropy rate of the Markov chain the table describes (framework/markov_tapes.py:178-187): with
// c = clip(p, 1e-280, 1) and the context sums  m[x] = sum_s c[x, s]:  sum_x m[x] * sum_s -(c/m) log(c/m).
// One thread per context; partial[b] = the block's share, added up in block order afterwards.
__global__ void __launch_bounds__(kThreads) entropy_partials_kernel(const double* __restrict__ y, uint64_t n_contexts,
                                                                    uint32_t A, double* __restrict__ partial) {
  __shared__ double smem[32];
  double s = 0.0;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_contexts; x += (uint64_t)gridDim.x * blockDim.x) {
    const double* row = y + x * A;
    double m = 0.0;
    for (uint32_t sym = 0; sym < A; ++sym) m += fmin(fmax(row[sym], 1e-280), 1.0);
    double h = 0.0;
    for (uint32_t sym = 0; sym < A; ++sym) {
      const double cond = fmin(fmax(row[sym], 1e-280), 1.0) / m;
      h += -cond * log(cond);
    }
    s += h * m;
  }
  const double r = block_sum(s, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

}  // namespace

struct Dop853 {
  std::shared_ptr<Model> keep;  // the structure stays alive as long as the solver
  Model* m = nullptr;
  PeerGroup* peer = nullptr;
  uint64_t n = 0;
  cudaStream_t st = nullptr;
  Dop853Tableau tab;
  double t = 0, t_old = 0, t_bound = 0, direction = 1, h_abs = 0, h_previous = 0;
  double rtol = 0, atol = 0, max_step = 0;
  double* ybuf[3] = {nullptr, nullptr, nullptr};
  int iy = 0, iy_old = 1, iy_new = 2;
  double* K[16];
  double* F[7];
  bool have_F_mem = false, dense_ready = false, f_in_last = false, have_step = false;
  bool fused = true;  // stage updates fused into the product kernel
  double* stage = nullptr;
  double* partial = nullptr;
  double* d_sums = nullptr;
  int64_t nfev = 0, n_accepted = 0, n_rejected = 0;
  int status = 0;  // 0 running, 1 finished, -1 failed

  double* y() { return ybuf[iy]; }
  double* y_old() { return ybuf[iy_old]; }
  double* y_new() { return ybuf[iy_new]; }
};

namespace {

// Solver vectors come from the library's stream-ordered pool (engine.h library_pool), which keeps
// up to its release threshold of freed memory: creating a solver per integration of a small problem
// then costs no driver allocation, which takes anything from microseconds to milliseconds per call
// depending on the box.
double* dvec(uint64_t n, cudaStream_t st) {
  return (double*)pool_alloc(std::max<uint64_t>(n, 1) * sizeof(double), st);
}

void fun(Dop853* s, const double* y, double* out) {
  if (s->peer) {  // all ranks together; the sum arrives in this rank's peer-visible result vector
    peer_rhs(*s->peer, *s->m, y, s->st);
    TAPES_CUDA_CHECK(cudaMemcpyAsync(out, s->peer->result.ptr[s->peer->rank], s->n * sizeof(double),
                                     cudaMemcpyDeviceToDevice, s->st));
  } else {
    rhs_device(*s->m, y, out, s->st);
  }
  s->nfev++;
}

Terms terms_of(Dop853* s, const double* coef, int count) {
  Terms t;
  for (int j = 0; j < count; ++j)
    if (coef[j] != 0.0) { t.coef[t.n] = coef[j]; t.vec[t.n] = s->K[j]; ++t.n; }
  return t;
}

void read_sums(Dop853* s, int n_sums, double* out) {
  final_sum_kernel<<<1, 1024, 0, s->st>>>(s->partial, kReduceBlocks, n_sums, s->d_sums);
  TAPES_CUDA_CHECK(cudaMemcpyAsync(out, s->d_sums, n_sums * sizeof(double), cudaMemcpyDeviceToHost, s->st));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(s->st));
}

// np.linalg.norm(x / scale) / sqrt(n) of (a - b) against scale(y)
double rms_norm(Dop853* s, const double* a, const double* b, const double* y) {
  scaled_norm_partials_kernel<<<kReduceBlocks, kThreads, 0, s->st>>>(a, b, y, s->rtol, s->atol, s->n, s->partial);
  double sum = 0;
  read_sums(s, 1, &sum);
  return std::sqrt(sum) / std::sqrt((double)s->n);
}

// scipy/integrate/_ivp/common.py: select_initial_step (order = error_estimator_order = 7)
double select_initial_step(Dop853* s) {
  const double interval = std::fabs(s->t_bound - s->t);
  if (interval == 0.0) return 0.0;
  double* f0 = s->K[0];
  const double d0 = rms_norm(s, s->y(), nullptr, s->y());
  const double d1 = rms_norm(s, f0, nullptr, s->y());
  double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  h0 = std::min(h0, interval);
  Terms t;
  t.n = 1; t.coef[0] = 1.0; t.vec[0] = f0;
  lincomb_kernel<<<grid_for(s->n, kThreads), kThreads, 0, s->st>>>(s->stage, s->y(), t, h0 * s->direction, s->n);
  fun(s, s->stage, s->K[1]);
  const double d2 = rms_norm(s, s->K[1], f0, s->y()) / h0;
  double h1;
  if (d1 <= 1e-15 && d2 <= 1e-15) h1 = std::max(1e-6, h0 * 1e-3);
  else h1 = std::pow(0.01 / std::max(d1, d2), 1.0 / (7 + 1));
  return std::min(std::min(100 * h0, h1), std::min(interval, s->max_step));
}

// The stage update that follows stage `done`: coefficients of K[0..done-1] plus the coefficient of
// the stage being computed, in ascending stage order like np.dot(K[:s].T, a[:s]).
StageUpdate update_after(Dop853* s, const double* coef, int done, double* target, double h) {
  StageUpdate up;
  for (int j = 0; j < done; ++j)
    if (coef[j] != 0.0) { up.coef[up.n] = coef[j]; up.vec[up.n] = s->K[j]; ++up.n; }
  up.coef_self = coef[done];
  up.y = s->y();
  up.stage = target;
  up.h = h;
  return up;
}

// rk_step + error norm for one attempted step of size h; leaves y_new and K[0..12].
double attempt(Dop853* s, double h) {
  const unsigned grid = grid_for(s->n, kThreads);
  if (s->fused) {
    // stage 1 argument from K[0] alone; afterwards every product writes the next argument itself
    Terms t1 = terms_of(s, s->tab.A[1], 1);
    lincomb_kernel<<<grid, kThreads, 0, s->st>>>(s->stage, s->y(), t1, h, s->n);
    for (int st = 1; st < 12; ++st) {
      const StageUpdate up = st < 11 ? update_after(s, s->tab.A[st + 1], st, s->stage, h)
                                     : update_after(s, s->tab.B, 11, s->y_new(), h);
      rhs_device_fused(*s->m, s->stage, s->K[st], up, s->st);
      s->nfev++;
    }
  } else {
    for (int st = 1; st < 12; ++st) {
      Terms t = terms_of(s, s->tab.A[st], st);
      lincomb_kernel<<<grid, kThreads, 0, s->st>>>(s->stage, s->y(), t, h, s->n);
      fun(s, s->stage, s->K[st]);
    }
    Terms tb = terms_of(s, s->tab.B, 12);
    lincomb_kernel<<<grid, kThreads, 0, s->st>>>(s->y_new(), s->y(), tb, h, s->n);
  }
  fun(s, s->y_new(), s->K[12]);
  Terms e5 = terms_of(s, s->tab.E5, 13), e3 = terms_of(s, s->tab.E3, 13);
  error_partials_kernel<<<kReduceBlocks, kThreads, 0, s->st>>>(e5, e3, s->y(), s->y_new(), s->rtol, s->atol, s->n,
                                                              s->partial);
  TAPES_CUDA_CHECK(cudaGetLastError());
  double sums[2];
  read_sums(s, 2, sums);
  const double err5 = sums[0], err3 = sums[1];
  if (err5 == 0.0 && err3 == 0.0) return 0.0;
  const double denom = err5 + 0.01 * err3;
  return std::fabs(h) * err5 / std::sqrt(denom * (double)s->n);
}

}  // namespace

Dop853* dop853_create(std::shared_ptr<Model> model, PeerGroup* peer, const Dop853Tableau& tab, const double* h_y0,
                      double t0, double t_bound, double rtol, double atol, double max_step, double first_step) {
  if (!model) throw std::runtime_error("null model");
  Model& m = *model;
  Dop853* s = new Dop853();
  try {
    s->keep = model;
    s->m = &m; s->peer = peer; s->n = m.n_states; s->st = m.stream; s->tab = tab;
    s->t = t0; s->t_old = t0; s->t_bound = t_bound;
    s->direction = t_bound != t0 ? (t_bound > t0 ? 1.0 : -1.0) : 1.0;
    // validate_tol: rtol below 100 eps is raised to it
    const double eps = std::numeric_limits<double>::epsilon();
    s->rtol = std::max(rtol, 100 * eps); s->atol = atol;
    s->max_step = max_step > 0 ? max_step : std::numeric_limits<double>::infinity();
    if (const char* f = std::getenv("TAPES_RK_FUSED")) s->fused = std::atoi(f) != 0;
    if (peer) s->fused = false;  // the stage update cannot ride on a product whose result is still partial
    for (int i = 0; i < 16; ++i) s->K[i] = nullptr;
    for (int i = 0; i < 7; ++i) s->F[i] = nullptr;
    for (int i = 0; i < 3; ++i) s->ybuf[i] = dvec(s->n, s->st);
    for (int i = 0; i < 16; ++i) s->K[i] = dvec(s->n, s->st);
    s->stage = dvec(s->n, s->st);
    s->partial = dvec(2 * kReduceBlocks, s->st);
    s->d_sums = dvec(4, s->st);
    TAPES_CUDA_CHECK(cudaMemcpyAsync(s->y(), h_y0, s->n * sizeof(double), cudaMemcpyHostToDevice, s->st));
    fun(s, s->y(), s->K[0]);  // self.f = self.fun(self.t, self.y)
    s->h_abs = first_step > 0 ? first_step : select_initial_step(s);
    TAPES_CUDA_CHECK(cudaStreamSynchronize(s->st));
  } catch (...) {
    dop853_destroy(s);
    throw;
  }
  return s;
}

void dop853_destroy(Dop853* s) {
  if (!s) return;
  for (double* p : s->ybuf) if (p) cudaFreeAsync(p, s->st);
  for (double* p : s->K) if (p) cudaFreeAsync(p, s->st);
  for (double* p : s->F) if (p) cudaFreeAsync(p, s->st);
  if (s->stage) cudaFreeAsync(s->stage, s->st);
  if (s->partial) cudaFreeAsync(s->partial, s->st);
  if (s->d_sums) cudaFreeAsync(s->d_sums, s->st);
  if (s->st) cudaStreamSynchronize(s->st);
  delete s;
}

int dop853_step(Dop853* s) {  // OdeSolver.step + RungeKutta._step_impl
  if (s->status != 0) return s->status;
  if (s->n == 0 || s->t == s->t_bound) {
    s->t_old = s->t; s->t = s->t_bound; s->status = 1;
    return s->status;
  }
  if (s->f_in_last) {  // K[0] = f: the derivative at the new point was the last stage
    std::swap(s->K[0], s->K[12]);
    s->f_in_last = false;
  }
  const double t = s->t;
  const double inf = std::numeric_limits<double>::infinity();
  const double min_step = 10 * std::fabs(std::nextafter(t, s->direction * inf) - t);
  double h_abs = s->h_abs;
  if (h_abs > s->max_step) h_abs = s->max_step;
  else if (h_abs < min_step) h_abs = min_step;
  bool accepted = false, rejected = false;
  double h = 0, t_new = t;
  const double MAX_FACTOR = 10, MIN_FACTOR = 0.2, SAFETY = 0.9, exponent = -1.0 / 8.0;
  while (!accepted) {
    if (h_abs < min_step) { s->status = -1; return s->status; }
    h = h_abs * s->direction;
    t_new = t + h;
    if (s->direction * (t_new - s->t_bound) > 0) t_new = s->t_bound;
    h = t_new - t;
    h_abs = std::fabs(h);
    const double err = attempt(s, h);
    if (err < 1) {
      double factor = err == 0 ? MAX_FACTOR : std::min(MAX_FACTOR, SAFETY * std::pow(err, exponent));
      if (rejected) factor = std::min(1.0, factor);
      h_abs *= factor;
      accepted = true;
      s->n_accepted++;
    } else {
      h_abs *= std::max(MIN_FACTOR, SAFETY * std::pow(err, exponent));
      rejected = true;
      s->n_rejected++;
    }
  }
  s->h_previous = h;
  // y_old <- y, y <- y_new
  const int old = s->iy_old;
  s->iy_old = s->iy; s->iy = s->iy_new; s->iy_new = old;
  s->t_old = t; s->t = t_new; s->h_abs = h_abs;
  s->f_in_last = true;
  s->dense_ready = false;
  s->have_step = true;
  if (s->direction * (s->t - s->t_bound) >= 0) s->status = 1;
  return s->status;
}

void dop853_dense_eval(Dop853* s, double t, double* d_out) {  // DOP853._dense_output_impl + _call_impl
  if (!s->have_step) throw std::runtime_error("dense output needs a completed step");
  const unsigned grid = grid_for(s->n, kThreads);
  if (!s->dense_ready) {
    if (!s->have_F_mem) {
      for (int i = 0; i < 7; ++i) s->F[i] = dvec(s->n, s->st);
      s->have_F_mem = true;
    }
    const double h = s->h_previous;
    for (int st = 13; st < 16; ++st) {
      Terms tt = terms_of(s, s->tab.A[st], st);
      lincomb_kernel<<<grid, kThreads, 0, s->st>>>(s->stage, s->y_old(), tt, h, s->n);
      fun(s, s->stage, s->K[st]);
    }
    dense_head_kernel<<<grid, kThreads, 0, s->st>>>(s->F[0], s->F[1], s->F[2], s->y(), s->y_old(), s->K[0], s->K[12],
                                                   h, s->n);
    for (int r = 0; r < 4; ++r) {
      Terms td = terms_of(s, s->tab.D[r], 16);
      scaled_sum_kernel<<<grid, kThreads, 0, s->st>>>(s->F[3 + r], td, h, s->n);
    }
    s->dense_ready = true;
  }
  SevenVecs fv;
  for (int i = 0; i < 7; ++i) fv.F[i] = s->F[i];
  const double x = (t - s->t_old) / s->h_previous;
  dense_eval_kernel<<<grid, kThreads, 0, s->st>>>(d_out, fv, s->y_old(), x, s->n);
  TAPES_CUDA_CHECK(cudaGetLastError());
}

const double* dop853_state(const Dop853* s) { return s->ybuf[s->iy]; }
double* dop853_dense_buffer(Dop853* s) { return s->stage; }
Model& dop853_model(Dop853* s) { return *s->m; }

void dop853_info(const Dop853* s, double out[6]) {
  out[0] = s->t; out[1] = s->t_old; out[2] = s->h_abs;
  out[3] = (double)s->nfev; out[4] = (double)s->n_accepted; out[5] = (double)s->n_rejected;
}

void observe_strided(Model& m, const double* d_y, const int64_t* offset, const int64_t* stride,
                     const int64_t* count, int64_t n_obs, double* h_out) {
  if (n_obs <= 0) return;
  for (int64_t o = 0; o < n_obs; ++o)
    if (offset[o] < 0 || stride[o] < 1 || count[o] < 0 ||
        (count[o] > 0 && (uint64_t)(offset[o] + (count[o] - 1) * stride[o]) >= m.n_states))
      throw std::runtime_error("observable outside the state table");
  // the set of sums is uploaded when it changes, not per call
  std::vector<int64_t> spec((size_t)(3 * n_obs));
  for (int64_t o = 0; o < n_obs; ++o) {
    spec[(size_t)o] = offset[o]; spec[(size_t)(n_obs + o)] = stride[o]; spec[(size_t)(2 * n_obs + o)] = count[o];
  }
  if (spec != m.obs_spec) {
    TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
    if ((size_t)n_obs > m.obs_capacity) {
      if (m.d_obs_spec) cudaFree(m.d_obs_spec);
      if (m.d_obs_out) cudaFree(m.d_obs_out);
      m.d_obs_spec = nullptr; m.d_obs_out = nullptr; m.obs_capacity = 0; m.obs_spec.clear();
      TAPES_CUDA_CHECK(cudaMalloc((void**)&m.d_obs_spec, 3 * (size_t)n_obs * sizeof(int64_t)));
      TAPES_CUDA_CHECK(cudaMalloc((void**)&m.d_obs_out, (size_t)n_obs * sizeof(double)));
      m.obs_capacity = (size_t)n_obs;
    }
    TAPES_CUDA_CHECK(cudaMemcpyAsync(m.d_obs_spec, spec.data(), spec.size() * sizeof(int64_t), cudaMemcpyHostToDevice,
                                     m.stream));
    TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));  // `spec` is pageable and about to change hands
    m.obs_spec.swap(spec);
  }
  int64_t longest = 0;
  for (int64_t o = 0; o < n_obs; ++o) longest = std::max(longest, count[o]);
  // about 2048 elements per block, at most 8 blocks per SM's worth over all observables
  int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((longest + 2047) / 2048, std::max<int64_t>(1, kReduceBlocks / n_obs)));
  if ((size_t)(n_obs * blocks) > m.obs_partial_capacity) {
    TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
    if (m.d_obs_partial) cudaFree(m.d_obs_partial);
    m.d_obs_partial = nullptr; m.obs_partial_capacity = 0;
    TAPES_CUDA_CHECK(cudaMalloc((void**)&m.d_obs_partial, (size_t)(n_obs * blocks) * sizeof(double)));
    m.obs_partial_capacity = (size_t)(n_obs * blocks);
  }
  observe_partials_kernel<<<dim3((unsigned)blocks, (unsigned)n_obs), kThreads, 0, m.stream>>>(
      d_y, m.d_obs_spec, m.d_obs_spec + n_obs, m.d_obs_spec + 2 * n_obs, m.d_obs_partial);
  observe_final_kernel<<<(unsigned)n_obs, 256, 0, m.stream>>>(m.d_obs_partial, (int)blocks, m.d_obs_out);
  TAPES_CUDA_CHECK(cudaGetLastError());
  TAPES_CUDA_CHECK(cudaMemcpyAsync(h_out, m.d_obs_out, (size_t)n_obs * sizeof(double), cudaMemcpyDeviceToHost, m.stream));
  TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
}

void observe_sequences(Model& m, const double* d_y, int64_t n_seq, const int64_t* seq_ptr, const int32_t* symbols,
                       double eps, double* h_out) {
  if (n_seq <= 0) return;
  const int k = m.k;
  const uint64_t A = (uint64_t)m.A;
  std::vector<int64_t> offset, stride, count, where_short;
  std::vector<int64_t> long_ptr(1, 0);
  std::vector<int32_t> long_symbols, where_long;
  for (int64_t o = 0; o < n_seq; ++o) {
    const int64_t len = seq_ptr[o + 1] - seq_ptr[o];
    if (len < 0) throw std::runtime_error("sequence offsets must ascend");
    for (int64_t c = 0; c < len; ++c)
      if (symbols[seq_ptr[o] + c] < 0 || (uint64_t)symbols[seq_ptr[o] + c] >= A)
        throw std::runtime_error("sequence symbol outside the alphabet");
    if (len <= k) {  // framework/markov_tapes.py:215-222: the last len axes fixed, the leading ones summed
      uint64_t off = 0;
      for (int64_t c = 0; c < len; ++c) off = off * A + (uint64_t)symbols[seq_ptr[o] + c];
      offset.push_back((int64_t)off); stride.push_back((int64_t)m.pow_a[len]); count.push_back((int64_t)m.pow_a[k - len]);
      where_short.push_back(o);
    } else {
      if (len - k + 1 > 4096) throw std::runtime_error("sequence too long");
      long_symbols.insert(long_symbols.end(), symbols + seq_ptr[o], symbols + seq_ptr[o + 1]);
      long_ptr.push_back((int64_t)long_symbols.size());
      where_long.push_back((int32_t)where_long.size());
    }
  }
  if (!where_short.empty()) {
    std::vector<double> sums(where_short.size());
    observe_strided(m, d_y, offset.data(), stride.data(), count.data(), (int64_t)where_short.size(), sums.data());
    for (size_t i = 0; i < where_short.size(); ++i) h_out[where_short[i]] = sums[i];
  }
  if (!where_long.empty()) {
    const size_t n_long = where_long.size();
    int64_t* d_ptr = (int64_t*)pool_alloc((n_long + 1) * sizeof(int64_t), m.stream);
    int32_t* d_sym = (int32_t*)pool_alloc(long_symbols.size() * sizeof(int32_t), m.stream);
    int32_t* d_where = (int32_t*)pool_alloc(n_long * sizeof(int32_t), m.stream);
    double* d_out = (double*)pool_alloc(n_long * sizeof(double), m.stream);
    std::vector<double> got(n_long);
    size_t windows = 1;
    for (size_t i = 0; i < n_long; ++i) windows = std::max<size_t>(windows, (size_t)(long_ptr[i + 1] - long_ptr[i]) - k + 1);
    try {
      TAPES_CUDA_CHECK(cudaMemcpyAsync(d_ptr, long_ptr.data(), (n_long + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, m.stream));
      TAPES_CUDA_CHECK(cudaMemcpyAsync(d_sym, long_symbols.data(), long_symbols.size() * sizeof(int32_t), cudaMemcpyHostToDevice, m.stream));
      TAPES_CUDA_CHECK(cudaMemcpyAsync(d_where, where_long.data(), n_long * sizeof(int32_t), cudaMemcpyHostToDevice, m.stream));
      observe_long_kernel<<<(unsigned)n_long, 128, windows * sizeof(double), m.stream>>>(d_y, d_ptr, d_sym, d_where, k, (uint32_t)A, eps, d_out);
      TAPES_CUDA_CHECK(cudaGetLastError());
      TAPES_CUDA_CHECK(cudaMemcpyAsync(got.data(), d_out, n_long * sizeof(double), cudaMemcpyDeviceToHost, m.stream));
      TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
    } catch (...) {
      cudaFreeAsync(d_ptr, m.stream); cudaFreeAsync(d_sym, m.stream); cudaFreeAsync(d_where, m.stream); cudaFreeAsync(d_out, m.stream);
      throw;
    }
    cudaFreeAsync(d_ptr, m.stream); cudaFreeAsync(d_sym, m.stream); cudaFreeAsync(d_where, m.stream); cudaFreeAsync(d_out, m.stream);
    size_t i = 0;
    for (int64_t o = 0; o < n_seq; ++o)
      if (seq_ptr[o + 1] - seq_ptr[o] > k) h_out[o] = got[i++];
  }
}

double markov_entropy(Model& m, const double* d_y) {
  if (m.k < 1) throw std::runtime_error("entropy needs cl_k >= 1");
  const uint64_t contexts = m.pow_a[m.k - 1];
  const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((contexts + kThreads - 1) / kThreads, (uint64_t)kReduceBlocks));
  double* d_partial = (double*)pool_alloc((size_t)blocks * sizeof(double), m.stream);
  double* d_sum = (double*)pool_alloc(sizeof(double), m.stream);
  double h = 0.0;
  try {
    entropy_partials_kernel<<<blocks, kThreads, 0, m.stream>>>(d_y, contexts, (uint32_t)m.A, d_partial);
    final_sum_kernel<<<1, 1024, 0, m.stream>>>(d_partial, blocks, 1, d_sum);
    TAPES_CUDA_CHECK(cudaGetLastError());
    TAPES_CUDA_CHECK(cudaMemcpyAsync(&h, d_sum, sizeof(double), cudaMemcpyDeviceToHost, m.stream));
    TAPES_CUDA_CHECK(cudaStreamSynchronize(m.stream));
  } catch (...) {
    cudaFreeAsync(d_partial, m.stream); cudaFreeAsync(d_sum, m.stream);
    throw;
  }
  cudaFreeAsync(d_partial, m.stream); cudaFreeAsync(d_sum, m.stream);
  return h;
}

}  // namespace tapes
