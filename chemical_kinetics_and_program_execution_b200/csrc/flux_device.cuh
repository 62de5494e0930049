// Device side of the sliced flux structure shared by the product kernels (flux.cu) and the
// single-launch right-hand side of small problems (engine.cu).
#pragma once

#include <cstdint>

namespace tapes {
namespace {

constexpr uint32_t kNone = 0xffffffffu;  // node ids stay below 2^31 - 1, so this is never an entry
constexpr uint32_t kSignBit = 0x80000000u;

// w_child = w_parent * ratio, pruned when the ratio is not > 0 (tm.scm:1316, 1350, 1373).
__device__ __forceinline__ double weight_from_ratio(double w_parent, double r) { return r > 0.0 ? w_parent * r : 0.0; }

// The flux of the right children, evaluated per prefix group (engine.h Model::out_ptr, in_ptr).
// out_sum null: the model has no prefix groups.
struct RightFlux {
  const double* out_sum = nullptr;   // per prefix: sum of the sums of the groups with that prefix
  const double* ratio = nullptr;     // right-extension ratio of every window
  const double* totals = nullptr;    // sums of all prefix groups
  const uint64_t* in_ptr = nullptr;  // per adjusted prefix: the groups that flow in there ...
  const uint2* in_pairs = nullptr;   // ... as (group number, the group's own prefix)
  uint32_t A = 1;
};

// dy/dt of `row` through right children: + the groups whose adjusted prefix is the row's (each child
// weight = group sum * ratio of the child's ORIGINAL window, tm.scm:1310-1318, added at the adjusted
// window, 1290), in list order, then - the row's own outflow (1288).  U list entries are in flight.
// Table indices stay below A^k < 2^32, so the index arithmetic is 32-bit.
template <int U>
__device__ __forceinline__ double right_flux(const RightFlux& f, uint64_t row) {
  if (!f.out_sum) return 0.0;
  const uint32_t row32 = (uint32_t)row;
  const uint32_t q = row32 / f.A, x = row32 - q * f.A;
  double in = 0.0;
  const uint64_t lo = f.in_ptr[q];
  const uint32_t n = (uint32_t)(f.in_ptr[q + 1] - lo);
  const uint2* list = f.in_pairs + lo;
  for (uint32_t e = 0; e < n; e += U) {
    uint2 pair[U];
    double t[U], r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) pair[u] = e + u < n ? list[e + u] : make_uint2(0u, 0u);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      t[u] = e + u < n ? f.totals[pair[u].x] : 0.0;
      r[u] = e + u < n ? f.ratio[pair[u].y * f.A + x] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) in += weight_from_ratio(t[u], r[u]);
  }
  return in - weight_from_ratio(f.out_sum[q], f.ratio[row32]);
}

// The pointers carry no __restrict__ here: the kernels that call this say what may alias (the
// product kernels declare their arguments const __restrict__ and get the read-only path; the
// single-launch kernel reads weights written earlier in the same launch and must not).
// dy/dt for one slice per warp.  Every lane sums its state's terms in the order runs, then
// columns.  The kernel is bound by memory latency x bandwidth (ncu, profiles/r01_f_*: 67 % of the
// HBM peak with 4 gathers per lane and one dependent load phase per batch), so U gathers are in
// flight per lane and the column words of the next batch are loaded while the current one is
// gathered.
template <int U>
__device__ __forceinline__ double slice_sum(const uint64_t* slice_ptr,
                                            const uint32_t* slice_runs,
                                            const uint32_t* words, const double* w,
                                            uint64_t s, unsigned lane) {
  const uint64_t at = slice_ptr[s], stop = slice_ptr[s + 1];
  const uint32_t n_runs = slice_runs[s];
  const uint32_t n_pairs = (n_runs + 1u) & ~1u;
  const uint2* runs = (const uint2*)(words + at);
  const uint32_t* cols = words + at + 2ull * n_pairs + lane;
  const uint32_t n_cols = (uint32_t)((stop - at - 2ull * n_pairs) >> 5);
  const uint32_t below = (1u << lane) - 1u;
  double acc = 0.0;

  uint2 mine = lane < n_runs ? runs[lane] : make_uint2(0u, 0u);
  uint32_t v[U];  // column words of the batch about to be gathered
#pragma unroll
  for (int u = 0; u < U; ++u) v[u] = (uint32_t)u < n_cols ? cols[32u * u] : kNone;

  for (uint32_t j0 = 0; j0 < n_runs; j0 += 32) {
    const uint32_t here = min(32u, n_runs - j0);
    for (uint32_t jj = 0; jj < here; jj += U) {  // lanes past `here` hold the empty run
      uint32_t first[U], mask[U];
      double x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        first[u] = __shfl_sync(0xffffffffu, mine.x, (jj + u) & 31);
        mask[u] = __shfl_sync(0xffffffffu, mine.y, (jj + u) & 31);
        if (32 % U != 0 && jj + u >= 32) mask[u] = 0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        x[u] = ((mask[u] >> lane) & 1u) ? w[(first[u] & ~kSignBit) + __popc(mask[u] & below)] : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u) acc += (first[u] & kSignBit) ? -x[u] : x[u];
    }
    if (j0 + 32 < n_runs) mine = j0 + 32 + lane < n_runs ? runs[j0 + 32 + lane] : make_uint2(0u, 0u);
  }
  for (uint32_t c0 = 0; c0 < n_cols; c0 += U) {
    uint32_t ahead[U];
    double x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) ahead[u] = c0 + U + u < n_cols ? cols[32u * (c0 + U + u)] : kNone;
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = v[u] != kNone ? w[v[u] & ~kSignBit] : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) acc += (v[u] & kSignBit) ? -x[u] : x[u];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ahead[u];
  }
  return acc;
}


}  // namespace
}  // namespace tapes
