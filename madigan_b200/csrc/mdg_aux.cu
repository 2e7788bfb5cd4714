// Kernels around the step path (sm_100a): Portfolio's derived accounting, the
// StackerDiscrete window materialiser with its normalisers, and per-slab episode
// statistics.  All HBM-bound streaming work; the window kernel stages one tile of
// the observation ring through shared memory to turn [slot][feat][env] rows into
// contiguous (env, k, feat) windows with coalesced traffic on both sides.
#include <float.h>
#include <math.h>
#include <stdio.h>

#include <stdlib.h>

#include "mdg_common.cuh"

namespace mdg {

constexpr int kBlockA = 128;

// ---------------------------------------------------------------------------
// derived accounting  (Portfolio.cpp:140-235, 243-252)
// ---------------------------------------------------------------------------
struct DerivedArgs {
  MdgParams P;
  MdgState S;
  MdgDerived D;
  int64_t N;
};

__global__ void __launch_bounds__(kBlockA) derived_kernel(const __grid_constant__ DerivedArgs a) {
  const int64_t N = a.N;
  const int64_t e = (int64_t)blockIdx.x * kBlockA + threadIdx.x;
  if (e >= N) return;
  const int na = a.P.n_assets;
  const MdgState& S = a.S;
  const MdgDerived& D = a.D;
  const double cash = S.cash[e];
  double av, ml, bms, se, um, bav;
  {
    const double l = S.ledger[e], p = S.price[e], m = S.mean_entry[e];
    const double mask = l < 0. ? 1. : 0.;
    av = l * p; ml = m * l; bms = S.borrowed[e]; se = l * (m * mask);
    um = fabs(l) * m; bav = l * (p * mask);
  }
  for (int j = 1; j < na; ++j) {
    const double l = S.ledger[(int64_t)j * N + e], p = S.price[(int64_t)j * N + e],
                 m = S.mean_entry[(int64_t)j * N + e];
    const double mask = l < 0. ? 1. : 0.;
    av = av + l * p; ml = ml + m * l; bms = bms + S.borrowed[(int64_t)j * N + e];
    se = se + l * (m * mask); um = um + fabs(l) * m; bav = bav + l * (p * mask);
  }
  const double equity = cash + av - bms;  // :211-213
  const double pnl = av - ml;             // :184-186
  const double balance = cash + se;       // :192-197
  if (D.equity) D.equity[e] = equity;
  if (D.asset_value) D.asset_value[e] = av;
  if (D.pnl) D.pnl[e] = pnl;
  if (D.balance) D.balance[e] = balance;
  if (D.available_margin) D.available_margin[e] = (balance + pnl) / a.P.required_margin;  // :229-231
  if (D.used_margin) D.used_margin[e] = a.P.required_margin * um;                          // :199-201
  if (D.borrowed_margin) D.borrowed_margin[e] = bms;
  if (D.borrowed_asset_value) D.borrowed_asset_value[e] = bav;  // :219-223
  if (D.risk) D.risk[e] = margin_call(cash, av, ml, bms, se, a.P.maintenance_margin) ? MDG_RISK_MARGIN_CALL : MDG_RISK_GREEN;
  const double w0 = (cash - bms) / equity;
  // abs-normalisers: sum |.| left to right (:144-148, :164-168)
  double sabs = 0., sabs_full = fabs(w0);
  for (int j = 0; j < na; ++j) {
    const double l = S.ledger[(int64_t)j * N + e], p = S.price[(int64_t)j * N + e];
    const double w = (l * p) / equity;
    sabs = (j == 0) ? fabs(w) : sabs + fabs(w);
    sabs_full = sabs_full + fabs(w);
  }
  if (D.ledger_normed_full) D.ledger_normed_full[e] = w0;
  if (D.ledger_abs_normed_full) D.ledger_abs_normed_full[e] = w0 / sabs_full;
  if (D.position_values_full) D.position_values_full[e] = cash - bms;
  if (D.ledger_full) D.ledger_full[e] = cash - bms;
  for (int j = 0; j < na; ++j) {
    const double l = S.ledger[(int64_t)j * N + e], p = S.price[(int64_t)j * N + e],
                 m = S.mean_entry[(int64_t)j * N + e];
    const double pv = l * p;
    const double w = pv / equity;
    if (D.position_values) D.position_values[(int64_t)j * N + e] = pv;          // :170-172
    if (D.pnl_positions) D.pnl_positions[(int64_t)j * N + e] = pv - m * l;      // :188-190
    if (D.ledger_normed) D.ledger_normed[(int64_t)j * N + e] = w;               // :140-142
    if (D.ledger_abs_normed) D.ledger_abs_normed[(int64_t)j * N + e] = w / sabs;
    if (D.ledger_normed_full) D.ledger_normed_full[(int64_t)(j + 1) * N + e] = w;  // :150-155
    if (D.ledger_abs_normed_full) D.ledger_abs_normed_full[(int64_t)(j + 1) * N + e] = w / sabs_full;
    if (D.position_values_full) D.position_values_full[(int64_t)(j + 1) * N + e] = pv;  // :174-178
    if (D.ledger_full) D.ledger_full[(int64_t)(j + 1) * N + e] = l;                      // :157-161
  }
}

// ---------------------------------------------------------------------------
// window materialiser  (utils/preprocessor.py:183-189 current_data, :53-107 normalisers)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double nan_to_num(double x) {  // np.nan_to_num
  if (x != x) return 0.;
  if (isinf(x)) return x > 0 ? DBL_MAX : -DBL_MAX;
  return x;
}

// Sum over the time axis of one column, in numpy's order (the reference's mean/std are numpy reductions):
// a (k, F>1) array reduced over axis 0 accumulates row by row -> a left-to-right fold; a (k, 1) array is
// coalesced to 1-D and summed with numpy's pairwise scheme (8 interleaved accumulators for n <= 128).
// MODE 0: x   1: (x-c)^2   2: x with NaN -> 0 (nansum)   3: (x-c)^2 with NaN -> 0
template <int MODE>
__device__ __forceinline__ double col_term(const double* col, int s, int sstride, double c) {
  const double x = col[s * sstride];
  if (MODE == 0) return x;
  if (MODE == 2) return (x == x) ? x : 0.;
  const double d = x - c;
  if (MODE == 3 && !(x == x)) return 0.;
  return d * d;
}
template <int MODE>
__device__ __forceinline__ double col_sum(const double* col, int n, int sstride, double c, bool pairwise) {
  if (!pairwise || n < 8 || n > 128) {
    double s = col_term<MODE>(col, 0, sstride, c);
    for (int i = 1; i < n; ++i) s = s + col_term<MODE>(col, i, sstride, c);
    return s;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = col_term<MODE>(col, j, sstride, c);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = r[j] + col_term<MODE>(col, i + j, sstride, c);
  }
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res = res + col_term<MODE>(col, i, sstride, c);
  return res;
}

struct WindowArgs {
  const double* ring;        // [k][F][N]
  const double* prefix;      // [N][k][F] rows of the last reset's history fill, or nullptr
  const int64_t* timestamp;  // [N] (nullptr: every row comes from the ring)
  const int64_t* reset_ts;   // [N]
  void* out;
  int64_t N;
  int F, k, head, n_valid, norm, out_dtype, layout, flat_prefix;
  int xform, Feff, Fout;  // Feff: features the normaliser sees; Fout: features written
  int envs;     // envs per block (blockDim.x)
  int sstride;  // smem stride between rows s (doubles)
  int estride;  // smem stride between envs   (doubles)
};

// blockDim = (envs, F): thread (x=e_local, y=f) owns one (env, feature) column.
__global__ void window_kernel(const __grid_constant__ WindowArgs a) {
  extern __shared__ double tile[];  // [envs][n_valid][F] with padded strides
  const int el = threadIdx.x, f = threadIdx.y;
  const int64_t e0 = (int64_t)blockIdx.x * a.envs;
  const int64_t e = e0 + el;
  const int nv = a.n_valid, F = a.F, k = a.k;
  const bool live = e < a.N;
  double* col = tile + (int64_t)el * a.estride + f;
  int slot0 = (a.head - (nv - 1)) % k;
  if (slot0 < 0) slot0 += k;

  // phase 1: coalesced ring reads (consecutive envs), column into smem.  Rows older than the env's last
  // reset are not in the ring: they come from the env-major prefix buffer (prices) or are flat (portfolio).
  if (live) {
    long long since = 0x7fffffff;  // ring rows written since the last reset (the newest row always is one)
    if (a.timestamp) since = a.timestamp[e] - a.reset_ts[e] + 1;
    // 16 independent loads in flight per thread (a one-load-per-iteration loop made the kernel latency-bound:
    // 1.1 ms per 65,536 x 64 x 16 window, profiles/r1_notes.md)
    constexpr int U = 16;
    const int Z = blockDim.z, z = threadIdx.z;  // the column's rows are dealt round-robin to Z threads
    const int isince = since > nv ? nv : (int)since;             // rows with age < isince come from the ring
    const double* rbase = a.ring + (int64_t)f * a.N + e;         // + slot * rstride
    const int64_t rstride = (int64_t)F * a.N;
    // prefix row of age `age`: k - 1 - (age - (since - 1)) = (k - 2 + since) - age
    const double* pbase = a.prefix ? a.prefix + ((int64_t)e * k + (k - 2 + isince)) * F + f : nullptr;
    const double flat = (f == 0) ? 1. : 0.;  // flat portfolio: ledgerNormedFull == [1, 0, ..., 0]
    for (int j0 = 0; z + Z * j0 < nv; j0 += U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = z + Z * (j0 + u);
        v[u] = 0.;
        if (s < nv) {
          const int age = nv - 1 - s;
          int slot = slot0 + s;
          if (slot >= k) slot -= k;
          if (age < isince) v[u] = rbase[slot * rstride];
          else if (pbase) v[u] = *(pbase - (int64_t)age * F);
          else v[u] = flat;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = z + Z * (j0 + u);
        if (s < nv) col[s * a.sstride] = v[u];
      }
    }
  }
  __syncthreads();
  const int Fe = a.Feff;
  if (a.xform == MDG_XFORM_PAIR_RATIO) {  // price[:, 0] / price[:, 1]  (preprocessor.py:314-315)
    if (live && f == 0 && threadIdx.z == 0)
      for (int s = 0; s < nv; ++s) col[s * a.sstride] = col[s * a.sstride] / col[s * a.sstride + 1];
    __syncthreads();
  }

  // phase 2: normalise in place.  The element-wise parts (logs, the final scale/shift) are spread over the Z
  // threads of a column; only the column statistics (sums in numpy's order) are one thread per column.
  // Divisions by a per-column constant are multiplications by its reciprocal, logs are mdg_math's fast_log
  // (observations carry the 1e-9 bar).
  {
    const int Z = blockDim.z, z = threadIdx.z;
    const bool mine = live && f < Fe;
    double* cpar = tile + (int64_t)a.envs * a.estride + ((int64_t)el * F + f) * 2;  // (shift, scale) per column
    // 2a: element-wise pre-transform
    if (mine && (a.norm == MDG_NORM_LOG || a.norm == MDG_NORM_LOG_STANDARD)) {
      for (int s = z; s < nv; s += Z) {
        double x = col[s * a.sstride];
        if (a.norm == MDG_NORM_LOG) x = (x != x) ? x : (x < 1e-5 ? 1e-5 : x);  // log(max(x, 1e-5))
        col[s * a.sstride] = fast_log(x);
      }
    }
    if (a.norm == MDG_NORM_LOG_STANDARD) __syncthreads();
    // 2b: column statistics
    if (mine && z == 0) {
      double shift = 0., scale = 1.;
      const bool pw = (Fe == 1);
      switch (a.norm) {
        case MDG_NORM_LOOKBACK:      // x / x[-1]
        case MDG_NORM_LOOKBACK_LOG:  // log(x / x[-1])
          scale = 1. / col[(nv - 1) * a.sstride];
          break;
        case MDG_NORM_STANDARD: {  // nan_to_num((x - mean(0)) / std(0))
          shift = col_sum<0>(col, nv, a.sstride, 0., pw) / nv;
          scale = 1. / sqrt(col_sum<1>(col, nv, a.sstride, shift, pw) / nv);
          break;
        }
        case MDG_NORM_LOG_STANDARD: {  // x = log(x); nan_to_num((x - nanmean) / nanstd)
          int cnt = 0;
          for (int s = 0; s < nv; ++s) cnt += (col[s * a.sstride] == col[s * a.sstride]) ? 1 : 0;
          shift = col_sum<2>(col, nv, a.sstride, 0., pw) / cnt;
          scale = 1. / sqrt(col_sum<3>(col, nv, a.sstride, shift, pw) / cnt);
          break;
        }
        default:
          break;
      }
      cpar[0] = shift; cpar[1] = scale;
    }
    const bool scaled = a.norm == MDG_NORM_LOOKBACK || a.norm == MDG_NORM_LOOKBACK_LOG ||
                        a.norm == MDG_NORM_STANDARD || a.norm == MDG_NORM_LOG_STANDARD;
    if (scaled) {
      __syncthreads();
      // 2c: element-wise finish
      if (mine) {
        const double shift = cpar[0], scale = cpar[1];
        const bool zs = a.norm == MDG_NORM_STANDARD || a.norm == MDG_NORM_LOG_STANDARD;
        for (int s = z; s < nv; s += Z) {
          double x = col[s * a.sstride];
          if (zs) {
            // (x - mean) / sd with sd == 0: 0/0 -> nan -> 0, c/0 -> +-inf -> +-DBL_MAX, as np.nan_to_num
            const double d = x - shift;
            x = nan_to_num(d * scale);
          } else {
            x = x * scale;
            if (a.norm == MDG_NORM_LOOKBACK_LOG) x = fast_log(x);
          }
          col[s * a.sstride] = x;
        }
      }
    }
  }
  if (a.norm == MDG_NORM_EXPANDING) {
    // x / _expanding_mean(x): the gufunc runs over the LAST axis (features), preprocessor.py:72-73,472-491
    const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    for (int r = tid; r < a.envs * nv; r += nthr) {
      const int re = r / nv, rs = r - re * nv;
      if (e0 + re >= a.N) continue;
      double* row = tile + (int64_t)re * a.estride + (int64_t)rs * a.sstride;
      double acc = (row[0] == row[0]) ? row[0] : 0.;
      int w = 1;
      double prev = row[0];
      row[0] = prev / acc;
      for (int i = 1; i < Fe; ++i) {
        const double x = row[i];
        if (x == x) { acc = acc + x; w += 1; }
        row[i] = x / (acc / w);
      }
    }
  }
  __syncthreads();

  // phase 3: contiguous write-out of the block's windows
  const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const int Fo = a.Fout;
  const int per_env = nv * Fo;
  const int rows = (int)((a.N - e0) < a.envs ? (a.N - e0) : a.envs);
  const bool f32 = a.out_dtype == MDG_DTYPE_F32, diff = a.xform == MDG_XFORM_RETURNS;
  if (a.layout == MDG_LAYOUT_NKF && nthr % Fo == 0) {
    // thread -> fixed feature, rows advance by nthr / Fo: consecutive threads write consecutive elements and no
    // index needs a division (a 64-bit div/mod per element made this phase instruction-bound)
    const int ff = tid % Fo, sstep = nthr / Fo;
    for (int re = 0; re < rows; ++re) {
      const double* tp = tile + (int64_t)re * a.estride + ff;
      const int64_t o0 = (e0 + re) * per_env + ff;
      for (int sr = tid / Fo; sr < nv; sr += sstep) {
        const double* q = tp + sr * a.sstride;
        const double v = diff ? q[1] - q[0] : q[0];  // np.diff(price): last axis (:330)
        if (f32) ((float*)a.out)[o0 + sr * Fo] = (float)v; else ((double*)a.out)[o0 + sr * Fo] = v;
      }
    }
  } else {
    for (int re = 0; re < rows; ++re) {
      for (int r = tid; r < per_env; r += nthr) {
        int sr, ff;
        if (a.layout == MDG_LAYOUT_NKF) { sr = r / Fo; ff = r - sr * Fo; } else { ff = r / nv; sr = r - ff * nv; }
        const double* q = tile + (int64_t)re * a.estride + (int64_t)sr * a.sstride + ff;
        const double v = diff ? q[1] - q[0] : q[0];
        const int64_t o = (e0 + re) * per_env + r;
        if (f32) ((float*)a.out)[o] = (float)v; else ((double*)a.out)[o] = v;
      }
    }
  }
}

__global__ void time_kernel(const int64_t* timestamp, int64_t N, int nv, int64_t* out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * nv) return;
  const int64_t e = idx / nv;
  const int s = (int)(idx - e * nv);
  out[idx] = timestamp[e] - (nv - 1 - s);  // every window row is exactly one generator tick
}

// ---------------------------------------------------------------------------
// episode statistics: one vector per slab, all-reduced across GPUs by the host (NCCL)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (!(v < __longlong_as_double((long long)assumed))) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (!(v > __longlong_as_double((long long)assumed))) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct StatsArgs {
  MdgParams P;
  MdgState S;
  MdgStepIO IO;
  int64_t N;
  double* out;
};

__global__ void stats_init_kernel(double* out, int n) {
  const int i = threadIdx.x;
  if (i < n) out[i] = (i == 3) ? INFINITY : (i == 4 ? -INFINITY : 0.);
}

// grid-stride, per-thread partials -> warp shuffles -> one atomic per warp per statistic
__global__ void __launch_bounds__(256) stats_kernel(const __grid_constant__ StatsArgs a) {
  const int64_t N = a.N;
  const int na = a.P.n_assets;
  double cnt = 0., seq = 0., ssq = 0., mn = INFINITY, mx = -INFINITY, srew = 0., scost = 0., ndone = 0.;
  double expo[MDG_MAX_ASSETS], held[MDG_MAX_ASSETS];
#pragma unroll
  for (int j = 0; j < MDG_MAX_ASSETS; ++j) { expo[j] = 0.; held[j] = 0.; }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (int64_t)gridDim.x * blockDim.x) {
    double av = 0., bms = 0.;
    for (int j = 0; j < na; ++j) {
      av = av + a.S.ledger[(int64_t)j * N + e] * a.S.price[(int64_t)j * N + e];
      bms = bms + a.S.borrowed[(int64_t)j * N + e];
    }
    const double eq = a.S.cash[e] + av - bms;
    cnt += 1.; seq += eq; ssq += eq * eq; mn = fmin(mn, eq); mx = fmax(mx, eq);
    if (a.IO.reward) srew += a.IO.reward[e];
    if (a.IO.done) ndone += a.IO.done[e] ? 1. : 0.;
#pragma unroll
    for (int j = 0; j < MDG_MAX_ASSETS; ++j) {
      if (j < na) {
        const double l = a.S.ledger[(int64_t)j * N + e];
        if (a.IO.trans_cost) scost += a.IO.trans_cost[(int64_t)j * N + e];
        expo[j] += fabs(l * a.S.price[(int64_t)j * N + e]) / eq;
        held[j] += (l != 0.) ? 1. : 0.;
      }
    }
  }
  const bool lead = (threadIdx.x & 31) == 0;
  cnt = warp_sum(cnt); seq = warp_sum(seq); ssq = warp_sum(ssq); srew = warp_sum(srew);
  scost = warp_sum(scost); ndone = warp_sum(ndone); mn = warp_min(mn); mx = warp_max(mx);
  if (lead) {
    atomicAdd(&a.out[0], cnt); atomicAdd(&a.out[1], seq); atomicAdd(&a.out[2], ssq);
    atomic_min_double(&a.out[3], mn); atomic_max_double(&a.out[4], mx);
    atomicAdd(&a.out[5], srew); atomicAdd(&a.out[6], scost); atomicAdd(&a.out[7], ndone);
  }
#pragma unroll
  for (int j = 0; j < MDG_MAX_ASSETS; ++j) {
    if (j < na) {
      const double x = warp_sum(expo[j]), h = warp_sum(held[j]);
      if (lead) {
        atomicAdd(&a.out[MDG_STATS_NSCALAR + j], x);
        atomicAdd(&a.out[MDG_STATS_NSCALAR + na + j], h);
      }
    }
  }
}

}  // namespace mdg

using namespace mdg;

extern "C" int mdg_derived(const MdgParams* P, const MdgState* S, const MdgDerived* D, const MdgLaunch* L) {
  if (!P || !S || !D || !L) return set_err(MDG_E_INVALID, "null argument");
  if (P->n_assets < 1 || P->n_assets > MDG_MAX_ASSETS) return set_err(MDG_E_UNSUPPORTED, "n_assets out of range");
  if (L->n_envs <= 0) return L->n_envs == 0 ? MDG_OK : set_err(MDG_E_INVALID, "n_envs < 0");
  DerivedArgs a;
  a.P = *P; a.S = *S; a.D = *D; a.N = L->n_envs;
  const unsigned grid = (unsigned)((a.N + kBlockA - 1) / kBlockA);
  derived_kernel<<<grid, kBlockA, 0, (cudaStream_t)L->stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_derived launch");
}

extern "C" int mdg_materialise_window(const MdgWindow* w) {
  if (!w || !w->ring || !w->out) return set_err(MDG_E_INVALID, "null window/ring/out");
  const int n_feats = w->n_feats, window = w->window, head = w->head, n_valid = w->n_valid;
  if (n_feats < 1 || n_feats > 64) return set_err(MDG_E_UNSUPPORTED, "n_feats must be in 1..64");
  if (window < 1 || head < 0 || head >= window || n_valid < 1 || n_valid > window)
    return set_err(MDG_E_INVALID, "bad window/head/n_valid");
  if (w->norm_type < MDG_NORM_NONE || w->norm_type > MDG_NORM_EXPANDING) return set_err(MDG_E_INVALID, "bad norm_type");
  if (w->out_dtype != MDG_DTYPE_F64 && w->out_dtype != MDG_DTYPE_F32) return set_err(MDG_E_INVALID, "bad out_dtype");
  if (w->out_layout != MDG_LAYOUT_NKF && w->out_layout != MDG_LAYOUT_NFK) return set_err(MDG_E_INVALID, "bad layout");
  if (w->timestamp && !w->reset_ts) return set_err(MDG_E_INVALID, "timestamp given without reset_ts");
  if (w->timestamp && !w->prefix && !w->flat_prefix) return set_err(MDG_E_INVALID, "no prefix source for rows older than the reset");
  if (w->transform < MDG_XFORM_NONE || w->transform > MDG_XFORM_RETURNS) return set_err(MDG_E_INVALID, "bad transform");
  if (w->transform == MDG_XFORM_PAIR_RATIO && n_feats != 2)
    return set_err(MDG_E_INVALID, "the pair-ratio window needs exactly 2 features");  // preprocessor.py:301 assert
  if (w->n_envs <= 0) return w->n_envs == 0 ? MDG_OK : set_err(MDG_E_INVALID, "n_envs < 0");
  if (w->transform == MDG_XFORM_RETURNS && n_feats == 1) return MDG_OK;  // np.diff of one feature: (k, 0), nothing to write
  WindowArgs a;
  a.xform = w->transform;
  a.Feff = (w->transform == MDG_XFORM_PAIR_RATIO) ? 1 : n_feats;
  a.Fout = (w->transform == MDG_XFORM_PAIR_RATIO) ? 1 : (w->transform == MDG_XFORM_RETURNS ? n_feats - 1 : n_feats);
  a.ring = w->ring; a.prefix = w->prefix; a.timestamp = w->timestamp; a.reset_ts = w->reset_ts;
  a.out = w->out; a.N = w->n_envs; a.F = n_feats; a.k = window; a.head = head; a.n_valid = n_valid;
  a.norm = w->norm_type; a.out_dtype = w->out_dtype; a.layout = w->out_layout; a.flat_prefix = w->flat_prefix;
  // envs per block: a power of two with about 64 columns per block and two threads per column -- small tiles
  // (35 KB at 64 x 16), six blocks per SM: measured best at 65,536 x 64 x 16 (profiles/window_cost.py)
  int envs = 32;
  while (envs > 1 && envs * n_feats > 64) envs >>= 1;
  static const int env_override = [] { const char* v = getenv("MDG_WIN_ENVS"); return v ? atoi(v) : 0; }();
  static const int z_override = [] { const char* v = getenv("MDG_WIN_Z"); return v ? atoi(v) : 0; }();
  if (env_override) envs = env_override;
  a.sstride = n_feats + 1;  // odd-ish row stride: conflict-free (N,F,k) reads
  for (;;) {
    int es = n_valid * a.sstride;
    es += ((2 - es) % 16 + 16) % 16;  // env stride == 2 (mod 16 doubles): conflict-free column writes
    a.estride = es;
    if (((size_t)envs * es + (size_t)envs * n_feats * 2) * sizeof(double) <= 96 * 1024 || envs == 1) break;
    envs >>= 1;
  }
  a.envs = envs;
  const size_t smem = ((size_t)envs * a.estride + (size_t)envs * n_feats * 2) * sizeof(double);
  if (smem > 200 * 1024) return set_err(MDG_E_UNSUPPORTED, "window tile does not fit shared memory");
  static size_t smem_allowed = 48 * 1024;  // raise the opt-in limit only when a larger tile is first needed
  if (smem > smem_allowed) {
    cudaError_t ce = cudaFuncSetAttribute(window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return cuda_err(ce, "window smem attr");
    smem_allowed = smem;
  }
  const unsigned grid = (unsigned)((w->n_envs + envs - 1) / envs);
  int zdim = 2;  // threads per column for the data-movement and element-wise phases
  if (z_override) zdim = z_override;
  window_kernel<<<grid, dim3(envs, n_feats, zdim), smem, (cudaStream_t)w->stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_materialise_window launch");
}

extern "C" int mdg_materialise_time(const int64_t* timestamp, int64_t n_envs, int32_t n_valid, int64_t* out,
                                    void* stream) {
  if (!timestamp || !out) return set_err(MDG_E_INVALID, "null timestamp/out");
  if (n_valid < 1) return set_err(MDG_E_INVALID, "bad n_valid");
  if (n_envs <= 0) return n_envs == 0 ? MDG_OK : set_err(MDG_E_INVALID, "n_envs < 0");
  const int64_t total = n_envs * n_valid;
  time_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(timestamp, n_envs, n_valid, out);
  return cuda_err(cudaGetLastError(), "mdg_materialise_time launch");
}

extern "C" int mdg_episode_stats(const MdgParams* P, const MdgState* S, const MdgStepIO* IO, const MdgLaunch* L,
                                 double* out) {
  if (!P || !S || !L || !out) return set_err(MDG_E_INVALID, "null argument");
  if (P->n_assets < 1 || P->n_assets > MDG_MAX_ASSETS) return set_err(MDG_E_UNSUPPORTED, "n_assets out of range");
  StatsArgs a;
  a.P = *P; a.S = *S;
  if (IO) a.IO = *IO; else memset(&a.IO, 0, sizeof(a.IO));
  a.N = L->n_envs; a.out = out;
  cudaStream_t st = (cudaStream_t)L->stream;
  stats_init_kernel<<<1, 64, 0, st>>>(out, MDG_STATS_NSCALAR + 2 * P->n_assets);
  if (a.N > 0) {
    int grid = (int)((a.N + 255) / 256);
    if (grid > 148 * 4) grid = 148 * 4;  // a few resident CTAs per SM, grid-stride over the slab
    stats_kernel<<<grid, 256, 0, st>>>(a);
  }
  return cuda_err(cudaGetLastError(), "mdg_episode_stats launch");
}
