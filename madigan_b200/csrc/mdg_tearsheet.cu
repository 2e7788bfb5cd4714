// On-device episode tearsheet (reference: madigan/utils/metrics.py:83-171 test_summary, helpers :334-421).
// One thread per env; [rows][N] accumulators, coalesced.  See include/madigan_b200.h (MdgTearsheet) for the contract.
#include <math.h>
#include <stdio.h>

#include "mdg_common.cuh"

namespace mdg {

struct TsResetArgs {
  MdgTearsheet ts;
  const uint8_t* mask;
};

__global__ void __launch_bounds__(256) tearsheet_reset_kernel(const __grid_constant__ TsResetArgs a) {
  const MdgTearsheet& t = a.ts;
  const int64_t N = t.n_envs, e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= N || (a.mask && !a.mask[e])) return;
  t.nsteps[e] = 0;
  t.active[e] = 1;
  t.sum_equity[e] = 0.; t.last_equity[e] = 0.; t.sum_reward[e] = 0.; t.sum_cost[e] = 0.;
  t.peak[e] = -INFINITY;
  t.min_valley[e] = INFINITY;
  for (int j = 0; j < t.n_assets; ++j) t.in_pos[(int64_t)j * N + e] = 0.;
  for (int j = 0; j < t.n_offsets; ++j) {
    t.ret_n[(int64_t)j * N + e] = 0;
    t.ret_mean[(int64_t)j * N + e] = 0.;
    t.ret_m2[(int64_t)j * N + e] = 0.;
    t.ret_down[(int64_t)j * N + e] = 0.;
  }
}

struct TsUpdateArgs {
  MdgTearsheet ts;
  MdgState S;
  MdgStepIO IO;
  int na;
};

__global__ void __launch_bounds__(256) tearsheet_update_kernel(const __grid_constant__ TsUpdateArgs a) {
  const MdgTearsheet& t = a.ts;
  const int64_t N = t.n_envs, e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= N || !t.active[e]) return;
  const int na = a.na, J = t.n_offsets;
  // equity of the portfolio as the step left it: cash + assetValue - borrowedMargin from the folds the step kernel
  // keeps in state (Portfolio.cpp:211-213; bit-identical to the derived-accounting kernel's equity)
  const double eq = a.S.cash[e] + a.S.folds[0 * N + e] - a.S.folds[2 * N + e];
  const int64_t n = t.nsteps[e];  // steps recorded before this one
  t.sum_equity[e] += eq;
  t.last_equity[e] = eq;
  t.sum_reward[e] += a.IO.reward[e];
  double cost = 0.;
  for (int j = 0; j < na; ++j) {
    cost += a.IO.trans_cost[(int64_t)j * N + e];
    if (a.S.ledger[(int64_t)j * N + e] != 0.) t.in_pos[(int64_t)j * N + e] += 1.;  // metrics.py:119
  }
  t.sum_cost[e] += cost;
  const double pk = fmax(t.peak[e], eq);  // expanding max incl. the current point (metrics.py:419)
  t.peak[e] = pk;
  t.min_valley[e] = fmin(t.min_valley[e], eq / pk);
  // log returns at offsets 2^j (metrics.py:367-411 with unit timestamps: out[r] = log(arr[r] / arr[r - tf]), r >= tf)
  const int R = J > 0 ? (1 << (J - 1)) : 0;
  for (int j = 0; j < J; ++j) {
    const int64_t tf = (int64_t)1 << j;
    if (n >= tf) {
      const double prev = t.eq_ring[(int64_t)((n - tf) & (R - 1)) * N + e];
      const double r = log(eq / prev);
      if (r == r) {  // nanmean / nanstd skip NaN
        const int64_t o = (int64_t)j * N + e;
        const int c = t.ret_n[o] + 1;
        const double d = r - t.ret_mean[o];
        const double m = t.ret_mean[o] + d / c;
        t.ret_n[o] = c;
        t.ret_mean[o] = m;
        t.ret_m2[o] += d * (r - m);
        if (r < 0.) t.ret_down[o] += r * r;
      }
    }
  }
  if (R > 0) t.eq_ring[(int64_t)(n & (R - 1)) * N + e] = eq;
  t.nsteps[e] = n + 1;
  if (a.IO.done && a.IO.done[e]) t.active[e] = 0;  // the step that reports done is the last one of the episode
}

struct TsSummaryArgs {
  MdgTearsheet ts;
  double* out;
};

__global__ void __launch_bounds__(256) tearsheet_summary_kernel(const __grid_constant__ TsSummaryArgs a) {
  const MdgTearsheet& t = a.ts;
  const int64_t N = t.n_envs, e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= N) return;
  const int na = t.n_assets, J = t.n_offsets;
  const int64_t n = t.nsteps[e];
  const double dn = (double)n;
  double* o = a.out + e;
  o[0 * N] = dn;
  o[1 * N] = t.sum_equity[e] / dn;
  o[2 * N] = n ? t.last_equity[e] : NAN;
  o[3 * N] = t.sum_reward[e] / dn;
  o[4 * N] = n ? t.min_valley[e] : NAN;
  o[5 * N] = t.sum_cost[e] / (dn * na);
  o[6 * N] = t.sum_cost[e];
  const int64_t max_tf = n ? (n - 1) / 10 : -1;  // (timestamps[-1] - timestamps[0]) // 10, metrics.py:108
  for (int j = 0; j < J; ++j) {
    const int64_t q = (int64_t)j * N + e;
    double mean = NAN, sharpe = NAN, sortino = NAN;
    if (((int64_t)1 << j) <= max_tf) {
      const int c = t.ret_n[q];
      mean = c ? t.ret_mean[q] : NAN;
      const double sd = sqrt(t.ret_m2[q] / (c - 1));  // np.nanstd(ddof=1)
      sharpe = (sd == 0.) ? 0. : mean / sd;           // metrics.py:345-348
      const double downside = t.ret_down[q] / (dn - 1);  // len(diff) counts the NaN head too, :359
      if (downside == 0.) sortino = (mean > 0.) ? 50. : 0.;  // :361-365
      else sortino = fmin(mean / downside, 50.);
    }
    o[(MDG_TS_NFIXED + 3 * j) * N] = mean;
    o[(MDG_TS_NFIXED + 3 * j + 1) * N] = sharpe;
    o[(MDG_TS_NFIXED + 3 * j + 2) * N] = sortino;
  }
  for (int i = 0; i < na; ++i) o[(MDG_TS_NFIXED + 3 * J + i) * N] = t.in_pos[(int64_t)i * N + e] / dn;
}

static int check_ts(const MdgTearsheet* t) {
  if (!t) return set_err(MDG_E_INVALID, "null tearsheet");
  if (t->n_assets < 1 || t->n_assets > MDG_MAX_ASSETS) return set_err(MDG_E_INVALID, "tearsheet: bad n_assets");
  if (t->n_offsets < 0 || t->n_offsets > MDG_TS_MAX_OFFSETS) return set_err(MDG_E_INVALID, "tearsheet: bad n_offsets");
  if (!t->nsteps || !t->active || !t->sum_equity || !t->last_equity || !t->sum_reward || !t->peak || !t->min_valley ||
      !t->sum_cost || !t->in_pos)
    return set_err(MDG_E_INVALID, "tearsheet: null accumulator");
  if (t->n_offsets > 0 && (!t->eq_ring || !t->ret_n || !t->ret_mean || !t->ret_m2 || !t->ret_down))
    return set_err(MDG_E_INVALID, "tearsheet: null return accumulator");
  return MDG_OK;
}

}  // namespace mdg

using namespace mdg;

extern "C" int mdg_tearsheet_reset(const MdgTearsheet* ts, const uint8_t* mask, void* stream) {
  int rc = check_ts(ts);
  if (rc) return rc;
  if (ts->n_envs <= 0) return ts->n_envs == 0 ? MDG_OK : set_err(MDG_E_INVALID, "n_envs < 0");
  TsResetArgs a;
  a.ts = *ts; a.mask = mask;
  tearsheet_reset_kernel<<<(unsigned)((ts->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_tearsheet_reset launch");
}

extern "C" int mdg_tearsheet_update(const MdgParams* P, const MdgState* S, const MdgStepIO* IO, const MdgTearsheet* ts,
                                    void* stream) {
  int rc = check_ts(ts);
  if (rc) return rc;
  if (!P || !S || !IO) return set_err(MDG_E_INVALID, "null params/state/io");
  if (!S->cash || !S->folds || !S->ledger || !IO->reward || !IO->trans_cost)
    return set_err(MDG_E_INVALID, "tearsheet update: null state/io pointer");
  if (P->n_assets != ts->n_assets) return set_err(MDG_E_INVALID, "tearsheet: n_assets mismatch");
  if (ts->n_envs <= 0) return ts->n_envs == 0 ? MDG_OK : set_err(MDG_E_INVALID, "n_envs < 0");
  TsUpdateArgs a;
  a.ts = *ts; a.S = *S; a.IO = *IO; a.na = P->n_assets;
  tearsheet_update_kernel<<<(unsigned)((ts->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_tearsheet_update launch");
}

extern "C" int mdg_tearsheet_summary(const MdgTearsheet* ts, double* out, void* stream) {
  int rc = check_ts(ts);
  if (rc) return rc;
  if (!out) return set_err(MDG_E_INVALID, "null out");
  if (ts->n_envs <= 0) return ts->n_envs == 0 ? MDG_OK : set_err(MDG_E_INVALID, "n_envs < 0");
  TsSummaryArgs a;
  a.ts = *ts; a.out = out;
  tearsheet_summary_kernel<<<(unsigned)((ts->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_tearsheet_summary launch");
}
