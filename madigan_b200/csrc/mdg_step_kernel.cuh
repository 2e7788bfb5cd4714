// Fused Env.step / Env.reset kernels for N independent environments (sm_100a).
//
// One thread per env; the whole portfolio (price, ledger, meanEntry, borrowedMargin
// for up to 16 assets) lives in registers, state tensors are [rows][N] so every
// load/store is one coalesced 256-byte run per warp.  One launch does what the
// reference does across Env.h:189-256, Broker.cpp:124-178, Portfolio.cpp:140-323,
// the DataSource.cpp getData family, offpolicy_q.py:140-164 and nstep_buffer.py:
//   transact (sequential over assets, risk-gated) -> generator tick -> equity,
//   reward, done -> newest observation-ring row -> agent reward -> shaped reward.
// HBM-bound integer/fp64 work: no tensor cores.
#pragma once
#include <math.h>
#include <stdio.h>

#include "mdg_common.cuh"

namespace mdg {

struct StepArgs {
  MdgParams P;
  MdgReward R;
  MdgState S;
  MdgStepIO IO;
  MdgLaunch L;
};

constexpr int kBlock = 128;

// ---------------------------------------------------------------------------
// reward shapers (utils/buffers/nstep_buffer.py), one scalar component
// ---------------------------------------------------------------------------
__device__ __forceinline__ double clip(double x, double lo, double hi) {
  if (x != x) return x;
  return x < lo ? lo : (x > hi ? hi : x);
}
constexpr double kEps32 = 1.1920928955078125e-07;  // np.finfo(np.float32).eps, nstep_buffer.py:20

__device__ __forceinline__ double dsr_value(double A, double B, double r) {  // :80-85
  const double dA = r - A, dB = r * r - B;
  const double v = B - A * A;
  return (B * dA - (A * dB) / 2) / (pow(v * v, 0.75) + kEps32);
}
__device__ __forceinline__ double ddr_value(double A, double B, double r) {  // :146-156
  if (r > 0.) return (r - A / 2) / (sqrt(B) + kEps32);
  return (B * (r - A / 2) - (A * (r * r)) / 2) / (pow(B, 1.5) + kEps32);
}

// entry j (0 = oldest) of this env's n-step buffer, component c
struct NStepView {
  const double* ring;  // [nstep][ra][N]
  int64_t N, e;
  int n, ra, c, base;  // base = physical slot of entry 0
  double newest;       // the entry added this step (also stored in the ring when n>1)
  int len;             // entries including the newest
  __device__ __forceinline__ double at(int j) const {
    if (j == len - 1) return newest;
    int slot = base + j;
    if (slot >= n) slot -= n;
    return ring[((int64_t)slot * ra + c) * N + e];
  }
};

// shaped reward of one pop over entries [first, len) ; updates A,B for DSR/DDR
static __device__ __noinline__ double shaper_pop(const MdgReward& R, const NStepView& v, int first, double& A,
                                          double& B) {
  const int n = v.len - first;
  const double* disc = R.discounts;
  switch (R.shaper) {
    case MDG_SHAPER_SUM:
    case MDG_SHAPER_COSINE: {  // :23-27, :182-204
      double s = 0.;
      for (int j = 0; j < n; ++j) s = s + disc[j] * v.at(first + j);
      return s;
    }
    case MDG_SHAPER_DSR: {  // :62-78
      double s = disc[0] * dsr_value(A, B, v.at(first));
      for (int j = 1; j < n; ++j) s = s + disc[j] * dsr_value(A, B, v.at(first + j));
      s = s / n;
      const double r0 = v.at(first), dA = r0 - A, dB = r0 * r0 - B;  // :87-91
      A += R.adaptation_rate * dA;
      B += R.adaptation_rate * dB;
      return clip(s, -1., 1.);
    }
    case MDG_SHAPER_DDR: {  // :128-162
      double s = disc[0] * ddr_value(A, B, v.at(first));
      for (int j = 1; j < n; ++j) s = s + disc[j] * ddr_value(A, B, v.at(first + j));
      s = s / n;
      const double r0 = v.at(first), dA = r0 - A;
      double m = r0 < 0. ? r0 : 0.;
      if (r0 != r0) m = r0;
      const double dB = m * m - B;
      A += R.adaptation_rate * dA;
      B += R.adaptation_rate * dB;
      return clip(s, -1., 1.);
    }
    case MDG_SHAPER_SHARPE: {  // :207-239
      if (n == 1) {
        double diff = v.at(first) - 0.;
        diff = (diff != 0.) ? diff : 0.;
        return diff / sqrt(diff * diff);
      }
      double sum = 0., ssq = 0.;
      for (int j = 0; j < n; ++j) {
        const double dj = (v.at(first + j) - 0.) * disc[j];
        if (j == 0) { sum = dj; ssq = dj * dj; } else { sum = sum + dj; ssq = ssq + dj * dj; }
      }
      const double num = sum / n;
      const double denom = sqrt(ssq / (n - 1));
      const double out = (denom != 0.) ? num / denom : 0.;
      return clip(.1 * out, -1., 1.);
    }
    case MDG_SHAPER_SORTINO_A: {  // :242-272
      const double ex = R.sortino_exp;
      if (n == 1) {
        const double diff = v.at(first) - 0.;
        const double downside = pow(pow(fabs(diff), ex), 1 / ex);
        return clip(0.1 * ((diff != 0.) ? diff / downside : 0.), -1., 1.);
      }
      double sum = 0., den = 0.;
      for (int j = 0; j < n; ++j) {
        const double dj = (v.at(first + j) - 0.) * disc[j];
        double down = dj < 0. ? dj : 0.;
        if (dj != dj) down = dj;
        if (down < -1.) down = -1.;
        const double t = pow(pow(fabs(down), ex) / (n - 1), 1 / ex);
        if (j == 0) { sum = dj; den = t; } else { sum = sum + dj; den = den + t; }
      }
      const double num = sum / n;
      const double zero_case = (num == 0.) ? 0. : 1.;
      const double normal = clip(.1 * (num / den), -1., 1.);
      return (den != 0.) ? normal : zero_case;
    }
    case MDG_SHAPER_SORTINO_B: {  // :276-312
      const double ex = R.sortino_exp;
      if (n == 1) {
        double diff = v.at(first) - 0.;
        if (diff < -1.) diff = -1.;
        if (diff < 0.) diff = -pow(-diff, 1 / ex);
        return clip(diff, -1., 1.);
      }
      double s = 0.;
      for (int j = 0; j < n; ++j) {
        double dj = (v.at(first + j) - 0.) * disc[j];
        if (dj < -1.) dj = -1.;
        if (dj < 0.) dj = -pow(-dj, 1 / ex);
        s = (j == 0) ? dj : s + dj;
      }
      return clip(s, -1., 1.);
    }
  }
  return 0.;
}

// ReplayBuffer.add (replay_buffer.py:68-80) + NStepBuffer.pop_nstep_sarsd (nstep_buffer.py:342-361)
// for component c of env e: add `raw`, pop once when full, drain on done.
__device__ __forceinline__ void shaper_add(const StepArgs& a, int64_t e, int c, int ra, double raw, bool done,
                                           int len_before, int& len_after, int& n_popped) {
  const MdgReward& R = a.R;
  const int64_t N = a.L.n_envs;
  const int n = R.nstep;
  double A = 0., B = 0.;
  const bool moments = (R.shaper == MDG_SHAPER_DSR || R.shaper == MDG_SHAPER_DDR);
  if (moments) {
    A = a.S.shaper_A[(int64_t)c * N + e];
    B = a.S.shaper_B[(int64_t)c * N + e];
  }
  NStepView v;
  v.ring = a.S.nstep_ring;
  v.N = N; v.e = e; v.n = n; v.ra = ra; v.c = c;
  v.newest = raw;
  v.len = len_before + 1;
  int base = a.L.nstep_pos - len_before;
  if (base < 0) base += n;
  v.base = base;
  if (n > 1) a.S.nstep_ring[((int64_t)a.L.nstep_pos * ra + c) * N + e] = raw;
  int first = 0, k = 0;
  if (v.len >= n) {
    a.IO.shaped_reward[((int64_t)k * ra + c) * N + e] = shaper_pop(R, v, first, A, B);
    ++first; ++k;
  }
  if (done) {
    while (first < v.len) {
      a.IO.shaped_reward[((int64_t)k * ra + c) * N + e] = shaper_pop(R, v, first, A, B);
      ++first; ++k;
    }
  }
  if (moments) {
    a.S.shaper_A[(int64_t)c * N + e] = A;
    a.S.shaper_B[(int64_t)c * N + e] = B;
  }
  len_after = v.len - first;
  n_popped = k;
}

// ---------------------------------------------------------------------------
// the step kernel
// ---------------------------------------------------------------------------
template <int CAP, bool EXACT>
__global__ void __launch_bounds__(kBlock) step_kernel(const __grid_constant__ StepArgs a) {
  constexpr int US = CAP | 1;  // odd row stride (in doubles): conflict-free per-thread rows
  __shared__ double s_units[kBlock * US];
  __shared__ double s_prev[CAP][kBlock];

  const MdgParams& P = a.P;
  const int64_t N = a.L.n_envs;
  const int na = EXACT ? CAP : P.n_assets;
  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * kBlock;
  const int64_t e = e0 + tid;
  const bool active = e < N;
  const int mode = a.L.mode;

  // stage this block's (rows, na) slice of the row-major units matrix through smem (coalesced)
  if (mode == MDG_MODE_MULTI) {
    const int64_t rows = (N - e0) < kBlock ? (N - e0) : kBlock;
    const int total = (int)rows * na;
    const double* src = a.IO.units + e0 * na;
    for (int idx = tid; idx < total; idx += kBlock) {
      const int r = idx / na, c = idx - r * na;
      s_units[r * US + c] = src[idx];
    }
  } else if (mode == MDG_MODE_SINGLE) {
    if (active) s_units[tid * US] = a.IO.units[e];
  }
  __syncthreads();
  if (!active) return;

  // ---- load state
  Port<CAP> q;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    if (EXACT || j < na) {
      q.price[j] = a.S.price[(int64_t)j * N + e];
      q.led[j] = a.S.ledger[(int64_t)j * N + e];
      q.mep[j] = a.S.mean_entry[(int64_t)j * N + e];
      q.bm[j] = a.S.borrowed[(int64_t)j * N + e];
    } else {
      q.price[j] = 0.; q.led[j] = 0.; q.mep[j] = 0.; q.bm[j] = 0.;
    }
  }
  q.cash = a.S.cash[e];
  const int64_t ts = a.S.timestamp[e];
  const bool shaping = (a.R.shaper != MDG_SHAPER_OFF) && (mode != MDG_MODE_HOLD);

  // ---- prevEq (Env.h:190,208,234) and prev position values (offpolicy_q.py:140-141)
  double prevEq;
  {
    double av, ml, bms, se;
    port_sums<CAP, EXACT>(q, na, av, ml, bms, se);
    prevEq = q.cash + av - bms;
  }
  if (shaping) {
#pragma unroll
    for (int j = 0; j < CAP; ++j)
      if (EXACT || j < na) s_prev[j][tid] = q.led[j] * q.price[j];
  }

  // ---- transactions, sequential over assets (Broker.cpp:144-158)
  bool bad_risk = false;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (EXACT || i < na) {
      double tp = 0., tu = 0., tc = 0.;
      int risk = MDG_RISK_GREEN;
      if (mode == MDG_MODE_MULTI) {
        risk = broker_transaction<CAP, EXACT>(q, na, P, i, s_units[tid * US + i], tp, tu, tc);
      } else if (mode == MDG_MODE_SINGLE && i == a.L.asset_idx) {
        risk = broker_transaction<CAP, EXACT>(q, na, P, i, s_units[tid * US], tp, tu, tc);
      }
      if (risk != MDG_RISK_GREEN && risk != MDG_RISK_INSUFF_MARGIN) bad_risk = true;
      if (mode != MDG_MODE_HOLD) {
        a.IO.trans_price[(int64_t)i * N + e] = tp;
        a.IO.trans_units[(int64_t)i * N + e] = tu;
        a.IO.trans_cost[(int64_t)i * N + e] = tc;
        a.IO.risk[(int64_t)i * N + e] = (uint8_t)risk;
      }
      if (shaping) s_units[tid * US + i] = tu * tp + tc;  // mar_diff, offpolicy_q.py:154-155
    }
  }
  if (mode != MDG_MODE_HOLD) {  // BrokerResponse.marginCall (Broker.cpp:135,156)
    double av, ml, bms, se;
    port_sums<CAP, EXACT>(q, na, av, ml, bms, se);
    a.IO.margin_call[e] = margin_call(q.cash, av, ml, bms, se, P.maintenance_margin) ? 1 : 0;
  }

  // ---- generator tick (DataSource.cpp getData family)
  {
    Draws d;
    d.init(a.IO, a.L, e, ts, 0, 0);
    double pair_mean = 0.;
#pragma unroll
    for (int i = 0; i < CAP; ++i) {
      if (EXACT || i < na) {
        const MdgAssetGen& g = P.gen[i];
        double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
        q.price[i] = gen_tick(g, q.price[i], gs, N, d, pair_mean);
        a.S.price[(int64_t)i * N + e] = q.price[i];
      }
    }
    a.S.timestamp[e] = ts + 1;
  }

  // ---- write back the ledger
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    if (EXACT || j < na) {
      a.S.ledger[(int64_t)j * N + e] = q.led[j];
      a.S.mean_entry[(int64_t)j * N + e] = q.mep[j];
      a.S.borrowed[(int64_t)j * N + e] = q.bm[j];
    }
  }
  a.S.cash[e] = q.cash;

  // ---- equity, reward, done (Env.h:192-198, 211-223, 237-249)
  double av, ml, bms, se;
  port_sums<CAP, EXACT>(q, na, av, ml, bms, se);
  const double currentEq = q.cash + av - bms;
  const double clampv = (mode == MDG_MODE_SINGLE) ? 0.01 : 0.3;
  a.IO.reward[e] = log(dmax(currentEq / prevEq, clampv));
  const bool mc = margin_call(q.cash, av, ml, bms, se, P.maintenance_margin);
  bool done = mc || (currentEq < 0.1 * P.init_cash);
  if (mode != MDG_MODE_HOLD) done = done || bad_risk;
  a.IO.done[e] = done ? 1 : 0;

  // ---- newest observation row: State(price, ledgerNormedFull, timestamp)  (Env.h:202,228,254)
  const int head = a.L.head;
  double cosv_pp = 0., cosv_qq = 0., cosv_pq = 0.;
  const bool cosine = shaping && a.R.shaper == MDG_SHAPER_COSINE;
  {
    const double w0 = (q.cash - bms) / currentEq;  // Portfolio.cpp:150-155
    a.IO.obs_port[((int64_t)head * (na + 1)) * N + e] = w0;
    if (cosine) {
      const double d0 = a.R.desired_portfolio[0];
      cosv_pp = w0 * w0; cosv_qq = d0 * d0; cosv_pq = w0 * d0;
    }
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      if (EXACT || j < na) {
        const double w = (q.led[j] * q.price[j]) / currentEq;
        a.IO.obs_price[((int64_t)head * na + j) * N + e] = q.price[j];
        a.IO.obs_port[((int64_t)head * (na + 1) + j + 1) * N + e] = w;
        if (cosine) {
          const double dj = a.R.desired_portfolio[j + 1];
          cosv_pp = cosv_pp + w * w; cosv_qq = cosv_qq + dj * dj; cosv_pq = cosv_pq + w * dj;
        }
      }
    }
    a.IO.obs_time[(int64_t)head * N + e] = ts + 1;
  }

  // ---- agent reward (offpolicy_q.py:152-164) and the n-step shaper
  if (shaping) {
    const int ra = a.R.reduce_rewards ? 1 : na;
    double extra = 0.;
    if (cosine) extra = a.R.cosine_temp * (cosv_pq / (sqrt(cosv_pp) * sqrt(cosv_qq)));  // nstep_buffer.py:173-191
    const int len_before = (a.R.nstep > 1) ? a.S.nstep_len[e] : 0;
    int len_after = 0, n_popped = 0;
    double rsum = 0.;
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      if (EXACT || j < na) {
        const double curVal = q.led[j] * q.price[j];
        double x = (curVal - s_prev[j][tid] - s_units[tid * US + j]) / prevEq;
        x += 1;
        const double r = log((x != x) ? x : ((x < .35) ? .35 : x));
        if (a.R.reduce_rewards) {
          rsum = (j == 0) ? r : rsum + r;
        } else {
          a.IO.agent_reward[(int64_t)j * N + e] = r;
          shaper_add(a, e, j, ra, cosine ? r + extra : r, done, len_before, len_after, n_popped);
        }
      }
    }
    if (a.R.reduce_rewards) {
      a.IO.agent_reward[e] = rsum;
      shaper_add(a, e, 0, 1, cosine ? rsum + extra : rsum, done, len_before, len_after, n_popped);
    }
    if (a.R.nstep > 1) a.S.nstep_len[e] = len_after;
    a.IO.n_popped[e] = n_popped;
  }
}

template <int CAP>
static inline int launch_step(const StepArgs& a, bool exact) {
  const int64_t N = a.L.n_envs;
  const unsigned grid = (unsigned)((N + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)a.L.stream;
  if (exact)
    step_kernel<CAP, true><<<grid, kBlock, 0, st>>>(a);
  else
    step_kernel<CAP, false><<<grid, kBlock, 0, st>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_step launch");
}

// one translation unit per capacity (mdg_step_inst.cu, -DMDG_CAP=n) so they compile in parallel
int launch_step_cap1(const StepArgs& a, bool exact);
int launch_step_cap2(const StepArgs& a, bool exact);
int launch_step_cap4(const StepArgs& a, bool exact);
int launch_step_cap8(const StepArgs& a, bool exact);
int launch_step_cap16(const StepArgs& a, bool exact);

}  // namespace mdg
