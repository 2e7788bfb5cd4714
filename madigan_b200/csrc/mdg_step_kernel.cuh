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
#define MDG_GEN_GENERIC 0  // per-asset generator table, non-inlined generator bodies (Composite, Synth, trends ...)
#define MDG_GEN_OUPAIRS 1  // every asset belongs to an OUPair (the headline workload): inlined pair update

// one term of the four folds for asset j
#define MDG_TERMS(j)                                         \
  const double t_av = q.led[j] * q.price[j];                 \
  const double t_ml = q.mep[j] * q.led[j];                   \
  const double t_se = (q.led[j] < 0.) ? t_ml : 0. * t_ml;    \
  const double t_bm = q.bm[j];

#ifndef MDG_MINB16
#define MDG_MINB16 2
#endif
template <int CAP> constexpr size_t step_smem_bytes() { return sizeof(double) * (size_t)kBlock * ((CAP | 1) + CAP + (3 * CAP + 1) / 2); }

template <int CAP, bool EXACT, int GENK>
__global__ void __launch_bounds__(kBlock, (CAP >= 16 ? MDG_MINB16 : 4)) step_kernel(const __grid_constant__ StepArgs a) {
  constexpr int US = CAP | 1;                 // odd row stride (doubles): conflict-free per-thread rows
  constexpr int NZ = (3 * CAP + 1) / 2;       // normal draws per tick, worst case (all OU pairs)
  // dynamic shared memory (57 KB at CAP=16, above the 48 KB static limit), see step_smem_bytes()
  extern __shared__ double smem_dyn[];
  double* s_units = smem_dyn;                                                   // [kBlock*US] units, later mar_diff (offpolicy_q.py:154)
  double (*s_prev)[kBlock] = reinterpret_cast<double (*)[kBlock]>(smem_dyn + kBlock * US);        // [CAP] prev position values, later reward numerators
  double (*s_z)[kBlock] = reinterpret_cast<double (*)[kBlock]>(smem_dyn + kBlock * US + CAP * kBlock);  // [NZ] this tick's normal draws

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

  // ---- load state (all loads issued before the RNG loop below, which hides their latency)
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

  // ---- this tick's normal draws (Philox4x32-10 + Box-Muller, or the injected stream)
  GenCtx ctx;
  ctx_init(ctx, &s_z[0][tid], kBlock, a.IO.uniforms, a.L, e, ts);
  fill_normals(&s_z[0][tid], kBlock, P.n_normals, a.IO.normals, ctx);

  // ---- prevEq (Env.h:190,208,234) and previous position values (offpolicy_q.py:140-141)
  double prevEq;
  {
    double av = 0., bms = 0.;
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      if (EXACT || j < na) {
        const double t = q.led[j] * q.price[j];
        if (j == 0) { av = t; bms = q.bm[0]; } else { av = av + t; bms = bms + q.bm[j]; }
        if (shaping) s_prev[j][tid] = t;
      }
    }
    prevEq = q.cash + av - bms;
  }

  // ---- transactions, sequential over assets (Broker.cpp:124-158, Portfolio.cpp:254-323).
  // Every accounting quantity is a left-to-right fold over the assets.  Asset i's risk gate sees
  // final values for assets < i and untouched values for assets >= i, so the fold is kept as a
  // running PREFIX over processed assets and only the suffix i..nA-1 is re-added: the same
  // operations in the same order as a full recomputation (bit-identical), at half the work.
  double pav = 0., pml = 0., pbm = 0., pse = 0.;
  bool bad_risk = false;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (EXACT || i < na) {
      double tp = 0., tu = 0., tc = 0.;
      int risk = MDG_RISK_GREEN;
      double units = 0.;
      if (mode == MDG_MODE_MULTI) units = s_units[tid * US + i];
      else if (mode == MDG_MODE_SINGLE && i == a.L.asset_idx) units = s_units[tid * US];
      if (units != 0.) {  // Broker.cpp:126 (NaN units do enter, as in the reference)
        const double price = q.price[i];
        double cur = q.led[i];
        const bool opposite = (signbit(units) != 0) != (signbit(cur) != 0);
        if (!opposite || units > -1 * cur) {  // Portfolio.cpp:257-258: only these orders are gated
          double av = pav, ml = pml, bms = pbm, se = pse;
#pragma unroll
          for (int j = i; j < CAP; ++j) {
            if (EXACT || j < na) {
              MDG_TERMS(j)
              if (j == 0) { av = t_av; ml = t_ml; bms = t_bm; se = t_se; }
              else { av = av + t_av; ml = ml + t_ml; bms = bms + t_bm; se = se + t_se; }
            }
          }
          const double pnl = av - ml;           // :184-186
          const double balance = q.cash + se;   // :192-197
          const double availableMargin = (balance + pnl) / P.required_margin;  // :229-231
          if (opposite) {
            const double excess = units + cur;
            if (availableMargin <= fabs(price * excess) || balance <= 0.) risk = MDG_RISK_INSUFF_MARGIN;
          } else if (margin_call(q.cash, av, ml, bms, se, P.maintenance_margin)) {
            risk = MDG_RISK_MARGIN_CALL;
          } else {
            const double cashAmount = price * units;
            if (availableMargin <= fabs(cashAmount) || balance <= 0.) risk = MDG_RISK_INSUFF_MARGIN;
          }
        }
        if (risk == MDG_RISK_GREEN) {
          // Broker::applySlippage / getTransactionCost  Broker.cpp:171-178
          const double slippage = (price * P.slippage_rel) + P.slippage_abs;
          const double transactionPrice = units < 0 ? (price - slippage) : (price + slippage);
          const double transactionCost = fabs(units * price) * P.tcost_rel + P.tcost_abs;
          tp = transactionPrice; tu = units; tc = transactionCost;
          // Portfolio::handleTransaction  Portfolio.cpp:284-323
          double mep = q.mep[i];
          if (opposite) {
            if (fabs(units) > fabs(cur)) {
              units += cur;
              q.cash += cur * transactionPrice;
              cur = 0.;
              mep = transactionPrice;
            }
          } else {
            mep += (transactionPrice - mep) * (units / (units + cur));
          }
          const double amount = transactionPrice * units;
          const double marginToUse = amount * P.required_margin;
          const double marginToBorrow = amount - marginToUse;
          double bm = q.bm[i];
          bm += marginToBorrow;
          q.cash -= (marginToUse + transactionCost);
          cur += units;
          if (fabs(cur) < 0.000001) {
            mep = 0.;
            if (bm > 0.) { q.cash -= bm; bm = 0.; }
          }
          if (bm < 0.) { q.cash -= bm; bm = 0.; }
          q.led[i] = cur; q.mep[i] = mep; q.bm[i] = bm;
        } else if (risk != MDG_RISK_INSUFF_MARGIN) {
          bad_risk = true;
        }
      }
      {  // extend the prefix folds with asset i's final values
        MDG_TERMS(i)
        if (i == 0) { pav = t_av; pml = t_ml; pbm = t_bm; pse = t_se; }
        else { pav = pav + t_av; pml = pml + t_ml; pbm = pbm + t_bm; pse = pse + t_se; }
      }
      if (mode != MDG_MODE_HOLD) {
        a.IO.trans_price[(int64_t)i * N + e] = tp;
        a.IO.trans_units[(int64_t)i * N + e] = tu;
        a.IO.trans_cost[(int64_t)i * N + e] = tc;
        a.IO.risk[(int64_t)i * N + e] = (uint8_t)risk;
        if (shaping) s_units[tid * US + i] = tu * tp + tc;  // mar_diff, offpolicy_q.py:154-155
      }
    }
  }
  // BrokerResponse.marginCall (Broker.cpp:135,156): Portfolio::checkRisk() after the last transaction
  if (mode != MDG_MODE_HOLD)
    a.IO.margin_call[e] = margin_call(q.cash, pav, pml, pbm, pse, P.maintenance_margin) ? 1 : 0;

  // ---- write back the ledger (prices follow after the tick)
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    if (EXACT || j < na) {
      a.S.ledger[(int64_t)j * N + e] = q.led[j];
      a.S.mean_entry[(int64_t)j * N + e] = q.mep[j];
      a.S.borrowed[(int64_t)j * N + e] = q.bm[j];
    }
  }
  a.S.cash[e] = q.cash;

  // ---- generator tick (DataSource.cpp getData family)
  if (GENK == MDG_GEN_OUPAIRS) {
#pragma unroll
    for (int p = 0; p < CAP / 2; ++p) {  // OUPair::getData, DataSource.cpp:1232-1240 (draw order rw, x0, x1)
      const MdgAssetGen& g0 = P.gen[2 * p];
      const MdgAssetGen& g1 = P.gen[2 * p + 1];
      double* mrow = a.S.gstate + (int64_t)g0.gslot * N + e;
      double m = *mrow;
      m += m * (draw_normal(ctx, g0.nslot_aux) * g0.p[2]);
      *mrow = m;
      q.price[2 * p] += (g0.p[0] * (m - q.price[2 * p])) + m * (draw_normal(ctx, g0.nslot) * g0.p[1]);
      q.price[2 * p + 1] += (g1.p[0] * (m - q.price[2 * p + 1])) + m * (draw_normal(ctx, g1.nslot) * g1.p[1]);
    }
  } else {
    double pair_mean = 0.;
#pragma unroll
    for (int i = 0; i < CAP; ++i) {
      if (EXACT || i < na) {
        const MdgAssetGen& g = P.gen[i];
        double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
        q.price[i] = gen_tick(g, q.price[i], gs, ctx, pair_mean);
      }
    }
  }
  a.S.timestamp[e] = ts + 1;

  // ---- equity, reward, done (Env.h:192-198, 211-223, 237-249); only the price-dependent fold changes
  double av = 0.;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    if (EXACT || j < na) {
      a.S.price[(int64_t)j * N + e] = q.price[j];
      const double t = q.led[j] * q.price[j];
      av = (j == 0) ? t : av + t;
    }
  }
  const double currentEq = q.cash + av - pbm;
  const double clampv = (mode == MDG_MODE_SINGLE) ? 0.01 : 0.3;
  a.IO.reward[e] = log(dmax(currentEq / prevEq, clampv));
  const bool mc = margin_call(q.cash, av, pml, pbm, pse, P.maintenance_margin);
  bool done = mc || (currentEq < 0.1 * P.init_cash);
  if (mode != MDG_MODE_HOLD) done = done || bad_risk;
  a.IO.done[e] = done ? 1 : 0;

  // ---- newest observation row: State(price, ledgerNormedFull, timestamp)  (Env.h:202,228,254).
  // Observations carry a 1e-9 bar (not bit-exactness), so the nA+1 divisions by equity are one
  // reciprocal and nA+1 multiplies.
  const int head = a.L.head;
  const double inv_eq = 1. / currentEq;
  double cosv_pp = 0., cosv_qq = 0., cosv_pq = 0.;
  const bool cosine = shaping && a.R.shaper == MDG_SHAPER_COSINE;
  {
    const double w0 = (q.cash - pbm) * inv_eq;  // Portfolio.cpp:150-155
    a.IO.obs_port[((int64_t)head * (na + 1)) * N + e] = w0;
    if (cosine) {
      const double d0 = a.R.desired_portfolio[0];
      cosv_pp = w0 * w0; cosv_qq = d0 * d0; cosv_pq = w0 * d0;
    }
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      if (EXACT || j < na) {
        const double cur_val = q.led[j] * q.price[j];
        const double w = cur_val * inv_eq;
        a.IO.obs_price[((int64_t)head * na + j) * N + e] = q.price[j];
        a.IO.obs_port[((int64_t)head * (na + 1) + j + 1) * N + e] = w;
        if (cosine) {
          const double dj = a.R.desired_portfolio[j + 1];
          cosv_pp = cosv_pp + w * w; cosv_qq = cosv_qq + dj * dj; cosv_pq = cosv_pq + w * dj;
        }
        // numerator of the agent reward: curr_val - prev_val - mar_diff (offpolicy_q.py:156)
        if (shaping) s_prev[j][tid] = cur_val - s_prev[j][tid] - s_units[tid * US + j];
      }
    }
    a.IO.obs_time[(int64_t)head * N + e] = ts + 1;
  }

  // ---- agent reward (offpolicy_q.py:152-164) and the n-step shaper; a rolled loop over assets
  if (shaping) {
    const int ra = a.R.reduce_rewards ? 1 : na;
    double extra = 0.;
    if (cosine) extra = a.R.cosine_temp * (cosv_pq / (sqrt(cosv_pp) * sqrt(cosv_qq)));  // nstep_buffer.py:173-191
    const int len_before = (a.R.nstep > 1) ? a.S.nstep_len[e] : 0;
    const double inv_prev = 1. / prevEq;
    int len_after = 0, n_popped = 0;
    double rsum = 0.;
#pragma unroll 1
    for (int j = 0; j < na; ++j) {
      double x = s_prev[j][tid] * inv_prev;
      x += 1;
      const double r = log((x != x) ? x : ((x < .35) ? .35 : x));
      if (a.R.reduce_rewards) {
        rsum = (j == 0) ? r : rsum + r;
      } else {
        a.IO.agent_reward[(int64_t)j * N + e] = r;
        shaper_add(a, e, j, ra, cosine ? r + extra : r, done, len_before, len_after, n_popped);
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

// host side: is every asset part of an OUPair laid out (role0, role1) consecutively?
static inline bool all_ou_pairs(const MdgParams& P) {
  if (P.n_assets % 2) return false;
  for (int i = 0; i < P.n_assets; i += 2) {
    const MdgAssetGen &g0 = P.gen[i], &g1 = P.gen[i + 1];
    if (g0.type != MDG_GEN_OUPAIR || g1.type != MDG_GEN_OUPAIR || g0.role != 0 || g1.role != 1 ||
        g1.partner != i || g0.gslot < 0)
      return false;
  }
  return true;
}

template <int CAP>
static inline int launch_step(const StepArgs& a, bool exact) {
  const int64_t N = a.L.n_envs;
  const unsigned grid = (unsigned)((N + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)a.L.stream;
  constexpr size_t smem = step_smem_bytes<CAP>();
  static bool attr_done = false;  // opt in to > 48 KB dynamic shared memory once per process
  if (!attr_done) {
    cudaError_t ce = cudaSuccess;
    if constexpr (CAP % 2 == 0)
      ce = cudaFuncSetAttribute(step_kernel<CAP, true, MDG_GEN_OUPAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(step_kernel<CAP, true, MDG_GEN_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(step_kernel<CAP, false, MDG_GEN_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return cuda_err(ce, "mdg_step smem attribute");
    attr_done = true;
  }
  if constexpr (CAP % 2 == 0) {
    if (exact && all_ou_pairs(a.P)) {
      step_kernel<CAP, true, MDG_GEN_OUPAIRS><<<grid, kBlock, smem, st>>>(a);
      return cuda_err(cudaGetLastError(), "mdg_step launch");
    }
  }
  if (exact)
    step_kernel<CAP, true, MDG_GEN_GENERIC><<<grid, kBlock, smem, st>>>(a);
  else
    step_kernel<CAP, false, MDG_GEN_GENERIC><<<grid, kBlock, smem, st>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_step launch");
}

// one translation unit per capacity (mdg_step_inst.cu, -DMDG_CAP=n) so they compile in parallel
int launch_step_cap1(const StepArgs& a, bool exact);
int launch_step_cap2(const StepArgs& a, bool exact);
int launch_step_cap4(const StepArgs& a, bool exact);
int launch_step_cap8(const StepArgs& a, bool exact);
int launch_step_cap16(const StepArgs& a, bool exact);

}  // namespace mdg
