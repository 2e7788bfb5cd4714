// Fused Env.step kernel for N independent environments (sm_100a).
//
// One launch does what the reference does across Env.h:189-256, Broker.cpp:124-178, Portfolio.cpp:140-323,
// the DataSource.cpp getData family, offpolicy_q.py:140-164 and nstep_buffer.py:
//   transact (sequential over assets, risk-gated) -> generator tick -> equity, reward, done ->
//   newest observation-ring row -> agent reward -> shaped reward -> (auto-reset) done list.
//
// Work decomposition ("one warp per generator group, one lane per env"):
//   * a block owns a tile of 32 envs; warp g of the block owns generator group g of those envs (an OU pair,
//     or a single asset of a Composite), lane = env.  State tensors are [rows][N], so every load/store of a
//     warp is one contiguous 256-byte run and every value is read once.
//   * everything that does not depend on earlier trades runs in parallel across the warps: state loads,
//     Philox + Box-Muller (the block generates each normal exactly once into shared memory), the generator
//     tick, the observation-ring stores.
//   * the ledger is inherently sequential over the assets of an env (cash and the risk gates of asset i see
//     the trades of assets < i, Broker.cpp:144-158): it runs as a CHAIN over the warps.  Warp g waits on a
//     named barrier for warp g-1's carry (cash, running sums, exact prefix folds; 13 doubles per env through
//     shared memory), transacts its assets with all 32 lanes busy, and hands the carry to warp g+1.
//   * the last warp holds the complete folds: it computes equity, reward, done, the n-step shaped reward, and
//     publishes 1/equity so that every warp writes its own ledgerNormedFull weights.
// Compared with one thread per env streaming over 16 assets (round 1: 13.8 warps per SM at 65,536 envs, one
// 8 k-instruction dependent chain per thread, latency-bound at 0.47 of the HBM roofline) this gives 8x the warps
// with ~1/6 of the per-thread chain and the same total instruction count.
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
  int n_groups;                   // warps per block: generator groups (OU pair = 1 group, other assets 1 each)
  int8_t leader[MDG_MAX_ASSETS];  // first asset of group g
  int* done_count;                // nullable (auto-reset): number of envs that finished this step ...
  int* done_list;                 // ... and their indices, appended by the last warp of each block
};

constexpr int kTile = 32;  // envs per block (one lane each)

// blocks per SM the two specialisations are compiled for (register cap = 65,536 / (threads * blocks))
#ifndef MDG_MINB_PAIRS
#define MDG_MINB_PAIRS 4    // 8 warps x 4 blocks: 64 registers
#endif
#ifndef MDG_MINB_GENERIC
#define MDG_MINB_GENERIC 2  // 16 warps x 2 blocks: 64 registers
#endif

// named barriers (ids 1..15; 0 is __syncthreads): producer warp arrives, consumer warp syncs, 64 threads
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

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
  // (v^2)^(3/4) = |v|^(3/2), as |v| sqrt|v| (2 ulp; pow() costs ~250 instructions per call)
  const double av = fabs(v);
  return (B * dA - (A * dB) / 2) / (av * sqrt(av) + kEps32);
}
__device__ __forceinline__ double ddr_value(double A, double B, double r) {  // :146-156
  if (r > 0.) return (r - A / 2) / (sqrt(B) + kEps32);
  return (B * (r - A / 2) - (A * (r * r)) / 2) / (B * sqrt(B) + kEps32);  // B^(3/2)
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
// for component c of env e: add `raw`, pop once when full, drain on done.  `valid` gates every store (lanes
// past the end of the slab run on a clamped env index).
static __device__ __noinline__ void shaper_add(const StepArgs& a, int64_t e, bool valid, int c, int ra, double raw,
                                               bool done, int len_before, int& len_after, int& n_popped) {
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
  if (n > 1 && valid) a.S.nstep_ring[((int64_t)a.L.nstep_pos * ra + c) * N + e] = raw;
  int first = 0, k = 0;
  if (v.len >= n) {
    const double s = shaper_pop(R, v, first, A, B);
    if (valid) a.IO.shaped_reward[((int64_t)k * ra + c) * N + e] = s;
    ++first; ++k;
  }
  if (done) {
#pragma unroll 1
    while (first < v.len) {
      const double s = shaper_pop(R, v, first, A, B);
      if (valid) a.IO.shaped_reward[((int64_t)k * ra + c) * N + e] = s;
      ++first; ++k;
    }
  }
  if (moments && valid) {
    a.S.shaper_A[(int64_t)c * N + e] = A;
    a.S.shaper_B[(int64_t)c * N + e] = B;
  }
  len_after = v.len - first;
  n_popped = k;
}

// ---------------------------------------------------------------------------
// the ledger
// ---------------------------------------------------------------------------
// rows of MdgState.folds
#define MDG_FOLD_AV 0
#define MDG_FOLD_ML 1
#define MDG_FOLD_BM 2
#define MDG_FOLD_SE 3
#define MDG_FOLD_G 4

// Exact risk gate of asset i (Portfolio::checkRisk(i, units), Portfolio.cpp:254-279): every accounting
// quantity is a left-to-right fold over the assets, exactly as in the oracle.  COLD path, called only when
// the cheap bound in tx_asset cannot decide.  (av..se) are the folds over the already processed
// assets 0..i-1 (final values); assets i.. are untouched so far -- the warps that own them are still waiting
// for the carry, and price/ledger rows are stored only after a warp's own transactions -- and are re-read
// from global memory.
static __device__ __noinline__ int exact_gate(const MdgState& S, int64_t N, int64_t e, int na, double cash,
                                              double reqM, double maintM, int i, double units, double av,
                                              double ml, double bms, double se) {
#pragma unroll 1
  for (int j = i; j < na; ++j) {
    const double l = S.ledger[(int64_t)j * N + e], p = S.price[(int64_t)j * N + e],
                 m = S.mean_entry[(int64_t)j * N + e], b = S.borrowed[(int64_t)j * N + e];
    const double t_se = l * (m * (l < 0. ? 1. : 0.));
    if (j == 0) { av = l * p; ml = m * l; bms = b; se = t_se; }
    else { av = av + l * p; ml = ml + m * l; bms = bms + b; se = se + t_se; }
  }
  const double price = S.price[(int64_t)i * N + e], cur = S.ledger[(int64_t)i * N + e];
  const bool opposite = (signbit(units) != 0) != (signbit(cur) != 0);
  const double pnl = av - ml;        // :184-186
  const double balance = cash + se;  // :192-197
  const double availableMargin = (balance + pnl) / reqM;  // :229-231
  if (opposite) {
    const double excess = units + cur;
    if (availableMargin <= fabs(price * excess) || balance <= 0.) return MDG_RISK_INSUFF_MARGIN;
    return MDG_RISK_GREEN;
  }
  if (margin_call(cash, av, ml, bms, se, maintM)) return MDG_RISK_MARGIN_CALL;
  if (availableMargin <= fabs(price * units) || balance <= 0.) return MDG_RISK_INSUFF_MARGIN;
  return MDG_RISK_GREEN;
}

// What the ledger chain hands from one warp to the next, per env.
//
// DECISION sums (approximate, updated by one precomputed delta per accepted order -- the gates' critical path):
//   X = balance + pnl (= availableMargin * requiredMargin), BAL = balance, E = equity at the old prices,
//   PNL = pnl, G = bound on the sum of the magnitudes of everything that went into them.
// EXACT values (the reference's own operation order, they ride along off the critical path): cash with every
// Portfolio::handleTransaction update applied in sequence, and the left-to-right folds of the final ledger.
struct StepAcc {
  double X, BAL, E, PNL, G;
  double cash;
  double pav, pml, pbm, pse;  // exact left-to-right folds over the processed assets, old prices
  double nav, gsum;           // exact fold of ledger*new price; magnitude sum for the next step
  double rprod;               // reduced agent reward: product of the clamped per-asset ratios so far
  int bad_risk;
};
// rows of the shared-memory carry
enum { C_X, C_BAL, C_E, C_PNL, C_G, C_CASH, C_PAV, C_PML, C_PBM, C_PSE, C_NAV, C_GSUM, C_RPROD, C_ROWS };

struct StepConsts {
  double reqM, maintM, band_scale;
  double g1, g2;  // magnitude-bound coefficients of one transaction
  bool reqM_ok;
  bool force_exact;  // MdgLaunch.flags & MDG_FLAG_FORCE_EXACT_GATE: every gate takes the exact (cold) path
};

// Rows of the per-thread staging column (shared memory): the fields of a planned order that are needed only AFTER
// its gate has decided.  They wait there instead of in registers so that the kernel fits the 64-register budget of
// 4 resident tiles per SM without local-memory spills (which overflow L1 and would put L2 latency on the chain).
enum { ST_TP, ST_TC, ST_C1, ST_C2, ST_C3, ST_NCUR, ST_NMEP, ST_NBM, ST_MEP, ST_BM, ST_ROWS };

// One order, everything that does NOT depend on the trades of earlier assets -- computed before the warp waits
// for the carry: the outcome Broker::handleTransaction would produce if the risk gate says green
// (Broker.cpp:124-142,171-178; Portfolio.cpp:284-323, same operations in the same order), the three cash
// updates of that outcome, and the deltas it would add to the decision sums.  What the gate itself needs stays in
// registers (Order); the rest goes to the staging column `st` (row stride `ss`):
//   tp, tc          transaction price (slippage applied) and cost
//   c1, c2, c3      cash += c1 (reversal through zero), cash -= c2 (margin + cost), cash -= c3 (margin returned)
//   n_cur/mep/bm    ledger, mean entry price, borrowed margin after the order;  mep, bm: before it
struct Order {
  double units;  // requested units (BrokerResponse.transactionUnits when accepted)
  double amt;    // |price * (units [+ ledger])|: the amount the gate compares with availableMargin
  double dX, dBAL, dE, dG;
  bool nz, gated, opposite, f1, f3;
};

__device__ __forceinline__ void plan_order(const MdgParams& P, const StepConsts& c, double price, double cur,
                                           double mep, double bm, double units, Order& o, double* st, int ss) {
  st[ST_MEP * ss] = mep;
  st[ST_BM * ss] = bm;
  o.units = units;
  o.nz = units != 0.;  // Broker.cpp:126 (NaN units do enter, as in the reference)
  o.opposite = (signbit(units) != 0) != (signbit(cur) != 0);
  o.gated = o.nz && (!o.opposite || units > -1 * cur);  // Portfolio.cpp:257-258: only these orders are gated
  o.amt = fabs(price * (o.opposite ? units + cur : units));
  // Broker::applySlippage / getTransactionCost  Broker.cpp:171-178
  const double slippage = (price * P.slippage_rel) + P.slippage_abs;
  const double transactionPrice = units < 0 ? (price - slippage) : (price + slippage);
  const double transactionCost = fabs(units * price) * P.tcost_rel + P.tcost_abs;
  st[ST_TP * ss] = transactionPrice;
  st[ST_TC * ss] = transactionCost;
  // Portfolio::handleTransaction  Portfolio.cpp:284-323
  const double prev_val = cur * price;
  const double o_ml = mep * cur, o_bm = bm;
  const double o_se = (cur < 0.) ? o_ml : 0.;
  double c1 = 0., c3 = 0.;
  o.f1 = false;
  if (o.opposite) {
    if (fabs(units) > fabs(cur)) {
      units += cur;
      o.f1 = true; c1 = cur * transactionPrice;
      cur = 0.;
      mep = transactionPrice;
    }
  } else {
    mep += (transactionPrice - mep) * (units / (units + cur));
  }
  const double amount = transactionPrice * units;
  const double marginToUse = amount * c.reqM;
  const double marginToBorrow = amount - marginToUse;
  bm += marginToBorrow;
  const double c2 = marginToUse + transactionCost;
  cur += units;
  o.f3 = false;
  if (fabs(cur) < 0.000001) {
    mep = 0.;
    if (bm > 0.) { o.f3 = true; c3 = bm; bm = 0.; }
  }
  if (bm < 0.) { o.f3 = true; c3 = bm; bm = 0.; }
  st[ST_C1 * ss] = c1; st[ST_C2 * ss] = c2; st[ST_C3 * ss] = c3;
  st[ST_NCUR * ss] = cur; st[ST_NMEP * ss] = mep; st[ST_NBM * ss] = bm;
  // deltas of the decision sums (any rounding is fine here: the gate trusts them only outside a 1e-9*G band)
  const double n_av = cur * price, n_ml = mep * cur;
  const double dAV = n_av - prev_val, dML = n_ml - o_ml, dBM = bm - o_bm;
  const double dSE = ((cur < 0.) ? n_ml : 0.) - o_se;
  const double dCash = (c1 - c2) - c3;
  o.dBAL = dCash + dSE;
  o.dX = o.dBAL + (dAV - dML);
  o.dE = (dCash + dAV) - dBM;
  // every new term's magnitude is at most its old magnitude (already in G) plus |units|*(|price|+|tp|), three
  // terms, plus the cost; with |tp| <= |price|(1+|slip_rel|)+|slip_abs| and |cost| <= |units price||tc_rel|+|tc_abs|
  // that is |units| * (|price| * g1 + g2) (+ |tc_abs|, added once per asset by warp 0)
  o.dG = fabs(o.units) * (fabs(price) * c.g1 + c.g2);
}

// The chain step of one asset: the risk gate (Portfolio::checkRisk(i, units), Portfolio.cpp:254-279) on the
// decision sums, then -- when green -- the planned outcome.  The gate compares folds over the whole portfolio
// with thresholds; the decision sums are rounded differently from the reference's fresh left-to-right folds
// (they differ by < 1e-13 * G), so a decision is taken from them only when it clears its threshold by 1e-9 * G.
// Otherwise -- a knife-edge, NaN/Inf, a non-positive required margin -- the exact folds decide (exact_gate).
// Decisions, and therefore ledgers, are bit-identical to the oracle's either way.
// Returns the risk code; cur/mep/bm are the asset's final values (the planned ones when the order executed).
__device__ __forceinline__ int gate_and_apply(const StepArgs& a, const StepConsts& c, StepAcc& A, const Order& o,
                                              const double* st, int ss, int64_t N, int64_t e, int na, int i,
                                              double price, double& cur, double& mep, double& bm) {
  int risk = MDG_RISK_GREEN;
  bool exec = false;
  if (o.nz) {
    if (o.gated) {
      const double band = c.band_scale * (A.G + o.amt);
      const double d1 = A.X - o.amt * c.reqM;  // availableMargin <= |amount|  <=>  d1 <= 0
      bool certain = c.reqM_ok && fabs(d1) > band && fabs(A.BAL) > band;
      int r_fast = (d1 <= 0. || A.BAL <= 0.) ? MDG_RISK_INSUFF_MARGIN : MDG_RISK_GREEN;
      if (!o.opposite) {  // Portfolio::checkRisk() first (:268), :243-252
        const double m = c.maintM * A.PNL;
        const double d3 = A.E + m, d4 = A.X + m;
        certain = certain && fabs(d3) > band && fabs(d4) > band;
        if (d3 <= 0. || d4 <= 0.) r_fast = MDG_RISK_MARGIN_CALL;
      }
      risk = (certain && !c.force_exact)
                 ? r_fast
                 : exact_gate(a.S, N, e, na, A.cash, c.reqM, c.maintM, i, o.units, A.pav, A.pml, A.pbm, A.pse);
    }
    if (risk == MDG_RISK_GREEN) {
      exec = true;
      A.X += o.dX; A.BAL += o.dBAL; A.E += o.dE; A.PNL += o.dX - o.dBAL; A.G += o.dG;
      if (o.f1) A.cash += st[ST_C1 * ss];  // Portfolio.cpp:296
      A.cash -= st[ST_C2 * ss];            // :311
      if (o.f3) A.cash -= st[ST_C3 * ss];  // :316-321
    } else if (risk != MDG_RISK_INSUFF_MARGIN) {
      A.bad_risk = 1;
    }
  }
  cur = exec ? st[ST_NCUR * ss] : cur;
  mep = st[(exec ? ST_NMEP : ST_MEP) * ss];
  bm = st[(exec ? ST_NBM : ST_BM) * ss];
  // exact folds of the final ledger, old prices (Portfolio.cpp:180-197,207-209)
  const double t_ml = mep * cur;
  const double t_se = (cur < 0.) ? t_ml : 0. * t_ml;
  if (i == 0) { A.pav = cur * price; A.pml = t_ml; A.pbm = bm; A.pse = t_se; }
  else { A.pav = A.pav + cur * price; A.pml = A.pml + t_ml; A.pbm = A.pbm + bm; A.pse = A.pse + t_se; }
  A.gsum += fabs(t_ml) + fabs(bm);
  return risk;
}

// dqn.py:165-178: centred action times (unit_size * availableMargin / price); action 0 closes an open position
__device__ __forceinline__ double action_units(int act, int half, double scale, double price, double cur) {
  if (act == 0) return (cur != 0.) ? -cur : 0.;
  return (double)(act - half) * (scale / price);
}

// One Philox block -> two standard normals (Box-Muller), the same (block, lane) addressing as draw_normal:
// slot s = block s>>1, lane s&1.
__device__ __forceinline__ void normal_block(uint32_t gid, uint32_t blk, uint32_t t_lo, uint32_t t_hi, uint32_t k0,
                                             uint32_t k1, double& z_lane0, double& z_lane1) {
  uint64_t x0, x1;
  philox4x32_10(gid, blk, t_lo, t_hi, k0, k1, x0, x1);
  const double u1 = ((double)(x0 >> 12) + 0.5) * 0x1.0p-52;
  const double u2 = (double)(x1 >> 11) * 0x1.0p-53;
  const double r = fast_sqrt_pos(-2.0 * fast_log_pos(u1));
  double sn, cs;
  fast_sincos_2pi(u2, sn, cs);
  z_lane0 = r * cs;
  z_lane1 = r * sn;
}

// Draw source of the step kernel's generic path: this step's normals sit in the block's shared-memory stash
// (column = lane), or in the injected stream (validation mode); uniforms (trend sources) on demand.
struct StepDraws {
  const double* z;         // stash column of this lane: z[slot * kTile]
  const double* normals;   // [n_normals][N] (validation mode) or nullptr
  const double* uniforms;  // [n_uniforms][N] (validation mode) or nullptr
  int64_t N, e, gstride;
  uint32_t gid, k0, k1, t_lo, t_hi;
};
__device__ __forceinline__ double draw_normal(StepDraws& c, int slot) {
  return c.normals ? c.normals[(int64_t)slot * c.N + c.e] : c.z[slot * kTile];
}
static __device__ __noinline__ double draw_uniform(StepDraws& c, int slot) {
  if (c.uniforms) return c.uniforms[(int64_t)slot * c.N + c.e];
  uint64_t x0, x1;
  philox4x32_10(c.gid, (1u << 16) | (uint32_t)(slot >> 1), c.t_lo, c.t_hi, c.k0, c.k1, x0, x1);
  return (double)(((slot & 1) ? x1 : x0) >> 11) * 0x1.0p-53;
}

// generator-state rows owned by asset i (see MdgAssetGen)
__device__ __forceinline__ int gen_state_rows(const MdgAssetGen& g) {
  if (g.gslot < 0) return 0;
  switch (g.type) {
    case MDG_GEN_TRENDYOU: return 4;
    case MDG_GEN_TRENDOU: return 3;
    case MDG_GEN_SIMPLETREND: return 2;
    default: return 1;
  }
}

template <int MAXW>
struct StepSmem {
  union {
    double z[kMaxNormals][kTile];               // this step's normals (dead once every warp has ticked) ...
    double cosp[2][MDG_MAX_ASSETS][kTile];      // ... then the PPC shaper's per-group partial sums
  };
  double carry[C_ROWS][kTile];
  double inv_prev[kTile], prev_eq[kTile], act_scale[kTile];  // written by warp 0 before the first barrier
  double inv_eq[kTile], w0[kTile];                           // written by the last warp
  int bad_risk[kTile];
  int done[kTile];
  double stage[2 * ST_ROWS][MAXW * 32];
};

template <bool PAIRS, bool ACTIONS, int MAXW>
__device__ __forceinline__ void step_body(const StepArgs& a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<MAXW>& sm = *reinterpret_cast<StepSmem<MAXW>*>(smem_raw);
  const int tid = threadIdx.x;
  const MdgParams& P = a.P;
  const MdgState& S = a.S;
  const int64_t N = a.L.n_envs;
  const int na = P.n_assets;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5, ng = a.n_groups;
  int64_t e = (int64_t)blockIdx.x * kTile + lane;
  const bool valid = e < N;
  if (!valid) e = N - 1;  // lanes past the end of the slab recompute the last env; every store is predicated
  const int mode = a.L.mode;
  const bool shaping = (a.R.shaper != MDG_SHAPER_OFF) && (mode != MDG_MODE_HOLD);
  const bool cosine = shaping && a.R.shaper == MDG_SHAPER_COSINE;
  const bool reduce = shaping && a.R.reduce_rewards;
  const bool last = g == ng - 1;
  const int head = a.L.head;

  // ---- this warp's group: assets i0 .. i0+cnt-1
  const int i0 = PAIRS ? 2 * g : a.leader[g];
  const int cnt = PAIRS ? 2 : ((P.gen[i0].type == MDG_GEN_OUPAIR) ? 2 : 1);
  double price[2], cur[2], mep[2], bm[2], units[2], newp[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q < cnt) {
      const int64_t o = (int64_t)(i0 + q) * N + e;
      price[q] = S.price[o]; cur[q] = S.ledger[o]; mep[q] = S.mean_entry[o]; bm[q] = S.borrowed[o];
    } else {
      price[q] = cur[q] = mep[q] = bm[q] = 0.;
    }
    units[q] = 0.;
  }
  const int64_t ts = S.timestamp[e];
  const bool by_actions = ACTIONS && mode == MDG_MODE_MULTI && a.IO.actions;
  int act[2] = {0, 0};
  if (by_actions) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (q < cnt) act[q] = a.IO.actions[e * na + i0 + q];
  } else if (mode == MDG_MODE_MULTI) {
    // the caller's units matrix is (N, nA) env-major (what an agent's network emits): an env's row is one
    // 128-byte line, read 8 or 16 bytes at a time by the warps of the block (L1 serves the re-reads)
    const double* urow = a.IO.units + e * na + i0;
    if (PAIRS && (na & 1) == 0 && (reinterpret_cast<uintptr_t>(a.IO.units) & 15) == 0) {
      const double2 u2 = *reinterpret_cast<const double2*>(urow);
      units[0] = u2.x; units[1] = u2.y;
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (q < cnt) units[q] = urow[q];
    }
  } else if (mode == MDG_MODE_SINGLE) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (q < cnt && i0 + q == a.L.asset_idx) units[q] = a.IO.units[e];
  }

  StepConsts c;
  c.reqM = P.required_margin;
  c.maintM = P.maintenance_margin;
  c.reqM_ok = c.reqM > 0. && c.reqM <= 1e6;
  c.band_scale = 1e-9 * (1. + fabs(c.maintM)) * (c.reqM > 1. ? c.reqM : 1.);
  c.g1 = 6. + 3. * fabs(P.slippage_rel) + fabs(P.tcost_rel);
  c.g2 = 3. * fabs(P.slippage_abs);
  c.force_exact = (a.L.flags & MDG_FLAG_FORCE_EXACT_GATE) != 0;

  // ---- warp 0 starts the chain from the state.  The folds of the incoming portfolio are exactly the folds this
  // kernel (or reset/init/refresh) computed at the end of the previous call -- same values, same left-to-right
  // order -- so they are carried in state instead of re-reading the whole portfolio before the first transaction.
  StepAcc A;
  if (g == 0) {
    A.cash = S.cash[e];
    const double rAV = S.folds[(int64_t)MDG_FOLD_AV * N + e], rML = S.folds[(int64_t)MDG_FOLD_ML * N + e];
    const double rBM = S.folds[(int64_t)MDG_FOLD_BM * N + e], rSE = S.folds[(int64_t)MDG_FOLD_SE * N + e];
    A.G = S.folds[(int64_t)MDG_FOLD_G * N + e] + na * fabs(P.tcost_abs);  // + the absolute cost of up to nA transactions
    A.BAL = A.cash + rSE;       // Portfolio.cpp:192-197
    A.PNL = rAV - rML;          // :184-186
    A.X = A.BAL + A.PNL;        // availableMargin * requiredMargin, :229-231
    A.E = A.cash + rAV - rBM;   // equity, :211-213  (= prevEq, Env.h:190,208,234)
    A.pav = A.pml = A.pbm = A.pse = A.nav = A.gsum = 0.;
    A.rprod = 1.;
    A.bad_risk = 0;
    sm.inv_prev[lane] = 1. / A.E;
    sm.prev_eq[lane] = A.E;
    // DQN.action_to_transaction (dqn.py:160-179) fused in front of the step: units from discrete actions and the
    // availableMargin of the incoming portfolio (Portfolio.cpp:229-231), one scale for every asset
    sm.act_scale[lane] = by_actions ? a.L.unit_size * (A.X / P.required_margin) : 0.;
  }

  // ---- this step's normals, each generated once per block: Philox block b by warp b mod ng.  The blocks are
  // independent of each other and of the ledger; they run while the state loads above are in flight.
  const uint32_t gid = (uint32_t)(a.L.env_offset + e);
  const uint32_t k0 = (uint32_t)a.L.seed, k1 = (uint32_t)(a.L.seed >> 32);
  const uint32_t t_lo = (uint32_t)(uint64_t)ts, t_hi = (uint32_t)((uint64_t)ts >> 32);
  if (!a.IO.normals) {
    const int nn = P.n_normals, nb = (nn + 1) >> 1;
#pragma unroll 1
    for (int b = g; b < nb; b += ng) {
      double za, zb;
      normal_block(gid, (uint32_t)b, t_lo, t_hi, k0, k1, za, zb);
      sm.z[2 * b][lane] = za;
      if (2 * b + 1 < nn) sm.z[2 * b + 1][lane] = zb;
    }
  }
  __syncthreads();

  // ---- generator tick (independent of the ledger: prices never depend on trades)
  if (PAIRS) {  // OUPair::getData, DataSource.cpp:1232-1240 (draw order rw, x0, x1)
    const MdgAssetGen& g0 = P.gen[i0];
    const MdgAssetGen& g1 = P.gen[i0 + 1];
    double* mrow = S.gstate + (int64_t)g0.gslot * N + e;
    double mean = *mrow;
    double z_rw, z0, z1;
    if (a.IO.normals) {
      z_rw = a.IO.normals[(int64_t)g0.nslot_aux * N + e];
      z0 = a.IO.normals[(int64_t)g0.nslot * N + e];
      z1 = a.IO.normals[(int64_t)g1.nslot * N + e];
    } else {
      z_rw = sm.z[3 * g][lane]; z0 = sm.z[3 * g + 1][lane]; z1 = sm.z[3 * g + 2][lane];
    }
    mean += mean * (z_rw * g0.p[2]);
    if (valid) *mrow = mean;
    newp[0] = price[0] + ((g0.p[0] * (mean - price[0])) + mean * (z0 * g0.p[1]));
    newp[1] = price[1] + ((g1.p[0] * (mean - price[1])) + mean * (z1 * g1.p[1]));
  } else {  // DataSource.cpp getData family, generator state of the group in registers
    StepDraws dr;
    dr.z = &sm.z[0][lane];
    dr.normals = a.IO.normals; dr.uniforms = a.IO.uniforms;
    dr.N = N; dr.e = e; dr.gstride = 1;
    dr.gid = gid; dr.k0 = k0; dr.k1 = k1; dr.t_lo = t_lo; dr.t_hi = t_hi;
    const MdgAssetGen& g0 = P.gen[i0];
    const int ngs = gen_state_rows(g0);
    double gsl[4] = {0., 0., 0., 0.};
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < ngs) gsl[r] = S.gstate[(int64_t)(g0.gslot + r) * N + e];
    double pair_mean = 0.;
    newp[0] = gen_tick(g0, price[0], gsl, dr, pair_mean);
    newp[1] = 0.;
    if (cnt == 2) newp[1] = gen_tick(P.gen[i0 + 1], price[1], gsl, dr, pair_mean);
    if (valid) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < ngs) S.gstate[(int64_t)(g0.gslot + r) * N + e] = gsl[r];
    }
  }
  if (valid) {  // State.price row (Env.h:202,228,254); the price STATE rows are stored after the transactions
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (q < cnt) a.IO.obs_price[((int64_t)head * na + i0 + q) * N + e] = newp[q];
  }

  // ---- plan this warp's orders: the outcome of each if its gate says green, and what it would add to the
  // decision sums -- everything that can be known before the trades of the earlier assets
  const double inv_prev = sm.inv_prev[lane];
  constexpr int SS = MAXW * 32;  // row stride of the staging columns
  Order ord[2];
  {
    const int act_half = a.L.action_atoms / 2;
    const double act_scale = by_actions ? sm.act_scale[lane] : 0.;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q < cnt) {
        if (by_actions) units[q] = action_units(act[q], act_half, act_scale, price[q], cur[q]);
        plan_order(P, c, price[q], cur[q], mep[q], bm[q], units[q], ord[q], &sm.stage[q * ST_ROWS][tid], SS);
      }
    }
  }

  // ---- the ledger chain: receive the carry from warp g-1, gate + apply this warp's orders, hand it on
  __syncwarp();
  if (g != 0) {
    bar_sync(g, 64);
    A.X = sm.carry[C_X][lane]; A.BAL = sm.carry[C_BAL][lane]; A.E = sm.carry[C_E][lane];
    A.PNL = sm.carry[C_PNL][lane]; A.G = sm.carry[C_G][lane];
    A.cash = sm.carry[C_CASH][lane];
    A.pav = sm.carry[C_PAV][lane]; A.pml = sm.carry[C_PML][lane]; A.pbm = sm.carry[C_PBM][lane];
    A.pse = sm.carry[C_PSE][lane];
    A.nav = sm.carry[C_NAV][lane]; A.gsum = sm.carry[C_GSUM][lane]; A.rprod = sm.carry[C_RPROD][lane];
    A.bad_risk = sm.bad_risk[lane];
  }
  int risk[2] = {MDG_RISK_GREEN, MDG_RISK_GREEN};
  double cur_val[2] = {0., 0.}, xr[2] = {1., 1.};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q < cnt) {
      const int i = i0 + q;
      const double* st = &sm.stage[q * ST_ROWS][tid];
      const double prev_val = cur[q] * price[q];  // offpolicy_q.py:141
      risk[q] = gate_and_apply(a, c, A, ord[q], st, SS, N, e, na, i, price[q], cur[q], mep[q], bm[q]);
      const bool done_tx = ord[q].nz && risk[q] == MDG_RISK_GREEN;
      // position value after the tick, fold of the new portfolio value
      const double cv = cur[q] * newp[q];
      cur_val[q] = cv;
      A.nav = (i == 0) ? cv : A.nav + cv;
      A.gsum += fabs(cv);
      if (shaping) {  // agent reward ratio (offpolicy_q.py:152-164): 1 + (cur - (prev + mar_diff)) / prevEq, floor .35
        const double pm = prev_val + (done_tx ? (ord[q].units * st[ST_TP * SS] + st[ST_TC * SS]) : 0.);
        double x = (cv - pm) * inv_prev;
        x += 1;
        x = (x != x) ? x : ((x < .35) ? .35 : x);
        xr[q] = x;
        // reduced reward: sum_j log(x_j) accumulated as the log of a product (each x_j is in [.35, ~1.x] and
        // nA <= 16, so it neither overflows nor underflows; the two differ by ~1e-15, rewards carry the 1e-9 bar)
        A.rprod = (i == 0) ? x : A.rprod * x;
      }
    }
  }
  __syncwarp();
  if (!last) {  // hand the carry to warp g+1
    sm.carry[C_X][lane] = A.X; sm.carry[C_BAL][lane] = A.BAL; sm.carry[C_E][lane] = A.E;
    sm.carry[C_PNL][lane] = A.PNL; sm.carry[C_G][lane] = A.G;
    sm.carry[C_CASH][lane] = A.cash;
    sm.carry[C_PAV][lane] = A.pav; sm.carry[C_PML][lane] = A.pml; sm.carry[C_PBM][lane] = A.pbm;
    sm.carry[C_PSE][lane] = A.pse;
    sm.carry[C_NAV][lane] = A.nav; sm.carry[C_GSUM][lane] = A.gsum; sm.carry[C_RPROD][lane] = A.rprod;
    sm.bad_risk[lane] = A.bad_risk;
    bar_arrive(g + 1, 64);
  }
  // ---- this warp's state and BrokerResponse rows
  if (valid) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q < cnt) {
        const int64_t o = (int64_t)(i0 + q) * N + e;
        const double* st = &sm.stage[q * ST_ROWS][tid];
        const bool done_tx = ord[q].nz && risk[q] == MDG_RISK_GREEN;
        if (done_tx) {
          S.ledger[o] = cur[q];
          S.mean_entry[o] = mep[q];
          S.borrowed[o] = bm[q];
        }
        S.price[o] = newp[q];
        if (mode != MDG_MODE_HOLD) {
          a.IO.trans_price[o] = done_tx ? st[ST_TP * SS] : 0.;
          a.IO.trans_units[o] = done_tx ? ord[q].units : 0.;
          a.IO.trans_cost[o] = done_tx ? st[ST_TC * SS] : 0.;
          a.IO.risk[o] = (uint8_t)risk[q];
        }
      }
    }
  }

  // ---- tail (last warp: it holds the complete folds): equity, reward, done (Env.h:192-198, 211-223, 237-249)
  const int ra = a.R.reduce_rewards ? 1 : na;
  int len_after = 0, n_popped = 0;
  if (last) {
    const double cash = A.cash, nav = A.nav, pml = A.pml, pbm = A.pbm, pse = A.pse, maintM = c.maintM;
    const double prevEq = sm.prev_eq[lane];
    const double currentEq = cash + nav - pbm;
    const double clampv = (mode == MDG_MODE_SINGLE) ? 0.01 : 0.3;
    const bool mc = margin_call(cash, nav, pml, pbm, pse, maintM);
    bool done = mc || (currentEq < 0.1 * P.init_cash);
    if (mode != MDG_MODE_HOLD) done = done || (A.bad_risk != 0);
    // State.portfolio row = ledgerNormedFull (Portfolio.cpp:150-155).  Observations carry a 1e-9 bar
    // (not bit-exactness): the nA+1 divisions by equity are one reciprocal and nA+1 multiplies.
    const double inv_eq = 1. / currentEq;
    const double w0 = (cash - pbm) * inv_eq;
    sm.inv_eq[lane] = inv_eq;
    sm.w0[lane] = w0;
    sm.done[lane] = done ? 1 : 0;
    if (valid) {
      // BrokerResponse.marginCall (Broker.cpp:135,156): Portfolio::checkRisk() after the last transaction
      if (mode != MDG_MODE_HOLD) a.IO.margin_call[e] = margin_call(cash, A.pav, pml, pbm, pse, maintM) ? 1 : 0;
      S.cash[e] = cash;
      S.timestamp[e] = ts + 1;
      S.folds[(int64_t)MDG_FOLD_AV * N + e] = nav;
      S.folds[(int64_t)MDG_FOLD_ML * N + e] = pml;
      S.folds[(int64_t)MDG_FOLD_BM * N + e] = pbm;
      S.folds[(int64_t)MDG_FOLD_SE * N + e] = pse;
      S.folds[(int64_t)MDG_FOLD_G * N + e] = fabs(cash) + A.gsum;
      a.IO.reward[e] = fast_log(dmax(currentEq / prevEq, clampv));
      a.IO.done[e] = done ? 1 : 0;
      a.IO.obs_port[((int64_t)head * (na + 1)) * N + e] = w0;
    }
    if (a.done_count) {  // auto-reset: finished envs append themselves to the reset list (warp-aggregated)
      const bool app = done && valid;
      const unsigned ballot = __ballot_sync(0xffffffffu, app);
      if (ballot) {
        const int leader_lane = __ffs(ballot) - 1;
        int base = 0;
        if (lane == leader_lane) base = atomicAdd(a.done_count, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader_lane);
        if (app) a.done_list[base + __popc(ballot & ((1u << lane) - 1u))] = (int)e;
      }
    }
    if (reduce && !cosine) {  // reduced agent reward -> n-step shaper, right here
      const double rsum = fast_log(A.rprod);
      const int len_before = (a.R.nstep > 1) ? S.nstep_len[e] : 0;
      if (valid) a.IO.agent_reward[e] = rsum;
      shaper_add(a, e, valid, 0, 1, rsum, done, len_before, len_after, n_popped);
      if (valid) {
        if (a.R.nstep > 1) S.nstep_len[e] = len_after;
        a.IO.n_popped[e] = n_popped;
      }
    }
  }
  __syncthreads();

  // ---- every warp: its assets' ledgerNormedFull weights; per-asset agent rewards
  const double inv_eq = sm.inv_eq[lane];
  double w[2] = {0., 0.};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q < cnt) {
      w[q] = cur_val[q] * inv_eq;
      if (valid) a.IO.obs_port[((int64_t)head * (na + 1) + i0 + q + 1) * N + e] = w[q];
    }
  }
  if (!shaping || (reduce && !cosine)) return;
  const bool done = sm.done[lane] != 0;
  double extra = 0.;
  if (cosine) {  // the PPC term needs the whole portfolio row (nstep_buffer.py:173-191)
    double pp = 0., pq = 0.;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q < cnt) {
        const double dj = a.R.desired_portfolio[i0 + q + 1];
        pp = (q == 0) ? w[q] * w[q] : pp + w[q] * w[q];
        pq = (q == 0) ? w[q] * dj : pq + w[q] * dj;
      }
    }
    sm.cosp[0][g][lane] = pp;
    sm.cosp[1][g][lane] = pq;
    __syncthreads();
    if (reduce && !last) return;
    const double w0 = sm.w0[lane], d0 = a.R.desired_portfolio[0];
    double spp = w0 * w0, sqq = d0 * d0, spq = w0 * d0;
    for (int h = 0; h < ng; ++h) {
      spp = spp + sm.cosp[0][h][lane];
      spq = spq + sm.cosp[1][h][lane];
    }
    for (int j = 0; j < na; ++j) {
      const double dj = a.R.desired_portfolio[j + 1];
      sqq = sqq + dj * dj;
    }
    extra = a.R.cosine_temp * (spq / (sqrt(spp) * sqrt(sqq)));
  }
  const int len_before = (a.R.nstep > 1) ? S.nstep_len[e] : 0;
  if (reduce) {  // cosine, reduced: last warp only
    const double rsum = fast_log(A.rprod);
    if (valid) a.IO.agent_reward[e] = rsum;
    shaper_add(a, e, valid, 0, 1, rsum + extra, done, len_before, len_after, n_popped);
  } else {  // per-asset agent rewards (offpolicy_q.py:152-164), one n-step buffer per asset
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q < cnt) {
        const double r = fast_log(xr[q]);
        if (valid) a.IO.agent_reward[(int64_t)(i0 + q) * N + e] = r;
        shaper_add(a, e, valid, i0 + q, ra, r + extra, done, len_before, len_after, n_popped);
      }
    }
  }
  if (!reduce) __syncthreads();  // per-asset: every warp has read nstep_len[e] before the last warp replaces it
  if (last && valid) {
    if (a.R.nstep > 1) S.nstep_len[e] = len_after;
    a.IO.n_popped[e] = n_popped;
  }
}

template <bool PAIRS, int MAXW, int MINB, bool ACTIONS>
__global__ void __launch_bounds__(MAXW * 32, MINB) step_kernel(const __grid_constant__ StepArgs a) {
  step_body<PAIRS, ACTIONS, MAXW>(a);
}

// host side: is every asset part of an OUPair laid out (role0, role1) with in-order noise slots?
static inline bool all_ou_pairs(const MdgParams& P) {
  if (P.n_assets % 2) return false;
  for (int i = 0; i < P.n_assets; i += 2) {
    const MdgAssetGen &g0 = P.gen[i], &g1 = P.gen[i + 1];
    const int s = 3 * (i / 2);
    if (g0.type != MDG_GEN_OUPAIR || g1.type != MDG_GEN_OUPAIR || g0.role != 0 || g1.role != 1 ||
        g1.partner != i || g0.gslot < 0 || g0.nslot_aux != s || g0.nslot != s + 1 || g1.nslot != s + 2)
      return false;
  }
  return P.n_normals == 3 * (P.n_assets / 2);
}

// generator groups of a parameter set: an OUPair (role 0 followed by its role 1) is one group, every other asset its own
static inline int fill_groups(const MdgParams& P, int8_t* leader) {
  int n = 0;
  for (int i = 0; i < P.n_assets; ++i)
    if (!(P.gen[i].type == MDG_GEN_OUPAIR && P.gen[i].role == 1)) leader[n++] = (int8_t)i;
  return n;
}

static inline int launch_step(StepArgs& a) {
  const int64_t N = a.L.n_envs;
  cudaStream_t st = (cudaStream_t)a.L.stream;
  const bool pairs = all_ou_pairs(a.P);
  a.n_groups = fill_groups(a.P, a.leader);
  if (pairs && a.n_groups != a.P.n_assets / 2) return set_err(MDG_E_INVALID, "inconsistent OUPair layout");
  for (int i = 0; i < a.P.n_assets; ++i) {  // a role-1 asset must directly follow its role-0 partner
    const MdgAssetGen& g = a.P.gen[i];
    if (g.type == MDG_GEN_OUPAIR && g.role == 1 &&
        (i == 0 || a.P.gen[i - 1].type != MDG_GEN_OUPAIR || a.P.gen[i - 1].role != 0))
      return set_err(MDG_E_INVALID, "OUPair assets must be adjacent (role 0, role 1)");
  }
  const unsigned grid = (unsigned)((N + kTile - 1) / kTile);
  const unsigned block = 32u * (unsigned)a.n_groups;
  const bool acts = a.L.mode == MDG_MODE_MULTI && a.IO.actions;
#define MDG_LAUNCH(PAIRS_, MAXW_, MINB_, ACT_)                                                                   \
  do {                                                                                                           \
    static const cudaError_t attr_ = cudaFuncSetAttribute(step_kernel<PAIRS_, MAXW_, MINB_, ACT_>,               \
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                                          (int)sizeof(StepSmem<MAXW_>));                         \
    if (attr_ != cudaSuccess) return cuda_err(attr_, "mdg_step shared-memory opt-in");                           \
    step_kernel<PAIRS_, MAXW_, MINB_, ACT_><<<grid, block, sizeof(StepSmem<MAXW_>), st>>>(a);                    \
  } while (0)
  if (pairs) { if (acts) MDG_LAUNCH(true, 8, MDG_MINB_PAIRS, true); else MDG_LAUNCH(true, 8, MDG_MINB_PAIRS, false); }
  else { if (acts) MDG_LAUNCH(false, 16, MDG_MINB_GENERIC, true); else MDG_LAUNCH(false, 16, MDG_MINB_GENERIC, false); }
#undef MDG_LAUNCH
  return cuda_err(cudaGetLastError(), "mdg_step launch");
}

}  // namespace mdg
