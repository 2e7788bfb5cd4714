// Fused Env.step kernel for N independent environments (sm_100a).
//
// One launch does what the reference does across Env.h:189-256, Broker.cpp:124-178, Portfolio.cpp:140-323,
// the DataSource.cpp getData family, offpolicy_q.py:140-164 and nstep_buffer.py:
//   transact (sequential over assets, risk-gated) -> generator tick -> equity, reward, done ->
//   newest observation-ring row -> agent reward -> shaped reward -> (auto-reset) done list.
//
// Work decomposition ("a tile of 32 envs per block; warp = generator group, lane = env"):
//   * a block owns a tile of 32 envs; warp g owns generator group g of those envs (an OU pair, or a single asset of
//     a Composite).  State tensors are [rows][N], so every load/store of a warp is one contiguous 256-byte run and
//     every value is read once.
//   * phase A, all warps in parallel -- everything that does not depend on earlier trades: state loads, Philox +
//     Box-Muller (each normal generated once per block into shared memory), the generator tick, the observation-ring
//     store, and the PLAN of every order: the outcome Broker::handleTransaction would produce if its risk gate says
//     green, as the handful of numbers the gate chain needs (amount, deltas of the running sums, cash updates).
//   * phase B, warp 0, lane = env -- the only inherently sequential part (cash and the risk gates of asset i see the
//     trades of assets < i, Broker.cpp:144-158): a loop over the assets that reads the plans from shared memory and
//     decides each gate from running sums: ~25 instructions per asset.
//   * phase C, all warps -- apply the decisions: final ledger rows, BrokerResponse rows, and each asset's terms of the
//     portfolio folds into shared memory.
//   * phase D, warp 0 -- the exact left-to-right folds (16 adds each), equity, reward, done, ledgerNormedFull row,
//     the n-step shaped reward, the auto-reset list.
// Compared with one thread per env streaming over 16 assets (round 1: 13.8 warps per SM at 65,536 envs, one
// 8 k-instruction dependent chain per thread, latency-bound at 0.47 of the HBM roofline) this gives 8x the threads,
// each with ~1/7 of the chain, and the same total instruction count.
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
#define MDG_MINB_PAIRS 3    // 9 warps x 3 blocks: 72 registers
#endif
#ifndef MDG_MINB_GENERIC
#define MDG_MINB_GENERIC 1  // 17 warps
#endif

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

struct StepConsts {
  double reqM, maintM, band_scale;
  double g1, g2;  // magnitude-bound coefficients of one transaction
  bool reqM_ok;
  bool force_exact;  // MdgLaunch.flags & MDG_FLAG_FORCE_EXACT_GATE: every gate takes the exact (cold) path
};

__device__ __forceinline__ StepConsts make_consts(const StepArgs& a) {
  const MdgParams& P = a.P;
  StepConsts c;
  c.reqM = P.required_margin;
  c.maintM = P.maintenance_margin;
  c.reqM_ok = c.reqM > 0. && c.reqM <= 1e6;
  c.band_scale = 1e-9 * (1. + fabs(c.maintM)) * (c.reqM > 1. ? c.reqM : 1.);
  c.g1 = 6. + 3. * fabs(P.slippage_rel) + fabs(P.tcost_rel);
  c.g2 = 3. * fabs(P.slippage_abs);
  c.force_exact = (a.L.flags & MDG_FLAG_FORCE_EXACT_GATE) != 0;
  return c;
}

// One order, everything that does NOT depend on the trades of earlier assets: the outcome
// Broker::handleTransaction would produce if the risk gate says green (Broker.cpp:124-142,171-178;
// Portfolio.cpp:284-323, same operations in the same order), the three cash updates of that outcome, and the
// deltas it would add to the decision sums of the gate chain.
enum { PF_NZ = 1, PF_GATED = 2, PF_OPP = 4, PF_F1 = 8, PF_F3 = 16 };
struct Plan {
  double amtR;              // |price * (units [+ ledger])| * requiredMargin: what the gate compares with balance + pnl
  double dX, dBAL, dE, dG;  // deltas of the decision sums
  double c1, c2, c3;        // cash += c1 (reversal through zero), cash -= c2 (margin + cost), cash -= c3 (margin returned)
  double ncur, nmep, nbm;   // ledger, mean entry price, borrowed margin after the order
  unsigned flags;           // PF_*
};

// Broker::applySlippage / getTransactionCost  Broker.cpp:171-178
__device__ __forceinline__ void order_price_cost(const MdgParams& P, double price, double units, double& tp,
                                                 double& tc) {
  const double slippage = (price * P.slippage_rel) + P.slippage_abs;
  tp = units < 0 ? (price - slippage) : (price + slippage);
  tc = fabs(units * price) * P.tcost_rel + P.tcost_abs;
}

__device__ __forceinline__ void plan_order(const MdgParams& P, const StepConsts& c, double price, double cur,
                                           double mep, double bm, double units, Plan& o) {
  const double units_req = units;
  const bool nz = units != 0.;  // Broker.cpp:126 (NaN units do enter, as in the reference)
  const bool opposite = (signbit(units) != 0) != (signbit(cur) != 0);
  const bool gated = nz && (!opposite || units > -1 * cur);  // Portfolio.cpp:257-258: only these orders are gated
  o.amtR = fabs(price * (opposite ? units + cur : units)) * c.reqM;
  double transactionPrice, transactionCost;
  order_price_cost(P, price, units, transactionPrice, transactionCost);
  // Portfolio::handleTransaction  Portfolio.cpp:284-323
  const double prev_val = cur * price;
  const double o_ml = mep * cur, o_bm = bm;
  const double o_se = (cur < 0.) ? o_ml : 0.;
  double c1 = 0., c3 = 0.;
  bool f1 = false, f3 = false;
  if (opposite) {
    if (fabs(units) > fabs(cur)) {
      units += cur;
      f1 = true; c1 = cur * transactionPrice;
      cur = 0.;
      mep = transactionPrice;
    }
  } else if (nz) {  // (a zero order on a flat position would divide 0 by 0; it never executes)
    mep += (transactionPrice - mep) * (units / (units + cur));
  }
  const double amount = transactionPrice * units;
  const double marginToUse = amount * c.reqM;
  const double marginToBorrow = amount - marginToUse;
  bm += marginToBorrow;
  const double c2 = marginToUse + transactionCost;
  cur += units;
  if (fabs(cur) < 0.000001) {
    mep = 0.;
    if (bm > 0.) { f3 = true; c3 = bm; bm = 0.; }
  }
  if (bm < 0.) { f3 = true; c3 = bm; bm = 0.; }
  o.c1 = c1; o.c2 = c2; o.c3 = c3;
  o.ncur = cur; o.nmep = mep; o.nbm = bm;
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
  // that is |units| * (|price| * g1 + g2) (+ |tc_abs|, added once per asset at the start of the chain)
  o.dG = fabs(units_req) * (fabs(price) * c.g1 + c.g2);
  o.flags = (nz ? PF_NZ : 0) | (gated ? PF_GATED : 0) | (opposite ? PF_OPP : 0) | (f1 ? PF_F1 : 0) | (f3 ? PF_F3 : 0);
}

// dqn.py:165-178: centred action times (unit_size * availableMargin / price); action 0 closes an open position
__device__ __forceinline__ double action_units(int act, int half, double scale, double price, double cur) {
  if (act == 0) return (cur != 0.) ? -cur : 0.;
  return (double)(act - half) * (scale / price);
}

// rows of the per-asset shared-memory records (each row = 32 lanes): the plan (phase A -> B), then, in the same
// storage, the asset's terms of the folds (phase C -> D)
enum { PR_AMTR, PR_DX, PR_DBAL, PR_DE, PR_DG, PR_C1, PR_C2, PR_C3, PR_ROWS };
enum { FR_AV, FR_ML, FR_BM, FR_SE, FR_CV, FR_XR };
// the folds of phase D: asset value and mean-entry value, borrowed margin, short entry value (old prices), asset
// value at the new prices, magnitude bound, product of the agent-reward ratios
enum { FD_AV, FD_ML, FD_BM, FD_SE, FD_NAV, FD_G, FD_RPROD, FD_ROWS };
// rows of the per-thread staging columns: the planned ledger values wait here for the gate's decision
enum { ST_NCUR, ST_NMEP, ST_NBM, ST_ROWS };

// COLD: the units of order j as phase A derived them (exact_gate re-plans executed orders)
static __device__ __noinline__ double order_units_cold(const StepArgs& a, int64_t e, int j, double act_scale,
                                                       double price, double cur) {
  const int na = a.P.n_assets;
  if (a.L.mode == MDG_MODE_MULTI) {
    if (a.IO.actions) return action_units(a.IO.actions[e * na + j], a.L.action_atoms / 2, act_scale, price, cur);
    return a.IO.units[e * na + j];
  }
  if (a.L.mode == MDG_MODE_SINGLE) return (j == a.L.asset_idx) ? a.IO.units[e] : 0.;
  return 0.;
}

// Exact risk gate of asset i (Portfolio::checkRisk(i, units), Portfolio.cpp:254-279): every accounting quantity is
// a fresh left-to-right fold over the assets, exactly as in the oracle.  COLD path, called from the gate chain
// (phase B) only when the cheap bound cannot decide.  Global memory still holds the portfolio as it was before the
// step (phase C stores the new rows); the orders of assets < i that executed (risk[j] green) are re-planned with
// the same arithmetic as phase A, so the values are the ones phase C will store.  `cash` is the exact cash after
// the trades of assets < i.
static __device__ __noinline__ int exact_gate(const StepArgs& a, int64_t e, int i, double cash, double act_scale,
                                              const unsigned char* flags_col, const unsigned char* risk_col) {
  const StepConsts c = make_consts(a);  // (cold: rebuilt here so that the callers' copy can stay in registers)
  const MdgState& S = a.S;
  const int64_t N = a.L.n_envs;
  const int na = a.P.n_assets;
  double av = 0., ml = 0., bms = 0., se = 0.;
#pragma unroll 1
  for (int j = 0; j < na; ++j) {
    double l = S.ledger[(int64_t)j * N + e], m = S.mean_entry[(int64_t)j * N + e], b = S.borrowed[(int64_t)j * N + e];
    const double p = S.price[(int64_t)j * N + e];
    if (j < i && (flags_col[j * kTile] & PF_NZ) && risk_col[j * kTile] == MDG_RISK_GREEN) {
      Plan o;
      plan_order(a.P, c, p, l, m, b, order_units_cold(a, e, j, act_scale, p, l), o);
      l = o.ncur; m = o.nmep; b = o.nbm;
    }
    const double t_se = l * (m * (l < 0. ? 1. : 0.));
    if (j == 0) { av = l * p; ml = m * l; bms = b; se = t_se; }
    else { av = av + l * p; ml = ml + m * l; bms = bms + b; se = se + t_se; }
  }
  const double price = S.price[(int64_t)i * N + e], cur = S.ledger[(int64_t)i * N + e];
  const double units = order_units_cold(a, e, i, act_scale, price, cur);
  const bool opposite = (signbit(units) != 0) != (signbit(cur) != 0);
  const double pnl = av - ml;        // :184-186
  const double balance = cash + se;  // :192-197
  const double availableMargin = (balance + pnl) / c.reqM;  // :229-231
  if (opposite) {
    const double excess = units + cur;
    if (availableMargin <= fabs(price * excess) || balance <= 0.) return MDG_RISK_INSUFF_MARGIN;
    return MDG_RISK_GREEN;
  }
  if (margin_call(cash, av, ml, bms, se, c.maintM)) return MDG_RISK_MARGIN_CALL;
  if (availableMargin <= fabs(price * units) || balance <= 0.) return MDG_RISK_INSUFF_MARGIN;
  return MDG_RISK_GREEN;
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

// COLD continuation of the gate chain from asset i0 on, for a lane whose fast gate could not decide: the same
// loop with exact_gate behind every uncertain gate.  Returns the exact cash after the last asset.
static __device__ __noinline__ double chain_cold(const StepArgs& a, int64_t e, int i0, double cash, double cX,
                                                 double cBAL, double cE, double cG, double act_scale,
                                                 const double* rec_col, const unsigned char* flags_col,
                                                 unsigned char* risk_col, int* bad_risk /* shared memory */) {
  const StepConsts c = make_consts(a);
  const int na = a.P.n_assets;
#pragma unroll 1
  for (int i = i0; i < na; ++i) {
    const unsigned f = flags_col[i * kTile];
    const double* r = rec_col + (size_t)i * PR_ROWS * kTile;
    int risk = MDG_RISK_GREEN;
    if (f & PF_NZ) {
      if (f & PF_GATED) {
        const double band = c.band_scale * (cG + r[PR_DG * kTile]);
        const double d1 = cX - r[PR_AMTR * kTile];
        bool certain = c.reqM_ok && fabs(d1) > band && fabs(cBAL) > band;
        int rf = (d1 <= 0. || cBAL <= 0.) ? MDG_RISK_INSUFF_MARGIN : MDG_RISK_GREEN;
        if (!(f & PF_OPP)) {
          const double m = c.maintM * (cX - cBAL);
          const double d3 = cE + m, d4 = cX + m;
          certain = certain && fabs(d3) > band && fabs(d4) > band;
          if (d3 <= 0. || d4 <= 0.) rf = MDG_RISK_MARGIN_CALL;
        }
        risk = (certain && !c.force_exact) ? rf : exact_gate(a, e, i, cash, act_scale, flags_col, risk_col);
      }
      if (risk == MDG_RISK_GREEN) {
        cX += r[PR_DX * kTile]; cBAL += r[PR_DBAL * kTile]; cE += r[PR_DE * kTile]; cG += r[PR_DG * kTile];
        if (f & PF_F1) cash += r[PR_C1 * kTile];  // Portfolio.cpp:296
        cash -= r[PR_C2 * kTile];                 // :311
        if (f & PF_F3) cash -= r[PR_C3 * kTile];  // :316-321
      } else if (risk != MDG_RISK_INSUFF_MARGIN) {
        *bad_risk = 1;
      }
    }
    risk_col[i * kTile] = (unsigned char)risk;
  }
  return cash;
}

template <int MAXW>
struct StepSmem {
  union {
    double rec[MDG_MAX_ASSETS][PR_ROWS][kTile];  // per-asset records: plans (A -> B), then fold terms (C -> D) ...
    double cosp[2][MDG_MAX_ASSETS][kTile];       // ... then the PPC shaper's per-group partial sums (phase E)
  };
  double z[kMaxNormals][kTile];        // this step's normals
  double stage[2 * ST_ROWS][MAXW * 32];
  double fold[FD_ROWS][kTile];               // the exact folds, one per group warp (phase D)
  double inv_prev[kTile], act_scale[kTile];  // written by the chain warp before the first barrier
  double cash[kTile];                        // exact cash after the last transaction (chain warp, end of phase B)
  double w0[kTile];                          // written by the chain warp in phase D
  int done[kTile];
  unsigned char flags[MDG_MAX_ASSETS][kTile];
  unsigned char risk[MDG_MAX_ASSETS][kTile];
};

// every thread of the block (group warps + the chain warp) meets at barrier 0
__device__ __forceinline__ void tile_sync() { __syncthreads(); }

// ---------------------------------------------------------------------------
// The chain warp (the block's last warp, lane = env): the sequential phases B and D.
// ---------------------------------------------------------------------------
template <bool ACTIONS, int MAXW>
__device__ __forceinline__ void chain_role(const StepArgs& a, StepSmem<MAXW>& sm, const StepConsts& c) {
  const MdgParams& P = a.P;
  const MdgState& S = a.S;
  const int64_t N = a.L.n_envs;
  const int na = P.n_assets;
  const int lane = threadIdx.x & 31, ng = a.n_groups;
  int64_t e = (int64_t)blockIdx.x * kTile + lane;
  const bool valid = e < N;
  if (!valid) e = N - 1;
  const int mode = a.L.mode;
  const bool shaping = (a.R.shaper != MDG_SHAPER_OFF) && (mode != MDG_MODE_HOLD);
  const bool cosine = shaping && a.R.shaper == MDG_SHAPER_COSINE;
  const bool reduce = shaping && a.R.reduce_rewards;
  const bool fast_tail = !shaping || (reduce && !cosine);
  const bool by_actions = ACTIONS && mode == MDG_MODE_MULTI && a.IO.actions;
  const int head = a.L.head;

  // ---- the chain starts from the state.  The folds of the incoming portfolio are exactly the folds this kernel
  // (or reset/init/refresh) computed at the end of the previous call -- same values, same left-to-right order --
  // so they are carried in state instead of re-reading the whole portfolio before the first transaction.
  // Decision sums (approximate, one precomputed delta per accepted order): X = balance + pnl (= availableMargin *
  // requiredMargin), BAL = balance, E = equity at the old prices, G = bound on the sum of the magnitudes of
  // everything that went into them; cash is exact (every Portfolio::handleTransaction update in the reference's order).
  const int64_t ts = S.timestamp[e];
  double cash = S.cash[e];
  double cX, cBAL, cE, cG;
  {
    const double rAV = S.folds[(int64_t)MDG_FOLD_AV * N + e], rML = S.folds[(int64_t)MDG_FOLD_ML * N + e];
    const double rBM = S.folds[(int64_t)MDG_FOLD_BM * N + e], rSE = S.folds[(int64_t)MDG_FOLD_SE * N + e];
    cG = S.folds[(int64_t)MDG_FOLD_G * N + e] + na * fabs(P.tcost_abs);  // + the absolute cost of up to nA transactions
    cBAL = cash + rSE;        // Portfolio.cpp:192-197
    cX = cBAL + (rAV - rML);  // availableMargin * requiredMargin, :184-186, :229-231
    cE = cash + rAV - rBM;    // equity, :211-213  (= prevEq, Env.h:190,208,234)
  }
  const double prevEq = cE;
  sm.inv_prev[lane] = 1. / cE;
  // DQN.action_to_transaction (dqn.py:160-179) fused in front of the step: units from discrete actions and the
  // availableMargin of the incoming portfolio (Portfolio.cpp:229-231), one scale for every asset
  const double act_scale = by_actions ? a.L.unit_size * (cX / P.required_margin) : 0.;
  if (by_actions) {
    sm.act_scale[lane] = act_scale;
    tile_sync();  // act_scale
  }
  // state of the reduced reward's n-step shaper: loaded here, used at the very end
  const bool moments = fast_tail && reduce && (a.R.shaper == MDG_SHAPER_DSR || a.R.shaper == MDG_SHAPER_DDR);
  const bool inline_shaper = fast_tail && reduce && a.R.nstep == 1 &&
                             (moments || a.R.shaper == MDG_SHAPER_SUM);  // the common case, without a call
  double shA = 0., shB = 0.;
  if (moments && inline_shaper) { shA = S.shaper_A[e]; shB = S.shaper_B[e]; }
  tile_sync();  // #2: plans

  // ---- phase B: the gate chain.  Portfolio::checkRisk(i, units), Portfolio.cpp:254-279, on the decision sums.
  // The gate compares folds over the whole portfolio with thresholds; the decision sums are rounded differently
  // from the reference's fresh left-to-right folds (they differ by < 1e-13 * G), so a decision is taken from them
  // only when it clears its threshold by 1e-9 * G: `certain`.  Otherwise -- a knife-edge, NaN/Inf, a non-positive
  // required margin -- the exact folds decide (exact_gate, cold).  Decisions, and therefore ledgers, are
  // bit-identical to the oracle's either way.  Branch-free apart from the cold call; the next asset's record is
  // loaded while this one is decided.
  int bad_risk = 0;
  sm.done[lane] = 0;  // chain_cold reports a margin call through this slot
  {
    // (the order's own magnitude |amount| is at most dG / 6, so G + dG bounds G + |amount|)
    int i = 0;
#pragma unroll 4
    for (; i < na; ++i) {
      const double* r = &sm.rec[i][0][lane];
      const unsigned f = sm.flags[i][lane];
      const double amtR = r[PR_AMTR * kTile], dX = r[PR_DX * kTile], dBAL = r[PR_DBAL * kTile], dE = r[PR_DE * kTile],
                   dG = r[PR_DG * kTile], c1 = r[PR_C1 * kTile], c2 = r[PR_C2 * kTile], c3 = r[PR_C3 * kTile];
      const bool nz = (f & PF_NZ) != 0, gated = (f & PF_GATED) != 0, opp = (f & PF_OPP) != 0;
      const double nG = cG + dG;
      const double band = c.band_scale * nG;
      const double d1 = cX - amtR;              // availableMargin <= |amount|  <=>  d1 <= 0
      const double m = c.maintM * (cX - cBAL);  // Portfolio::checkRisk() first (:268), :243-252
      const double d3 = cE + m, d4 = cX + m;
      const bool ok1 = (fabs(d1) > band) & (fabs(cBAL) > band);
      const bool ok2 = (fabs(d3) > band) & (fabs(d4) > band);
      const bool certain = c.reqM_ok & ok1 & (opp | ok2) & !c.force_exact;
      if (gated & !certain) break;  // this lane leaves the fast chain: chain_cold continues from asset i
      const bool insuff = (d1 <= 0.) | (cBAL <= 0.);
      const bool mcall = !opp & ((d3 <= 0.) | (d4 <= 0.));
      const bool rej = gated & (insuff | mcall);
      const int risk = rej ? (mcall ? MDG_RISK_MARGIN_CALL : MDG_RISK_INSUFF_MARGIN) : MDG_RISK_GREEN;
      if (nz & !rej) {
        cX += dX; cBAL += dBAL; cE += dE; cG = nG;
        if (f & PF_F1) cash += c1;  // Portfolio.cpp:296
        cash -= c2;                 // :311
        if (f & PF_F3) cash -= c3;  // :316-321
      }
      bad_risk |= (gated & mcall) ? 1 : 0;
      sm.risk[i][lane] = (unsigned char)risk;
    }
    if (i < na)
      cash = chain_cold(a, e, i, cash, cX, cBAL, cE, cG, act_scale, &sm.rec[0][0][lane], &sm.flags[0][lane],
                        &sm.risk[0][lane], &sm.done[lane]);
    sm.cash[lane] = cash;
    bad_risk |= sm.done[lane];
  }
  tile_sync();  // #3: decisions, normals
  tile_sync();  // #4: fold terms

  tile_sync();  // #5: folds

  // ---- phase D: equity, reward, done (Env.h:192-198, 211-223, 237-249) from the exact folds
  int len_after = 0, n_popped = 0;
  const double pav = sm.fold[FD_AV][lane], pml = sm.fold[FD_ML][lane], pbm = sm.fold[FD_BM][lane],
               pse = sm.fold[FD_SE][lane], nav = sm.fold[FD_NAV][lane], gsum = sm.fold[FD_G][lane];
  const double rprod = shaping ? sm.fold[FD_RPROD][lane] : 1.;
  const double maintM = c.maintM;
  const double currentEq = cash + nav - pbm;
  const double clampv = (mode == MDG_MODE_SINGLE) ? 0.01 : 0.3;
  const bool mc = margin_call(cash, nav, pml, pbm, pse, maintM);
  bool done = mc || (currentEq < 0.1 * P.init_cash);
  if (mode != MDG_MODE_HOLD) done = done || (bad_risk != 0);
  // State.portfolio row = ledgerNormedFull (Portfolio.cpp:150-155).  Observations carry a 1e-9 bar
  // (not bit-exactness): the nA+1 divisions by equity are one reciprocal and nA+1 multiplies.
  const double w0 = (cash - pbm) * (1. / currentEq);
  if (!fast_tail) {
    sm.w0[lane] = w0;
    sm.done[lane] = done ? 1 : 0;
  }
  if (valid) {
    // BrokerResponse.marginCall (Broker.cpp:135,156): Portfolio::checkRisk() after the last transaction
    if (mode != MDG_MODE_HOLD) a.IO.margin_call[e] = margin_call(cash, pav, pml, pbm, pse, maintM) ? 1 : 0;
    S.cash[e] = cash;
    S.timestamp[e] = ts + 1;
    S.folds[(int64_t)MDG_FOLD_AV * N + e] = nav;
    S.folds[(int64_t)MDG_FOLD_ML * N + e] = pml;
    S.folds[(int64_t)MDG_FOLD_BM * N + e] = pbm;
    S.folds[(int64_t)MDG_FOLD_SE * N + e] = pse;
    S.folds[(int64_t)MDG_FOLD_G * N + e] = fabs(cash) + gsum;
    a.IO.done[e] = done ? 1 : 0;
    a.IO.obs_port[((int64_t)head * (na + 1)) * N + e] = w0;
    a.IO.reward[e] = fast_log(dmax(currentEq / prevEq, clampv));
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
  if (fast_tail) {
    if (reduce) {  // reduced agent reward -> n-step shaper, right here
      const double rsum = fast_log(rprod);
      if (valid) a.IO.agent_reward[e] = rsum;
      if (inline_shaper) {  // nstep == 1: add, pop at once (shaper_add / shaper_pop with len == n == 1, same arithmetic)
        double sh;
        const double d0 = a.R.discounts[0];
        if (a.R.shaper == MDG_SHAPER_DSR) {         // nstep_buffer.py:62-91
          sh = d0 * dsr_value(shA, shB, rsum);
          sh = sh / 1;
          const double dA = rsum - shA, dB = rsum * rsum - shB;
          shA += a.R.adaptation_rate * dA;
          shB += a.R.adaptation_rate * dB;
          sh = clip(sh, -1., 1.);
        } else if (a.R.shaper == MDG_SHAPER_DDR) {  // :128-162
          sh = d0 * ddr_value(shA, shB, rsum);
          sh = sh / 1;
          const double dA = rsum - shA;
          double mneg = rsum < 0. ? rsum : 0.;
          if (rsum != rsum) mneg = rsum;
          const double dB = mneg * mneg - shB;
          shA += a.R.adaptation_rate * dA;
          shB += a.R.adaptation_rate * dB;
          sh = clip(sh, -1., 1.);
        } else {                                     // sum_default :23-27
          sh = 0.;
          sh = sh + d0 * rsum;
        }
        if (valid) {
          a.IO.shaped_reward[e] = sh;
          a.IO.n_popped[e] = 1;
          if (moments) { S.shaper_A[e] = shA; S.shaper_B[e] = shB; }
        }
      } else {
        const int len_before = (a.R.nstep > 1) ? S.nstep_len[e] : 0;
        shaper_add(a, e, valid, 0, 1, rsum, done, len_before, len_after, n_popped);
        if (valid) {
          if (a.R.nstep > 1) S.nstep_len[e] = len_after;
          a.IO.n_popped[e] = n_popped;
        }
      }
    }
    return;
  }
  tile_sync();  // #6: w0, done; the records are dead (cosp may overwrite them)

  // ---- phase E: the PPC term of a reduced reward; the n-step bookkeeping of per-asset rewards
  const int len_before = (a.R.nstep > 1) ? S.nstep_len[e] : 0;
  if (cosine) tile_sync();  // cosp partial sums
  if (reduce) {             // cosine, reduced
    const double d0 = a.R.desired_portfolio[0];
    double spp = w0 * w0, sqq = d0 * d0, spq = w0 * d0;
    for (int h = 0; h < ng; ++h) {
      spp = spp + sm.cosp[0][h][lane];
      spq = spq + sm.cosp[1][h][lane];
    }
    for (int j = 0; j < na; ++j) {
      const double dj = a.R.desired_portfolio[j + 1];
      sqq = sqq + dj * dj;
    }
    const double extra = a.R.cosine_temp * (spq / (sqrt(spp) * sqrt(sqq)));
    const double rsum = fast_log(rprod);
    if (valid) a.IO.agent_reward[e] = rsum;
    shaper_add(a, e, valid, 0, 1, rsum + extra, done, len_before, len_after, n_popped);
  } else {  // per-asset buffers all move in lockstep (ReplayBuffer.add + pop, nstep_buffer.py:342-361)
    const int n = a.R.nstep;
    int len = len_before + 1, first = 0;
    if (len >= n) { ++first; ++n_popped; }
    if (done) { n_popped += len - first; first = len; }
    len_after = len - first;
    tile_sync();  // every group warp has read nstep_len[e]
  }
  if (valid) {
    if (a.R.nstep > 1) S.nstep_len[e] = len_after;
    a.IO.n_popped[e] = n_popped;
  }
}

// ---------------------------------------------------------------------------
// A group warp (warp g = generator group g, lane = env): the parallel phases A and C.
// ---------------------------------------------------------------------------
template <bool PAIRS, bool ACTIONS, int MAXW>
__device__ __forceinline__ void group_role(const StepArgs& a, StepSmem<MAXW>& sm, const StepConsts& c) {
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
  const bool fast_tail = !shaping || (reduce && !cosine);  // nothing left for the group warps after phase C
  const int head = a.L.head;

  // ---- this warp's group: assets i0 .. i0+cnt-1
  const int i0 = PAIRS ? 2 * g : a.leader[g];
  const int cnt = PAIRS ? 2 : ((P.gen[i0].type == MDG_GEN_OUPAIR) ? 2 : 1);
  double price[2], cur[2], mep[2], bm[2], units[2];
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
  double mean = 0.;  // OUPair: the pair's shared mean
  if (PAIRS) mean = S.gstate[(int64_t)P.gen[i0].gslot * N + e];
  const bool by_actions = ACTIONS && mode == MDG_MODE_MULTI && a.IO.actions;
  int act[2] = {0, 0};
  if (by_actions) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (q < cnt) act[q] = a.IO.actions[e * na + i0 + q];
  } else if (mode == MDG_MODE_MULTI) {
    // the caller's units matrix is (N, nA) env-major (what an agent's network emits): an env's row is one
    // 128-byte line, read 8 or 16 bytes at a time by the warps of the block
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
  if (by_actions) tile_sync();  // act_scale

  // ---- phase A: plan this warp's orders
  constexpr int SS = MAXW * 32;  // row stride of the staging columns
  {
    const int act_half = a.L.action_atoms / 2;
    const double act_scale = by_actions ? sm.act_scale[lane] : 0.;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q < cnt) {
        if (by_actions) units[q] = action_units(act[q], act_half, act_scale, price[q], cur[q]);
        Plan o;
        plan_order(P, c, price[q], cur[q], mep[q], bm[q], units[q], o);
        double* r = &sm.rec[i0 + q][0][lane];
        r[PR_AMTR * kTile] = o.amtR;
        r[PR_DX * kTile] = o.dX; r[PR_DBAL * kTile] = o.dBAL; r[PR_DE * kTile] = o.dE; r[PR_DG * kTile] = o.dG;
        r[PR_C1 * kTile] = o.c1; r[PR_C2 * kTile] = o.c2; r[PR_C3 * kTile] = o.c3;
        sm.flags[i0 + q][lane] = (unsigned char)o.flags;
        double* st = &sm.stage[q * ST_ROWS][tid];
        st[ST_NCUR * SS] = o.ncur; st[ST_NMEP * SS] = o.nmep; st[ST_NBM * SS] = o.nbm;
      }
    }
  }
  tile_sync();  // #2: plans, inv_prev

  // ---- while the chain warp decides the gates: this step's normals, each generated once per block into shared
  // memory (Philox block b by warp b mod ng).  Independent of the ledger: prices never depend on trades.
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
  tile_sync();  // #3: decisions, normals

  // ---- phase C: generator tick
  double newp[2];
  if (PAIRS) {  // OUPair::getData, DataSource.cpp:1232-1240 (draw order rw, x0, x1)
    const MdgAssetGen& g0 = P.gen[i0];
    const MdgAssetGen& g1 = P.gen[i0 + 1];
    double* mrow = S.gstate + (int64_t)g0.gslot * N + e;
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

  // ---- phase C: apply the decisions.  State, observation and BrokerResponse rows; the asset's terms of the folds
  const double inv_prev = sm.inv_prev[lane];
  double cur_val[2] = {0., 0.}, xr[2] = {1., 1.};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q < cnt) {
      const int i = i0 + q;
      const int64_t o = (int64_t)i * N + e;
      const int risk = sm.risk[i][lane];
      const bool exec = (sm.flags[i][lane] & PF_NZ) && risk == MDG_RISK_GREEN;
      const double* st = &sm.stage[q * ST_ROWS][tid];
      const double prev_val = cur[q] * price[q];  // offpolicy_q.py:141
      double tp = 0., tc = 0.;
      if (exec) {
        order_price_cost(P, price[q], units[q], tp, tc);
        cur[q] = st[ST_NCUR * SS]; mep[q] = st[ST_NMEP * SS]; bm[q] = st[ST_NBM * SS];
      }
      if (valid) {
        if (exec) {
          S.ledger[o] = cur[q];
          S.mean_entry[o] = mep[q];
          S.borrowed[o] = bm[q];
        }
        S.price[o] = newp[q];
        a.IO.obs_price[((int64_t)head * na + i) * N + e] = newp[q];  // State.price row (Env.h:202,228,254)
        if (mode != MDG_MODE_HOLD) {
          a.IO.trans_price[o] = tp;
          a.IO.trans_units[o] = exec ? units[q] : 0.;
          a.IO.trans_cost[o] = tc;
          a.IO.risk[o] = (uint8_t)risk;
        }
      }
      // terms of the exact folds of the final ledger, old prices (Portfolio.cpp:180-197,207-209) and new prices
      const double t_ml = mep[q] * cur[q];
      const double t_se = (cur[q] < 0.) ? t_ml : 0. * t_ml;
      const double cv = cur[q] * newp[q];  // position value after the tick
      cur_val[q] = cv;
      double* r = &sm.rec[i][0][lane];
      r[FR_AV * kTile] = cur[q] * price[q];
      r[FR_ML * kTile] = t_ml;
      r[FR_BM * kTile] = bm[q];
      r[FR_SE * kTile] = t_se;
      r[FR_CV * kTile] = cv;
      if (shaping) {  // agent reward ratio (offpolicy_q.py:152-164): 1 + (cur - (prev + mar_diff)) / prevEq, floor .35
        const double pm = prev_val + (exec ? (units[q] * tp + tc) : 0.);
        double x = (cv - pm) * inv_prev;
        x += 1;
        x = (x != x) ? x : ((x < .35) ? .35 : x);
        xr[q] = x;
        r[FR_XR * kTile] = x;
      }
    }
  }
  tile_sync();  // #4: fold terms

  // ---- phase D: the exact left-to-right folds over the assets (Portfolio.cpp:180-197,207-213), one fold per warp
#pragma unroll 1
  for (int k = g; k < FD_ROWS; k += ng) {
    double v;
    if (k == FD_G) {  // bound on the magnitudes that go into the next step's decision sums (any order)
      v = 0.;
#pragma unroll 4
      for (int i = 0; i < na; ++i)
        v += fabs(sm.rec[i][FR_ML][lane]) + fabs(sm.rec[i][FR_BM][lane]) + fabs(sm.rec[i][FR_CV][lane]);
    } else if (k == FD_RPROD) {
      // reduced reward: sum_j log(x_j) accumulated as the log of a product (each x_j is in [.35, ~1.x] and
      // nA <= 16, so it neither overflows nor underflows; the two differ by ~1e-15, rewards carry the 1e-9 bar)
      v = 1.;
      if (shaping) {
        v = sm.rec[0][FR_XR][lane];
#pragma unroll 4
        for (int i = 1; i < na; ++i) v = v * sm.rec[i][FR_XR][lane];
      }
    } else {
      const int row = (k == FD_NAV) ? FR_CV : k;  // FD_AV..FD_SE == FR_AV..FR_SE
      v = sm.rec[0][row][lane];
#pragma unroll 4
      for (int i = 1; i < na; ++i) v = v + sm.rec[i][row][lane];
    }
    sm.fold[k][lane] = v;
  }
  tile_sync();  // #5: folds

  // State.portfolio row = ledgerNormedFull (Portfolio.cpp:150-155).  Observations carry a 1e-9 bar (not
  // bit-exactness): the nA+1 divisions by equity are one reciprocal and nA+1 multiplies.
  const double inv_eq = 1. / (sm.cash[lane] + sm.fold[FD_NAV][lane] - sm.fold[FD_BM][lane]);
  double w[2] = {0., 0.};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q < cnt) {
      w[q] = cur_val[q] * inv_eq;
      if (valid) a.IO.obs_port[((int64_t)head * (na + 1) + i0 + q + 1) * N + e] = w[q];
    }
  }
  if (fast_tail) return;
  tile_sync();  // #6: w0, done; the records are dead (cosp may overwrite them)

  // ---- phase E: per-asset agent rewards, the PPC term
  const int ra = a.R.reduce_rewards ? 1 : na;
  const bool done = sm.done[lane] != 0;
  double extra = 0.;
  const int len_before = (!reduce && a.R.nstep > 1) ? S.nstep_len[e] : 0;
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
    tile_sync();  // cosp partial sums
    if (reduce) return;  // the chain warp finishes a reduced reward
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
  // per-asset agent rewards (offpolicy_q.py:152-164), one n-step buffer per asset
  int len_after = 0, n_popped = 0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q < cnt) {
      const double r = fast_log(xr[q]);
      if (valid) a.IO.agent_reward[(int64_t)(i0 + q) * N + e] = r;
      shaper_add(a, e, valid, i0 + q, ra, r + extra, done, len_before, len_after, n_popped);
    }
  }
  tile_sync();  // every group warp has read nstep_len[e] before the chain warp replaces it
}

template <bool PAIRS, bool ACTIONS, int MAXW>
__device__ __forceinline__ void step_body(const StepArgs& a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<MAXW>& sm = *reinterpret_cast<StepSmem<MAXW>*>(smem_raw);
  const StepConsts c = make_consts(a);
#ifndef MDG_NO_CHAIN
  if ((int)(threadIdx.x >> 5) == a.n_groups) { chain_role<ACTIONS, MAXW>(a, sm, c); return; }
#endif
#ifndef MDG_NO_GROUP
  group_role<PAIRS, ACTIONS, MAXW>(a, sm, c);
#endif
}

template <bool PAIRS, int MAXW, int MINB, bool ACTIONS>
__global__ void __launch_bounds__((MAXW + 1) * 32, MINB) step_kernel(const __grid_constant__ StepArgs a) {
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
  const unsigned block = 32u * (unsigned)(a.n_groups + 1);  // group warps + the chain warp
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
