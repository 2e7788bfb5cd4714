// Warp-specialised fused step for the all-OU-pairs workload (the headline configuration), sm_100a.
//
// Why: the streaming kernel (mdg_step_kernel.cuh) runs one thread per env, and at 65,536 envs that is
// only ~14 warps per SM.  Its profile (profiles/step_kernel_r1e.txt) shows ~35 % of the instructions in
// Philox + Box-Muller and reward logs -- work with NO serial dependence -- executed inside the one
// per-env dependent chain.  Here a CTA owns 128 envs and 512 threads:
//   * ledger warps (0-3), one thread per env: the genuinely serial part -- risk gate and transaction of
//     asset i see the cash/folds left by assets < i (Broker.cpp:144-158) -- and the per-env folds;
//   * worker warps (4-15): everything per (env, Philox block) / (env, pair) / (env, asset): the normal
//     draws, the OU-pair ticks, state + ring-row stores, portfolio weights and agent-reward logs.
// The two roles exchange per-env rows through shared memory and meet at named barriers; `setmaxnreg`
// moves registers from the workers (40) to the ledger threads (136).  Arithmetic, operation order and
// therefore results are identical to the streaming kernel's.
#pragma once
#include "mdg_step_kernel.cuh"

namespace mdg {

constexpr int kPE = 128;                 // envs per CTA
constexpr int kPT = 512;                 // threads per CTA
constexpr int kPW = kPT - kPE;           // worker threads
constexpr int kPairsMaxAssets = MDG_MAX_ASSETS;

__device__ __forceinline__ void bar_all(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kPT) : "memory"); }

// shared-memory rows, each [kPE] doubles
struct PairsSmem {
  double* z;    // [3*np]   normal draws of this tick
  double* cur;  // [na]     final ledger units
  double* px;   // [na]     old price
  double* pm;   // [na]     prev value + mar_diff, later the agent reward r_j
  double* cv;   // [na]     position value after the tick, later the portfolio weight
  double* env;  // [4]      inv_eq, inv_prev, mean scratch...
};
__host__ __device__ constexpr size_t pairs_smem_bytes(int na) {
  return sizeof(double) * (size_t)kPE * (size_t)(3 * (na / 2) + 4 * na + 4);
}

__global__ void __launch_bounds__(kPT, 2) step_pairs_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ double smem_raw[];
  const MdgParams& P = a.P;
  const MdgState& S = a.S;
  const int64_t N = a.L.n_envs;
  const int na = P.n_assets, np = na >> 1, nn = 3 * np;
  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * kPE;
  const int mode = a.L.mode;
  const bool shaping = (a.R.shaper != MDG_SHAPER_OFF) && (mode != MDG_MODE_HOLD);
  const bool cosine = shaping && a.R.shaper == MDG_SHAPER_COSINE;
  const int head = a.L.head;
  PairsSmem sm;
  sm.z = smem_raw;
  sm.cur = sm.z + (size_t)nn * kPE;
  sm.px = sm.cur + (size_t)na * kPE;
  sm.pm = sm.px + (size_t)na * kPE;
  sm.cv = sm.pm + (size_t)na * kPE;
  sm.env = sm.cv + (size_t)na * kPE;

  if (tid < kPE) {
    // =========================== ledger role: one thread per env ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int el = tid;
    const int64_t e = e0 + el;
    const bool live = e < N;
    StepAcc A;
    StepConsts c;
    c.reqM = P.required_margin;
    c.maintM = P.maintenance_margin;
    c.reqM_ok = c.reqM > 0. && c.reqM <= 1e6;
    c.band_scale = 1e-9 * (1. + fabs(c.maintM)) * (c.reqM > 1. ? c.reqM : 1.);
    int64_t ts = 0;
    double prevEq = 1.;
    A.cash = 0.; A.rAV = A.rML = A.rBM = A.rSE = A.G = 0.;
    A.pav = A.pml = A.pbm = A.pse = A.nav = A.gsum = 0.;
    A.bad_risk = false;
    if (live) {
      const double* urow = a.IO.units ? a.IO.units + (mode == MDG_MODE_MULTI ? e * na : e) : nullptr;
      A.cash = S.cash[e];
      ts = S.timestamp[e];
      A.rAV = S.folds[(int64_t)MDG_FOLD_AV * N + e];
      A.rML = S.folds[(int64_t)MDG_FOLD_ML * N + e];
      A.rBM = S.folds[(int64_t)MDG_FOLD_BM * N + e];
      A.rSE = S.folds[(int64_t)MDG_FOLD_SE * N + e];
      A.G = S.folds[(int64_t)MDG_FOLD_G * N + e];
      prevEq = A.cash + A.rAV - A.rBM;  // Env.h:190,208,234
      // ---- phase A: gates and transactions, sequential over assets; next asset's state prefetched
      double n_price = S.price[e], n_cur = S.ledger[e], n_mep = S.mean_entry[e], n_bm = S.borrowed[e];
      double n_units = (mode == MDG_MODE_MULTI) ? urow[0] : 0.;
#pragma unroll 1
      for (int i = 0; i < na; ++i) {
        const double price = n_price;
        double cur = n_cur, mep = n_mep, bm = n_bm, units = n_units;
        if (i + 1 < na) {
          n_price = S.price[(int64_t)(i + 1) * N + e];
          n_cur = S.ledger[(int64_t)(i + 1) * N + e];
          n_mep = S.mean_entry[(int64_t)(i + 1) * N + e];
          n_bm = S.borrowed[(int64_t)(i + 1) * N + e];
          n_units = (mode == MDG_MODE_MULTI) ? urow[i + 1] : 0.;
        }
        if (mode == MDG_MODE_SINGLE) units = (i == a.L.asset_idx) ? urow[0] : 0.;
        double tp, tu, tc, prev_val;
        int risk;
        tx_asset(a, c, A, N, e, na, i, price, cur, mep, bm, units, tp, tu, tc, risk, prev_val);
        sm.cur[i * kPE + el] = cur;
        sm.px[i * kPE + el] = price;
        sm.pm[i * kPE + el] = prev_val + (tu * tp + tc);  // prev_val + mar_diff, offpolicy_q.py:153-156
      }
      // BrokerResponse.marginCall (Broker.cpp:135,156): Portfolio::checkRisk() after the last transaction
      if (mode != MDG_MODE_HOLD)
        a.IO.margin_call[e] = margin_call(A.cash, A.pav, A.pml, A.pbm, A.pse, c.maintM) ? 1 : 0;
      S.cash[e] = A.cash;
    }
    bar_all(1);  // ledger rows ready <-> normals ready (the workers read timestamp[e] before this barrier)
    if (live) S.timestamp[e] = ts + 1;
    bar_all(2);  // workers have ticked: cv rows ready
    bool done = false;
    if (live) {
      // ---- phase C: fold of the new position values, equity, reward, done (Env.h:192-198, 211-223, 237-249)
      double nav = 0., g = 0.;
#pragma unroll 4
      for (int j = 0; j < na; ++j) {
        const double v = sm.cv[j * kPE + el];
        nav = (j == 0) ? v : nav + v;
        g += fabs(v);
      }
      S.folds[(int64_t)MDG_FOLD_AV * N + e] = nav;
      S.folds[(int64_t)MDG_FOLD_ML * N + e] = A.pml;
      S.folds[(int64_t)MDG_FOLD_BM * N + e] = A.pbm;
      S.folds[(int64_t)MDG_FOLD_SE * N + e] = A.pse;
      S.folds[(int64_t)MDG_FOLD_G * N + e] = fabs(A.cash) + A.gsum + g;
      const double currentEq = A.cash + nav - A.pbm;
      const double clampv = (mode == MDG_MODE_SINGLE) ? 0.01 : 0.3;
      a.IO.reward[e] = fast_log(dmax(currentEq / prevEq, clampv));
      const bool mc = margin_call(A.cash, nav, A.pml, A.pbm, A.pse, c.maintM);
      done = mc || (currentEq < 0.1 * P.init_cash);
      if (mode != MDG_MODE_HOLD) done = done || A.bad_risk;
      a.IO.done[e] = done ? 1 : 0;
      const double inv_eq = 1. / currentEq;
      sm.env[0 * kPE + el] = inv_eq;
      sm.env[1 * kPE + el] = 1. / prevEq;
      const double w0 = (A.cash - A.pbm) * inv_eq;  // Portfolio.cpp:150-155
      a.IO.obs_port[((int64_t)head * (na + 1)) * N + e] = w0;
      sm.env[2 * kPE + el] = w0;
    }
    bar_all(3);  // per-env scalars ready
    if (!shaping) return;
    bar_all(4);  // workers: weights in cv, agent rewards in pm
    if (live) {
      // ---- phase E: left-to-right folds of the agent rewards / cosine terms, n-step shaper
      const int ra = a.R.reduce_rewards ? 1 : na;
      const int len_before = (a.R.nstep > 1) ? S.nstep_len[e] : 0;
      int len_after = 0, n_popped = 0;
      double extra = 0.;
      if (cosine) {  // nstep_buffer.py:173-191
        const double w0 = sm.env[2 * kPE + el], d0 = a.R.desired_portfolio[0];
        double pp = w0 * w0, qq = d0 * d0, pq = w0 * d0;
#pragma unroll 1
        for (int j = 0; j < na; ++j) {
          const double w = sm.cv[j * kPE + el], dj = a.R.desired_portfolio[j + 1];
          pp = pp + w * w; qq = qq + dj * dj; pq = pq + w * dj;
        }
        extra = a.R.cosine_temp * (pq / (sqrt(pp) * sqrt(qq)));
      }
      if (a.R.reduce_rewards) {
        double rsum = 0.;
#pragma unroll 4
        for (int j = 0; j < na; ++j) {
          const double r = sm.pm[j * kPE + el];
          rsum = (j == 0) ? r : rsum + r;
        }
        a.IO.agent_reward[e] = rsum;
        shaper_add(a, e, 0, 1, cosine ? rsum + extra : rsum, done, len_before, len_after, n_popped);
      } else {
#pragma unroll 1
        for (int j = 0; j < na; ++j) {
          const double r = sm.pm[j * kPE + el];
          shaper_add(a, e, j, ra, cosine ? r + extra : r, done, len_before, len_after, n_popped);
        }
      }
      if (a.R.nstep > 1) S.nstep_len[e] = len_after;
      a.IO.n_popped[e] = n_popped;
    }
  } else {
    // =========================== worker role: lanes over envs, tasks over rows ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int w = tid - kPE;
    const uint32_t k0 = (uint32_t)a.L.seed, k1 = (uint32_t)(a.L.seed >> 32);
    // ---- phase 0: this tick's normals, one task per (Philox block, env)
    if (a.IO.normals) {
      for (int task = w; task < nn * kPE; task += kPW) {
        const int s = task / kPE, el = task - s * kPE;
        const int64_t e = e0 + el;
        if (e < N) sm.z[s * kPE + el] = a.IO.normals[(int64_t)s * N + e];
      }
    } else {
      const int nb = (nn + 1) >> 1;
      for (int task = w; task < nb * kPE; task += kPW) {
        const int b = task / kPE, el = task - b * kPE;
        const int64_t e = e0 + el;
        if (e < N) {
          const unsigned long long tick = (unsigned long long)S.timestamp[e];
          uint64_t x0, x1;
          philox4x32_10((uint32_t)(a.L.env_offset + e), (uint32_t)b, (uint32_t)tick, (uint32_t)(tick >> 32), k0, k1,
                        x0, x1);
          const double u1 = ((double)(x0 >> 12) + 0.5) * 0x1.0p-52;
          const double u2 = (double)(x1 >> 11) * 0x1.0p-53;
          const double r = sqrt(-2.0 * fast_log_pos(u1));
          double sn, cs;
          fast_sincos_2pi(u2, sn, cs);
          sm.z[(2 * b) * kPE + el] = r * cs;
          if (2 * b + 1 < nn) sm.z[(2 * b + 1) * kPE + el] = r * sn;
        }
      }
    }
    bar_all(1);
    // ---- phase B: OUPair::getData (DataSource.cpp:1232-1240, draw order rw, x0, x1), one task per (pair, env)
    for (int task = w; task < np * kPE; task += kPW) {
      const int p = task / kPE, el = task - p * kPE;
      const int64_t e = e0 + el;
      if (e < N) {
        const MdgAssetGen& g0 = P.gen[2 * p];
        const MdgAssetGen& g1 = P.gen[2 * p + 1];
        double* mrow = S.gstate + (int64_t)g0.gslot * N + e;
        double m = *mrow;
        m += m * (sm.z[g0.nslot_aux * kPE + el] * g0.p[2]);
        *mrow = m;
        const double p0 = sm.px[(2 * p) * kPE + el], p1 = sm.px[(2 * p + 1) * kPE + el];
        const double newp0 = p0 + ((g0.p[0] * (m - p0)) + m * (sm.z[g0.nslot * kPE + el] * g0.p[1]));
        const double newp1 = p1 + ((g1.p[0] * (m - p1)) + m * (sm.z[g1.nslot * kPE + el] * g1.p[1]));
        S.price[(int64_t)(2 * p) * N + e] = newp0;
        S.price[(int64_t)(2 * p + 1) * N + e] = newp1;
        a.IO.obs_price[((int64_t)head * na + 2 * p) * N + e] = newp0;  // State.price row (Env.h:202,228,254)
        a.IO.obs_price[((int64_t)head * na + 2 * p + 1) * N + e] = newp1;
        sm.cv[(2 * p) * kPE + el] = sm.cur[(2 * p) * kPE + el] * newp0;
        sm.cv[(2 * p + 1) * kPE + el] = sm.cur[(2 * p + 1) * kPE + el] * newp1;
      }
    }
    bar_all(2);
    bar_all(3);
    // ---- phase D: portfolio weights (one reciprocal per env, 1e-9 bar) and agent rewards, per (asset, env)
    for (int task = w; task < na * kPE; task += kPW) {
      const int j = task / kPE, el = task - j * kPE;
      const int64_t e = e0 + el;
      if (e < N) {
        const double cur_val = sm.cv[j * kPE + el];
        const double wgt = cur_val * sm.env[0 * kPE + el];
        a.IO.obs_port[((int64_t)head * (na + 1) + j + 1) * N + e] = wgt;
        if (shaping) {
          double x = (cur_val - sm.pm[j * kPE + el]) * sm.env[1 * kPE + el];
          x += 1;
          const double r = fast_log((x != x) ? x : ((x < .35) ? .35 : x));  // offpolicy_q.py:156-161
          sm.pm[j * kPE + el] = r;
          sm.cv[j * kPE + el] = wgt;
          if (!a.R.reduce_rewards) a.IO.agent_reward[(int64_t)j * N + e] = r;
        }
      }
    }
    if (shaping) bar_all(4);
  }
}

static inline int launch_step_pairs(const StepArgs& a) {
  const int64_t N = a.L.n_envs;
  const unsigned grid = (unsigned)((N + kPE - 1) / kPE);
  const size_t smem = pairs_smem_bytes(a.P.n_assets);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(step_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)pairs_smem_bytes(kPairsMaxAssets));
    if (ce != cudaSuccess) return cuda_err(ce, "mdg_step smem attribute");
    attr_done = true;
  }
  step_pairs_kernel<<<grid, kPT, smem, (cudaStream_t)a.L.stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_step launch");
}

}  // namespace mdg
