// extern "C" entry points of the step path + the (rare, non-unrolled) reset / init kernels.
#include <stdlib.h>

#include "mdg_step_kernel.cuh"

namespace mdg {

char* err_buf() {
  thread_local static char buf[512] = "";
  return buf;
}

// ---------------------------------------------------------------------------
// reset: Env::reset (Env.h:181-187) + StackerDiscrete.initialize_history (preprocessor.py:191-194)
// ---------------------------------------------------------------------------
struct ResetArgs {
  MdgParams P;
  MdgState S;
  MdgStepIO IO;
  MdgLaunch L;
  const uint8_t* mask;
  int fill_ticks;
  int clear_nstep;
};

// Envs that reset are rare and scattered, and each must fast-forward fill_ticks generator ticks
// (A17: a reset consumes k ticks).  One thread per env would leave 31 lanes idle for 64 ticks in
// almost every warp, so the block (a) compacts the indices of its resetting envs in shared memory
// and (b) spreads (generator group, env) items over ALL its threads: an OU pair or a single
// asset is an independent stochastic process, and Philox draws are addressed by (env, tick, slot),
// so any lane can compute any group's path.
constexpr int kResetBlock = 256;

__global__ void __launch_bounds__(kResetBlock) reset_kernel(const __grid_constant__ ResetArgs a) {
  __shared__ int s_list[kResetBlock];
  __shared__ long long s_ts[kResetBlock];
  __shared__ int s_leader[MDG_MAX_ASSETS];
  __shared__ int s_count, s_nlead;
  const MdgParams& P = a.P;
  const int64_t N = a.L.n_envs;
  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * kResetBlock;
  const int na = P.n_assets;
  const int k = a.L.window;
  const int fill = a.fill_ticks;
  if (tid == 0) {
    s_count = 0;
    int n = 0;
    for (int i = 0; i < na; ++i)
      if (!(P.gen[i].type == MDG_GEN_OUPAIR && P.gen[i].role == 1)) s_leader[n++] = i;
    s_nlead = n;
  }
  __syncthreads();
  {  // (a) warp-aggregated compaction of the resetting envs of this block
    const int64_t e = e0 + tid;
    const bool flag = (e < N) && (!a.mask || a.mask[e]);
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    int base = 0;
    if ((tid & 31) == 0 && ballot) base = atomicAdd(&s_count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) s_list[base + __popc(ballot & ((1u << (tid & 31)) - 1u))] = tid;
  }
  __syncthreads();
  const int nd = s_count;
  if (nd == 0) return;
  const double cash = P.init_cash;
  const double eq = cash + 0. - 0.;  // flat portfolio: equity == cash
  // per-env scalars: fresh Broker/Account/Portfolio (Env.h:150-165), cash weight and timestamps of the rows
  for (int it = tid; it < nd; it += kResetBlock) {
    const int64_t e = e0 + s_list[it];
    if (a.clear_nstep && a.S.nstep_len) a.S.nstep_len[e] = 0;  // offpolicy_q.py:94
    a.S.cash[e] = cash;
    if (a.S.folds) {  // flat portfolio: every fold is a sum of zeros
      a.S.folds[(int64_t)MDG_FOLD_AV * N + e] = 0.;
      a.S.folds[(int64_t)MDG_FOLD_ML * N + e] = 0.;
      a.S.folds[(int64_t)MDG_FOLD_BM * N + e] = 0.;
      a.S.folds[(int64_t)MDG_FOLD_SE * N + e] = 0.;
      a.S.folds[(int64_t)MDG_FOLD_G * N + e] = fabs(cash);
    }
    const long long ts = a.S.timestamp[e];
    s_ts[it] = ts;
    a.S.timestamp[e] = ts + fill;
    a.S.reset_ts[e] = ts + fill;
    a.IO.obs_port[((int64_t)a.L.head * (na + 1)) * N + e] = (cash - 0.) / eq;  // newest row = current state
  }
  __syncthreads();
  // (b) items ordered leader-major so that neighbouring lanes run the same generator type
  const int nitems = nd * s_nlead;
  for (int item = tid; item < nitems; item += kResetBlock) {
    const int l = item / nd, d = item - l * nd;
    const int64_t e = e0 + s_list[d];
    const int i0 = s_leader[l];
    const int cnt = (P.gen[i0].type == MDG_GEN_OUPAIR) ? 2 : 1;
    double pr[2];
    for (int c = 0; c < cnt; ++c) {  // dataSource_->reset() and the empty ledger
      const MdgAssetGen& g = P.gen[i0 + c];
      double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
      pr[c] = gen_reset(g, a.S.price[(int64_t)(i0 + c) * N + e], gs, N);
      a.S.ledger[(int64_t)(i0 + c) * N + e] = 0.;
      a.S.mean_entry[(int64_t)(i0 + c) * N + e] = 0.;
      a.S.borrowed[(int64_t)(i0 + c) * N + e] = 0.;
    }
    LazyDraws dr;
    dr.N = N; dr.e = e; dr.gstride = N;
    dr.gid = (uint32_t)(a.L.env_offset + e);
    dr.k0 = (uint32_t)a.L.seed; dr.k1 = (uint32_t)(a.L.seed >> 32);
    const long long ts0 = s_ts[d];
#pragma unroll 1
    for (int t = 0; t < fill; ++t) {
      const long long tick = ts0 + t;
      dr.t_lo = (uint32_t)(unsigned long long)tick;
      dr.t_hi = (uint32_t)((unsigned long long)tick >> 32);
      dr.cached_block = -1;
      dr.normals = a.IO.normals ? a.IO.normals + (int64_t)t * P.n_normals * N : nullptr;
      dr.uniforms = a.IO.uniforms ? a.IO.uniforms + (int64_t)t * P.n_uniforms * N : nullptr;
      double pair_mean = 0.;
#pragma unroll 1
      for (int c = 0; c < cnt; ++c) {
        const MdgAssetGen& g = P.gen[i0 + c];
        double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
        pr[c] = gen_tick(g, pr[c], gs, dr, pair_mean);
        a.IO.pre_price[((int64_t)e * k + (k - fill + t)) * na + i0 + c] = pr[c];
        if (t == fill - 1) {  // newest row = current state, also in the ring
          a.IO.obs_price[((int64_t)a.L.head * na + i0 + c) * N + e] = pr[c];
          a.IO.obs_port[((int64_t)a.L.head * (na + 1) + i0 + c + 1) * N + e] = (0. * pr[c]) / eq;
        }
      }
    }
    for (int c = 0; c < cnt; ++c) a.S.price[(int64_t)(i0 + c) * N + e] = pr[c];
  }
}

// ---------------------------------------------------------------------------
// reset with a workspace: three packed kernels instead of one block-local kernel.
//   scan : one thread per env; flagged envs append themselves to a global list (warp-aggregated
//          atomics), get a fresh portfolio and DataSource::reset().
//   rng  : one thread per (listed env, tick, Philox block) -> the expensive, dependence-free part
//          (Philox + Box-Muller) runs perfectly packed at full occupancy, however scattered the
//          resetting envs are; normals go to an L2-resident scratch [listed env][tick][slot].
//   recur: one thread per (listed env, generator group): the serial recurrence over the ticks with
//          the generator state in registers / local memory, reading the scratch, writing ring rows.
// The host launches ceil(N / cap) rng+recur passes without knowing how many envs reset; passes beyond
// the list end exit at once.
// ---------------------------------------------------------------------------
struct ResetWsArgs {
  MdgParams P;
  MdgState S;
  MdgStepIO IO;
  MdgLaunch L;
  const uint8_t* mask;
  int fill_ticks;
  int clear_nstep;
  int* count;       // workspace: number of listed envs
  int* list;        // workspace: env index of each listed env
  double* scratch;  // workspace: [cap][fill_ticks][n_normals]
  int cap;          // listed envs per pass
  int pass;
};

__global__ void __launch_bounds__(256) reset_scan_kernel(const __grid_constant__ ResetWsArgs a) {
  const MdgParams& P = a.P;
  const int64_t N = a.L.n_envs;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool flag = (e < N) && (!a.mask || a.mask[e]);
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  if (!ballot) return;
  int base = 0;
  if (lane == (__ffs(ballot) - 1)) base = atomicAdd(a.count, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (!flag) return;
  a.list[base + __popc(ballot & ((1u << lane) - 1u))] = (int)e;
  const int na = P.n_assets;
  const int fill = a.fill_ticks;
  const double cash = P.init_cash;
  const double eq = cash + 0. - 0.;  // flat portfolio: equity == cash
  if (a.clear_nstep && a.S.nstep_len) a.S.nstep_len[e] = 0;  // offpolicy_q.py:94
  a.S.cash[e] = cash;
  if (a.S.folds) {  // flat portfolio: every fold is a sum of zeros
    a.S.folds[(int64_t)MDG_FOLD_AV * N + e] = 0.;
    a.S.folds[(int64_t)MDG_FOLD_ML * N + e] = 0.;
    a.S.folds[(int64_t)MDG_FOLD_BM * N + e] = 0.;
    a.S.folds[(int64_t)MDG_FOLD_SE * N + e] = 0.;
    a.S.folds[(int64_t)MDG_FOLD_G * N + e] = fabs(cash);
  }
  const long long ts = a.S.timestamp[e];
  a.S.timestamp[e] = ts + fill;  // the later kernels recover ts as timestamp - fill
  a.S.reset_ts[e] = ts + fill;
  a.IO.obs_port[((int64_t)a.L.head * (na + 1)) * N + e] = (cash - 0.) / eq;  // newest row = current state
  // the per-asset part (DataSource::reset(), zeroed ledger rows) is done by reset_recur_kernel, which has one
  // thread per (env, generator group) instead of one serial 16-asset loop of dependent loads per flagged env
}

__global__ void __launch_bounds__(256) reset_rng_kernel(const __grid_constant__ ResetWsArgs a) {
  const int count = *a.count;
  const int first = a.pass * a.cap;
  if (first >= count) return;
  const int here = (count - first) < a.cap ? (count - first) : a.cap;
  const int nn = a.P.n_normals, fill = a.fill_ticks;
  const int64_t N = a.L.n_envs;
  if (a.IO.normals) {  // validation mode: copy the injected stream [tick][slot][env]
    const int64_t total = (int64_t)here * fill * nn;
    for (int64_t item = (int64_t)blockIdx.x * 256 + threadIdx.x; item < total; item += (int64_t)gridDim.x * 256) {
      const int s = (int)(item % nn);
      const int64_t r = item / nn;
      const int t = (int)(r % fill), d = (int)(r / fill);
      const int64_t e = a.list[first + d];
      a.scratch[item] = a.IO.normals[((int64_t)t * nn + s) * N + e];
    }
    return;
  }
  const int nb = (nn + 1) >> 1;
  const uint32_t k0 = (uint32_t)a.L.seed, k1 = (uint32_t)(a.L.seed >> 32);
  // one block walks whole envs (grid-stride); inside an env, thread j = tick * nb + block.  All index
  // arithmetic is 32-bit with a multiply-high reciprocal (64-bit div/mod cost more than the Box-Muller).
  const uint32_t per_env = (uint32_t)(fill * nb);
  const uint32_t magic = (uint32_t)((0x100000000ull + (uint32_t)nb - 1) / (uint32_t)nb);  // ceil(2^32 / nb)
  for (int d = blockIdx.x; d < here; d += gridDim.x) {
    const int64_t e = a.list[first + d];
    const uint32_t gid = (uint32_t)(a.L.env_offset + e);
    const unsigned long long tick0 = (unsigned long long)(a.S.timestamp[e] - fill);
    double* zenv = a.scratch + (int64_t)d * fill * nn;
    for (uint32_t j = threadIdx.x; j < per_env; j += 256) {
      const uint32_t t = (nb == 1) ? j : __umulhi(j, magic);  // j / nb, exact for j < 2^16 (magic overflows for nb == 1)
      const uint32_t b = j - t * (uint32_t)nb;
      const unsigned long long tick = tick0 + t;
      uint64_t x0, x1;
      philox4x32_10(gid, b, (uint32_t)tick, (uint32_t)(tick >> 32), k0, k1, x0, x1);
      const double u1 = ((double)(x0 >> 12) + 0.5) * 0x1.0p-52;
      const double u2 = (double)(x1 >> 11) * 0x1.0p-53;
      const double r2 = fast_sqrt_pos(-2.0 * fast_log_pos(u1));
      double sn, cs;
      fast_sincos_2pi(u2, sn, cs);
      double* z = zenv + t * (uint32_t)nn + 2 * b;
      if ((nn & 1) == 0) {
        *reinterpret_cast<double2*>(z) = make_double2(r2 * cs, r2 * sn);  // rows of an even length stay 16-byte aligned
      } else {
        z[0] = r2 * cs;
        if (2 * (int)b + 1 < nn) z[1] = r2 * sn;
      }
    }
  }
}

__global__ void __launch_bounds__(128) reset_recur_kernel(const __grid_constant__ ResetWsArgs a) {
  __shared__ int s_leader[MDG_MAX_ASSETS];
  __shared__ int s_nlead;
  const MdgParams& P = a.P;
  const int na = P.n_assets;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < na; ++i)
      if (!(P.gen[i].type == MDG_GEN_OUPAIR && P.gen[i].role == 1)) s_leader[n++] = i;
    s_nlead = n;
  }
  __syncthreads();
  const int count = *a.count;
  const int first = a.pass * a.cap;
  if (first >= count) return;
  const int here = (count - first) < a.cap ? (count - first) : a.cap;
  const int nlead = s_nlead;
  const int nn = P.n_normals, fill = a.fill_ticks, k = a.L.window;
  const int64_t N = a.L.n_envs;
  const double cash = P.init_cash;
  const double eq = cash + 0. - 0.;
  const bool eq_plain = eq > 0. && eq < 1e300;  // then (0*p)/eq == 0*p exactly
  const uint32_t total = (uint32_t)here * (uint32_t)nlead;  // < 2^31 * 16 would overflow: here <= cap <= 65536
  const uint32_t lmagic = (uint32_t)((0x100000000ull + (uint32_t)nlead - 1) / (uint32_t)nlead);
  // item = d * nlead + l: the groups of one env are neighbouring lanes -> they read one contiguous scratch row
  for (uint32_t item = blockIdx.x * 128 + threadIdx.x; item < total; item += gridDim.x * 128) {
    const int d = (nlead == 1) ? (int)item : (int)__umulhi(item, lmagic);  // item / nlead
    const int l = (int)(item - (uint32_t)d * (uint32_t)nlead);
    const int64_t e = a.list[first + d];
    const int i0 = s_leader[l];
    const int cnt = (P.gen[i0].type == MDG_GEN_OUPAIR) ? 2 : 1;
    // generator state of this group in registers / local memory for the whole fast-forward
    double pr[2], gsl[4];
    const int gslot0 = P.gen[i0].gslot;
    const int ngs = gslot0 < 0 ? 0
                    : (P.gen[i0].type == MDG_GEN_TRENDYOU ? 4 : P.gen[i0].type == MDG_GEN_TRENDOU ? 3
                       : P.gen[i0].type == MDG_GEN_SIMPLETREND ? 2 : 1);
    for (int r = 0; r < ngs; ++r) gsl[r] = a.S.gstate[(int64_t)(gslot0 + r) * N + e];
    for (int c = 0; c < cnt; ++c) pr[c] = a.S.price[(int64_t)(i0 + c) * N + e];
    for (int c = 0; c < cnt; ++c) {  // dataSource_->reset(); fresh Broker/Account/Portfolio (Env.h:150-165)
      pr[c] = gen_reset(P.gen[i0 + c], pr[c], gsl, 1);
      a.S.ledger[(int64_t)(i0 + c) * N + e] = 0.;
      a.S.mean_entry[(int64_t)(i0 + c) * N + e] = 0.;
      a.S.borrowed[(int64_t)(i0 + c) * N + e] = 0.;
    }
    const long long ts0 = a.S.timestamp[e] - fill;
    const double* zrow = a.scratch + (int64_t)d * fill * nn;
    double* prow = a.IO.pre_price + ((int64_t)e * k + (k - fill)) * na + i0;
    if (P.gen[i0].type == MDG_GEN_OUPAIR) {
      // OUPair::getData (DataSource.cpp:1232-1240) inlined; the next tick's three normals are loaded while
      // the current tick is computed (the recurrence itself is ~10 dependent flops per tick)
      const MdgAssetGen& g0 = P.gen[i0];
      const MdgAssetGen& g1 = P.gen[i0 + 1];
      const double theta0 = g0.p[0], phi0 = g0.p[1], noise = g0.p[2], theta1 = g1.p[0], phi1 = g1.p[1];
      const int s_rw = g0.nslot_aux, s_0 = g0.nslot, s_1 = g1.nslot;
      double m = gsl[0];
      // chunks of 8 ticks: 24 independent L2 loads in flight, then 8 short dependent updates
      constexpr int TC = 8;
#pragma unroll 1
      for (int t0 = 0; t0 < fill; t0 += TC) {
        double zr[TC], za[TC], zb[TC];
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          if (t0 + u < fill) {
            const double* zn = zrow + (int64_t)(t0 + u) * nn;
            zr[u] = zn[s_rw]; za[u] = zn[s_0]; zb[u] = zn[s_1];
          }
        }
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          if (t0 + u < fill) {
            m += m * (zr[u] * noise);
            pr[0] += (theta0 * (m - pr[0])) + m * (za[u] * phi0);
            pr[1] += (theta1 * (m - pr[1])) + m * (zb[u] * phi1);
            double* dst = prow + (int64_t)(t0 + u) * na;
            if ((na & 1) == 0) {  // 16-byte aligned: one vector store, the env's pairs fill whole sectors together
              *reinterpret_cast<double2*>(dst) = make_double2(pr[0], pr[1]);
            } else {
              dst[0] = pr[0];
              dst[1] = pr[1];
            }
          }
        }
      }
      gsl[0] = m;
    } else {
      TickDraws dr;
      dr.N = N; dr.e = e; dr.gstride = 1;
      dr.gid = (uint32_t)(a.L.env_offset + e);
      dr.k0 = (uint32_t)a.L.seed; dr.k1 = (uint32_t)(a.L.seed >> 32);
#pragma unroll 1
      for (int t = 0; t < fill; ++t) {
        const long long tick = ts0 + t;
        dr.t_lo = (uint32_t)(unsigned long long)tick;
        dr.t_hi = (uint32_t)((unsigned long long)tick >> 32);
        dr.z = zrow + (int64_t)t * nn;
        dr.uniforms = a.IO.uniforms ? a.IO.uniforms + (int64_t)t * P.n_uniforms * N : nullptr;
        double pair_mean = 0.;
        pr[0] = gen_tick(P.gen[i0], pr[0], gsl, dr, pair_mean);
        prow[(int64_t)t * na] = pr[0];
      }
    }
    for (int c = 0; c < cnt; ++c) {  // newest row = current state, also in the ring
      a.IO.obs_price[((int64_t)a.L.head * na + i0 + c) * N + e] = pr[c];
      a.IO.obs_port[((int64_t)a.L.head * (na + 1) + i0 + c + 1) * N + e] = eq_plain ? 0. * pr[c] : (0. * pr[c]) / eq;
    }
    for (int r = 0; r < ngs; ++r) a.S.gstate[(int64_t)(gslot0 + r) * N + e] = gsl[r];
    for (int c = 0; c < cnt; ++c) a.S.price[(int64_t)(i0 + c) * N + e] = pr[c];
  }
}

// constructor state (Env.h:139-165 before the first tick)
struct InitArgs {
  MdgParams P;
  MdgReward R;
  MdgState S;
  MdgLaunch L;
};
__global__ void __launch_bounds__(kBlock) init_kernel(const __grid_constant__ InitArgs a) {
  const int64_t N = a.L.n_envs;
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= N) return;
  const int na = a.P.n_assets;
  for (int i = 0; i < na; ++i) {
    const MdgAssetGen& g = a.P.gen[i];
    double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
    a.S.price[(int64_t)i * N + e] = gen_start(g, gs, N);
    a.S.ledger[(int64_t)i * N + e] = 0.;
    a.S.mean_entry[(int64_t)i * N + e] = 0.;
    a.S.borrowed[(int64_t)i * N + e] = 0.;
  }
  a.S.cash[e] = a.P.init_cash;
  a.S.timestamp[e] = 0;
  if (a.S.reset_ts) a.S.reset_ts[e] = 0;
  if (a.S.folds) {
    for (int r = 0; r < 4; ++r) a.S.folds[(int64_t)r * N + e] = 0.;
    a.S.folds[(int64_t)MDG_FOLD_G * N + e] = fabs(a.P.init_cash);
  }
  const int ra = a.R.reduce_rewards ? 1 : na;
  for (int c = 0; c < ra; ++c) {
    if (a.S.shaper_A) a.S.shaper_A[(int64_t)c * N + e] = 0.;
    if (a.S.shaper_B) a.S.shaper_B[(int64_t)c * N + e] = 0.;
  }
  if (a.S.nstep_len) a.S.nstep_len[e] = 0;
}

// exact left-to-right folds of the portfolio as stored (after external writes into the state tensors)
struct FoldArgs {
  MdgParams P;
  MdgState S;
  int64_t N;
};
__global__ void __launch_bounds__(kBlock) refresh_folds_kernel(const __grid_constant__ FoldArgs a) {
  const int64_t N = a.N;
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= N) return;
  double av = 0., ml = 0., bms = 0., se = 0., g = fabs(a.S.cash[e]);
  for (int j = 0; j < a.P.n_assets; ++j) {
    const double l = a.S.ledger[(int64_t)j * N + e], p = a.S.price[(int64_t)j * N + e],
                 m = a.S.mean_entry[(int64_t)j * N + e], b = a.S.borrowed[(int64_t)j * N + e];
    const double t_ml = m * l;
    const double t_se = (l < 0.) ? t_ml : 0. * t_ml;
    if (j == 0) { av = l * p; ml = t_ml; bms = b; se = t_se; }
    else { av = av + l * p; ml = ml + t_ml; bms = bms + b; se = se + t_se; }
    g += fabs(l * p) + fabs(t_ml) + fabs(b);
  }
  a.S.folds[(int64_t)MDG_FOLD_AV * N + e] = av;
  a.S.folds[(int64_t)MDG_FOLD_ML * N + e] = ml;
  a.S.folds[(int64_t)MDG_FOLD_BM * N + e] = bms;
  a.S.folds[(int64_t)MDG_FOLD_SE * N + e] = se;
  a.S.folds[(int64_t)MDG_FOLD_G * N + e] = g;
}

static int check_common(const MdgParams* P, const MdgLaunch* L) {
  if (!P || !L) return set_err(MDG_E_INVALID, "null params/launch");
  if (P->n_assets < 1 || P->n_assets > MDG_MAX_ASSETS)
    return set_err(MDG_E_UNSUPPORTED, "n_assets must be in 1..MDG_MAX_ASSETS (thread-per-env kernels)");
  if (L->n_envs < 0) return set_err(MDG_E_INVALID, "n_envs < 0");
  if (L->n_envs > (int64_t)2147483647 * kBlock) return set_err(MDG_E_UNSUPPORTED, "n_envs too large");
  return MDG_OK;
}

}  // namespace mdg

using namespace mdg;

extern "C" int mdg_step(const MdgParams* P, const MdgReward* R, const MdgState* S, const MdgStepIO* IO,
                        const MdgLaunch* L) {
  int rc = check_common(P, L);
  if (rc) return rc;
  if (!S || !IO) return set_err(MDG_E_INVALID, "null state/io");
  if (!S->folds) return set_err(MDG_E_INVALID, "state.folds is null");
  if (L->window < 1 || L->head < 0 || L->head >= L->window) return set_err(MDG_E_INVALID, "bad window/head");
  if (L->mode < MDG_MODE_HOLD || L->mode > MDG_MODE_SINGLE) return set_err(MDG_E_INVALID, "bad mode");
  const bool by_actions = L->mode == MDG_MODE_MULTI && IO->actions;
  if (L->mode != MDG_MODE_HOLD && !IO->units && !by_actions) return set_err(MDG_E_INVALID, "units is null");
  if (by_actions && (L->action_atoms < 1 || L->action_atoms > 127))
    return set_err(MDG_E_INVALID, "action_atoms must be in [1, 127] with io.actions");
  if (L->mode == MDG_MODE_SINGLE && (L->asset_idx < 0 || L->asset_idx >= P->n_assets))
    return set_err(MDG_E_INVALID, "asset index out of range");  // std::out_of_range -> IndexError
  if (L->n_envs == 0) return MDG_OK;
  StepArgs a;
  a.P = *P;
  if (R) a.R = *R; else { memset(&a.R, 0, sizeof(a.R)); a.R.shaper = MDG_SHAPER_OFF; a.R.nstep = 1; }
  a.S = *S;
  a.IO = *IO;
  a.L = *L;
  if (a.R.shaper != MDG_SHAPER_OFF) {
    if (a.R.nstep < 1 || a.R.nstep > MDG_MAX_NSTEP) return set_err(MDG_E_INVALID, "nstep out of range");
    if (!IO->agent_reward || !IO->shaped_reward || !IO->n_popped)
      return set_err(MDG_E_INVALID, "shaper on but agent_reward/shaped_reward/n_popped is null");
    if ((a.R.shaper == MDG_SHAPER_DSR || a.R.shaper == MDG_SHAPER_DDR) && (!S->shaper_A || !S->shaper_B))
      return set_err(MDG_E_INVALID, "DSR/DDR need shaper_A/shaper_B");
    if (a.R.nstep > 1 && (!S->nstep_ring || !S->nstep_len))
      return set_err(MDG_E_INVALID, "nstep>1 needs nstep_ring/nstep_len");
    if (L->nstep_pos < 0 || L->nstep_pos >= a.R.nstep) return set_err(MDG_E_INVALID, "bad nstep_pos");
  }
  return launch_step(a);
}

extern "C" int mdg_reset(const MdgParams* P, const MdgState* S, const MdgStepIO* IO, const MdgLaunch* L,
                         const uint8_t* mask, int fill_ticks, int clear_nstep) {
  int rc = check_common(P, L);
  if (rc) return rc;
  if (!S || !IO) return set_err(MDG_E_INVALID, "null state/io");
  if (!S->reset_ts || !IO->pre_price) return set_err(MDG_E_INVALID, "state.reset_ts / io.pre_price is null");
  if (L->window < 1 || L->head < 0 || L->head >= L->window) return set_err(MDG_E_INVALID, "bad window/head");
  if (fill_ticks < 1) fill_ticks = 1;
  if (fill_ticks > L->window) return set_err(MDG_E_INVALID, "fill_ticks > window");
  if (L->n_envs == 0) return MDG_OK;
  ResetArgs a;
  a.P = *P; a.S = *S; a.IO = *IO; a.L = *L;
  a.mask = mask; a.fill_ticks = fill_ticks; a.clear_nstep = clear_nstep;
  const unsigned grid = (unsigned)((L->n_envs + kResetBlock - 1) / kResetBlock);
  reset_kernel<<<grid, kResetBlock, 0, (cudaStream_t)L->stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_reset launch");
}

extern "C" int64_t mdg_reset_workspace_bytes(const MdgParams* P, int64_t n_envs, int fill_ticks) {
  if (!P || n_envs < 0) return -1;
  if (fill_ticks < 1) fill_ticks = 1;
  // header (count) + list + scratch for min(N, 65536) listed envs per pass
  const int64_t cap = n_envs < 65536 ? n_envs : 65536;
  return 256 + 4 * ((n_envs + 63) / 64 * 64) + 8 * cap * (int64_t)fill_ticks * (P->n_normals > 0 ? P->n_normals : 1);
}

extern "C" int mdg_reset_ws(const MdgParams* P, const MdgState* S, const MdgStepIO* IO, const MdgLaunch* L,
                            const uint8_t* mask, int fill_ticks, int clear_nstep, void* workspace,
                            int64_t workspace_bytes) {
  if (!workspace) return mdg_reset(P, S, IO, L, mask, fill_ticks, clear_nstep);
  int rc = check_common(P, L);
  if (rc) return rc;
  if (!S || !IO) return set_err(MDG_E_INVALID, "null state/io");
  if (!S->reset_ts || !IO->pre_price) return set_err(MDG_E_INVALID, "state.reset_ts / io.pre_price is null");
  if (L->window < 1 || L->head < 0 || L->head >= L->window) return set_err(MDG_E_INVALID, "bad window/head");
  if (fill_ticks < 1) fill_ticks = 1;
  if (fill_ticks > L->window) return set_err(MDG_E_INVALID, "fill_ticks > window");
  const int64_t N = L->n_envs;
  if (N == 0) return MDG_OK;
  if (N > 2147483647) return set_err(MDG_E_UNSUPPORTED, "n_envs too large for the reset list");
  const int nn = P->n_normals > 0 ? P->n_normals : 1;
  const int64_t list_bytes = 4 * ((N + 63) / 64 * 64);
  const int64_t per_env = 8 * (int64_t)fill_ticks * nn;
  const int64_t cap64 = (workspace_bytes - 256 - list_bytes) / per_env;
  if (cap64 < 1) return set_err(MDG_E_INVALID, "reset workspace too small (see mdg_reset_workspace_bytes)");
  ResetWsArgs a;
  a.P = *P; a.S = *S; a.IO = *IO; a.L = *L;
  a.mask = mask; a.fill_ticks = fill_ticks; a.clear_nstep = clear_nstep;
  a.count = (int*)workspace;
  a.list = (int*)((char*)workspace + 256);
  a.scratch = (double*)((char*)workspace + 256 + list_bytes);
  a.cap = (int)(cap64 < N ? cap64 : N);
  a.pass = 0;
  cudaStream_t st = (cudaStream_t)L->stream;
  cudaError_t ce = cudaMemsetAsync(a.count, 0, sizeof(int), st);
  if (ce != cudaSuccess) return cuda_err(ce, "mdg_reset_ws memset");
  reset_scan_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(a);
  const int passes = (int)((N + a.cap - 1) / a.cap);
  const int nb = (P->n_normals + 1) / 2 > 0 ? (P->n_normals + 1) / 2 : 1;
  // profiling knob (profiles/breakdown.py): stop after the scan (1) or the rng kernel (2); results are then wrong
  static const int phases = [] { const char* v = getenv("MDG_RESET_PHASES"); return v ? atoi(v) : 3; }();
  for (int p = 0; p < passes && phases >= 2; ++p) {
    a.pass = p;
    // grids sized for a full pass but capped: the kernels are grid-stride and exit at once past the list end
    unsigned g1 = (unsigned)(a.cap < 148 * 8 ? a.cap : 148 * 8);  // one block per listed env, grid-stride
    (void)nb;
    reset_rng_kernel<<<g1 ? g1 : 1, 256, 0, st>>>(a);
    if (phases < 3) continue;
    int64_t rec_items = (int64_t)a.cap * P->n_assets;
    unsigned g2 = (unsigned)((rec_items + 127) / 128 < 148 * 16 ? (rec_items + 127) / 128 : 148 * 16);
    reset_recur_kernel<<<g2 ? g2 : 1, 128, 0, st>>>(a);
  }
  return cuda_err(cudaGetLastError(), "mdg_reset_ws launch");
}

extern "C" int mdg_step_autoreset(const MdgParams* P, const MdgReward* R, const MdgState* S, const MdgStepIO* IO,
                                  const MdgLaunch* L, int fill_ticks, int clear_nstep, void* workspace,
                                  int64_t workspace_bytes) {
  int rc = mdg_step(P, R, S, IO, L);
  if (rc) return rc;
  MdgStepIO io = *IO;  // the reset draws from Philox (no injected stream) and takes no units
  io.units = nullptr; io.normals = nullptr; io.uniforms = nullptr; io.actions = nullptr;
  return mdg_reset_ws(P, S, &io, L, IO->done, fill_ticks, clear_nstep, workspace, workspace_bytes);
}

extern "C" int mdg_init_state(const MdgParams* P, const MdgReward* R, const MdgState* S, const MdgLaunch* L) {
  int rc = check_common(P, L);
  if (rc) return rc;
  if (!S) return set_err(MDG_E_INVALID, "null state");
  if (L->n_envs == 0) return MDG_OK;
  InitArgs a;
  a.P = *P;
  if (R) a.R = *R; else { memset(&a.R, 0, sizeof(a.R)); a.R.nstep = 1; }
  a.S = *S; a.L = *L;
  const unsigned grid = (unsigned)((L->n_envs + kBlock - 1) / kBlock);
  init_kernel<<<grid, kBlock, 0, (cudaStream_t)L->stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_init_state launch");
}

extern "C" int mdg_refresh_folds(const MdgParams* P, const MdgState* S, const MdgLaunch* L) {
  int rc = check_common(P, L);
  if (rc) return rc;
  if (!S || !S->folds) return set_err(MDG_E_INVALID, "null state/folds");
  if (L->n_envs == 0) return MDG_OK;
  FoldArgs a;
  a.P = *P; a.S = *S; a.N = L->n_envs;
  const unsigned grid = (unsigned)((L->n_envs + kBlock - 1) / kBlock);
  refresh_folds_kernel<<<grid, kBlock, 0, (cudaStream_t)L->stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_refresh_folds launch");
}

extern "C" int mdg_abi_version(void) { return MDG_ABI_VERSION; }
extern "C" const char* mdg_last_error(void) { return err_buf(); }
extern "C" int mdg_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(MdgAssetGen);
    case 1: return (int)sizeof(MdgParams);
    case 2: return (int)sizeof(MdgReward);
    case 3: return (int)sizeof(MdgState);
    case 4: return (int)sizeof(MdgStepIO);
    case 5: return (int)sizeof(MdgLaunch);
    case 6: return (int)sizeof(MdgDerived);
    case 7: return (int)sizeof(MdgWindow);
    case 8: return (int)sizeof(MdgReplay);
    case 9: return (int)sizeof(MdgReplayBatch);
  }
  return -1;
}
