// extern "C" entry points of the step path + the (rare, non-unrolled) reset / init kernels.
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
    for (int t = 0; t < fill; ++t) {
      int slot = (a.L.head - (fill - 1 - t)) % k;
      if (slot < 0) slot += k;
      a.IO.obs_port[((int64_t)slot * (na + 1)) * N + e] = (cash - 0.) / eq;
      a.IO.obs_time[(int64_t)slot * N + e] = ts + t + 1;
    }
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
    dr.N = N; dr.e = e;
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
      int slot = (a.L.head - (fill - 1 - t)) % k;
      if (slot < 0) slot += k;
      double pair_mean = 0.;
#pragma unroll 1
      for (int c = 0; c < cnt; ++c) {
        const MdgAssetGen& g = P.gen[i0 + c];
        double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
        pr[c] = gen_tick(g, pr[c], gs, dr, pair_mean);
        a.IO.obs_price[((int64_t)slot * na + i0 + c) * N + e] = pr[c];
        a.IO.obs_port[((int64_t)slot * (na + 1) + i0 + c + 1) * N + e] = (0. * pr[c]) / eq;
      }
    }
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
  if (L->mode != MDG_MODE_HOLD && !IO->units) return set_err(MDG_E_INVALID, "units is null");
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
  const int na = P->n_assets;
  (void)na;
  return launch_step(a);
}

extern "C" int mdg_reset(const MdgParams* P, const MdgState* S, const MdgStepIO* IO, const MdgLaunch* L,
                         const uint8_t* mask, int fill_ticks, int clear_nstep) {
  int rc = check_common(P, L);
  if (rc) return rc;
  if (!S || !IO) return set_err(MDG_E_INVALID, "null state/io");
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
  }
  return -1;
}
