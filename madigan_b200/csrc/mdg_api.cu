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

__global__ void __launch_bounds__(kBlock) reset_kernel(const __grid_constant__ ResetArgs a) {
  __shared__ double s_z[kMaxNormals][kBlock];
  const MdgParams& P = a.P;
  const int64_t N = a.L.n_envs;
  const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (e >= N) return;
  if (a.mask && !a.mask[e]) return;
  const int na = P.n_assets;
  const int k = a.L.window;
  if (a.clear_nstep && a.S.nstep_len) a.S.nstep_len[e] = 0;  // offpolicy_q.py:94
  // dataSource_->reset(); fresh Broker/Account/Portfolio (Env.h:150-165)
  for (int i = 0; i < na; ++i) {
    const MdgAssetGen& g = P.gen[i];
    double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
    a.S.price[(int64_t)i * N + e] = gen_reset(g, a.S.price[(int64_t)i * N + e], gs, N);
    a.S.ledger[(int64_t)i * N + e] = 0.;
    a.S.mean_entry[(int64_t)i * N + e] = 0.;
    a.S.borrowed[(int64_t)i * N + e] = 0.;
  }
  const double cash = P.init_cash;
  a.S.cash[e] = cash;
  int64_t ts = a.S.timestamp[e];
  double* zcol = &s_z[0][threadIdx.x];
#pragma unroll 1
  for (int t = 0; t < a.fill_ticks; ++t) {
    GenCtx ctx;
    ctx_init(ctx, zcol, kBlock, a.IO.uniforms ? a.IO.uniforms + (int64_t)t * P.n_uniforms * N : nullptr, a.L, e, ts);
    fill_normals(zcol, kBlock, P.n_normals, a.IO.normals ? a.IO.normals + (int64_t)t * P.n_normals * N : nullptr, ctx);
    double pair_mean = 0.;
    int slot = (a.L.head - (a.fill_ticks - 1 - t)) % k;
    if (slot < 0) slot += k;
    // flat portfolio: equity == cash, ledgerNormedFull == [cash/equity, 0*price/equity ...]
    const double eq = cash + 0. - 0.;
    a.IO.obs_port[((int64_t)slot * (na + 1)) * N + e] = (cash - 0.) / eq;
#pragma unroll 1
    for (int i = 0; i < na; ++i) {
      const MdgAssetGen& g = P.gen[i];
      double* gs = a.S.gstate + (int64_t)(g.gslot < 0 ? 0 : g.gslot) * N + e;
      const double pr = gen_tick(g, a.S.price[(int64_t)i * N + e], gs, ctx, pair_mean);
      a.S.price[(int64_t)i * N + e] = pr;
      a.IO.obs_price[((int64_t)slot * na + i) * N + e] = pr;
      a.IO.obs_port[((int64_t)slot * (na + 1) + i + 1) * N + e] = (0. * pr) / eq;
    }
    ts += 1;
    a.IO.obs_time[(int64_t)slot * N + e] = ts;
  }
  a.S.timestamp[e] = ts;
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
  const int ra = a.R.reduce_rewards ? 1 : na;
  for (int c = 0; c < ra; ++c) {
    if (a.S.shaper_A) a.S.shaper_A[(int64_t)c * N + e] = 0.;
    if (a.S.shaper_B) a.S.shaper_B[(int64_t)c * N + e] = 0.;
  }
  if (a.S.nstep_len) a.S.nstep_len[e] = 0;
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
  if (na == 1) return launch_step_cap1(a, true);
  if (na == 2) return launch_step_cap2(a, true);
  if (na <= 4) return launch_step_cap4(a, na == 4);
  if (na <= 8) return launch_step_cap8(a, na == 8);
  return launch_step_cap16(a, na == 16);
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
  const unsigned grid = (unsigned)((L->n_envs + kBlock - 1) / kBlock);
  reset_kernel<<<grid, kBlock, 0, (cudaStream_t)L->stream>>>(a);
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
