// Streaming reward normalisers for N envs (sm_100a): environments/reward_normalization.pyx:14-272.
//
// One thread per env; every piece of state is an [N] row (the queue is a ring [window][N]), so each access of a
// warp is one contiguous run.  HBM-bound scalar work (a few dozen flops and ~10 words per env-step).  The
// arithmetic follows the reference's Cython statement by statement (compiled with -fmad=false; IEEE sqrt and
// division), so the fixed-window kinds agree with the CPU oracle bit for bit; SHARPE_EWMA goes through pow().
#include <math.h>
#include <stdio.h>

#include "mdg_common.cuh"

namespace mdg {

struct RnArgs {
  MdgRewardNorm R;
  const double* reward;
  const uint8_t* mask;
  double* out;
};

__device__ __forceinline__ void rn_reset_env(const MdgRewardNorm& R, int64_t e) {
  if (R.size) R.size[e] = 0;    // while not buffer.empty(): buffer.pop()   :86-87
  if (R.front) R.front[e] = 0;
  if (R.count) R.count[e] = 0;
  if (R.mean_est) R.mean_est[e] = 0.;
  if (R.ssq) R.ssq[e] = 0.;
  if (R.kind == MDG_RN_SHARPE_EWMA) {  // :240-248
    R.ewma[e] = 0.; R.ewma_old[e] = 0.; R.ewssq_old[e] = 0.; R.ewssq[e] = 0.;
    R.w1[e] = 1.; R.w2[e] = 1.;
  }
}

__global__ void __launch_bounds__(256) rn_reset_kernel(const __grid_constant__ RnArgs a) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= a.R.n_envs) return;
  if (a.mask && !a.mask[e]) return;
  rn_reset_env(a.R, e);
}

__global__ void __launch_bounds__(256) rn_stream_kernel(const __grid_constant__ RnArgs a) {
  const MdgRewardNorm& R = a.R;
  const int64_t N = R.n_envs;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= N) return;
  const double reward = a.reward[e];
  const bool fresh = a.mask && a.mask[e];
  double out = 0.;
  switch (R.kind) {
    case MDG_RN_NULL:
      out = reward;  // :49-58
      break;
    case MDG_RN_SHARPE_FIXED:
    case MDG_RN_SORTINO_A: {  // :61-145
      const int window = R.window;
      int size = fresh ? 0 : R.size[e], front = fresh ? 0 : R.front[e];
      double mean = fresh ? 0. : R.mean_est[e], ssq = fresh ? 0. : R.ssq[e];
      if (size == window) {  // tail_adjust :111-118
        const double remove = R.buffer[(int64_t)front * N + e];
        front = (front + 1 == window) ? 0 : front + 1;
        size -= 1;
        const double delt = remove - mean;
        mean -= delt / size;
        ssq -= (delt * (remove - mean));
      }
      {  // head_add :103-109
        int slot = front + size;
        if (slot >= window) slot -= window;
        R.buffer[(int64_t)slot * N + e] = reward;
        size += 1;
        const double delt = reward - mean;
        mean += delt / size;
        ssq += delt * (reward - mean);
      }
      R.size[e] = size; R.front[e] = front; R.mean_est[e] = mean; R.ssq[e] = ssq;
      if (size > 1) {
        double v = reward / sqrt((ssq + 1e-8) / size);
        if (R.kind == MDG_RN_SORTINO_A && v < 0) v = -1 * (v * v);
        out = v;
      }
      break;
    }
    case MDG_RN_SORTINO_B:
    case MDG_RN_SORTINO_C: {  // :148-218
      const int window = R.window;
      int size = fresh ? 0 : R.size[e], front = fresh ? 0 : R.front[e];
      unsigned count = fresh ? 0u : (unsigned)R.count[e];
      double mean = fresh ? 0. : R.mean_est[e], ssq = fresh ? 0. : R.ssq[e];
      double delt = reward - mean;  // update :189-202: only below-mean rewards enter the estimate AND the queue
      if (delt < 0) {
        count += 1;
        mean += delt / count;
        ssq += delt * (reward - mean);
        if (window == size) {
          count -= 1;
          const double remove = R.buffer[(int64_t)front * N + e];
          front = (front + 1 == window) ? 0 : front + 1;
          size -= 1;
          delt = remove - mean;
          mean -= delt / count;
          ssq -= (delt * (remove - mean));
        }
        int slot = front + size;
        if (slot >= window) slot -= window;
        R.buffer[(int64_t)slot * N + e] = reward;
        size += 1;
      }
      R.size[e] = size; R.front[e] = front; R.count[e] = (int32_t)count; R.mean_est[e] = mean; R.ssq[e] = ssq;
      if (size > 1) {
        double v = reward / sqrt((ssq + 1e-8) / count);
        if (R.kind == MDG_RN_SORTINO_B && v < 0) v = -1 * (v * v);
        out = v;
      }
      break;
    }
    case MDG_RN_SHARPE_EWMA: {  // :221-272
      const double alpha = R.alpha;
      int count = fresh ? 0 : R.count[e];
      double ewma = fresh ? 0. : R.ewma[e], ewma_old = fresh ? 0. : R.ewma_old[e];
      double ewssq_old = fresh ? 0. : R.ewssq_old[e];
      double w1 = fresh ? 1. : R.w1[e], w2 = fresh ? 1. : R.w2[e];
      count += 1;
      const double pw = pow(1 - alpha, (double)count);
      w1 += pw;
      w2 += pw * pw;
      const double ewma_prev = ewma;
      ewma_old = ewma_old * (1 - alpha) + reward;
      ewma = ewma_old / w1;
      ewssq_old = ewssq_old * (1 - alpha) + ((reward - ewma) * (reward - ewma_prev));
      const double ewssq = ewssq_old / (w1 - w2 / w1);
      R.count[e] = count; R.ewma[e] = ewma; R.ewma_old[e] = ewma_old; R.ewssq_old[e] = ewssq_old;
      R.ewssq[e] = ewssq; R.w1[e] = w1; R.w2[e] = w2;
      if (count > 1) out = reward / sqrt(ewssq);
      break;
    }
  }
  a.out[e] = out;
}

static int rn_check(const MdgRewardNorm* rn) {
  if (!rn) return set_err(MDG_E_INVALID, "null reward normaliser");
  if (rn->n_envs < 0) return set_err(MDG_E_INVALID, "n_envs < 0");
  switch (rn->kind) {
    case MDG_RN_NULL:
      return MDG_OK;
    case MDG_RN_SHARPE_FIXED:
    case MDG_RN_SORTINO_A:
    case MDG_RN_SORTINO_B:
    case MDG_RN_SORTINO_C:
      // window 1 divides by an empty queue's size in the reference (a ZeroDivisionError inside a cdef function)
      if (rn->window < 2) return set_err(MDG_E_INVALID, "reward normaliser window must be >= 2");
      if (!rn->buffer || !rn->size || !rn->front || !rn->mean_est || !rn->ssq)
        return set_err(MDG_E_INVALID, "fixed-window reward normaliser: null state");
      if ((rn->kind == MDG_RN_SORTINO_B || rn->kind == MDG_RN_SORTINO_C) && !rn->count)
        return set_err(MDG_E_INVALID, "SortinoFixedWindowB/C: null count");
      return MDG_OK;
    case MDG_RN_SHARPE_EWMA:
      if (!rn->count || !rn->ewma || !rn->ewma_old || !rn->ewssq_old || !rn->ewssq || !rn->w1 || !rn->w2)
        return set_err(MDG_E_INVALID, "SharpeEWMA: null state");
      return MDG_OK;
  }
  return set_err(MDG_E_INVALID, "unknown reward normaliser kind");
}

}  // namespace mdg

using namespace mdg;

extern "C" int mdg_reward_norm_reset(const MdgRewardNorm* rn, const uint8_t* mask, void* stream) {
  int rc = rn_check(rn);
  if (rc) return rc;
  if (rn->n_envs == 0 || rn->kind == MDG_RN_NULL) return MDG_OK;
  RnArgs a;
  a.R = *rn; a.reward = nullptr; a.mask = mask; a.out = nullptr;
  rn_reset_kernel<<<(unsigned)((rn->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_reward_norm_reset launch");
}

extern "C" int mdg_reward_norm_stream(const MdgRewardNorm* rn, const double* reward, const uint8_t* reset_mask,
                                      double* out, void* stream) {
  int rc = rn_check(rn);
  if (rc) return rc;
  if (!reward || !out) return set_err(MDG_E_INVALID, "null reward/out");
  if (rn->n_envs == 0) return MDG_OK;
  RnArgs a;
  a.R = *rn; a.reward = reward; a.mask = reset_mask; a.out = out;
  rn_stream_kernel<<<(unsigned)((rn->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "mdg_reward_norm_stream launch");
}
