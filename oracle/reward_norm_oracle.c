/*
 * reward_norm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's streaming reward normalisers
 * (madigan/environments/reward_normalization.pyx), one env at a time, on the same
 * state layout as the CUDA path (MdgRewardNorm with HOST pointers, rows of length N).
 * Pinned against the reference's own Cython module through tests/golden/reward_norm.npz
 * (tests/golden/make_golden_reward_norm.py cythonizes the .pyx in a temp directory).
 * Only tests/ may call it.
 */
#include <math.h>
#include <stddef.h>

#include "../include/madigan_b200.h"

/* queue<double> of one env on the ring storage buffer[slot * N + e] */
typedef struct {
  double *buf;
  int64_t N, e;
  int window, size, front;
} Queue;
static void q_push(Queue *q, double v) { /* queue.push */
  int slot = q->front + q->size;
  if (slot >= q->window) slot -= q->window;
  q->buf[(int64_t)slot * q->N + q->e] = v;
  q->size += 1;
}
static double q_front(const Queue *q) { return q->buf[(int64_t)q->front * q->N + q->e]; }
static void q_pop(Queue *q) {
  q->front = (q->front + 1 == q->window) ? 0 : q->front + 1;
  q->size -= 1;
}

static void reset_env(const MdgRewardNorm *R, int64_t e) {
  if (R->size) R->size[e] = 0; /* :86-87, :168-169 */
  if (R->front) R->front[e] = 0;
  if (R->count) R->count[e] = 0;
  if (R->mean_est) R->mean_est[e] = 0.;
  if (R->ssq) R->ssq[e] = 0.;
  if (R->kind == MDG_RN_SHARPE_EWMA) { /* :240-248 */
    R->ewma[e] = 0.; R->ewma_old[e] = 0.; R->ewssq_old[e] = 0.; R->ewssq[e] = 0.;
    R->w1[e] = 1.; R->w2[e] = 1.;
  }
}

void orc_reward_norm_reset(const MdgRewardNorm *R, const uint8_t *mask) {
  for (int64_t e = 0; e < R->n_envs; ++e)
    if (!mask || mask[e]) reset_env(R, e);
}

/* SharpeFixedWindow.stream / SortinoFixedWindowA.stream  (reward_normalization.pyx:92-118, :137-145) */
static double sharpe_fixed(const MdgRewardNorm *R, int64_t e, double reward, int sortino_a) {
  Queue q = {R->buffer, R->n_envs, e, R->window, R->size[e], R->front[e]};
  double mean_est = R->mean_est[e], ssq = R->ssq[e];
  if (q.size == R->window) { /* tail_adjust :111-118 */
    const double remove = q_front(&q);
    q_pop(&q);
    const double delt = remove - mean_est;
    mean_est -= delt / q.size;
    ssq -= (delt * (remove - mean_est));
  }
  { /* head_add :103-109 */
    q_push(&q, reward);
    const double delt = reward - mean_est;
    mean_est += delt / q.size;
    ssq += delt * (reward - mean_est);
  }
  R->size[e] = q.size; R->front[e] = q.front; R->mean_est[e] = mean_est; R->ssq[e] = ssq;
  if (q.size <= 1) return 0.;
  reward = reward / sqrt((ssq + 1e-8) / q.size);
  if (sortino_a && reward < 0) return -1 * (reward * reward);
  return reward;
}

/* SortinoFixedWindowB.stream / C.stream (:177-218); update :189-202 */
static double sortino_bc(const MdgRewardNorm *R, int64_t e, double reward, int square_negative) {
  Queue q = {R->buffer, R->n_envs, e, R->window, R->size[e], R->front[e]};
  double mean_est = R->mean_est[e], ssq = R->ssq[e];
  unsigned int count = (unsigned int)R->count[e];
  double delt = reward - mean_est;
  if (delt < 0) {
    count += 1;
    mean_est += delt / count;
    ssq += delt * (reward - mean_est);
    if (R->window == q.size) {
      count -= 1;
      const double remove = q_front(&q);
      q_pop(&q);
      delt = remove - mean_est;
      mean_est -= delt / count;
      ssq -= (delt * (remove - mean_est));
    }
    q_push(&q, reward);
  }
  R->size[e] = q.size; R->front[e] = q.front; R->count[e] = (int32_t)count;
  R->mean_est[e] = mean_est; R->ssq[e] = ssq;
  if (q.size <= 1) return 0.;
  reward = reward / sqrt((ssq + 1e-8) / count);
  if (square_negative && reward < 0) return -1 * (reward * reward);
  return reward;
}

/* SharpeEWMA.stream / update (:250-272) */
static double sharpe_ewma(const MdgRewardNorm *R, int64_t e, double value) {
  const double alpha = R->alpha;
  int count = R->count[e] + 1;
  const double pw = pow(1 - alpha, (double)count);
  const double w1 = R->w1[e] + pw;
  const double w2 = R->w2[e] + pw * pw;
  const double ewma_prev = R->ewma[e];
  const double ewma_old = R->ewma_old[e] * (1 - alpha) + value;
  const double ewma = ewma_old / w1;
  const double ewssq_old = R->ewssq_old[e] * (1 - alpha) + ((value - ewma) * (value - ewma_prev));
  const double ewssq = ewssq_old / (w1 - w2 / w1);
  R->count[e] = count; R->w1[e] = w1; R->w2[e] = w2; R->ewma[e] = ewma; R->ewma_old[e] = ewma_old;
  R->ewssq_old[e] = ewssq_old; R->ewssq[e] = ewssq;
  if (count <= 1) return 0.;
  return value / sqrt(ewssq);
}

void orc_reward_norm_stream(const MdgRewardNorm *R, const double *reward, const uint8_t *reset_mask, double *out) {
  for (int64_t e = 0; e < R->n_envs; ++e) {
    if (reset_mask && reset_mask[e]) reset_env(R, e);
    switch (R->kind) {
      case MDG_RN_NULL: out[e] = reward[e]; break; /* :49-58 */
      case MDG_RN_SHARPE_FIXED: out[e] = sharpe_fixed(R, e, reward[e], 0); break;
      case MDG_RN_SORTINO_A: out[e] = sharpe_fixed(R, e, reward[e], 1); break;
      case MDG_RN_SORTINO_B: out[e] = sortino_bc(R, e, reward[e], 1); break;
      case MDG_RN_SORTINO_C: out[e] = sortino_bc(R, e, reward[e], 0); break;
      case MDG_RN_SHARPE_EWMA: out[e] = sharpe_ewma(R, e, reward[e]); break;
      default: out[e] = 0.;
    }
  }
}
