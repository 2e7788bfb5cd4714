/*
 * mdg_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, one env at a time) of the reference's Env.step hot
 * path, used only as the checker in tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under
 * madigan_b200/ may include, link or call it.
 *
 * It reuses the POD config structs of include/madigan_b200.h so that the checker
 * and the CUDA path are driven by byte-identical configuration.
 */
#ifndef MDG_ORACLE_H_
#define MDG_ORACLE_H_

#include "../include/madigan_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_GSTATE 320 /* 4 rows per trend asset; 4K(+1+T) per SINEDYNAMIC(TREND) asset */
#define ORC_MAX_CTOR_U (3 * MDG_MAX_SINE_COMPONENTS)

typedef struct OrcEnv {
  MdgParams P;
  MdgReward R;
  /* Portfolio (Portfolio.h:111-123) */
  double price[MDG_MAX_ASSETS];
  double ledger[MDG_MAX_ASSETS];
  double mep[MDG_MAX_ASSETS];
  double bm[MDG_MAX_ASSETS];
  double cash;
  /* generator */
  double gstate[ORC_MAX_GSTATE];
  int64_t timestamp;
  /* free-running RNG identity */
  uint64_t seed;
  int64_t gid;
  /* reward shaper (nstep_buffer.py) */
  double A[MDG_MAX_ASSETS], B[MDG_MAX_ASSETS];
  double ring[MDG_MAX_NSTEP][MDG_MAX_ASSETS]; /* oldest first */
  int32_t ring_len;
  /* test hook (single-asset SINEDYNAMIC* sources): canonical uniforms to use instead of the Philox
   * constructor / reset() draws, slot 3c + {freq, mu, amp}; n_ctor_u = 0 -> Philox */
  int32_t n_ctor_u;
  double ctor_u[ORC_MAX_CTOR_U];
} OrcEnv;

typedef struct OrcStepOut {
  double price[MDG_MAX_ASSETS];    /* State.price                 */
  double port[MDG_MAX_ASSETS + 1]; /* State.portfolio             */
  int64_t timestamp;               /* State.timestamp             */
  double reward;
  uint8_t done;
  double tp[MDG_MAX_ASSETS], tu[MDG_MAX_ASSETS], tc[MDG_MAX_ASSETS];
  uint8_t risk[MDG_MAX_ASSETS];
  uint8_t margin_call;
  double agent_reward[MDG_MAX_ASSETS];
  double shaped[MDG_MAX_NSTEP][MDG_MAX_ASSETS];
  int32_t n_popped;
} OrcStepOut;

/* Philox4x32-10 (Salmon et al. 2011) and the draw conventions shared with the kernels */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_draw_normal(uint64_t seed, int64_t gid, int64_t tick, int slot);
double orc_draw_uniform(uint64_t seed, int64_t gid, int64_t tick, int slot);

/* ---- single env ---- */
void orc_init(OrcEnv *e, const MdgParams *P, const MdgReward *R, uint64_t seed, int64_t gid);
/* as orc_init, with injected constructor uniforms (see OrcEnv.ctor_u) */
void orc_init_inject(OrcEnv *e, const MdgParams *P, const MdgReward *R, uint64_t seed, int64_t gid,
                     const double *ctor_u, int n_ctor_u);
void orc_set_ctor_uniforms(OrcEnv *e, const double *u, int n);
void orc_tick(OrcEnv *e, const double *normals, const double *uniforms);
void orc_reset(OrcEnv *e, const double *normals, const double *uniforms, OrcStepOut *out);
void orc_step(OrcEnv *e, int mode, const double *units, int asset_idx, const double *normals,
              const double *uniforms, OrcStepOut *out);
void orc_shaper_feed(OrcEnv *e, const double *raw, int ra, const double *port, int done, OrcStepOut *out);
/* Portfolio primitives, exposed for the reference's ledger known-answer tests */
int orc_check_risk(const OrcEnv *e);
int orc_check_risk_asset(const OrcEnv *e, int i, double units);
void orc_handle_transaction(OrcEnv *e, int i, double tp, double units, double cost);
int orc_broker_transaction(OrcEnv *e, int i, double units, double *tp, double *tu, double *tc);
double orc_equity(const OrcEnv *e);
double orc_asset_value(const OrcEnv *e);
double orc_pnl(const OrcEnv *e);
double orc_balance(const OrcEnv *e);
double orc_available_margin(const OrcEnv *e);
double orc_used_margin(const OrcEnv *e);
double orc_borrowed_margin(const OrcEnv *e);
double orc_borrowed_asset_value(const OrcEnv *e);
void orc_ledger_normed(const OrcEnv *e, double *out);
void orc_ledger_normed_full(const OrcEnv *e, double *out);
void orc_ledger_abs_normed(const OrcEnv *e, double *out);
void orc_ledger_abs_normed_full(const OrcEnv *e, double *out);

/* ---- batch of independent envs, tensors laid out as the CUDA path ([rows][N]) ---- */
typedef struct OrcBatch OrcBatch;
OrcBatch *orc_batch_create(int64_t n, const MdgParams *P, const MdgReward *R, uint64_t seed,
                           int64_t env_offset);
void orc_batch_destroy(OrcBatch *b);
OrcEnv *orc_batch_env(OrcBatch *b, int64_t i);
void orc_batch_export_state(const OrcBatch *b, const MdgState *host_state);
/* same argument meaning as mdg_step / mdg_reset, but every pointer is a HOST pointer */
void orc_action_units(const OrcEnv *e, const int8_t *actions, int action_atoms, double unit_size, double *units);
void orc_weight_units(const OrcEnv *e, const float *weights, double *units);
void orc_batch_step(OrcBatch *b, const MdgStepIO *io, const MdgLaunch *launch, int threads);
void orc_batch_reset(OrcBatch *b, const MdgStepIO *io, const MdgLaunch *launch,
                     const uint8_t *mask, int fill_ticks, int clear_nstep, int threads);
void orc_batch_derived(const OrcBatch *b, const MdgDerived *out);

#ifdef __cplusplus
}
#endif
#endif
